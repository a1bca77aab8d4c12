// Tensor-core engine of the fused 2-layer MLP (bf16 storage, 64-wide sources):
//   gather prologue (cp.async, 128-byte-swizzled smem tiles)  ->  tcgen05.mma (M=128, N=64, K=64 per
//   source, fp32 accumulators in TMEM)  ->  epilogue from TMEM (bias, SiLU, second GEMM operand
//   written back to smem)  ->  tcgen05.mma  ->  epilogue (bias, LayerNorm, residual / edge update).
//
// One CTA = 256 threads = one 128-row tile at a time, persistent over tiles, 2 CTAs per SM.  In the
// epilogues warp w owns TMEM lanes 32*(w%4).. (rows) and columns 32*(w/4).. : two threads per row,
// 32 columns each (LayerNorm needs one exchange of partial sums between them).  The gather of tile
// t+1 is issued as soon as the first GEMM of tile t has drained the X buffer, so it overlaps both
// epilogues and the second GEMM.
//
// Same inputs / outputs / saved activations as the SIMT engine (mlp2.cu): the two are
// interchangeable per call and cross-checked in tests/test_gpu_kernels.py.
#include "common.cuh"
#include "tc_common.cuh"

namespace p4c {
namespace tc {

constexpr int TM = 128;
constexpr int NTHREADS = 256;
constexpr int HID = P4C_HIDDEN;
constexpr int XBLK = TM * 128;   // bytes of one 128x64 bf16 tile
constexpr int WBLK = HID * 128;  // bytes of one 64x64 bf16 tile
constexpr uint32_t TMEM_COLS = 128;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

struct FwdSmem {
  uint8_t X[P4C_MAX_SRC][XBLK];
  uint8_t A2[XBLK];
  uint8_t W1[P4C_MAX_SRC][WBLK];  // [n = hidden unit][k]  K-major, swizzled
  uint8_t W2[WBLK];               // [n = output unit][k]
  float b1[HID], b2[HID], gamma[HID], beta[HID];
  uint32_t row_off[2][P4C_MAX_SRC][TM];  // source row offset / 8 elements; ~0u = invalid row
  int32_t orow[2][TM];                   // row in the output-like arrays; -1 = invalid
  float2 ln_part[2][TM];  // per-row partial (sum, sum of squares) of each column half
  uint64_t mbar[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void row_info(const p4c_mlp_t& m, long long tile_row0, long long total_rows,
                                         uint32_t (*row_off)[TM], int32_t* orow, int tid) {
  if (tid >= TM) return;
  const long long g = tile_row0 + tid;
  long long o = -1;
  long long b = 0;
  int rr = 0;
  if (g < total_rows) {
    b = g / m.rows;
    rr = static_cast<int>(g - b * m.rows);
    o = b * (m.full_rows > 0 ? m.full_rows : m.rows) + rr;
  }
  orow[tid] = static_cast<int32_t>(o);
  for (int s = 0; s < m.nsrc; ++s) {
    uint32_t off = kNoRow;
    if (o >= 0) {
      const long long idx = m.src[s].index ? static_cast<long long>(m.src[s].index[rr]) : static_cast<long long>(rr);
      off = static_cast<uint32_t>((b * m.src[s].batch_stride + idx * m.src[s].ld) >> 3);
    }
    row_off[s][tid] = off;
  }
}

// 128 rows x 64 bf16 per source -> swizzled smem; 8 consecutive threads fetch one 128-byte row
__device__ __forceinline__ void issue_gather(const p4c_mlp_t& m, const uint32_t (*row_off)[TM], uint8_t (*X)[XBLK],
                                             int tid) {
  for (int s = 0; s < m.nsrc; ++s) {
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(m.src[s].ptr);
    const uint32_t xs = smem_u32(X[s]);
#pragma unroll
    for (int i = 0; i < (TM * 8) / NTHREADS; ++i) {
      const int idx = i * NTHREADS + tid;
      const int r = idx >> 3, j = idx & 7;
      const uint32_t off = row_off[s][r];
      const __nv_bfloat16* src = off != kNoRow ? base + (static_cast<long long>(off) << 3) + j * 8 : base;
      cp_async16(xs + sw128_offset(r, j), src, off != kNoRow ? 16u : 0u);
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 2)
mlp2_tc_fwd_kernel(const p4c_mlp_t m, const __nv_bfloat16* __restrict__ residual, __nv_bfloat16* __restrict__ out,
                   __nv_bfloat16* __restrict__ out2, __nv_bfloat16* __restrict__ save_h1,
                   float* __restrict__ save_stats) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(
      smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, hsel = warp >> 2;  // TMEM lane quadrant (rows) / column half
  const int row = quad * 32 + lane, col0 = hsel * 32;
  const long long total_rows = static_cast<long long>(m.rows) * m.batch;
  const long long ntiles = (total_rows + TM - 1) / TM;
  const int OUT = m.out_features;
  const int nsrc = m.nsrc;

  // ---- one-time set-up: weights -> bf16 swizzled tiles, vectors, barriers, TMEM
  for (int i = tid; i < HID * m.k1; i += NTHREADS) {
    const int c = i / m.k1, k = i - c * m.k1;
    const int kb = k >> 6, kk = k & 63;
    *reinterpret_cast<__nv_bfloat16*>(S.W1[kb] + sw128_offset(c, kk >> 3) + (kk & 7) * 2) =
        __float2bfloat16_rn(m.w1[i]);
  }
  for (int i = tid; i < HID * HID; i += NTHREADS) {
    const int c = i >> 6, k = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(S.W2 + sw128_offset(c, k >> 3) + (k & 7) * 2) =
        __float2bfloat16_rn(c < OUT ? m.w2[c * HID + k] : 0.f);
  }
  if (tid < HID) {
    S.b1[tid] = m.b1[tid];
    S.b2[tid] = tid < OUT ? m.b2[tid] : 0.f;
    S.gamma[tid] = m.gamma ? m.gamma[tid] : 1.f;
    S.beta[tid] = m.beta ? m.beta[tid] : 0.f;
  }
  if (tid == 0) {
    mbar_init(&S.mbar[0], 1);
    mbar_init(&S.mbar[1], 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&S.tmem_base, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t tmem_row = tmem + (static_cast<uint32_t>(quad * 32) << 16) + col0;
  constexpr uint32_t IDESC = make_idesc_bf16(TM, HID, 0, 0);

  long long tile = blockIdx.x;
  if (tile < ntiles) {
    row_info(m, tile * TM, total_rows, S.row_off[0], S.orow[0], tid);
    __syncthreads();
    issue_gather(m, S.row_off[0], S.X, tid);
  }
  cp_async_commit();

  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int p = it & 1;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();  // X(tile) complete and visible to the tensor core

    // ---- GEMM 1: D1[128x64] = X[128 x 64*nsrc] * W1^T
    if (tid == 0) {
      tc_fence_after();
      for (int s = 0; s < nsrc; ++s) {
        const uint32_t xa = smem_u32(S.X[s]), wa = smem_u32(S.W1[s]);
#pragma unroll
        for (int j = 0; j < 4; ++j)  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
          mma_bf16_ss(tmem, make_desc_sw128(xa + j * 32, 16, 1024), make_desc_sw128(wa + j * 32, 16, 1024),
                      IDESC, (s | j) ? 1u : 0u);
      }
      mma_commit(&S.mbar[0]);
    }
    // meanwhile: addressing of the next tile
    const long long next = tile + gridDim.x;
    if (next < ntiles) row_info(m, next * TM, total_rows, S.row_off[p ^ 1], S.orow[p ^ 1], tid);
    mbar_wait(&S.mbar[0], parity);
    tc_fence_after();
    __syncthreads();  // row_off[p^1] visible; X is free again
    if (next < ntiles) issue_gather(m, S.row_off[p ^ 1], S.X, tid);
    cp_async_commit();

    // ---- epilogue 1: h1 = D1 + b1 (saved), a1 = SiLU(h1) -> A2 (bf16, swizzled)
    const long long orow = S.orow[p][row];
    {
      float v[32];
      tmem_ld32(tmem_row, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c0 = col0 + j * 8;
        float h[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = v[j * 8 + q] + S.b1[c0 + q];
        uint4 hb;
        hb.x = pack_bf16x2(h[0], h[1]); hb.y = pack_bf16x2(h[2], h[3]);
        hb.z = pack_bf16x2(h[4], h[5]); hb.w = pack_bf16x2(h[6], h[7]);
        if (save_h1) {
          if (orow >= 0) *reinterpret_cast<uint4*>(save_h1 + orow * HID + c0) = hb;
          // backward recomputes from the stored (rounded) value
          float2 t;
          t = unpack_bf16x2(hb.x); h[0] = t.x; h[1] = t.y;
          t = unpack_bf16x2(hb.y); h[2] = t.x; h[3] = t.y;
          t = unpack_bf16x2(hb.z); h[4] = t.x; h[5] = t.y;
          t = unpack_bf16x2(hb.w); h[6] = t.x; h[7] = t.y;
        }
        uint4 ab;
        ab.x = pack_bf16x2(fast_silu(h[0]), fast_silu(h[1])); ab.y = pack_bf16x2(fast_silu(h[2]), fast_silu(h[3]));
        ab.z = pack_bf16x2(fast_silu(h[4]), fast_silu(h[5])); ab.w = pack_bf16x2(fast_silu(h[6]), fast_silu(h[7]));
        *reinterpret_cast<uint4*>(S.A2 + sw128_offset(row, c0 >> 3)) = ab;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- GEMM 2: D2[128x64] = A2[128x64] * W2^T
    if (tid == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(S.A2), wa = smem_u32(S.W2);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mma_bf16_ss(tmem + HID, make_desc_sw128(aa + j * 32, 16, 1024), make_desc_sw128(wa + j * 32, 16, 1024),
                    IDESC, j ? 1u : 0u);
      mma_commit(&S.mbar[1]);
    }
    mbar_wait(&S.mbar[1], parity);
    tc_fence_after();

    // ---- epilogue 2: y = D2 + b2, LayerNorm (two threads per row), residual, stores
    {
      float y[32];
      tmem_ld32(tmem_row + HID, y);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) y[q] += S.b2[col0 + q];
      if (m.layer_norm) {
        float ps = 0.f, pq = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) {
          ps += y[q];
          pq = fmaf(y[q], y[q], pq);
        }
        S.ln_part[hsel][row] = make_float2(ps, pq);
        __syncthreads();
        const float2 lo = S.ln_part[0][row], hi = S.ln_part[1][row];  // same order in both threads
        const float mean = (lo.x + hi.x) * (1.0f / HID);
        const float var = fmaxf((lo.y + hi.y) * (1.0f / HID) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + m.ln_eps);
#pragma unroll
        for (int q = 0; q < 32; ++q) y[q] = (y[q] - mean) * rstd * S.gamma[col0 + q] + S.beta[col0 + q];
        if (save_stats && orow >= 0 && hsel == 0) {
          save_stats[2 * orow] = mean;
          save_stats[2 * orow + 1] = rstd;
        }
      }
      if (orow >= 0) {
        if (OUT == HID) {
          const __nv_bfloat16* e_row =
              out2 ? reinterpret_cast<const __nv_bfloat16*>(m.src[0].ptr) +
                         (static_cast<long long>(S.row_off[p][0][row]) << 3) + col0
                   : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float* yy = y + j * 8;
            const int c0 = col0 + j * 8;
            if (out2) {
              const uint4 e = *reinterpret_cast<const uint4*>(e_row + j * 8);
              const float2 e0 = unpack_bf16x2(e.x), e1 = unpack_bf16x2(e.y), e2 = unpack_bf16x2(e.z),
                           e3 = unpack_bf16x2(e.w);
              uint4 o;
              o.x = pack_bf16x2(e0.x + yy[0], e0.y + yy[1]); o.y = pack_bf16x2(e1.x + yy[2], e1.y + yy[3]);
              o.z = pack_bf16x2(e2.x + yy[4], e2.y + yy[5]); o.w = pack_bf16x2(e3.x + yy[6], e3.y + yy[7]);
              *reinterpret_cast<uint4*>(out2 + orow * HID + c0) = o;
            }
            if (residual) {
              const uint4 e = *reinterpret_cast<const uint4*>(residual + orow * HID + c0);
              const float2 e0 = unpack_bf16x2(e.x), e1 = unpack_bf16x2(e.y), e2 = unpack_bf16x2(e.z),
                           e3 = unpack_bf16x2(e.w);
              yy[0] += e0.x; yy[1] += e0.y; yy[2] += e1.x; yy[3] += e1.y;
              yy[4] += e2.x; yy[5] += e2.y; yy[6] += e3.x; yy[7] += e3.y;
            }
            uint4 o;
            o.x = pack_bf16x2(yy[0], yy[1]); o.y = pack_bf16x2(yy[2], yy[3]);
            o.z = pack_bf16x2(yy[4], yy[5]); o.w = pack_bf16x2(yy[6], yy[7]);
            *reinterpret_cast<uint4*>(out + orow * HID + c0) = o;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int c = col0 + q;
            if (c < OUT) {
              float v = y[q];
              if (residual) v += __bfloat162float(residual[orow * OUT + c]);
              out[orow * OUT + c] = __float2bfloat16_rn(v);
            }
          }
        }
      }
    }
    tc_fence_before();
  }

  cp_async_wait<0>();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc

bool tc_forward_supported(const p4c_mlp_t* m) {
  if (m->dtype != P4C_BF16) return false;
  for (int s = 0; s < m->nsrc; ++s)
    if (m->src[s].width != P4C_HIDDEN || (m->src[s].ld % 8) != 0 || (m->src[s].batch_stride % 8) != 0 ||
        (reinterpret_cast<uintptr_t>(m->src[s].ptr) % 16) != 0)
      return false;
  return true;
}

int launch_mlp2_tc_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2, void* save_h1,
                           float* save_stats, cudaStream_t st) {
  using namespace tc;
  const long long total = static_cast<long long>(m->rows) * m->batch;
  const long long ntiles = (total + TM - 1) / TM;
  const int grid = static_cast<int>(ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs);
  const size_t smem = sizeof(FwdSmem) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp2_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mlp_forward(tc): cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    configured = true;
  }
  mlp2_tc_fwd_kernel<<<grid, NTHREADS, smem, st>>>(
      *m, static_cast<const __nv_bfloat16*>(residual), static_cast<__nv_bfloat16*>(out),
      static_cast<__nv_bfloat16*>(out2), static_cast<__nv_bfloat16*>(save_h1), save_stats);
  P4C_CHECK_LAUNCH("mlp_forward(tc)");
  return P4C_OK;
}

}  // namespace p4c
