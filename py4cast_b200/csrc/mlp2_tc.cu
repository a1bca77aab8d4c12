// Tensor-core engine of the fused 2-layer MLP — FORWARD (bf16 storage):
//   gather prologue (cp.async into 128-byte-swizzled smem tiles; or an in-prologue segment sum)
//   -> tcgen05.mma (M=128, N=64, K=64 per K-block, fp32 accumulators in TMEM)
//   -> epilogue from TMEM (bias, SiLU, second GEMM operand written back to smem)
//   -> tcgen05.mma -> epilogue (bias, LayerNorm, residual / edge update, optional fused
//      aggregation over groups of consecutive rows).
//
// One CTA = 256 threads = one 128-row tile at a time, persistent over tiles, 2 CTAs per SM.  In the
// epilogues warp w owns TMEM lanes 32*(w%4).. (rows) and columns 32*(w/4).. : two threads per row,
// 32 columns each (LayerNorm needs one exchange of partial sums between them).  The gather of tile
// t+1 is issued as soon as the first GEMM of tile t has drained the X buffer, so it overlaps both
// epilogues and the second GEMM.
//
// Layer-1 input = up to 3 K-blocks of 64 columns (tc_common.cuh KBlocks): 64-wide sources (plain,
// gathered through an index, or segment sums of another tensor) and dense sources of any width whose
// row stride is a multiple of 8 elements (the padding must be finite: it is multiplied by zero
// weights).  Weights come from a pre-packed bf16 image (p4c_pack_weights) with 16-byte copies, or
// are converted in the kernel when no image is given.
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
  float vec[4][HID];              // b1, b2, gamma, beta (image order)
  uint32_t row_off[2][P4C_MAX_SRC][TM];  // block row offset / 8 elements (segment source: first
                                         // edge of the segment); ~0u = invalid row
  uint32_t seg_hi[2][TM];                // segment source: one past the last edge
  uint32_t seg_base[2][TM];              // segment source: sample offset / 8 elements
  float seg_sc[2][TM];                   // segment source: row scale (1/deg for mean aggregation)
  int32_t orow[2][TM];                   // row in the output-like arrays; -1 = invalid
  float2 ln_part[2][TM];  // per-row partial (sum, sum of squares) of each column half
  uint64_t mbar[2];
  uint32_t tmem_base;
};

struct FwdArgs {
  p4c_mlp_t m;
  KBlocks kb;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  __nv_bfloat16* save_h1;
  float* save_stats;
  // fused aggregation (optional): rows come in groups of `agg_group` consecutive rows (a power of
  // two <= 32 dividing the sample's row count) whose y are summed into aggr_out[row / agg_group]
  __nv_bfloat16* aggr_out;
  int agg_group;
  float agg_scale;
  int zero_x;  // some block fetches fewer than 8 chunks per row: clear the tiles once
};

__device__ __forceinline__ void row_info(const FwdArgs& a, long long tile_row0, long long total_rows,
                                         FwdSmem& S, int p, int tid) {
  if (tid >= TM) return;
  const p4c_mlp_t& m = a.m;
  const long long g = tile_row0 + tid;
  long long o = -1;
  long long b = 0;
  int rr = 0;
  if (g < total_rows) {
    b = g / m.rows;
    rr = static_cast<int>(g - b * m.rows);
    o = b * (m.full_rows > 0 ? m.full_rows : m.rows) + rr;
  }
  S.orow[p][tid] = static_cast<int32_t>(o);
  for (int k = 0; k < a.kb.nblk; ++k) {
    const p4c_src_t& src = m.src[a.kb.src[k]];
    uint32_t off = kNoRow;
    if (src.seg_rowptr) {
      uint32_t hi = 0, sbase = 0;
      float sc = 1.f;
      if (o >= 0) {
        off = static_cast<uint32_t>(src.seg_rowptr[rr]);
        hi = static_cast<uint32_t>(src.seg_rowptr[rr + 1]);
        sbase = static_cast<uint32_t>((b * src.batch_stride) >> 3);
        if (src.seg_scale) sc = src.seg_scale[rr];
      }
      S.seg_hi[p][tid] = hi;
      S.seg_base[p][tid] = sbase;
      S.seg_sc[p][tid] = sc;
    } else if (o >= 0) {
      const long long idx = src.index ? static_cast<long long>(src.index[rr]) : static_cast<long long>(rr);
      off = static_cast<uint32_t>((b * src.batch_stride + idx * src.ld + a.kb.col0[k]) >> 3);
    }
    S.row_off[p][k][tid] = off;
  }
}

// 128 rows x 64 bf16 per block -> swizzled smem; 8 consecutive threads handle one 128-byte row
__device__ __forceinline__ void issue_gather(const FwdArgs& a, FwdSmem& S, int p, int tid) {
  for (int k = 0; k < a.kb.nblk; ++k) {
    const p4c_src_t& src = a.m.src[a.kb.src[k]];
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(src.ptr);
    const uint32_t xs = smem_u32(S.X[k]);
    if (src.seg_rowptr) {
      // InteractionNet.aggregate fused into the prologue (fixed summation order)
      seg_sum_tile<NTHREADS>(base, src.seg_perm, S.row_off[p][k], S.seg_hi[p], S.seg_base[p], S.seg_sc[p],
                             S.orow[p], S.X[k], reinterpret_cast<__nv_bfloat16*>(src.seg_out), tid);
      continue;
    }
    const int nch = a.kb.nchunks[k];
#pragma unroll
    for (int i = 0; i < (TM * 8) / NTHREADS; ++i) {
      const int idx = i * NTHREADS + tid;
      const int r = idx >> 3, j = idx & 7;
      if (j < nch) {
        const uint32_t off = S.row_off[p][k][r];
        const __nv_bfloat16* sp = off != kNoRow ? base + (static_cast<long long>(off) << 3) + j * 8 : base;
        cp_async16(xs + sw128_offset(r, j), sp, off != kNoRow ? 16u : 0u);
      }
    }
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) mlp2_tc_fwd_kernel(const __grid_constant__ FwdArgs a) {
  extern __shared__ unsigned char smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(
      smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const p4c_mlp_t& m = a.m;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, hsel = warp >> 2;  // TMEM lane quadrant (rows) / column half
  const int row = quad * 32 + lane, col0 = hsel * 32;
  const long long total_rows = static_cast<long long>(m.rows) * m.batch;
  const long long ntiles = (total_rows + TM - 1) / TM;
  const int OUT = m.out_features;
  const int nblk = a.kb.nblk;
  __nv_bfloat16* const out = a.out;
  __nv_bfloat16* const out2 = a.out2;
  __nv_bfloat16* const save_h1 = a.save_h1;
  const __nv_bfloat16* const residual = a.residual;

  // ---- one-time set-up: weights (image: 16-byte async copies), barriers, TMEM
  if (a.zero_x) {
    uint4* xz = reinterpret_cast<uint4*>(S.X);
    for (int i = tid; i < nblk * XBLK / 16; i += NTHREADS) xz[i] = make_uint4(0, 0, 0, 0);
  }
  if (m.packed) {
    const uint8_t* img = reinterpret_cast<const uint8_t*>(m.packed);
    for (int i = tid; i < nblk * (WBLK / 16); i += NTHREADS)
      cp_async16(smem_u32(S.W1) + i * 16, img + PK_W1 + i * 16, 16u);
    for (int i = tid; i < WBLK / 16; i += NTHREADS) cp_async16(smem_u32(S.W2) + i * 16, img + PK_W2 + i * 16, 16u);
    if (tid < 64) cp_async16(smem_u32(S.vec) + tid * 16, img + PK_VEC + tid * 16, 16u);
  } else {
    for (int k = 0; k < nblk; ++k) {
      const int ks = a.kb.kstart[k], kv = a.kb.kvalid[k];
      for (int i = tid; i < HID * HID; i += NTHREADS) {
        const int c = i >> 6, kk = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(S.W1[k] + sw128_offset(c, kk >> 3) + (kk & 7) * 2) =
            __float2bfloat16_rn(kk < kv ? m.w1[static_cast<long long>(c) * m.k1 + ks + kk] : 0.f);
      }
    }
    for (int i = tid; i < HID * HID; i += NTHREADS) {
      const int c = i >> 6, k = i & 63;
      *reinterpret_cast<__nv_bfloat16*>(S.W2 + sw128_offset(c, k >> 3) + (k & 7) * 2) =
          __float2bfloat16_rn(c < OUT ? m.w2[c * HID + k] : 0.f);
    }
    if (tid < HID) {
      S.vec[0][tid] = m.b1[tid];
      S.vec[1][tid] = tid < OUT ? m.b2[tid] : 0.f;
      S.vec[2][tid] = m.gamma ? m.gamma[tid] : 1.f;
      S.vec[3][tid] = m.beta ? m.beta[tid] : 0.f;
    }
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
  const float* const b1 = S.vec[0];
  const float* const b2 = S.vec[1];
  const float* const gamma = S.vec[2];
  const float* const beta = S.vec[3];

  long long tile = blockIdx.x;
  if (tile < ntiles) {
    row_info(a, tile * TM, total_rows, S, 0, tid);
    __syncthreads();
    issue_gather(a, S, 0, tid);
  }
  cp_async_commit();

  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int p = it & 1;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();  // X(tile) (and, first time, the weights) complete and visible to the tensor core

    // ---- GEMM 1: D1[128x64] = X[128 x 64*nblk] * W1^T
    if (tid == 0) {
      tc_fence_after();
      for (int k = 0; k < nblk; ++k) {
        const uint32_t xa = smem_u32(S.X[k]), wa = smem_u32(S.W1[k]);
        const int nk = (a.kb.kvalid[k] + 15) >> 4;  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
        for (int j = 0; j < nk; ++j)
          mma_bf16_ss(tmem, make_desc_sw128(xa + j * 32, 16, 1024), make_desc_sw128(wa + j * 32, 16, 1024),
                      IDESC, (k | j) ? 1u : 0u);
      }
      mma_commit(&S.mbar[0]);
    }
    // meanwhile: addressing of the next tile
    const long long next = tile + gridDim.x;
    if (next < ntiles) row_info(a, next * TM, total_rows, S, p ^ 1, tid);
    mbar_wait(&S.mbar[0], parity);
    tc_fence_after();
    __syncthreads();  // row_off[p^1] visible; X is free again
    if (next < ntiles) issue_gather(a, S, p ^ 1, tid);
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
        for (int q = 0; q < 8; ++q) h[q] = v[j * 8 + q] + b1[c0 + q];
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
      for (int q = 0; q < 32; ++q) y[q] += b2[col0 + q];
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
        for (int q = 0; q < 32; ++q) y[q] = (y[q] - mean) * rstd * gamma[col0 + q] + beta[col0 + q];
        if (a.save_stats && orow >= 0 && hsel == 0) {
          a.save_stats[2 * orow] = mean;
          a.save_stats[2 * orow + 1] = rstd;
        }
      }
      if (OUT == HID) {
        if (orow >= 0 && (out || out2)) {
          const __nv_bfloat16* e_row =
              out2 ? reinterpret_cast<const __nv_bfloat16*>(m.src[0].ptr) +
                         (static_cast<long long>(S.row_off[p][0][row]) << 3) + col0
                   : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float yy[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) yy[q] = y[j * 8 + q];
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
            if (out) {
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
          }
        }
        if (a.aggr_out) {
          // fused aggregation: sum the messages of `agg_group` consecutive rows (= lanes) with a
          // fixed butterfly (fp32, before any rounding); invalid rows contribute zero
          const int gsz = a.agg_group;
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            float v = orow >= 0 ? y[q] : 0.f;
            for (int off = 1; off < gsz; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            y[q] = v * a.agg_scale;
          }
          if (orow >= 0 && (lane & (gsz - 1)) == 0) {
            __nv_bfloat16* dst = a.aggr_out + (orow / gsz) * HID + col0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 o;
              o.x = pack_bf16x2(y[j * 8 + 0], y[j * 8 + 1]); o.y = pack_bf16x2(y[j * 8 + 2], y[j * 8 + 3]);
              o.z = pack_bf16x2(y[j * 8 + 4], y[j * 8 + 5]); o.w = pack_bf16x2(y[j * 8 + 6], y[j * 8 + 7]);
              *reinterpret_cast<uint4*>(dst + j * 8) = o;
            }
          }
        }
      } else if (orow >= 0) {
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
    tc_fence_before();
  }

  cp_async_wait<0>();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc

static bool tc_src_ok(const p4c_src_t& s) {
  if ((s.ld % 8) != 0 || (s.batch_stride % 8) != 0 || (reinterpret_cast<uintptr_t>(s.ptr) % 16) != 0) return false;
  if ((s.index || s.seg_rowptr) && s.width != P4C_HIDDEN) return false;
  if (s.seg_rowptr && (s.index || s.ld != P4C_HIDDEN)) return false;
  if (s.seg_out && (reinterpret_cast<uintptr_t>(s.seg_out) % 16) != 0) return false;
  return true;
}

bool tc_kblocks(const p4c_mlp_t* m, tc::KBlocks* kb) {
  int widths[P4C_MAX_SRC] = {0, 0, 0};
  for (int s = 0; s < m->nsrc; ++s) widths[s] = m->src[s].width;
  return tc::make_kblocks(m->nsrc, widths, kb);
}

bool tc_forward_supported(const p4c_mlp_t* m) {
  if (m->dtype != P4C_BF16) return false;
  tc::KBlocks kb;
  if (!tc_kblocks(m, &kb)) return false;
  int nseg = 0;
  for (int s = 0; s < m->nsrc; ++s) {
    if (!tc_src_ok(m->src[s])) return false;
    nseg += m->src[s].seg_rowptr != nullptr;
  }
  if (nseg > 1) return false;
  if (m->packed && (reinterpret_cast<uintptr_t>(m->packed) % 16) != 0) return false;
  return true;
}

int launch_mlp2_tc_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2, void* save_h1,
                           float* save_stats, void* aggr_out, int agg_group, float agg_scale, cudaStream_t st) {
  using namespace tc;
  FwdArgs a;
  a.m = *m;
  tc_kblocks(m, &a.kb);
  a.residual = static_cast<const __nv_bfloat16*>(residual);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.out2 = static_cast<__nv_bfloat16*>(out2);
  a.save_h1 = static_cast<__nv_bfloat16*>(save_h1);
  a.save_stats = save_stats;
  a.aggr_out = static_cast<__nv_bfloat16*>(aggr_out);
  a.agg_group = agg_group;
  a.agg_scale = agg_scale;
  a.zero_x = 0;
  for (int k = 0; k < a.kb.nblk; ++k) a.zero_x |= a.kb.nchunks[k] < 8;
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
  mlp2_tc_fwd_kernel<<<grid, NTHREADS, smem, st>>>(a);
  P4C_CHECK_LAUNCH("mlp_forward(tc)");
  return P4C_OK;
}

}  // namespace p4c
