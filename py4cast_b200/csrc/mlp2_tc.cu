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
  float2 ln_part[4][TM];  // per-row partial (sum, sum of squares) of each column group
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
  int tile_rows;  // valid rows per 128-row tile (16..128)
  long long* trace;  // developer aid (p4c_set_trace_buffer): SM clock at the phases of CTA 0's first tile
};

static long long* g_trace = nullptr;
#ifdef P4C_PHASE_TRACE  // make EXTRA=-DP4C_PHASE_TRACE
#define P4C_TRACE(i)                                   \
  do {                                                 \
    if (a.trace && blockIdx.x == 0 && tid == 0 && it == 0) a.trace[i] = clock64(); \
  } while (0)
#else
#define P4C_TRACE(i) do { (void)it; } while (0)
#endif

__device__ __forceinline__ void row_info(const FwdArgs& a, long long tile_row0, long long total_rows,
                                         FwdSmem& S, int p, int tid) {
  if (tid >= TM) return;
  const p4c_mlp_t& m = a.m;
  const long long g = tile_row0 + tid;
  long long o = -1;
  long long b = 0;
  int rr = 0;
  if (tid < a.tile_rows && g < total_rows) {
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
template <int NT, bool SEG>
__device__ __forceinline__ void issue_gather(const FwdArgs& a, FwdSmem& S, int p, int tid) {
  for (int k = 0; k < a.kb.nblk; ++k) {
    const p4c_src_t& src = a.m.src[a.kb.src[k]];
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(src.ptr);
    const uint32_t xs = smem_u32(S.X[k]);
    if (SEG && src.seg_rowptr) {
      // InteractionNet.aggregate fused into the prologue (fixed summation order)
#ifdef P4C_PHASE_TRACE
      long long* const seg_trace = (a.trace && blockIdx.x == 0 && tid == 0) ? a.trace : nullptr;
#else
      long long* const seg_trace = nullptr;
#endif
      seg_sum_tile<NT>(base, src.seg_perm, S.row_off[p][k], S.seg_hi[p], S.seg_base[p], S.seg_sc[p],
                       S.orow[p], S.X[k], reinterpret_cast<__nv_bfloat16*>(src.seg_out), tid, a.tile_rows, seg_trace);
      continue;
    }
    const int nch = a.kb.nchunks[k];
#pragma unroll
    for (int i = 0; i < (TM * 8) / NT; ++i) {
      const int idx = i * NT + tid;
      const int r = idx >> 3, j = idx & 7;
      if (j < nch) {
        const uint32_t off = S.row_off[p][k][r];
        const __nv_bfloat16* sp = off != kNoRow ? base + (static_cast<long long>(off) << 3) + j * 8 : base;
        cp_async16(xs + sw128_offset(r, j), sp, off != kNoRow ? 16u : 0u);
      }
    }
  }
}

__device__ __forceinline__ uint4 pack8f(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
  float2 t;
  t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}

// NT threads = NT/128 threads per row, CPT = 8192/NT columns each (256 -> 32, 512 -> 16).
// SEG: the instantiation that can sum segments in its prologue (it needs the registers of the
// 256-thread shape); plain / gathered sources run with 512 threads (twice the warps to hide latency).
template <int NT, bool SEG>
__global__ void __launch_bounds__(NT, 2) mlp2_tc_fwd_kernel(const __grid_constant__ FwdArgs a) {
  constexpr int CPT = TM * HID / NT;  // columns per thread
  constexpr int NCG = NT / TM;        // column groups (threads per row)
  extern __shared__ unsigned char smem_raw[];
  // 128-byte swizzle needs 1024-byte aligned tiles
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(
      smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const p4c_mlp_t& m = a.m;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // uniform for the compiler
  const int quad = warp & 3, cg = warp >> 2;  // TMEM lane quadrant (rows) / column group
  const int row = quad * 32 + lane, col0 = cg * CPT;
  const long long total_rows = static_cast<long long>(m.rows) * m.batch;
  const int TR = a.tile_rows;
  const long long ntiles = (total_rows + TR - 1) / TR;
  const int OUT = m.out_features;
  const int nblk = a.kb.nblk;
  __nv_bfloat16* const out = a.out;
  __nv_bfloat16* const out2 = a.out2;
  __nv_bfloat16* const save_h1 = a.save_h1;
  const __nv_bfloat16* const residual = a.residual;

  // ---- one-time set-up: weights (image: 16-byte async copies), barriers, TMEM
  {
    const int it = 0;
    P4C_TRACE(0);
  }
  if (a.zero_x) {
    uint4* xz = reinterpret_cast<uint4*>(S.X);
    for (int i = tid; i < nblk * XBLK / 16; i += NT) xz[i] = make_uint4(0, 0, 0, 0);
  }
  if (m.packed) {
    const uint8_t* img = reinterpret_cast<const uint8_t*>(m.packed);
    for (int i = tid; i < nblk * (WBLK / 16); i += NT)
      cp_async16(smem_u32(S.W1) + i * 16, img + PK_W1 + i * 16, 16u);
    for (int i = tid; i < WBLK / 16; i += NT) cp_async16(smem_u32(S.W2) + i * 16, img + PK_W2 + i * 16, 16u);
    if (tid < 64) cp_async16(smem_u32(S.vec) + tid * 16, img + PK_VEC + tid * 16, 16u);
  } else {
    for (int k = 0; k < nblk; ++k) {
      const int ks = a.kb.kstart[k], kv = a.kb.kvalid[k];
      for (int i = tid; i < HID * HID; i += NT) {
        const int c = i >> 6, kk = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(S.W1[k] + sw128_offset(c, kk >> 3) + (kk & 7) * 2) =
            __float2bfloat16_rn(kk < kv ? m.w1[static_cast<long long>(c) * m.k1 + ks + kk] : 0.f);
      }
    }
    for (int i = tid; i < HID * HID; i += NT) {
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
  const float* const b1 = S.vec[0] + col0;
  const float* const b2 = S.vec[1] + col0;
  const float* const gamma = S.vec[2] + col0;
  const float* const beta = S.vec[3] + col0;

  long long tile = blockIdx.x;
  {
    const int it = 0;
    P4C_TRACE(1);
  }
  pdl_launch_dependents();
  if (tile < ntiles) row_info(a, tile * TR, total_rows, S, 0, tid);  // graph constants only
  pdl_wait();  // the previous kernel's outputs (this kernel's sources) are complete and visible
  if (tile < ntiles) {
    __syncthreads();
    {
      const int it = 0;
      P4C_TRACE(2);
    }
    issue_gather<NT, SEG>(a, S, 0, tid);
  }
  cp_async_commit();
  {
    const int it = 0;
    P4C_TRACE(3);
  }

  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int p = it & 1;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    cp_async_wait<0>();
    fence_proxy_async();
    __syncthreads();  // X(tile) (and, first time, the weights) complete and visible to the tensor core
    P4C_TRACE(4);

    // ---- GEMM 1: D1[128x64] = X[128 x 64*nblk] * W1^T
    if (warp == 0) {
      tc_fence_after();
      for (int k = 0; k < nblk; ++k) {
        const uint32_t xa = smem_u32(S.X[k]), wa = smem_u32(S.W1[k]);
        const int nk = (a.kb.kvalid[k] + 15) >> 4;  // UMMA_K = 16 bf16 = 32 bytes along the swizzled row
        for (int j = 0; j < nk; ++j)
          mma_bf16_ss_elect(tmem, make_desc_sw128(xa + j * 32, 16, 1024), make_desc_sw128(wa + j * 32, 16, 1024),
                      IDESC, (k | j) ? 1u : 0u);
      }
      mma_commit_elect(&S.mbar[0]);
    }
    // meanwhile: addressing of the next tile
    const long long next = tile + gridDim.x;
    if (next < ntiles) row_info(a, next * TR, total_rows, S, p ^ 1, tid);
    mbar_wait(&S.mbar[0], parity);
    tc_fence_after();
    __syncthreads();  // row_off[p^1] visible; X is free again
    P4C_TRACE(5);
    if (next < ntiles) issue_gather<NT, SEG>(a, S, p ^ 1, tid);
    cp_async_commit();

    // ---- epilogue 1: h1 = D1 + b1 (saved), a1 = SiLU(h1) -> A2 (bf16, swizzled)
    const long long orow = S.orow[p][row];
    {
      float v[CPT];
      tmem_ld<CPT>(tmem_row, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CPT / 8; ++j) {
        float h[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = v[j * 8 + q] + b1[j * 8 + q];
        const uint4 hb = pack8f(h);
        if (save_h1) {
          if (orow >= 0) *reinterpret_cast<uint4*>(save_h1 + orow * HID + col0 + j * 8) = hb;
          unpack8f(hb, h);  // backward recomputes from the stored (rounded) value
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) h[q] = fast_silu(h[q]);
        *reinterpret_cast<uint4*>(S.A2 + sw128_offset(row, (col0 >> 3) + j)) = pack8f(h);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    P4C_TRACE(6);

    // ---- GEMM 2: D2[128x64] = A2[128x64] * W2^T
    if (warp == 0) {
      tc_fence_after();
      const uint32_t aa = smem_u32(S.A2), wa = smem_u32(S.W2);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mma_bf16_ss_elect(tmem + HID, make_desc_sw128(aa + j * 32, 16, 1024), make_desc_sw128(wa + j * 32, 16, 1024),
                    IDESC, j ? 1u : 0u);
      mma_commit_elect(&S.mbar[1]);
    }
    // residual / edge rows of this tile: fetched while GEMM 2 runs (consumed at the very end)
    uint4 res_pre[CPT / 8];
    if (OUT == HID && orow >= 0 && (residual || out2)) {
      const __nv_bfloat16* rp =
          out2 ? reinterpret_cast<const __nv_bfloat16*>(m.src[0].ptr) + (static_cast<long long>(S.row_off[p][0][row]) << 3) + col0
               : residual + orow * HID + col0;
#pragma unroll
      for (int j = 0; j < CPT / 8; ++j) res_pre[j] = *reinterpret_cast<const uint4*>(rp + j * 8);
    }
    mbar_wait(&S.mbar[1], parity);
    tc_fence_after();
    P4C_TRACE(7);

    // ---- epilogue 2: y = D2 + b2, LayerNorm (NCG threads per row), residual, stores
    {
      float y[CPT];
      tmem_ld<CPT>(tmem_row + HID, y);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < CPT; ++q) y[q] += b2[q];
      if (m.layer_norm) {
        float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent chains
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          ps[q & 3] += y[q];
          pq[q & 3] = fmaf(y[q], y[q], pq[q & 3]);
        }
        S.ln_part[cg][row] = make_float2((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]));
        __syncthreads();
        float sx = 0.f, sq = 0.f;
#pragma unroll
        for (int c = 0; c < NCG; ++c) {  // same order in every thread of the row
          const float2 t = S.ln_part[c][row];
          sx += t.x;
          sq += t.y;
        }
        const float mean = sx * (1.0f / HID);
        const float var = fmaxf(sq * (1.0f / HID) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + m.ln_eps);
        const float nm = -mean * rstd;
#pragma unroll
        for (int q = 0; q < CPT; ++q) y[q] = fmaf(fmaf(y[q], rstd, nm), gamma[q], beta[q]);
        if (a.save_stats && orow >= 0 && cg == 0) {
          *reinterpret_cast<float2*>(a.save_stats + 2 * orow) = make_float2(mean, rstd);
        }
      }
      if (OUT == HID) {
        if (orow >= 0 && (out || out2)) {
          // out2 = src0 row + y (edge update); out = y (+ residual).  One of residual / out2 per
          // call site (InteractionNet: edge MLP has out2, node MLP has residual).
#pragma unroll
          for (int j = 0; j < CPT / 8; ++j) {
            float yy[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) yy[q] = y[j * 8 + q];
            const int c0 = col0 + j * 8;
            if (out2) {
              float e[8];
              unpack8f(res_pre[j], e);
#pragma unroll
              for (int q = 0; q < 8; ++q) e[q] += yy[q];
              *reinterpret_cast<uint4*>(out2 + orow * HID + c0) = pack8f(e);
            }
            if (out) {
              if (residual) {
                float e[8];
                if (out2) unpack8f(*reinterpret_cast<const uint4*>(residual + orow * HID + c0), e);
                else unpack8f(res_pre[j], e);
#pragma unroll
                for (int q = 0; q < 8; ++q) yy[q] += e[q];
              }
              *reinterpret_cast<uint4*>(out + orow * HID + c0) = pack8f(yy);
            }
          }
        }
        if (a.aggr_out) {
          // fused aggregation: sum the messages of `agg_group` consecutive rows (= lanes) with a
          // transposing butterfly: at every step a lane hands one half of its columns to its
          // partner and adds the partner's other half, so after log2(group) steps each lane holds
          // CPT/group columns summed over the group (fp32, fixed order; invalid rows add zero).
          const int gsz = a.agg_group;
          if (orow < 0) {
#pragma unroll
            for (int q = 0; q < CPT; ++q) y[q] = 0.f;
          }
          int cbase = 0;  // first column (relative to col0) this lane ends up with
#define P4C_AGG_STEP(OFF, N)                                                     \
          if ((N) >= 2 && gsz > OFF) {                                           \
            const bool up = (lane & OFF) != 0;                                   \
            _Pragma("unroll") for (int i = 0; i < (N) / 2; ++i) {                \
              const float send = up ? y[i] : y[i + (N) / 2];                     \
              const float keep = up ? y[i + (N) / 2] : y[i];                     \
              y[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);             \
            }                                                                    \
            if (up) cbase += (N) / 2;                                            \
          }
          P4C_AGG_STEP(1, CPT)
          P4C_AGG_STEP(2, CPT / 2)
          P4C_AGG_STEP(4, CPT / 4)
          P4C_AGG_STEP(8, CPT / 8)
#undef P4C_AGG_STEP
          if (orow >= 0) {
            const int ncol = CPT / gsz;
            __nv_bfloat16* dst = a.aggr_out + (orow / gsz) * HID + col0 + cbase;
            const float sc = a.agg_scale;
#pragma unroll
            for (int q = 0; q < CPT; ++q) y[q] *= sc;
            if (ncol >= 8) {
#pragma unroll
              for (int j = 0; j < CPT / 8; ++j)
                if (j < ncol / 8) *reinterpret_cast<uint4*>(dst + j * 8) = pack8f(y + j * 8);
            } else if (ncol == 4) {
              uint2 o;
              o.x = pack_bf16x2(y[0], y[1]); o.y = pack_bf16x2(y[2], y[3]);
              *reinterpret_cast<uint2*>(dst) = o;
            } else {
              *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(y[0], y[1]);
            }
          }
        }
      } else if (orow >= 0) {
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
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
    P4C_TRACE(8);
  }

  cp_async_wait<0>();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
  {
    const int it = 0;
    P4C_TRACE(9);
  }
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

void tc_set_trace(long long* buf) { tc::g_trace = buf; }
long long* tc_get_trace() { return tc::g_trace; }

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
  if (ws_forward_supported(m, residual, out2, aggr_out, agg_group) &&
      (m->out_features == P4C_HIDDEN || (out && reinterpret_cast<uintptr_t>(out) % 16 == 0)))
    return launch_mlp2_ws_forward(m, residual, out, out2, save_h1, save_stats, aggr_out, agg_group, agg_scale, st);
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
  a.tile_rows = m->tile_rows > 0 ? m->tile_rows : TM;
  a.trace = g_trace;
  if (aggr_out && a.tile_rows != TM) {
    set_error("mlp_forward_aggregate: tile_rows must be 128 (or 0)");
    return P4C_EINVAL;
  }
  const long long total = static_cast<long long>(m->rows) * m->batch;
  const long long ntiles = (total + a.tile_rows - 1) / a.tile_rows;
  const int grid = static_cast<int>(ntiles < tc_max_ctas() ? ntiles : tc_max_ctas());
  const size_t smem = sizeof(FwdSmem) + 1024;
  bool has_seg = false;
  for (int s = 0; s < m->nsrc; ++s) has_seg |= m->src[s].seg_rowptr != nullptr;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp2_tc_fwd_kernel<512, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp2_tc_fwd_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mlp_forward(tc): cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    configured = true;
  }
  if (has_seg)
    launch_pdl(mlp2_tc_fwd_kernel<256, true>, dim3(grid), dim3(256), smem, st, a);
  else
    launch_pdl(mlp2_tc_fwd_kernel<512, false>, dim3(grid), dim3(512), smem, st, a);
  P4C_CHECK_LAUNCH("mlp_forward(tc)");
  return P4C_OK;
}

}  // namespace p4c
