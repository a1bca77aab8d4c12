// Tensor-core engine of the fused 2-layer MLP — BACKWARD (bf16 storage, 64-wide sources).
//
// Per 128-row tile (TPR threads per row, TPR = 2 or 4: warp w owns TMEM lanes 32*(w%4).. = rows and the
// 64/TPR columns starting at (64/TPR)*(w/4); TPR = 2 is the default, the 4-threads-per-row shape runs
// 32 warps per SM at 64 registers and measured slower), five groups of tcgen05 MMAs with fp32 accumulators in
// TMEM (256 columns per CTA: 192 persistent weight-gradient columns + one 64-column scratch):
//   G1  h2   = a1 · W2^T                      (recomputed; LayerNorm input)      -> scratch
//   G2  da1  = dh2 · W2                                                           -> scratch
//   G3  dW2 += a1^T · dh2        (A, B MN-major: the [row][col] tiles read "transposed")
//   G4  dW1 += X^T · dh1         (two M=128 accumulators: sources 0-1, source 2)
//   G5  dX_s = dh1 · W1[:, s]    (per identity-indexed source)                    -> scratch
// The weight-gradient accumulators stay in TMEM over all tiles of the CTA and are written once,
// as per-CTA partials that mlp2_wgrad_reduce_kernel sums in a fixed order (no float atomics).
// An MN-major M=128 operand spans two consecutive 16 KB smem blocks; when only 64 rows of the
// product are wanted (source 2, a1) the second block is whatever tile follows in shared memory
// and the corresponding 64 TMEM lanes are simply never read.
// Bias / LayerNorm-affine gradients are column sums over rows = over threads: a 31-shuffle
// "transpose-reduce" butterfly per 32 columns, in a fixed order.
#include "common.cuh"
#include "tc_common.cuh"

namespace p4c {
namespace tc {

constexpr int TM = 128;
constexpr int HID = P4C_HIDDEN;
constexpr int BLK = TM * 128;    // 16 KB: one 128x64 bf16 tile
constexpr int WBLK = HID * 128;  // 8 KB: one 64x64 bf16 tile
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t COL_DW1A = 0, COL_DW1B = 64, COL_DW2 = 128, COL_SCRATCH = 192;

// small control block placed after the tiles (kept tiny: two CTAs must fit in one SM)
struct BwdCtl {
  float b2[HID], gamma[HID];
  uint32_t row_off[P4C_MAX_SRC][TM];     // source row offset in units of 8 elements; ~0u = invalid row
  int32_t orow[2][TM];                   // row in the output-like arrays; -1 = invalid row
  float2 ln_part[2][TM];                 // LayerNorm-backward row sums (exchange between the threads of a row)
  uint64_t mbar[4];
  uint32_t tmem_base;
};
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

// bf16x2 -> two floats with integer ops (exact)
__device__ __forceinline__ void unpack2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xFFFF0000u);
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  unpack2(u.x, f[0], f[1]); unpack2(u.y, f[2], f[3]);
  unpack2(u.z, f[4], f[5]); unpack2(u.w, f[6], f[7]);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ---- packed bf16x2 activation math (two elements per instruction, one MUFU op per pair) -------------
// P4C_BWD_PACKED_DH1 (default on): dh1 = da1 * SiLU'(h1) as one bf16x2 product per pair (da1 is
// rounded to bf16 first; the result is bf16 storage anyway).
// P4C_BWD_PACKED_SILU (default OFF): SiLU and SiLU' of the saved pre-activation evaluated in bf16x2
// as well.  Measured on B200: another 0.8 % of the step, but the extra rounding of the recomputed a1
// raises the worst parameter-gradient error of a model rollout from 1.3e-2 to 1.6e-2 (bar: 2e-2),
// so it stays off.
#ifndef P4C_BWD_PACKED_DH1
#define P4C_BWD_PACKED_DH1 1
#endif
#ifndef P4C_BWD_PACKED_SILU
#define P4C_BWD_PACKED_SILU 0
#endif
// h (bf16x2) -> a = h * sigmoid(h), g = sigmoid(h) * (1 + h * (1 - sigmoid(h)))
__device__ __forceinline__ void bf2_silu_and_grad(uint32_t h, uint32_t& a, uint32_t& g) {
  constexpr uint32_t HALF = 0x3F003F00u, NEG1 = 0xBF80BF80u;  // (0.5, 0.5), (-1, -1)
  const uint32_t sg = bf2_fma(bf2_tanh(bf2_mul(h, HALF)), HALF, HALF);
  a = bf2_mul(h, sg);
  const uint32_t u = bf2_fma(a, NEG1, h);  // h - h * sg
  g = bf2_fma(sg, u, sg);
}
__device__ __forceinline__ uint32_t bf2_from_floats(float lo, float hi) { return pack_bf16x2(lo, hi); }

struct BwdArgs {
  p4c_mlp_t m;
  p4c_mlp_bwd_t g;
  KBlocks kb;
  int ndx;         // K-blocks whose input gradient is produced (<= 2)
  int dx_blk[2];   // their block ids
  int zero_x;      // some block fetches fewer than 8 chunks per row: clear the tiles once
  // where a row of dy_a lives: 0 = generic (index / scale looked up per tile), 1 = same gather
  // index and batch stride as K-block `dy_blk` (its row offset is reused: no index round trip),
  // 2 = identity rows laid out like the outputs (dy_a + orow * out_features)
  int dy_mode, dy_blk;
  int ablate;        // developer aid (P4C_BWD_ABLATE, timing experiments only: results are WRONG when non-zero)
  long long* trace;  // developer aid (p4c_set_trace_buffer): SM clock at the phases of CTA 0's first tile, slots 16..31
};
#ifdef P4C_PHASE_TRACE  // make EXTRA=-DP4C_PHASE_TRACE
#define P4C_BWD_TRACE(i)                                                             \
  do {                                                                               \
    if (args.trace && blockIdx.x == 0 && tid == 0 && (it == 0 || it == 2))           \
      args.trace[16 + (it ? 16 : 0) + (i)] = clock64();                              \
  } while (0)
#else
#define P4C_BWD_TRACE(i) do { (void)it; } while (0)
#endif

// Addressing of one tile row (thread tid < TM owns row tid): output row and, per K-block, the source
// row.  The gather-index loads are issued here (bwd_row_addr) and their values are first touched
// when the offsets are computed and stored (bwd_row_store, a phase later), so that the load
// latency hides behind whatever runs in between instead of stalling the top of the tile.
struct RowAddr {
  int32_t orow;
  int32_t b;                   // sample of the row
  int32_t idx[P4C_MAX_SRC];    // source row per K-block (gathered or identity)
};
// Tile order: tile T = (row block T / batch, sample T % batch): the samples of one row block run
// back to back on neighbouring CTAs, so a K-block shared by the batch (static edge embeddings,
// batch_stride 0: 168 MB for m2g, more than L2) comes from HBM once per step instead of once per sample.
__device__ __forceinline__ RowAddr bwd_row_addr(const p4c_mlp_t& m, const KBlocks& kb, long long tile, int tid) {
  RowAddr a;
  a.orow = -1;
  a.b = 0;
#pragma unroll
  for (int k = 0; k < P4C_MAX_SRC; ++k) a.idx[k] = 0;
  const uint32_t blk = static_cast<uint32_t>(tile) / static_cast<uint32_t>(m.batch);
  const int rr = static_cast<int>(blk) * TM + tid;
  if (tid < TM && rr < m.rows) {
    const int b = static_cast<int>(static_cast<uint32_t>(tile) - blk * static_cast<uint32_t>(m.batch));
    a.b = b;
    a.orow = static_cast<int32_t>(static_cast<long long>(b) * (m.full_rows > 0 ? m.full_rows : m.rows) + rr);
#pragma unroll
    for (int k = 0; k < P4C_MAX_SRC; ++k) {
      if (k < kb.nblk) {
        const int32_t* ix = m.src[kb.src[k]].index;
        a.idx[k] = ix ? __ldg(ix + rr) : rr;
      }
    }
  }
  return a;
}
__device__ __forceinline__ void bwd_row_store(const RowAddr& a, const p4c_mlp_t& m, const KBlocks& kb,
                                              uint32_t (*row_off)[TM], int tid) {
  if (tid >= TM) return;
#pragma unroll
  for (int k = 0; k < P4C_MAX_SRC; ++k)
    if (k < kb.nblk) {
      const p4c_src_t& src = m.src[kb.src[k]];
      row_off[k][tid] = a.orow < 0 ? kNoRow
                                   : static_cast<uint32_t>((static_cast<long long>(a.b) * src.batch_stride +
                                                            static_cast<long long>(a.idx[k]) * src.ld + kb.col0[k]) >> 3);
    }
}

// next tile's saved pre-activation rows -> the (free) A1 tile, 16-byte chunks, coalesced
template <int NT>
__device__ __forceinline__ void bwd_h1_prefetch(const __nv_bfloat16* h1p, const int32_t* orow, uint8_t* A1, int tid) {
  const uint32_t a1 = smem_u32(A1);
#pragma unroll
  for (int i = 0; i < (TM * 8) / NT; ++i) {
    const int idx = i * NT + tid;
    const int r = idx >> 3, j = idx & 7;
    const long long o = orow[r];
    cp_async16(a1 + sw128_offset(r, j), o >= 0 ? h1p + o * HID + j * 8 : h1p, o >= 0 ? 16u : 0u);
  }
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" : : "l"(p)); }
template <int NT>
__device__ __forceinline__ void bwd_gather(const p4c_mlp_t& m, const KBlocks& kb, const uint32_t (*row_off)[TM],
                                           uint8_t* xbase, int tid) {
  for (int k = 0; k < kb.nblk; ++k) {
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(m.src[kb.src[k]].ptr);
    const uint32_t xs = smem_u32(xbase + k * BLK);
    const int nch = kb.nchunks[k];
#pragma unroll
    for (int i = 0; i < (TM * 8) / NT; ++i) {
      const int idx = i * NT + tid;
      const int r = idx >> 3, j = idx & 7;
      if (j < nch) {
        const uint32_t off = row_off[k][r];
        const __nv_bfloat16* src = off != kNoRow ? base + (static_cast<long long>(off) << 3) + j * 8 : base;
        cp_async16(xs + sw128_offset(r, j), src, off != kNoRow ? 16u : 0u);
      }
    }
  }
}

// shared-memory map (all tiles 1024-byte aligned):
//   [X0 .. X(nblk-1)] [A1] [DH] [W2] [W2T] [W1T x ndx] [ONES] [BwdCtl]
// ONES: two 8-row swizzle atoms of bf16 1.0 -- the second 64-row half of an MN-major M=128 A operand
// is pointed at it (leading-dimension offset of the descriptor), so that lanes 64..127 of the dW2 /
// dW1B accumulators hold the column sums of dh2 / dh1 = the bias gradients, for free.
constexpr int ONES_BYTES = 2048;
__host__ __device__ inline size_t bwd_smem_bytes(int nblk, int ndx) {
  return static_cast<size_t>(nblk + 2) * BLK + static_cast<size_t>(2 + ndx) * WBLK + ONES_BYTES + sizeof(BwdCtl);
}

template <int TPR>
__global__ void __launch_bounds__(TM * TPR, 2) mlp2_tc_bwd_kernel(const __grid_constant__ BwdArgs args) {
  constexpr int NTHREADS = TM * TPR;  // TPR threads per tile row
  constexpr int CPT = HID / TPR;      // columns per thread
  constexpr int NJ = CPT / 8;         // 16-byte chunks per thread
  extern __shared__ unsigned char smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const p4c_mlp_t& m = args.m;
  const p4c_mlp_bwd_t& g = args.g;
  const KBlocks& kb = args.kb;
  const int ndx = args.ndx;
  const int nblk = kb.nblk;
  uint8_t* Xs = base;
  uint8_t* A1 = base + nblk * BLK;
  uint8_t* DH = A1 + BLK;
  uint8_t* W2s = DH + BLK;
  uint8_t* W2T = W2s + WBLK;
  uint8_t* W1T = W2T + WBLK;
  uint8_t* ONES = W1T + ndx * WBLK;
  BwdCtl& C = *reinterpret_cast<BwdCtl*>(ONES + ONES_BYTES);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // uniform for the compiler
  const int quad = warp & 3, hsel = warp >> 2;  // TMEM lane quadrant (rows) / column group
  const int row = quad * 32 + lane, col0 = hsel * CPT;
  const int quad_bar = 1 + quad;  // named barrier of the TPR warps that share these rows
  const long long ntiles = static_cast<long long>((m.rows + TM - 1) / TM) * m.batch;  // row blocks x samples
  const int OUT = m.out_features;
  const bool ln = m.layer_norm != 0;
  const bool need_w = g.partials_mode != 0;  // deferred reduction: every weight gradient is produced
  const bool need_w1 = need_w || g.g_w1 != nullptr, need_w2 = need_w || g.g_w2 != nullptr;

  // ---- one-time set-up
  {
    const int it = 0;
    P4C_BWD_TRACE(0);
  }
  if (args.zero_x) {
    uint4* xz = reinterpret_cast<uint4*>(Xs);
    for (int i = tid; i < nblk * BLK / 16; i += NTHREADS) xz[i] = make_uint4(0, 0, 0, 0);
  }
  if (m.packed) {
    const uint8_t* img = reinterpret_cast<const uint8_t*>(m.packed);
    for (int i = tid; i < 2 * (WBLK / 16); i += NTHREADS)  // W2, W2T are adjacent in the image and in smem
      cp_async16(smem_u32(W2s) + i * 16, img + PK_W2 + i * 16, 16u);
    for (int d = 0; d < ndx; ++d)
      for (int i = tid; i < WBLK / 16; i += NTHREADS)
        cp_async16(smem_u32(W1T + d * WBLK) + i * 16, img + PK_W1T + args.dx_blk[d] * WBLK + i * 16, 16u);
    if (tid < 16) cp_async16(smem_u32(C.b2) + tid * 16, img + PK_VEC + 256 + tid * 16, 16u);
    else if (tid < 32) cp_async16(smem_u32(C.gamma) + (tid - 16) * 16, img + PK_VEC + 512 + (tid - 16) * 16, 16u);
  } else {
    for (int i = tid; i < HID * HID; i += NTHREADS) {
      const int c = i >> 6, k = i & 63;
      const __nv_bfloat16 w = __float2bfloat16_rn(c < OUT ? m.w2[c * HID + k] : 0.f);
      *reinterpret_cast<__nv_bfloat16*>(W2s + sw128_offset(c, k >> 3) + (k & 7) * 2) = w;  // [n=c][K=k]
      *reinterpret_cast<__nv_bfloat16*>(W2T + sw128_offset(k, c >> 3) + (c & 7) * 2) = w;  // [n=k][K=c]
    }
    for (int d = 0; d < ndx; ++d) {
      const int ks = kb.kstart[args.dx_blk[d]], kv = kb.kvalid[args.dx_blk[d]];
      for (int i = tid; i < HID * HID; i += NTHREADS) {
        const int c = i >> 6, k = i & 63;  // W1T_blk[n=k][K=c] = w1[c][kstart + k]
        *reinterpret_cast<__nv_bfloat16*>(W1T + d * WBLK + sw128_offset(k, c >> 3) + (c & 7) * 2) =
            __float2bfloat16_rn(k < kv ? m.w1[static_cast<long long>(c) * m.k1 + ks + k] : 0.f);
      }
    }
    if (tid < HID) {
      C.b2[tid] = tid < OUT ? m.b2[tid] : 0.f;
      C.gamma[tid] = m.gamma ? m.gamma[tid] : 1.f;
    }
  }
  if (tid < ONES_BYTES / 16)
    reinterpret_cast<uint4*>(ONES)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  cp_async_commit();  // group 0: the weight image (empty when converted in the kernel)
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&C.mbar[i], 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&C.tmem_base, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = C.tmem_base;
  {
    const int it = 0;
    P4C_BWD_TRACE(1);
  }
  const uint32_t tmem_row = tmem + (static_cast<uint32_t>(quad * 32) << 16) + col0;
  constexpr uint32_t IDESC_KK = make_idesc_bf16(TM, HID, 0, 0);  // A, B K-major
  constexpr uint32_t IDESC_MM = make_idesc_bf16(TM, HID, 1, 1);  // A, B MN-major (transposed reads)

  const int k1pad = (m.k1 + 31) / 32 * 32;
  const long long pf = static_cast<long long>(k1pad) * HID + HID * HID + 4 * HID;
  float* const part = g.wgrad_partials + static_cast<long long>(blockIdx.x) * pf;
  float* const pvec = part + static_cast<long long>(k1pad) * HID + HID * HID;  // db1, db2, dgamma, dbeta
  const bool padd = g.partials_mode == 2;  // add to what earlier calls left in this CTA's block
  if (padd) {  // pull this CTA's block towards L2: it is read once, at the very end
    for (long long o = static_cast<long long>(tid) * 32; o < pf; o += NTHREADS * 32) prefetch_l2(part + o);
  }

  const __nv_bfloat16* h1p = reinterpret_cast<const __nv_bfloat16*>(g.h1);
  const __nv_bfloat16* dya = reinterpret_cast<const __nv_bfloat16*>(g.dy_a);
  const __nv_bfloat16* dyb = reinterpret_cast<const __nv_bfloat16*>(g.dy_b);
  __nv_bfloat16* dh1_out = reinterpret_cast<__nv_bfloat16*>(g.dh1_out);

  // per-thread column-sum accumulators (column col0 + lane, rows of this warp): dgamma, dbeta
  // (db1, db2 are accumulated by the tensor core: ONES lanes of dW1B / dW2)
  float vacc[2] = {0.f, 0.f};

  long long tile = blockIdx.x;
  pdl_launch_dependents();
  {
    const RowAddr ra = bwd_row_addr(m, kb, tile, tid);  // graph constants only
    if (tid < TM) C.orow[0][tid] = ra.orow;
    bwd_row_store(ra, m, kb, C.row_off, tid);
  }
  pdl_wait();  // the previous kernel's outputs are complete and visible
  __syncthreads();
  {
    const int it = 0;
    P4C_BWD_TRACE(2);
  }
  bwd_h1_prefetch<NTHREADS>(h1p, C.orow[0], A1, tid);
  cp_async_commit();  // group 1: h1 of the first tile
  bwd_gather<NTHREADS>(m, kb, C.row_off, Xs, tid);
  cp_async_commit();  // group 2: X of the first tile
  // commit order from here on: ..., X(t), h1(t+1), X(t+1), ...: wait<1> at the top of tile t leaves
  // X(t) in flight; before G4 of tile t only X(t) is outstanding (wait<0>), h1(t+1) is issued after it.

  for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
    const int p = it & 1;
    const uint32_t parity = static_cast<uint32_t>(it & 1);
    const uint32_t acc = it ? 1u : 0u;  // weight-gradient accumulators: overwrite on the first tile
    const long long next = tile + gridDim.x;

    // weight image and h1(tile) have landed (X(tile) may still be in flight); the barrier makes
    // the h1 chunks copied by other threads visible and ends the previous tile (A1 / DH / scratch
    // are free)
    cp_async_wait<1>();
    tc_fence_before();
    __syncthreads();
    P4C_BWD_TRACE(3);
    const long long orow_true = C.orow[p][row];  // -1 = row beyond the problem
    const long long orow = (args.ablate & 4) ? -1 : orow_true;  // (ablation: no per-row global reads / writes)

    // addressing of the next tile: index loads issued now, consumed in phase f
    RowAddr nra = {};
    nra.orow = -1;
    if (next < ntiles) {
      nra = bwd_row_addr(m, kb, next, tid);
      if (tid < TM) C.orow[p ^ 1][tid] = nra.orow;
    }

    // ---- a. issue every global load of the tile at once: dy_a (gathered, scaled), dy_b, LN stats
    const __nv_bfloat16* dy_row = dya;  // first element of this row of dy_a
    float dy_sc = 1.f;
    if (orow >= 0) {
      if (args.dy_mode == 1) {
        dy_row = dya + (static_cast<long long>(C.row_off[args.dy_blk][row]) << 3);
      } else if (args.dy_mode == 2) {
        dy_row = dya + orow * OUT;
      } else {
        const uint32_t blk = static_cast<uint32_t>(tile) / static_cast<uint32_t>(m.batch);
        const long long bb = static_cast<uint32_t>(tile) - blk * static_cast<uint32_t>(m.batch);
        const int rr = static_cast<int>(blk) * TM + row;
        const long long ia = g.dy_a_index ? static_cast<long long>(g.dy_a_index[rr]) : static_cast<long long>(rr);
        if (g.dy_a_scale) dy_sc = g.dy_a_scale[ia];
        dy_row = dya + bb * g.dy_a_batch_stride + ia * OUT;
      }
    }
    uint4 dyr[NJ];
    uint4 dybr[NJ];
    float2 lnst = make_float2(0.f, 0.f);  // (mean, rstd)
    const bool wide = OUT == HID;
    if (wide) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        dyr[j] = orow >= 0 ? *reinterpret_cast<const uint4*>(dy_row + col0 + j * 8) : make_uint4(0, 0, 0, 0);
        dybr[j] = (orow >= 0 && dyb) ? *reinterpret_cast<const uint4*>(dyb + orow * HID + col0 + j * 8)
                                     : make_uint4(0, 0, 0, 0);
      }
    }
    if (ln && orow >= 0) lnst = *reinterpret_cast<const float2*>(g.stats + 2 * orow);

    // h1 part-row from the A1 tile: a1 = SiLU(h1) written back in place, SiLU'(h1) kept in
    // registers (packed bf16) for phase e -- one sigmoid (one MUFU op) serves both
    uint4 h1r[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      uint4* slot = reinterpret_cast<uint4*>(A1 + sw128_offset(row, (col0 >> 3) + j));
#if P4C_BWD_PACKED_SILU
      const uint4 hv = *slot;
      uint4 av;
      bf2_silu_and_grad(hv.x, av.x, h1r[j].x);
      bf2_silu_and_grad(hv.y, av.y, h1r[j].y);
      bf2_silu_and_grad(hv.z, av.z, h1r[j].z);
      bf2_silu_and_grad(hv.w, av.w, h1r[j].w);
      *slot = av;
#else
      float h[8], gs[8];
      unpack8(*slot, h);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float sg = fast_sigmoid(h[q]);
        gs[q] = fmaf(sg, fmaf(-h[q], sg, h[q]), sg);  // s (1 + h (1 - s))
        h[q] *= sg;
      }
      *slot = pack8(h);
      h1r[j] = pack8(gs);
#endif
    }
    // dy half-row -> packed bf16 registers (dy_a scaled, + dy_b)
    if (wide) {
      if (dy_sc != 1.f || dyb) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float d[8], e[8];
          unpack8(dyr[j], d);
          unpack8(dybr[j], e);
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] = fmaf(d[q], dy_sc, e[q]);
          dyr[j] = pack8(d);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int c0 = col0 + j * 8;
        float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (orow >= 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int c = c0 + q;
            if (c < OUT) {
              float v = __bfloat162float(dy_row[c]) * dy_sc;
              if (dyb) v += __bfloat162float(dyb[orow * OUT + c]);
              d[q] = v;
            }
          }
        }
        dyr[j] = pack8(d);
      }
    }

    if (ln) {
      const float mean = lnst.x, rstd = lnst.y;
      // ---- b. G1: h2 = a1 W2^T -> scratch
      P4C_BWD_TRACE(4);
      if (!(args.ablate & 32)) fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && !(args.ablate & 8)) {
        tc_fence_after();
        const uint32_t aa = smem_u32(A1), wa = smem_u32(W2s);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_bf16_ss_elect(tmem + COL_SCRATCH, make_desc_sw128(aa + j * 32, 16, 1024),
                      make_desc_sw128(wa + j * 32, 16, 1024), IDESC_KK, j ? 1u : 0u);
        mma_commit_elect(&C.mbar[0]);
      }
      if (!(args.ablate & 8)) mbar_wait(&C.mbar[0], parity);
      tc_fence_after();
      P4C_BWD_TRACE(5);
      // ---- c. LayerNorm backward: xhat from the recomputed h2, row sums exchanged between the
      //         two threads of a row, then dh2 (kept in dyr), dgamma / dbeta column sums
      float xh[CPT];
      tmem_ld<CPT>(tmem_row + COL_SCRATCH, xh);
      tmem_ld_wait();
      float s1 = 0.f, s2 = 0.f;
      const float nmr = -mean * rstd;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float d[8];
        unpack8(dyr[j], d);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = col0 + j * 8 + q;
          xh[j * 8 + q] = fmaf(xh[j * 8 + q] + C.b2[c], rstd, nmr);  // (h2 - mean) * rstd
          const float dxh = d[q] * C.gamma[c];
          s1 += dxh;
          s2 = fmaf(dxh, xh[j * 8 + q], s2);
        }
      }
      // row sums over the TPR threads of a row (same association order in every thread); only the
      // TPR warps that share the rows synchronise (named barrier)
      if (TPR == 4) {
        if (hsel >= 2) C.ln_part[hsel - 2][row] = make_float2(s1, s2);
        asm volatile("bar.sync %0, %1;" : : "r"(quad_bar), "n"(32 * TPR) : "memory");
        if (hsel < 2) {
          const float2 o = C.ln_part[hsel][row];
          C.ln_part[hsel][row] = make_float2(s1 + o.x, s2 + o.y);
        }
      } else {
        C.ln_part[hsel][row] = make_float2(s1, s2);
      }
      asm volatile("bar.sync %0, %1;" : : "r"(quad_bar), "n"(32 * TPR) : "memory");
      {
        const float2 lo = C.ln_part[0][row], hi = C.ln_part[1][row];
        s1 = (lo.x + hi.x) * (1.0f / HID);
        s2 = (lo.y + hi.y) * (1.0f / HID);
      }
      const float rs1 = -s1 * rstd, rs2 = -s2 * rstd;
      {  // dbeta contribution: column sums of dy (before dyr is overwritten with dh2)
        float t1[CPT];
#pragma unroll
        for (int j = 0; j < NJ; ++j) unpack8(dyr[j], t1 + j * 8);
        if (!(args.ablate & 16)) vacc[1] += warp_colsum<CPT>(t1, lane);
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        float d[8], o[8];
        unpack8(dyr[j], d);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = col0 + j * 8 + q;
          const float x = xh[j * 8 + q];
          const float dxh = d[q] * C.gamma[c];
          o[q] = fmaf(x, rs2, fmaf(dxh, rstd, rs1));  // rstd * (dxh - s1 - x * s2)
          xh[j * 8 + q] = d[q] * x;  // dgamma contribution
        }
        dyr[j] = pack8(o);  // dyr now holds dh2
      }
      if (!(args.ablate & 16)) vacc[0] += warp_colsum<CPT>(xh, lane); else vacc[0] += xh[0] + xh[CPT - 1];
    }
    // dh2 -> DH tile (db2 = its column sums: ONES lanes of G3)
#pragma unroll
    for (int j = 0; j < NJ; ++j) *reinterpret_cast<uint4*>(DH + sw128_offset(row, (col0 >> 3) + j)) = dyr[j];
    if (!(args.ablate & 32)) fence_proxy_async();
    tc_fence_before();
    __syncthreads();

    // ---- d. G2: da1 = dh2 W2 -> scratch ;  G3: dW2 += a1^T dh2
    P4C_BWD_TRACE(6);
    if (warp == 0 && !(args.ablate & 8)) {
      tc_fence_after();
      const uint32_t da = smem_u32(DH), wt = smem_u32(W2T), a1 = smem_u32(A1);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mma_bf16_ss_elect(tmem + COL_SCRATCH, make_desc_sw128(da + j * 32, 16, 1024),
                    make_desc_sw128(wt + j * 32, 16, 1024), IDESC_KK, j ? 1u : 0u);
      {
        const uint32_t ones = smem_u32(ONES);
#pragma unroll
        for (int j = 0; j < 8; ++j)  // K = 128 rows, 16 per MMA = 2 groups of 8 rows = 2048 bytes
          mma_bf16_ss_elect(tmem + COL_DW2, make_desc_sw128(a1 + j * 2048, ones - (a1 + j * 2048), 1024),
                      make_desc_sw128(da + j * 2048, BLK, 1024), IDESC_MM, (acc | j) ? 1u : 0u);
      }
      mma_commit_elect(&C.mbar[1]);
    }
    if (!(args.ablate & 8)) mbar_wait(&C.mbar[1], parity);
    tc_fence_after();
    P4C_BWD_TRACE(7);
    // ---- e. dh1 = da1 * SiLU'(h1) -> DH tile (overwrites dh2; db1 = its column sums, ONES lanes of G4)
    {
      float v[CPT];
      tmem_ld<CPT>(tmem_row + COL_SCRATCH, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
#if P4C_BWD_PACKED_DH1
        uint4 packed = pack8(v + j * 8);  // da1 -> bf16, times SiLU'(h1) in bf16x2
        packed.x = bf2_mul(packed.x, h1r[j].x);
        packed.y = bf2_mul(packed.y, h1r[j].y);
        packed.z = bf2_mul(packed.z, h1r[j].z);
        packed.w = bf2_mul(packed.w, h1r[j].w);
        if (orow < 0) packed = make_uint4(0, 0, 0, 0);
#else
        float h[8];
        unpack8(h1r[j], h);  // SiLU'(h1)
#pragma unroll
        for (int q = 0; q < 8; ++q) v[j * 8 + q] = orow >= 0 ? v[j * 8 + q] * h[q] : 0.f;
        const uint4 packed = pack8(v + j * 8);
#endif
        *reinterpret_cast<uint4*>(DH + sw128_offset(row, (col0 >> 3) + j)) = packed;
      }
    }
    P4C_BWD_TRACE(8);
    // X(tile) landed (issued one tile ago).  Nothing else of this thread is in flight here: the
    // proxy fence (MEMBAR.ALL.CTA in SASS) drains every outstanding global access of the thread, so
    // the dh1 stores and the next tile's h1 prefetch are issued AFTER it, under the G4 / G5 MMAs.
    cp_async_wait<0>();
    if (!(args.ablate & 32)) fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    P4C_BWD_TRACE(9);

    // ---- f. G4: dW1 += X^T dh1 ; G5(0): dX = dh1 W1[:, s0]
    if (warp == 0 && !(args.ablate & 8)) {
      tc_fence_after();
      const uint32_t da = smem_u32(DH), xa = smem_u32(Xs);
      {
        const uint32_t ones = smem_u32(ONES);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          mma_bf16_ss_elect(tmem + COL_DW1A, make_desc_sw128(xa + j * 2048, BLK, 1024),
                      make_desc_sw128(da + j * 2048, BLK, 1024), IDESC_MM, (acc | j) ? 1u : 0u);
        // dW1B: lanes 0..63 = X2^T dh1 (three K-blocks) or ones; lanes 64..127 = 1^T dh1 = db1
        const uint32_t x2 = nblk == 3 ? xa + 2 * BLK : ones;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t ab = nblk == 3 ? x2 + j * 2048 : ones;
          mma_bf16_ss_elect(tmem + COL_DW1B, make_desc_sw128(ab, ones - ab, 1024),
                      make_desc_sw128(da + j * 2048, BLK, 1024), IDESC_MM, (acc | j) ? 1u : 0u);
        }
      }
      if (ndx > 0) {
        const uint32_t wt = smem_u32(W1T);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_bf16_ss_elect(tmem + COL_SCRATCH, make_desc_sw128(da + j * 32, 16, 1024),
                      make_desc_sw128(wt + j * 32, 16, 1024), IDESC_KK, j ? 1u : 0u);
      }
      mma_commit_elect(&C.mbar[2]);
    }
    // meanwhile: dh1 to global memory (this thread's own chunks, read back from the DH tile), the
    // next tile's h1 rows into the A1 tile (free: G1, G3 complete), its source offsets (index
    // loads issued at the top of this tile) and L2 prefetch of its identity-indexed per-row inputs
    if (dh1_out && orow >= 0) {
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        *reinterpret_cast<uint4*>(dh1_out + orow * HID + col0 + j * 8) =
            *reinterpret_cast<const uint4*>(DH + sw128_offset(row, (col0 >> 3) + j));
    }
    if (next < ntiles && !(args.ablate & 2)) bwd_h1_prefetch<NTHREADS>(h1p, C.orow[p ^ 1], A1, tid);
    cp_async_commit();
    if (next < ntiles) {
      bwd_row_store(nra, m, kb, C.row_off, tid);
      const long long no = C.orow[p ^ 1][row];
      if (no >= 0) {
        if (args.dy_mode == 2) prefetch_l2(dya + no * OUT + col0);
        if (dyb) prefetch_l2(dyb + no * OUT + col0);
        if (ln && hsel == 0) prefetch_l2(g.stats + 2 * no);
      }
    }
    if (!(args.ablate & 8)) mbar_wait(&C.mbar[2], parity);
    tc_fence_after();
    P4C_BWD_TRACE(10);
    __syncthreads();  // next row_off visible; X is free (DH still feeds a possible second dX)
    if (next < ntiles && !(args.ablate & 1)) bwd_gather<NTHREADS>(m, kb, C.row_off, Xs, tid);
    cp_async_commit();

    // ---- g. dX stores (one or two identity-indexed sources)
    for (int d = 0; d < ndx; ++d) {
      if (d == 1) {
        tc_fence_before();
        __syncthreads();  // everyone has read the scratch of dX(0)
        if (warp == 0) {
          tc_fence_after();
          const uint32_t da = smem_u32(DH), wt = smem_u32(W1T + WBLK);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_bf16_ss_elect(tmem + COL_SCRATCH, make_desc_sw128(da + j * 32, 16, 1024),
                        make_desc_sw128(wt + j * 32, 16, 1024), IDESC_KK, j ? 1u : 0u);
          mma_commit_elect(&C.mbar[3]);
        }
        mbar_wait(&C.mbar[3], parity);
        tc_fence_after();
      }
      const int blk = args.dx_blk[d];
      const int s = kb.src[blk];
      __nv_bfloat16* dxp = reinterpret_cast<__nv_bfloat16*>(g.dx[s]) + kb.col0[blk];
      const int ld = g.dx_ld[s];
      const int nch = kb.nchunks[blk];
      const bool add_dy = (blk == 0 && g.dx0_add_dy != 0);
      float v[CPT];
      tmem_ld<CPT>(tmem_row + COL_SCRATCH, v);
      tmem_ld_wait();
      if (orow >= 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          float* o = v + j * 8;
          const int c0 = col0 + j * 8;
          if ((c0 >> 3) >= nch) continue;  // beyond the (padded) width of this block
          __nv_bfloat16* dst = dxp + orow * ld + c0;
          if (g.dx_accumulate[s]) {
            float e[8];
            unpack8(*reinterpret_cast<const uint4*>(dst), e);
#pragma unroll
            for (int q = 0; q < 8; ++q) o[q] += e[q];
          }
          if (add_dy) {
            // residual / edge-update pass-through: d src0 += dy (dy_b only when dx0_add_dy == 1)
            float e[8];
            if (g.dx0_add_dy == 2) {
              unpack8(*reinterpret_cast<const uint4*>(dy_row + c0), e);
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] = fmaf(e[q], dy_sc, o[q]);
            }
            if (dyb) {
              unpack8(*reinterpret_cast<const uint4*>(dyb + orow * HID + c0), e);
#pragma unroll
              for (int q = 0; q < 8; ++q) o[q] += e[q];
            }
          }
          *reinterpret_cast<uint4*>(dst) = pack8(o);
        }
      }
    }
    P4C_BWD_TRACE(11);
  }
  tc_fence_before();
  __syncthreads();  // last tile done (the loop's own end-of-tile barrier sits at the top of the next tile)

  // ---- per-CTA partials: weight and bias gradients from TMEM, LayerNorm-affine gradients from the
  //      warp sums.  In accumulate mode the block's previous content is fetched with all loads in
  //      flight at once (the lines were pulled into L2 when the kernel started).
  cp_async_wait<0>();
  tc_fence_after();
  {
    // this thread's rows of the partial block: dW1 (blocks 0/1), dW1 (block 2), dW2, and for the
    // thread that owns TMEM lane 64: db1, db2
    float* dstA = nullptr;
    float* dstB = nullptr;
    float* dstC = nullptr;
    {
      const int blk = row >> 6, kk = row & 63;
      if (need_w1 && blk < nblk && kk < kb.kvalid[blk]) dstA = part + static_cast<long long>(kb.kstart[blk] + kk) * HID + col0;
      if (need_w1 && nblk == 3 && row < kb.kvalid[2]) dstB = part + static_cast<long long>(kb.kstart[2] + row) * HID + col0;
      if (need_w2 && row < HID) dstC = part + static_cast<long long>(k1pad) * HID + static_cast<long long>(row) * HID + col0;
      if (row == HID) {
        dstB = pvec + col0;        // db1
        dstC = pvec + HID + col0;  // db2
      }
    }
    constexpr int NQ = CPT / 4;
    float4 eA[NQ], eB[NQ], eC[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      eA[q] = (padd && dstA) ? *reinterpret_cast<const float4*>(dstA + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      eB[q] = (padd && dstB) ? *reinterpret_cast<const float4*>(dstB + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
      eC[q] = (padd && dstC) ? *reinterpret_cast<const float4*>(dstC + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    auto put32 = [&](float* dst, const float (&v)[CPT], const float4 (&e)[NQ]) {
      if (!dst) return;
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        *reinterpret_cast<float4*>(dst + 4 * q) =
            make_float4(v[4 * q] + e[q].x, v[4 * q + 1] + e[q].y, v[4 * q + 2] + e[q].z, v[4 * q + 3] + e[q].w);
    };
    float v[CPT];
    tmem_ld<CPT>(tmem_row + COL_DW1A, v);  // dW1[k][c] for K-blocks 0,1 (TMEM lane = 64 * block + k)
    tmem_ld_wait();
    put32(dstA, v, eA);
    tmem_ld<CPT>(tmem_row + COL_DW1B, v);
    tmem_ld_wait();
    put32(dstB, v, eB);
    tmem_ld<CPT>(tmem_row + COL_DW2, v);
    tmem_ld_wait();
    put32(dstC, v, eC);
  }
  {
    const int it = 0;
    P4C_BWD_TRACE(12);
  }
  // cross-warp sums in a fixed order; the A1 tile is free by now and serves as scratch
  float (*red)[2][32] = reinterpret_cast<float (*)[2][32]>(A1);  // [warp][vector][lane]
#pragma unroll
  for (int v = 0; v < 2; ++v) red[warp][v][lane] = vacc[v];
  __syncthreads();
  if (tid < HID) {
    // column tid lives in the warps w0..w0+3 of its column group, lane l (see warp_colsum)
    const int w0 = (tid / CPT) * 4, l = (tid % CPT) * (32 / CPT);
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float sum = (red[w0][v][l] + red[w0 + 1][v][l]) + (red[w0 + 2][v][l] + red[w0 + 3][v][l]);
      float* d = pvec + (2 + v) * HID + tid;  // dgamma, dbeta
      *d = padd ? *d + sum : sum;
    }
  }
  tc_fence_before();
  __syncthreads();
  {
    const int it = 0;
    P4C_BWD_TRACE(13);
  }
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tc

bool tc_kblocks(const p4c_mlp_t* m, tc::KBlocks* kb);

static int tc_bwd_dx_blocks(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, const tc::KBlocks& kb, int* blks) {
  int ndx = 0;
  for (int k = 0; k < kb.nblk; ++k)
    if (g->dx[kb.src[k]]) {
      if (ndx < 2) blks[ndx] = k;
      ++ndx;
    }
  return ndx;
}

bool tc_backward_supported(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g) {
  if (!tc_forward_supported(m)) return false;
  tc::KBlocks kb;
  tc_kblocks(m, &kb);
  for (int s = 0; s < m->nsrc; ++s) {
    if (m->src[s].seg_rowptr) return false;  // the backward reads the saved segment sums as plain rows
    if (g->dx[s] && (g->dx_ld[s] % 8 != 0 || (reinterpret_cast<uintptr_t>(g->dx[s]) % 16) != 0)) return false;
  }
  int blks[2];
  if (tc_bwd_dx_blocks(m, g, kb, blks) > 2) return false;
  if (m->out_features == P4C_HIDDEN) {
    if (g->dy_a_batch_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(g->dy_a) % 16) != 0) return false;
    if (g->dy_b && (reinterpret_cast<uintptr_t>(g->dy_b) % 16) != 0) return false;
  }
  if (g->dh1_out && (reinterpret_cast<uintptr_t>(g->dh1_out) % 16) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(g->h1) % 16) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(g->wgrad_partials) % 16) != 0) return false;
  return true;
}

int tc_backward_grid(const p4c_mlp_t* m) {
  const long long total = static_cast<long long>(m->rows) * m->batch;
  (void)total;
  const long long ntiles = static_cast<long long>((m->rows + tc::TM - 1) / tc::TM) * m->batch;
  return static_cast<int>(ntiles < tc_max_ctas() ? ntiles : tc_max_ctas());
}

int launch_mlp2_tc_backward(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, cudaStream_t st) {
  using namespace tc;
  BwdArgs a;
  a.m = *m;
  a.g = *g;
  tc_kblocks(m, &a.kb);
  a.dx_blk[0] = a.dx_blk[1] = 0;
  a.ndx = tc_bwd_dx_blocks(m, g, a.kb, a.dx_blk);
  a.zero_x = 0;
  for (int k = 0; k < a.kb.nblk; ++k) a.zero_x |= a.kb.nchunks[k] < 8;
  a.dy_mode = 0;
  a.dy_blk = 0;
  a.trace = tc_get_trace();
  {
    static int ablate = -1;
    if (ablate < 0) {
      const char* e = getenv("P4C_BWD_ABLATE");
      ablate = e ? atoi(e) : 0;
    }
    a.ablate = ablate;
  }
  const long long full = m->full_rows > 0 ? m->full_rows : m->rows;
  if (!g->dy_a_scale) {
    if (!g->dy_a_index && (m->batch == 1 || g->dy_a_batch_stride == full * m->out_features)) {
      // identity rows; with full_rows the caller pre-offsets dy_a like the other output-like arrays
      a.dy_mode = 2;
    } else if (g->dy_a_index && m->out_features == P4C_HIDDEN) {
      for (int k = 0; k < a.kb.nblk; ++k) {
        const p4c_src_t& sk = m->src[a.kb.src[k]];
        if (sk.index == g->dy_a_index && sk.ld == P4C_HIDDEN && a.kb.col0[k] == 0 &&
            sk.batch_stride == g->dy_a_batch_stride) {
          a.dy_mode = 1;
          a.dy_blk = k;
          break;
        }
      }
    }
  }
  const size_t smem = bwd_smem_bytes(a.kb.nblk, a.ndx) + 1024;
  // two threads per tile row (a 4-threads-per-row shape, 32 warps per SM at 64 registers, measured 8 % slower in
  // round 1 and was removed)
  constexpr int tpr = 2;
  auto kernel = mlp2_tc_bwd_kernel<2>;
  static size_t configured[1] = {0};
  if (smem > configured[0]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mlp_backward(tc): cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    configured[0] = smem;
  }
  launch_pdl(kernel, dim3(tc_backward_grid(m)), dim3(TM * tpr), smem, st, a);
  P4C_CHECK_LAUNCH("mlp_backward(tc)");
  return P4C_OK;
}

}  // namespace p4c
