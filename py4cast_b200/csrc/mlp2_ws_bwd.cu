// Warp-specialised, software-pipelined BACKWARD of the fused 2-layer MLP (bf16 storage, LayerNorm,
// 64 outputs) for the large edge / grid problems: the operator of mlp2_tc_bwd_kernel (mlp2_tc_bwd.cu)
// re-organised like the pipelined forward (mlp2_ws.cu), and RECOMPUTING the layer-1 pre-activation
// h1 = X W1^T + b1 and the LayerNorm statistics from the gathered sources instead of reading what the
// forward saved: the forward of such a call writes neither h1 nor the statistics (2 x 128 B + 8 B per
// row of HBM traffic less, e.g. 5.4 GB per m2g InteractionNet call at batch 16).
//
// One persistent CTA per SM, 24 warps (21 working), two tile CONTEXTS (tile t of the CTA runs in context t & 1):
//   warps 16-19 producer   cp.async gather of the up-to-three 64-column K-blocks of tile t into the X tiles
//                          of its context (four complete 128-byte rows per instruction), gather indices
//                          one tile ahead, L2 prefetch of the next tile's rows.
//   warps 20-21 MMA        one warp per context (one elected lane issues), in the order of the context's chain:
//        G0  h1  = X W1^T                               -> chain scratch of the context (TMEM)
//        G1  h2  = a1 W2^T            (LayerNorm input)  -> chain scratch
//        G2  da1 = dh2 W2             (W2 tile read MN-major) -> chain scratch
//        G3  dW2 += [a1 | 1]^T dh2    (lanes 64.. = db2)
//        G4  dW1 += [X0 | X1]^T dh1,  [X2 | 1]^T dh1   (lanes 64.. = db1)
//        G5  dX_s = dh1 W1[:, block s] (W1 tiles read MN-major; up to two products) -> dX scratch
//   warps 0-7 / 8-15       compute group of context 0 / 1; warp w owns TMEM lane quadrant w % 4 (rows) and
//                          column half (w / 4) % 2: thread = (tile row, 32 columns):
//        a  h1 (+b1, rounded to bf16 as the forward stored it) -> a1 = SiLU(h1) -> A1 tile; SiLU'(h1)
//           stays in registers (packed bf16)
//        c  LayerNorm backward from the recomputed h2: ONE exchange of four row sums between the two
//           threads of a row (sum h2, sum h2^2, sum g dy, sum g dy h2) gives mean, rstd and both
//           projection terms; dh2 -> DH tile; dgamma / dbeta column sums (transposing butterflies)
//        e  dh1 = da1 * SiLU'(h1) -> DH tile (and, coalesced, to HBM for the by-sender pass)
//        g  dX epilogues: identity-indexed sources are stored; for receiver-sorted edge sets of uniform
//           in-degree the receiver gradient is summed over the rows of a group (butterfly) and
//           accumulated into d_rec — replaces one p4c_seg_linear pass over dh1.
//   While one group waits for its MMAs the other runs its element-wise phase; no CTA-wide barrier
//   inside the tile loop (64-thread named barriers and mbarriers only).
//
// TMEM (512 columns): dW1A 64 | dW1B 64 | dW2 64 | chain scratch 2 x 64 | dX scratch 2 x 64.
// Shared memory: 2 x (3 X + A1 + DH tiles of 16 KB) + W1 (24 KB) + W2 (8 KB) + ONES + control = ~204 KB.
// Weight gradients: per-CTA partial blocks in the layout of the classic engine (fixed-order reduction
// by wgrad_reduce, no float atomics): results are bit-reproducible run to run.
//
// Reference semantics: autograd of mfai / neural-lam `make_mlp` inside `InteractionNet`
// (SURVEY.md §3.2, K1-K6), reached from py4cast/lightning.py:586-596.
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

#ifndef P4C_WB_ABLATE
#define P4C_WB_ABLATE 0  // developer aid (scripts/ablate_ws_bwd.sh): 1 no column sums, 2 no dh1 store, 4 no dX epilogue
#endif

namespace p4c {
namespace tc {

constexpr int WB_TM = 128;
constexpr int WB_HID = P4C_HIDDEN;
constexpr int WB_BLK = WB_TM * 128;    // 16 KB: one 128 x 64 bf16 tile
constexpr int WB_WBLK = WB_HID * 128;  // 8 KB: one 64 x 64 bf16 tile
constexpr int WB_CWARPS = 16, WB_PWARPS = 4;
constexpr int WB_MMA_WARP = WB_CWARPS + WB_PWARPS;
constexpr int WB_THREADS = (WB_CWARPS + WB_PWARPS + 4) * 32;  // warps 21-23 only pad the MMA warp's warpgroup
// registers: the launch bound (768 threads) caps the kernel at 80 per thread; the compute warpgroups
// raise their share to 96, the producer / MMA warpgroups drop to 40 / 56: 32 x (16 x 96 + 4 x 40 + 4 x 56) = 768 x 80,
// the pool setmaxnreg redistributes is what the CTA was launched with, not the whole register file
constexpr int WB_REGS_COMPUTE = 96, WB_REGS_PRODUCER = 40, WB_REGS_MMA = 56;
constexpr uint32_t WB_TMEM_COLS = 512;
constexpr uint32_t WB_COL_DW1A = 0, WB_COL_DW1B = 64, WB_COL_DW2 = 128, WB_COL_CH = 192, WB_COL_DX = 320, WB_COL_BS = 448;
constexpr uint32_t kWbNoRow = 0xFFFFFFFFu;
constexpr int WB_ONES_BYTES = 2048;

struct WbSmem {
  // X ring: 2 stages of 3 K-blocks, or 3 stages of <= 2 (stage s at X + s * stage_blocks * 16 KB); the K-blocks of
  // a stage are adjacent (M = 128 operands span two of them)
  uint8_t X[2 * P4C_MAX_SRC][WB_BLK];
  uint8_t A1[2][WB_BLK];
  uint8_t DH[2][WB_BLK];
  uint8_t W1[P4C_MAX_SRC][WB_WBLK];
  uint8_t W2[WB_WBLK];
  uint8_t WBIAS[WB_WBLK];  // bias tile of the weight image (PK_BIAS): b1 in K-slice 0, b2 in K-slice 1
  uint8_t ONES[WB_ONES_BYTES];
  float vec[4][WB_HID];  // b1, b2, gamma, beta
  float4 ln_part[2][2][WB_TM];  // [group][column half][row]
  uint64_t x_full[3], x_empty[3], h1_full[2], a1_full[2], h2_full[2], dh2_full[2], da1_full[2], dh1_full[2],
      dx_full[2];
  uint64_t dxs_empty[2];
  uint64_t g3_turn[2], g4_turn[2];  // weight-gradient products of a context's tile issued (MMA warp -> MMA warp)
  uint64_t bs_empty;                // the batch-summed dX of a row block was read (compute group -> MMA warps)
  uint32_t tmem_base;
};

static_assert(sizeof(WbSmem) + 1024 <= 232448, "pipelined backward: shared memory beyond the 227 KB a CTA can opt into");

struct WbArgs {
  p4c_mlp_t m;
  p4c_mlp_bwd_t g;
  KBlocks kb;
  int nslot;         // dX products per tile (<= 2)
  int slot_blk[2];   // K-block of each
  int slot_kind[2];  // 0: stored to g.dx[source] (identity rows), 1: group sum accumulated into g.dx_group_out
  int zero_x;
  unsigned long long div_magic;  // ceil(2^64 / batch): row block = umul64hi(tile, magic)
  uint32_t batch, ntiles, full_rows;
  int total_blocks;  // partial blocks of this MLP's workspace (blocks >= gridDim.x are zeroed in mode 1)
  // identity-indexed K-blocks arrive by TMA (bit k of tma_mask; tma_b[k] = 0 for a source shared by the batch)
  int tma_mask;
  int tma_b[P4C_MAX_SRC];
  alignas(64) CUtensorMap tmap[P4C_MAX_SRC];
  // dh1 rows of a tile leave by ONE TMA store of the DH tile, issued by the context's MMA warp once the tile is
  // complete (instead of 4 x (shuffle, 16-byte shared-memory read, global store) per compute thread)
  alignas(64) CUtensorMap tmap_dh1;
  int dh1_tma;
  // dx0_batch_sum: every CTA owns a contiguous range of row blocks and runs the samples of a block back to back
  // (tile t of the CTA = sample t % batch of block blk_lo + t / batch): dh1 W1[:, block 0] of the samples is
  // accumulated in TMEM (64 columns) and stored once per block
  int bsum;
  uint32_t nblocks;  // row blocks of the problem
  long long* trace;  // developer aid (p4c_set_trace_buffer): SM clock at the phases of CTA 0's tile 2, see scripts/trace_ws_bwd.py
};
// slots 0.. compute group 0 (warp 0), 16.. MMA lane (context 0), 32.. producer warp 0
// Phase traces are compiled in only with -DP4C_PHASE_TRACE (make EXTRA=-DP4C_PHASE_TRACE; the trace_* / bwd_micro
// scripts print zeros otherwise): the ~30 predicated clock reads per tile cost 1.5 % of the m2g edge backward.
#ifdef P4C_PHASE_TRACE
#define P4C_WB_TRACE(cond, i)                                   \
  do {                                                          \
    if (a.trace && blockIdx.x == 0 && (cond)) a.trace[i] = clock64(); \
  } while (0)
#else
#define P4C_WB_TRACE(cond, i) do { (void)sizeof(cond); } while (0)
#endif
#define P4C_WB_KTRACE(i) P4C_WB_TRACE(tid == 0, i)  // kernel-level timeline (slots 48..), scripts/trace_ws_bwd_small.py

__device__ __forceinline__ bool wb_mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ unsigned long long wb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wb_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" : : "l"(p)); }

__global__ void __launch_bounds__(WB_THREADS, 1) mlp2_ws_bwd_kernel(const __grid_constant__ WbArgs a) {
  extern __shared__ unsigned char smem_raw[];
  WbSmem& S = *reinterpret_cast<WbSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const p4c_mlp_t& m = a.m;
  const p4c_mlp_bwd_t& g = a.g;
  // the warp index is broadcast from lane 0 so that the compiler KNOWS it is warp-uniform: everything the MMA
  // warps derive from it (context, tile addresses, descriptors) then lives in uniform registers
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nblk = a.kb.nblk;
  const uint32_t nbatch = a.batch;
  const uint32_t rows = static_cast<uint32_t>(m.rows);
  const int nslot = a.nslot;
  const bool bsum = a.bsum != 0;
  // tile t of this CTA -> (row block, sample): round-robin over the CTAs, row-block-major / sample-minor (the
  // samples of a block run back to back on neighbouring CTAs: a K-block shared by the batch comes from HBM once),
  // or, for dx0_batch_sum, a contiguous range of row blocks per CTA with the samples of a block back to back
  const uint32_t blk_lo =
      bsum ? static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x) * a.nblocks / gridDim.x) : 0u;
  auto tile_coords = [&](uint32_t t, uint32_t& blk, uint32_t& b) {
    const uint32_t tile = bsum ? t : blockIdx.x + t * gridDim.x;
    const uint32_t q = nbatch == 1 ? tile
                                   : static_cast<uint32_t>(__umul64hi(static_cast<unsigned long long>(tile), a.div_magic));
    b = tile - q * nbatch;
    blk = blk_lo + q;
  };
  P4C_WB_KTRACE(48);  // kernel-level timeline (slots 48..): entry, set-up done, first tile, last tile, partials, exit
#ifdef P4C_PHASE_TRACE
  if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[56] = static_cast<long long>(wb_globaltimer());
#endif
  // X ring depth: three stages when a stage is at most two K-blocks (the gather of tile t + 2 then runs while
  // tiles t, t + 1 compute), two otherwise (shared memory); dX scratch: one per context when a tile has a
  // single dX product, else one pair shared by the contexts (TMEM columns)
  const uint32_t xns = nblk <= 2 ? 3u : 2u, xsb = nblk <= 2 ? 2u : 3u;
  const bool dx_shared = nslot > 1;

  // ---- set-up: weight image, ONES, barriers, TMEM
  if (a.zero_x) {
    uint4* xz = reinterpret_cast<uint4*>(S.X);
    for (int i = tid; i < 2 * P4C_MAX_SRC * WB_BLK / 16; i += WB_THREADS) xz[i] = make_uint4(0, 0, 0, 0);
  }
  {
    const uint8_t* img = reinterpret_cast<const uint8_t*>(m.packed);
    for (int i = tid; i < nblk * (WB_WBLK / 16); i += WB_THREADS)
      cp_async16(smem_u32(S.W1) + i * 16, img + PK_W1 + i * 16, 16u);
    for (int i = tid; i < WB_WBLK / 16; i += WB_THREADS) cp_async16(smem_u32(S.W2) + i * 16, img + PK_W2 + i * 16, 16u);
    for (int i = tid; i < WB_WBLK / 16; i += WB_THREADS) cp_async16(smem_u32(S.WBIAS) + i * 16, img + PK_BIAS + i * 16, 16u);
    if (tid < 64) cp_async16(smem_u32(S.vec) + tid * 16, img + PK_VEC + tid * 16, 16u);
    cp_async_commit();
  }
  if (tid < WB_ONES_BYTES / 16)
    reinterpret_cast<uint4*>(S.ONES)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (tid == 0) {
    for (int c = 0; c < 3; ++c) {
      mbar_init(&S.x_full[c], WB_PWARPS);
      mbar_init(&S.x_empty[c], 1);
    }
    for (int c = 0; c < 2; ++c) {
      mbar_init(&S.h1_full[c], 1);
      mbar_init(&S.a1_full[c], 8);
      mbar_init(&S.h2_full[c], 1);
      mbar_init(&S.dh2_full[c], 8);
      mbar_init(&S.da1_full[c], 1);
      mbar_init(&S.dh1_full[c], 8);
      mbar_init(&S.dx_full[c], 1);
    }
    mbar_init(&S.dxs_empty[0], 8);
    mbar_init(&S.dxs_empty[1], 8);
    mbar_init(&S.bs_empty, 8);
    for (int c = 0; c < 2; ++c) {
      mbar_init(&S.g3_turn[c], 1);
      mbar_init(&S.g4_turn[c], 1);
    }
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == WB_MMA_WARP) tmem_alloc(&S.tmem_base, WB_TMEM_COLS);
  cp_async_wait<0>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t ntl =
      bsum ? (static_cast<uint32_t>(static_cast<unsigned long long>(blockIdx.x + 1) * a.nblocks / gridDim.x) - blk_lo) * nbatch
           : (a.ntiles > blockIdx.x ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u);
  pdl_launch_dependents();
  P4C_WB_KTRACE(49);
#ifdef P4C_PHASE_TRACE
  if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[55] = ntl;
#endif

  float vacc0 = 0.f, vacc1 = 0.f;  // compute warps: dgamma / dbeta column sums (column col0 + lane, rows of the warp)

  if (warp >= WB_CWARPS && warp < WB_MMA_WARP) {
    // =============================== producer ===============================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" : : "n"(WB_REGS_PRODUCER));
    const int pw = warp - WB_CWARPS;
    const int sub = lane >> 3, j = lane & 7;  // lane -> (row within a group of 4, 16-byte chunk)
    int32_t idx_next[P4C_MAX_SRC];
    uint32_t b_next = 0, blk_next = 0;
    bool valid_next = false;
    auto load_indices = [&](uint32_t t) {  // gather indices of tile t (graph constants)
      uint32_t blk;
      tile_coords(t, blk, b_next);
      blk_next = blk;
      const uint32_t rr = blk * WB_TM + pw * 32 + lane;
      valid_next = t < ntl && rr < rows;
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k) {
        idx_next[k] = static_cast<int32_t>(rr);
        if (k < nblk && valid_next) {
          const int32_t* ix = m.src[a.kb.src[k]].index;
          if (ix) idx_next[k] = ix[rr];
        }
      }
    };
    auto row_offsets = [&](uint32_t (&off)[P4C_MAX_SRC]) {  // in units of 8 elements
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k) {
        off[k] = kWbNoRow;
        if (k < nblk && valid_next) {
          const p4c_src_t& src = m.src[a.kb.src[k]];
          off[k] = static_cast<uint32_t>(
              (static_cast<long long>(b_next) * src.batch_stride + static_cast<long long>(idx_next[k]) * src.ld +
               a.kb.col0[k]) >> 3);
        }
      }
    };
    load_indices(0);
    pdl_wait();  // everything above reads only constants of the step (weight image, graph indices)
    uint32_t off[P4C_MAX_SRC];
    row_offsets(off);
    for (uint32_t t = 0; t < ntl; ++t) {
      const uint32_t c = t % xns, n = t / xns;
      const bool trc = t == 2 && pw == 0 && lane == 0;
      P4C_WB_TRACE(trc, 32);
      const uint32_t b_cur = b_next, row0 = blk_next * WB_TM;
      load_indices(t + 1);
      if (n > 0) mbar_wait_backoff<100, true>(&S.x_empty[c], (n - 1) & 1u);
      P4C_WB_TRACE(trc, 33);
      if (a.tma_mask && pw == 0 && lane == 0) {  // whole tiles by TMA: bytes counted on the stage's barrier
#pragma unroll
        for (int k = 0; k < P4C_MAX_SRC; ++k)
          if (k < nblk && (a.tma_mask >> k & 1)) {
            mbar_expect_tx(&S.x_full[c], WB_BLK);
            tma_load_3d(smem_u32(S.X[c * xsb + k]), &a.tmap[k], &S.x_full[c], a.kb.col0[k], static_cast<int>(row0),
                        a.tma_b[k] ? static_cast<int>(b_cur) : 0);
          }
      }
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k) {
        if (k < nblk && !(a.tma_mask >> k & 1)) {
          const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(m.src[a.kb.src[k]].ptr);
          const uint32_t xs = smem_u32(S.X[c * xsb + k]);
          const bool fetch = j < a.kb.nchunks[k];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + sub;
            const uint32_t o = __shfl_sync(0xffffffffu, off[k], r);
            if (fetch) {
              const bool ok = o != kWbNoRow;
              cp_async16(xs + sw128_offset(pw * 32 + r, j), ok ? base + (static_cast<size_t>(o) << 3) + j * 8 : base,
                         ok ? 16u : 0u);
            }
          }
        }
      }
      cp_async_commit();
      P4C_WB_TRACE(trc, 34);
      // next tile: row offsets (its index loads were issued above) and an L2 prefetch of its rows
      row_offsets(off);
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k)
        if (k < nblk && off[k] != kWbNoRow)
          wb_prefetch_l2(reinterpret_cast<const __nv_bfloat16*>(m.src[a.kb.src[k]].ptr) + (static_cast<size_t>(off[k]) << 3));
      cp_async_wait<0>();
      P4C_WB_TRACE(trc, 35);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.x_full[c]);
    }
  } else if (warp >= WB_MMA_WARP) {
    // ================================= MMA ==================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" : : "n"(WB_REGS_MMA));
    if (warp < WB_MMA_WARP + 2) {
      // Two MMA warps, one per context (warp-uniform control flow: all 32 lanes run the loop, one elected lane
      // issues).  Every product is a run of MMAs whose operand descriptors advance by a constant (K-major: 32
      // bytes = 2 address units per K = 16; MN-major: 16 tile rows = 2048 bytes = 128 units): one base
      // descriptor per operand, one 64-bit add per MMA.
      // The weight-gradient products (G3, G4) of tile t are issued only after those of tile t - 1 have
      // COMPLETED (the other warp commits them to g3_turn / g4_turn): the fp32 accumulation order, and with it
      // every bit of the result, is the same in every run.  The first product issued into an accumulator
      // overwrites it.
      constexpr uint32_t IDESC_KK = make_idesc_bf16(WB_TM, WB_HID, 0, 0);  // A, B K-major
      constexpr uint32_t IDESC_KM = make_idesc_bf16(WB_TM, WB_HID, 0, 1);  // A K-major, B MN-major
      constexpr uint32_t IDESC_MM = make_idesc_bf16(WB_TM, WB_HID, 1, 1);  // A, B MN-major
      constexpr uint32_t IDESC_MK = make_idesc_bf16(WB_TM, WB_HID, 1, 0);  // A MN-major (all ones), B K-major
      constexpr long long KSTEP = 2, MNSTEP = 128;
      // [tile rows | ONES] as an MN-major M = 128 operand: the address moves on, the ONES block does not
      constexpr long long ONES_STEP = MNSTEP - (MNSTEP << 16);
      const uint32_t c = warp - WB_MMA_WARP;
      const uint32_t ones = smem_u32(S.ONES), a1 = smem_u32(S.A1[c]), da = smem_u32(S.DH[c]);
      const uint32_t ch = tmem + WB_COL_CH + c * WB_HID;
      const uint32_t dxc = tmem + WB_COL_DX + (dx_shared ? 0u : c * WB_HID);
      const uint64_t w2_k = make_desc_sw128(smem_u32(S.W2), 16, 1024);
      const uint64_t w2_mn = make_desc_sw128(smem_u32(S.W2), WB_WBLK, 1024);
      const uint64_t a1_k = make_desc_sw128(a1, 16, 1024), a1_ones = make_desc_sw128(a1, ones - a1, 1024);
      const uint64_t dh_k = make_desc_sw128(da, 16, 1024), dh_mn = make_desc_sw128(da, WB_BLK, 1024);
      // biases inside the products: [1 ... 1] (M = 128 x K = 16 of ones, both 64-wide M blocks at the ONES rows)
      // times K-slice 0 (b1 hi, lo) / 1 (b2 hi, lo) of the bias tile
      const uint64_t ones_mn = make_desc_sw128(ones, 0, 1024), wb_k = make_desc_sw128(smem_u32(S.WBIAS), 16, 1024);
      for (uint32_t n = 0, t = c; t < ntl; ++n, t += 2) {
        const uint32_t par = n & 1u;
        const uint32_t xst = t % xns, xpar = (t / xns) & 1u;  // X ring stage of the tile
        const uint32_t xa = smem_u32(S.X[xst * xsb]);
        // turn of tile t - 1 (the other context's tile n - 1 for context 0, tile n for context 1)
        const uint32_t tpar = (c == 0 ? n - 1u : n) & 1u;
        const bool trc = t == 2;

        // ---- G0: h1 = X W1^T (the forward's GEMM 1, same operand order)
        mbar_wait_backoff<20>(&S.x_full[xst], xpar);
        P4C_WB_TRACE(trc && lane == 0, 21);
        tc_fence_after();
        {
          uint32_t acc = 0u;
          for (int k = 0; k < nblk; ++k) {
            uint64_t xd = make_desc_sw128(xa + k * WB_BLK, 16, 1024), wd = make_desc_sw128(smem_u32(S.W1[k]), 16, 1024);
            const int nk = (a.kb.kvalid[k] + 15) >> 4;
            for (int q = 0; q < nk; ++q, xd += KSTEP, wd += KSTEP, acc = 1u) mma_bf16_ss_elect(ch, xd, wd, IDESC_KK, acc);
          }
          mma_bf16_ss_elect(ch, ones_mn, wb_k, IDESC_MK, acc);  // + b1
        }
        mma_commit_elect(&S.h1_full[c]);
        P4C_WB_TRACE(trc && lane == 0, 16);

        // ---- G1: h2 = a1 W2^T
        mbar_wait_backoff<20>(&S.a1_full[c], par);
        P4C_WB_TRACE(trc && lane == 0, 22);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) mma_bf16_ss_elect(ch, a1_k + q * KSTEP, w2_k + q * KSTEP, IDESC_KK, q ? 1u : 0u);
        mma_bf16_ss_elect(ch, ones_mn, wb_k + KSTEP, IDESC_MK, 1u);  // + b2
        mma_commit_elect(&S.h2_full[c]);
        P4C_WB_TRACE(trc && lane == 0, 17);

        // ---- G2: da1 = dh2 W2 (W2 tile [o][h]: K = o along the rows, N = h contiguous); G3: dW2 += [a1 | 1]^T dh2
        mbar_wait_backoff<20>(&S.dh2_full[c], par);
        P4C_WB_TRACE(trc && lane == 0, 23);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < 4; ++q) mma_bf16_ss_elect(ch, dh_k + q * KSTEP, w2_mn + q * MNSTEP, IDESC_KM, q ? 1u : 0u);
        if (t > 0) {
          mbar_wait_backoff<20>(&S.g3_turn[c ^ 1u], tpar);
          tc_fence_after();
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)  // K = 128 tile rows, 16 per MMA
          mma_bf16_ss_elect(tmem + WB_COL_DW2, a1_ones + q * ONES_STEP, dh_mn + q * MNSTEP, IDESC_MM, (t | q) ? 1u : 0u);
        mma_commit_elect(&S.g3_turn[c]);  // COMPLETION, not issue: products of different issuing warps are not
                                          // executed in issue order (measured: dW2 varied in its last bits)
        mma_commit_elect(&S.da1_full[c]);
        P4C_WB_TRACE(trc && lane == 0, 18);

        // ---- G4: dW1 (K-blocks 0, 1) += [X0 | X1]^T dh1; dW1 (K-block 2), db1: [X2 | 1]^T dh1 or [1 | 1]^T dh1
        mbar_wait_backoff<20>(&S.dh1_full[c], par);
        P4C_WB_TRACE(trc && lane == 0, 24);
        if (a.dh1_tma && lane == 0) {  // the DH tile holds dh1 (fenced for the async proxy by its writers)
          uint32_t blk, bs;
          tile_coords(t, blk, bs);
          tma_store_3d(&a.tmap_dh1, da, 0, static_cast<int>(blk * WB_TM), static_cast<int>(bs));
          tma_store_commit();
        }
        __syncwarp();
        if (t > 0) mbar_wait_backoff<20>(&S.g4_turn[c ^ 1u], tpar);
        tc_fence_after();
        {
          const uint64_t x01 = make_desc_sw128(xa, WB_BLK, 1024);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            mma_bf16_ss_elect(tmem + WB_COL_DW1A, x01 + q * MNSTEP, dh_mn + q * MNSTEP, IDESC_MM, (t | q) ? 1u : 0u);
          const uint64_t x2o = nblk == 3 ? make_desc_sw128(xa + 2 * WB_BLK, ones - (xa + 2 * WB_BLK), 1024)
                                         : make_desc_sw128(ones, 0, 1024);
          const long long x2step = nblk == 3 ? ONES_STEP : 0ll;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            mma_bf16_ss_elect(tmem + WB_COL_DW1B, x2o + q * x2step, dh_mn + q * MNSTEP, IDESC_MM, (t | q) ? 1u : 0u);
        }
        if (!bsum) mma_commit_elect(&S.g4_turn[c]);
        mma_commit_elect(&S.x_empty[xst]);
        P4C_WB_TRACE(trc && lane == 0, 19);

        // ---- G5: dX products into the dX scratch: the context's own (free once its previous tile was read) or
        //      the pair shared by the contexts (free once tile t - 1 was read)
        if (nslot > 0) {
          if (dx_shared) {
            if (t > 0) mbar_wait_backoff<20>(&S.dxs_empty[0], (t - 1) & 1u);
          } else if (n > 0) {
            mbar_wait_backoff<20>(&S.dxs_empty[c], (n - 1) & 1u);
          }
          tc_fence_after();
          for (int s = 0; s < nslot; ++s) {
            const uint64_t wd = make_desc_sw128(smem_u32(S.W1[a.slot_blk[s]]), WB_WBLK, 1024);
#pragma unroll
            for (int q = 0; q < 4; ++q)  // W1 tile [n][c]: K = n along the rows, N = c contiguous
              mma_bf16_ss_elect(dxc + s * WB_HID, dh_k + q * KSTEP, wd + q * MNSTEP, IDESC_KM, q ? 1u : 0u);
          }
        }
        // ---- batch sum: BS (+)= dh1 W1[:, block 0], accumulated over the samples of the row block (tiles of both
        //      contexts: the turn barrier, committed after this product, keeps them in tile order; the first sample
        //      overwrites once the previous block's sum was read)
        if (bsum) {
          const uint32_t run = nbatch == 1 ? t : static_cast<uint32_t>(__umul64hi(static_cast<unsigned long long>(t), a.div_magic));
          const uint32_t bs = t - run * nbatch;
          if (bs == 0 && run > 0) mbar_wait_backoff<20>(&S.bs_empty, (run - 1) & 1u);
          tc_fence_after();
          const uint64_t wd = make_desc_sw128(smem_u32(S.W1[0]), WB_WBLK, 1024);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            mma_bf16_ss_elect(tmem + WB_COL_BS, dh_k + q * KSTEP, wd + q * MNSTEP, IDESC_KM, (bs | q) ? 1u : 0u);
          mma_commit_elect(&S.g4_turn[c]);
        }
        // the compute group overwrites the DH tile after dx_full (dX staging) or after a later commit of this warp
        // (next tile's dh2): the TMA store must have read it by then
        if (a.dh1_tma && lane == 0) tma_store_wait_read();
        __syncwarp();
        mma_commit_elect(&S.dx_full[c]);
        P4C_WB_TRACE(trc && lane == 0, 20);
      }
      if (a.dh1_tma && lane == 0) tma_store_wait_all();
    }
  } else {
    // ================================ compute ===============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" : : "n"(WB_REGS_COMPUTE));
    const int grp = warp >> 3, quad = warp & 3, hsel = (warp >> 2) & 1;
    const int row = quad * 32 + lane, col0 = hsel * 32;
    const int sub = lane >> 2, j = hsel * 4 + (lane & 3);  // store mapping: row 8 i + sub of the warp, chunk j
    const uint32_t full = a.full_rows;
    const uint32_t a1s = smem_u32(S.A1[grp]), dhs = smem_u32(S.DH[grp]);
    const uint32_t tm_lane = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t tm_ch = tm_lane + WB_COL_CH + grp * WB_HID + col0;
    const uint32_t tm_dx = tm_lane + WB_COL_DX + (dx_shared ? 0u : grp * WB_HID) + col0;
    const int pair_bar = 1 + grp * 4 + quad;  // named barrier of the two warps that share these rows
    const float* const gamma = S.vec[2] + col0;
    const __nv_bfloat16* const dya = reinterpret_cast<const __nv_bfloat16*>(g.dy_a);
    const __nv_bfloat16* const dyb = reinterpret_cast<const __nv_bfloat16*>(g.dy_b);
    __nv_bfloat16* const dh1_out = reinterpret_cast<__nv_bfloat16*>(g.dh1_out);
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    pdl_wait();
    P4C_WB_KTRACE(50);
    for (uint32_t t = grp; t < ntl; t += 2) {
      const uint32_t par = (t >> 1) & 1u;
      uint32_t blk, bsamp;
      tile_coords(t, blk, bsamp);
      const uint32_t rr = blk * WB_TM + row;
      const bool valid = rr < rows;
      const int32_t orow = valid ? static_cast<int32_t>(bsamp * full + rr) : -1;  // row in the output-like arrays

      // dy part-row: only PREFETCHED (L2) here.  The loads themselves are issued after the proxy fence of
      // phase a: fence.proxy.async is a MEMBAR.ALL.CTA in SASS and waits for every outstanding global access
      // of the thread, i.e. an HBM-latency load in flight would stall the hand-over of a1 to the MMA warp
      // (measured: phase a 1 750 -> 3 900 cycles when dy are identity rows of a tensor larger than L2).
      const __nv_bfloat16* dy_row = dya;
      uint4 dyr[4];
      if (valid) {
        const long long ia = g.dy_a_index ? static_cast<long long>(__ldg(g.dy_a_index + rr)) : static_cast<long long>(rr);
        dy_row = dya + static_cast<long long>(bsamp) * g.dy_a_batch_stride + ia * WB_HID + col0;
        wb_prefetch_l2(dy_row);
        if (dyb) wb_prefetch_l2(dyb + static_cast<long long>(orow) * WB_HID + col0);
      }

      const bool trc = t == 2 && tid == 0;
      P4C_WB_TRACE(trc, 0);
      // ---- a. h1 -> a1 = SiLU(h1) -> A1 tile; SiLU'(h1) -> registers
      mbar_wait_suspend(&S.h1_full[grp], par);
      P4C_WB_TRACE(trc, 1);
      tc_fence_after();
      uint4 gs[4];
      {
        float v[32];
        tmem_ld32(tm_ch, v);
        tmem_ld_wait();
        P4C_WB_TRACE(trc, 13);
        tc_fence_before();
        const f32x2 half2 = f2_dup(0.5f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float h[8];
          const uint4 hb = ws_pack8(v + c * 8);  // b1 came with the product; the forward kept h1 in bf16: same rounding
          ws_unpack8(hb, h);
          float sg8[8];
#pragma unroll
          for (int q = 0; q < 8; q += 2) {  // two elements per instruction (fp32 pairs): the forward's operations
            const f32x2 hp = f2_pack(h[q], h[q + 1]);
            float t0, t1;
            f2_unpack(f2_mul(hp, half2), t0, t1);
            asm("tanh.approx.f32 %0, %0;" : "+f"(t0));
            asm("tanh.approx.f32 %0, %0;" : "+f"(t1));
            const f32x2 sg = f2_fma(half2, f2_pack(t0, t1), half2);  // sigmoid
            const f32x2 ap = f2_mul(hp, sg);                         // a1 = SiLU(h1)
            f2_unpack(f2_fma(sg, f2_sub(hp, ap), sg), sg8[q], sg8[q + 1]);  // s (1 + h (1 - s))
            f2_unpack(ap, h[q], h[q + 1]);
          }
          sts128(a1s + sw128_offset(row, hsel * 4 + c), ws_pack8(h));
          gs[c] = ws_pack8(sg8);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.a1_full[grp]);
      P4C_WB_TRACE(trc, 2);

      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) dyr[c] = *reinterpret_cast<const uint4*>(dy_row + c * 8);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) dyr[c] = zero4;
      }
      // dy += dy_b (identity rows; edge-update / residual branches)
      if (dyb && valid) {
        const __nv_bfloat16* dyb_row = dyb + static_cast<long long>(orow) * WB_HID + col0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float d[8], e[8];
          ws_unpack8(dyr[c], d);
          ws_unpack8(*reinterpret_cast<const uint4*>(dyb_row + c * 8), e);
#pragma unroll
          for (int q = 0; q < 8; ++q) d[q] += e[q];
          dyr[c] = ws_pack8(d);
        }
      }

      // ---- c. LayerNorm backward from the recomputed h2
      mbar_wait_suspend(&S.h2_full[grp], par);
      P4C_WB_TRACE(trc, 3);
      tc_fence_after();
      {
        float xh[32];
        tmem_ld32(tm_ch, xh);
        tmem_ld_wait();
        P4C_WB_TRACE(trc, 14);
        tc_fence_before();
        // fp32 pairs: two elements per instruction; the four row sums are kept as (even, odd) column pairs
        f32x2 X2[16];
        f32x2 s1p = 0ull, s2p = 0ull, s3p = 0ull, s4p = 0ull;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float d[8];
          ws_unpack8(dyr[c], d);
          const float4 ga = *reinterpret_cast<const float4*>(gamma + c * 8), gb = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
          const f32x2 gp[4] = {f2_pack(ga.x, ga.y), f2_pack(ga.z, ga.w), f2_pack(gb.x, gb.y), f2_pack(gb.z, gb.w)};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f32x2 x = f2_pack(xh[c * 8 + 2 * q], xh[c * 8 + 2 * q + 1]);  // h2 (b2 came with the product)
            X2[c * 4 + q] = x;
            const f32x2 dg = f2_mul(f2_pack(d[2 * q], d[2 * q + 1]), gp[q]);
            s1p = f2_add(s1p, x);
            s2p = f2_fma(x, x, s2p);
            s3p = f2_add(s3p, dg);
            s4p = f2_fma(dg, x, s4p);
          }
        }
        const float s1 = f2_hsum(s1p), s2 = f2_hsum(s2p), s3 = f2_hsum(s3p), s4 = f2_hsum(s4p);
        S.ln_part[grp][hsel][row] = make_float4(s1, s2, s3, s4);
        asm volatile("bar.sync %0, 64;" : : "r"(pair_bar) : "memory");
        const float4 p0 = S.ln_part[grp][0][row], p1 = S.ln_part[grp][1][row];  // same order in both threads
        const float mean = (p0.x + p1.x) * (1.0f / WB_HID);
        const float var = fmaxf((p0.y + p1.y) * (1.0f / WB_HID) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + m.ln_eps);
        const float t3 = p0.z + p1.z, t4 = p0.w + p1.w;
        // dh2 = rstd (g dy - mean_c(g dy) - xhat mean_c(g dy xhat)),  xhat = (h2 - mean) rstd
        const float nmr = -mean * rstd;
        const float k1 = -(t3 * (1.0f / WB_HID)) * rstd;
        const float k2 = -(rstd * (t4 - mean * t3) * (1.0f / WB_HID)) * rstd;
        const f32x2 rstd2 = f2_dup(rstd), nmr2 = f2_dup(nmr), k12 = f2_dup(k1), k22 = f2_dup(k2);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float d[8], o[8];
          ws_unpack8(dyr[c], d);
          const float4 ga = *reinterpret_cast<const float4*>(gamma + c * 8), gb = *reinterpret_cast<const float4*>(gamma + c * 8 + 4);
          const f32x2 gp[4] = {f2_pack(ga.x, ga.y), f2_pack(ga.z, ga.w), f2_pack(gb.x, gb.y), f2_pack(gb.z, gb.w)};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const f32x2 dp = f2_pack(d[2 * q], d[2 * q + 1]);
            const f32x2 xhat = f2_fma(X2[c * 4 + q], rstd2, nmr2);
            const f32x2 dg = f2_mul(dp, gp[q]);
            f2_unpack(f2_fma(xhat, k22, f2_fma(dg, rstd2, k12)), o[2 * q], o[2 * q + 1]);
            f2_unpack(f2_mul(dp, xhat), xh[c * 8 + 2 * q], xh[c * 8 + 2 * q + 1]);  // dgamma contribution
          }
          sts128(dhs + sw128_offset(row, hsel * 4 + c), ws_pack8(o));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.dh2_full[grp]);
        P4C_WB_TRACE(trc, 4);
        // column sums over the rows of the warp while the tensor core runs G2 / G3
#if !(P4C_WB_ABLATE & 1)
        vacc0 += warp_colsum<32>(xh, lane);
#pragma unroll
        for (int c = 0; c < 4; ++c) ws_unpack8(dyr[c], xh + c * 8);
        vacc1 += warp_colsum<32>(xh, lane);
#else
        vacc0 += xh[lane & 31];
#endif
        P4C_WB_TRACE(trc, 5);
      }

      // ---- e. dh1 = da1 * SiLU'(h1) -> DH tile (overwrites dh2: G2, G3 are complete), then to HBM
      mbar_wait_suspend(&S.da1_full[grp], par);
      P4C_WB_TRACE(trc, 6);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(tm_ch, v);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 p = ws_pack8(v + c * 8);  // da1 -> bf16, times SiLU'(h1) in bf16x2
          p.x = bf2_mul(p.x, gs[c].x);
          p.y = bf2_mul(p.y, gs[c].y);
          p.z = bf2_mul(p.z, gs[c].z);
          p.w = bf2_mul(p.w, gs[c].w);
          if (!valid) p = zero4;
          sts128(dhs + sw128_offset(row, hsel * 4 + c), p);
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.dh1_full[grp]);
      P4C_WB_TRACE(trc, 7);
      if (dh1_out && !a.dh1_tma && !(P4C_WB_ABLATE & 2)) {  // coalesced: lane -> 16-byte chunk j of row 8 i + sub of this warp's rows
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + sub;
          const int32_t orr = __shfl_sync(0xffffffffu, orow, r);
          const uint4 u = lds128(dhs + sw128_offset(quad * 32 + r, j));
          if (orr >= 0) *reinterpret_cast<uint4*>(dh1_out + static_cast<long long>(orr) * WB_HID + j * 8) = u;
        }
      }

      // the group-summed receiver gradient is accumulated in place: pull this lane's slice of the row into L1
      // while G4 / G5 run (the read-modify-write below would otherwise expose the load latency)
      const int gsz = g.dx_group;
      int cbase = 0;
      uint4 grp_old = zero4;
      if (g.dx_group_out) {
        if (gsz > 1 && (lane & 1)) cbase += 16;
        if (gsz > 2 && (lane & 2)) cbase += 8;
        if (gsz > 4 && (lane & 4)) cbase += 4;
        if (valid) {
          const __nv_bfloat16* old = reinterpret_cast<const __nv_bfloat16*>(g.dx_group_out) +
                                     static_cast<long long>(orow / gsz) * WB_HID + col0 + cbase;
          // groups of 4 / 8 rows: the lane's slice is one 16 / 8-byte word -- LOADED now (it is only this thread's
          // to update), so that the read-modify-write after the MMAs does not wait for it; wider slices are
          // pulled into L1
          if (gsz == 4) grp_old = *reinterpret_cast<const uint4*>(old);
          else if (gsz == 8) *reinterpret_cast<uint2*>(&grp_old) = *reinterpret_cast<const uint2*>(old);
          else asm volatile("prefetch.global.L1 [%0];" : : "l"(old));
        }
      }
      // ---- g. dX epilogues (G4 / G5 complete: the DH and X tiles of the context are free afterwards)
      P4C_WB_TRACE(trc, 8);
      mbar_wait_suspend(&S.dx_full[grp], par);
      P4C_WB_TRACE(trc, 9);
      tc_fence_after();
      if (bsum && bsamp == nbatch - 1) {
        // last sample of the row block: the batch-summed gradient of the shared src[0] rows (every product into BS
        // completed before this tile's dx_full); released first, then staged like a stored dX slot
        float v[32];
        tmem_ld32(tm_lane + WB_COL_BS + col0, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.bs_empty);
#pragma unroll
        for (int c = 0; c < 4; ++c) sts128(dhs + sw128_offset(row, hsel * 4 + c), ws_pack8(v + c * 8));
        __syncwarp();
        __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(g.dx0_batch_sum);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = quad * 32 + 8 * i + sub;
          const uint32_t grow = blk * WB_TM + r;
          const uint4 u = lds128(dhs + sw128_offset(r, j));
          if (grow < rows) *reinterpret_cast<uint4*>(base + static_cast<long long>(grow) * WB_HID + j * 8) = u;
        }
        __syncwarp();
      }
      for (int s = 0; s < ((P4C_WB_ABLATE & 4) ? 0 : nslot); ++s) {
        float v[32];
        tmem_ld32(tm_dx + s * WB_HID, v);
        tmem_ld_wait();
        P4C_WB_TRACE(trc && s == 0, 11);
        const int kblk = a.slot_blk[s];
        if (a.slot_kind[s] == 0) {
          const int src = a.kb.src[kblk];
          const int ld = g.dx_ld[src];
          const int nch = a.kb.nchunks[kblk];
          const bool add_dy = kblk == 0 && g.dx0_add_dy != 0;
          if (!g.dx_accumulate[src]) {
            // bf16 rows staged in the (free) DH tile, then stored with complete 64-byte row segments per
            // lane group; the residual pass-through d src0 += dy comes from the registers that still hold dy
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float* o = v + c * 8;
              float e[8];
              if (add_dy && valid) {
                if (g.dx0_add_dy == 2) {
                  if (dyb) ws_unpack8(*reinterpret_cast<const uint4*>(dy_row + c * 8), e);  // dyr holds dy_a + dy_b
                  else ws_unpack8(dyr[c], e);
#pragma unroll
                  for (int q = 0; q < 8; ++q) o[q] += e[q];
                }
                if (dyb) {
                  ws_unpack8(*reinterpret_cast<const uint4*>(dyb + static_cast<long long>(orow) * WB_HID + col0 + c * 8), e);
#pragma unroll
                  for (int q = 0; q < 8; ++q) o[q] += e[q];
                }
              }
              sts128(dhs + sw128_offset(row, hsel * 4 + c), ws_pack8(o));
            }
            __syncwarp();
            __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(g.dx[src]) + a.kb.col0[kblk];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 8 * i + sub;
              const int32_t orr = __shfl_sync(0xffffffffu, orow, r);
              const uint4 u = lds128(dhs + sw128_offset(quad * 32 + r, j));
              if (orr >= 0 && j < nch) *reinterpret_cast<uint4*>(base + static_cast<long long>(orr) * ld + j * 8) = u;
            }
            __syncwarp();
          } else if (valid) {
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.dx[src]) + a.kb.col0[kblk] +
                                 static_cast<long long>(orow) * ld + col0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (hsel * 4 + c >= nch) continue;  // beyond the (padded) width of this block
              float* o = v + c * 8;
              float e[8];
              if (g.dx_accumulate[src]) {
                ws_unpack8(*reinterpret_cast<const uint4*>(dst + c * 8), e);
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] += e[q];
              }
              if (add_dy) {  // residual / edge-update pass-through: d src0 += dy (dy_b only when dx0_add_dy == 1)
                if (g.dx0_add_dy == 2) {
                  ws_unpack8(*reinterpret_cast<const uint4*>(dy_row + c * 8), e);
#pragma unroll
                  for (int q = 0; q < 8; ++q) o[q] += e[q];
                }
                if (dyb) {
                  ws_unpack8(*reinterpret_cast<const uint4*>(dyb + static_cast<long long>(orow) * WB_HID + col0 + c * 8), e);
#pragma unroll
                  for (int q = 0; q < 8; ++q) o[q] += e[q];
                }
              }
              *reinterpret_cast<uint4*>(dst + c * 8) = ws_pack8(o);
            }
          }
        } else {
          // sum over the `dx_group` consecutive rows (= lanes) that gather the same source row:
          // transposing butterfly (fp32, fixed order; rows beyond the problem hold zeros)
#define P4C_WB_GRP_STEP(OFF, N)                                              \
          if (gsz > OFF) {                                                   \
            const bool up = (lane & OFF) != 0;                               \
            _Pragma("unroll") for (int i = 0; i < (N) / 2; ++i) {            \
              const float send = up ? v[i] : v[i + (N) / 2];                 \
              const float keep = up ? v[i + (N) / 2] : v[i];                 \
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);         \
            }                                                                \
          }
          P4C_WB_GRP_STEP(1, 32)
          P4C_WB_GRP_STEP(2, 16)
          P4C_WB_GRP_STEP(4, 8)
#undef P4C_WB_GRP_STEP
          if (valid) {
            const int ncol = 32 / gsz;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.dx_group_out) +
                                 static_cast<long long>(orow / gsz) * WB_HID + col0 + cbase;
            if (ncol >= 8) {
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (c < ncol / 8) {
                  float e[8];
                  ws_unpack8(gsz == 4 ? grp_old : *reinterpret_cast<const uint4*>(dst + c * 8), e);
#pragma unroll
                  for (int q = 0; q < 8; ++q) e[q] += v[c * 8 + q];
                  *reinterpret_cast<uint4*>(dst + c * 8) = ws_pack8(e);
                }
              }
            } else {  // group of 8: four columns per lane
              const uint2 u = make_uint2(grp_old.x, grp_old.y);
              float e0, e1, e2, e3;
              ws_unpack2(u.x, e0, e1);
              ws_unpack2(u.y, e2, e3);
              uint2 o;
              o.x = pack_bf16x2(e0 + v[0], e1 + v[1]);
              o.y = pack_bf16x2(e2 + v[2], e3 + v[3]);
              *reinterpret_cast<uint2*>(dst) = o;
            }
          }
        }
      }
      P4C_WB_TRACE(trc, 12);
      if (nslot > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.dxs_empty[dx_shared ? 0 : grp]);
      }
      P4C_WB_TRACE(trc, 10);
#ifdef P4C_PHASE_TRACE
      P4C_WB_TRACE(tid == 0 && t < 12, 58 + (t >> 1));  // end of the group's tiles 0, 2, 4, ...
#endif
    }
  }

  // ---- every tile of the CTA is done (each group waited for the commit that follows its last MMAs)
  P4C_WB_KTRACE(51);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  P4C_WB_KTRACE(52);

  // per-CTA partials: weight and bias gradients from TMEM (layout of the classic engine)
  const int k1pad = (m.k1 + 31) / 32 * 32;
  const long long pf = static_cast<long long>(k1pad) * WB_HID + WB_HID * WB_HID + 4 * WB_HID;
  float* const part = g.wgrad_partials + static_cast<long long>(blockIdx.x) * pf;
  float* const pvec = part + static_cast<long long>(k1pad) * WB_HID + WB_HID * WB_HID;  // db1, db2, dgamma, dbeta
  const bool padd = g.partials_mode == 2;
  if (warp < 8) {
    const bool need_w = g.partials_mode != 0;
    const bool need_w1 = need_w || g.g_w1 != nullptr, need_w2 = need_w || g.g_w2 != nullptr;
    const int quad = warp & 3, hsel = warp >> 2;
    const int row = quad * 32 + lane, col0 = hsel * 32;
    const uint32_t tm_row = tmem + (static_cast<uint32_t>(quad * 32) << 16) + col0;
    float* dstA = nullptr;
    float* dstB = nullptr;
    float* dstC = nullptr;
    {
      const int blk = row >> 6, kk = row & 63;
      if (need_w1 && blk < nblk && blk < 2 && kk < a.kb.kvalid[blk])
        dstA = part + static_cast<long long>(a.kb.kstart[blk] + kk) * WB_HID + col0;
      if (need_w1 && nblk == 3 && row < a.kb.kvalid[2]) dstB = part + static_cast<long long>(a.kb.kstart[2] + row) * WB_HID + col0;
      if (need_w2 && row < WB_HID) dstC = part + static_cast<long long>(k1pad) * WB_HID + static_cast<long long>(row) * WB_HID + col0;
      if (row == WB_HID) {
        dstB = pvec + col0;           // db1
        dstC = pvec + WB_HID + col0;  // db2
      }
    }
    auto put32 = [&](float* dst, const float (&v)[32]) {
      if (!dst) return;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float4 e = padd ? *reinterpret_cast<const float4*>(dst + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(dst + 4 * q) =
            make_float4(v[4 * q] + e.x, v[4 * q + 1] + e.y, v[4 * q + 2] + e.z, v[4 * q + 3] + e.w);
      }
    };
    float v[32];
    tmem_ld32(tm_row + WB_COL_DW1A, v);  // dW1[k][c] for K-blocks 0, 1 (TMEM lane = 64 * block + k)
    tmem_ld_wait();
    put32(dstA, v);
    tmem_ld32(tm_row + WB_COL_DW1B, v);
    tmem_ld_wait();
    put32(dstB, v);
    tmem_ld32(tm_row + WB_COL_DW2, v);
    tmem_ld_wait();
    put32(dstC, v);
  } else if (warp < WB_CWARPS && g.partials_mode == 1) {
    // the workspace of this MLP may hold more blocks than this grid writes (sized for the classic
    // engine's grid): they are summed by the deferred reduction, so they must read as zeros
    const int t8 = tid - 8 * 32;
    for (int bidx = blockIdx.x + gridDim.x; bidx < a.total_blocks; bidx += gridDim.x) {
      float4* z = reinterpret_cast<float4*>(g.wgrad_partials + static_cast<long long>(bidx) * pf);
      for (long long i = t8; i < pf / 4; i += 8 * 32) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  // dgamma / dbeta: cross-warp sums in a fixed order (the A1 tile of context 0 serves as scratch)
  float (*red)[2][32] = reinterpret_cast<float (*)[2][32]>(S.A1[0]);  // [warp][vector][lane]
  if (warp < WB_CWARPS) {
    red[warp][0][lane] = vacc0;
    red[warp][1][lane] = vacc1;
  }
  __syncthreads();
  if (tid < WB_HID) {
    // column tid: column half h = tid / 32 -> warps with (w / 4) % 2 == h, i.e. w = 4 h + q (group 0) and 8 + 4 h + q
    const int h = tid >> 5, l = tid & 31, w0 = 4 * h;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const float s0 = (red[w0][v][l] + red[w0 + 1][v][l]) + (red[w0 + 2][v][l] + red[w0 + 3][v][l]);
      const float s1 = (red[w0 + 8][v][l] + red[w0 + 9][v][l]) + (red[w0 + 10][v][l] + red[w0 + 11][v][l]);
      float* d = pvec + (2 + v) * WB_HID + tid;  // dgamma, dbeta
      *d = padd ? *d + (s0 + s1) : (s0 + s1);
    }
  }
  P4C_WB_KTRACE(53);
  tc_fence_before();
  __syncthreads();
  if (warp == WB_MMA_WARP) tmem_dealloc(tmem, WB_TMEM_COLS);
  P4C_WB_KTRACE(54);
#ifdef P4C_PHASE_TRACE
  if (a.trace && blockIdx.x == 0 && tid == 0) a.trace[57] = static_cast<long long>(wb_globaltimer());
#endif
}

}  // namespace tc

bool tc_kblocks(const p4c_mlp_t* m, tc::KBlocks* kb);

static long long g_wb_min_tiles = -1;
static long long g_wb_launches = 0;
static long long g_wb_bsum_launches = 0;
static long long wb_min_tiles() {
  if (g_wb_min_tiles < 0) {
    const char* e = getenv("P4C_WS_BWD_MIN_TILES");
    g_wb_min_tiles = (e && atoll(e) > 0) ? atoll(e) : 2ll * kNumSMs;
  }
  return g_wb_min_tiles;
}

// Shape part of the decision (a function of the problem alone, so that the forward can know whether
// the backward will recompute h1): bf16, LayerNorm, 64 outputs, packed weights, plain (non-segment)
// sources, at least wb_min_tiles() tiles.
bool ws_backward_shape_ok(const p4c_mlp_t* m) {
  if (!tc_engine_enabled() || !tc_forward_supported(m)) return false;
  if (!m->packed || m->out_features != P4C_HIDDEN || !m->layer_norm) return false;
  for (int s = 0; s < m->nsrc; ++s)
    if (m->src[s].seg_rowptr) return false;
  const long long full = m->full_rows > 0 ? m->full_rows : m->rows;
  const long long total = static_cast<long long>(m->rows) * m->batch;
  if (total >= (1ll << 31) - 256 * 1024 || full * m->batch >= (1ll << 31)) return false;
  const long long ntiles = static_cast<long long>((m->rows + 127) / 128) * m->batch;
  if (ntiles < wb_min_tiles()) return false;
  // at most two dX products per tile: every K-block of an identity-indexed source may need one, the
  // gathered sources share one (dx_group_out) -- decided here, on the shape alone, so that the answer
  // the forward gets (save h1 or not) cannot be contradicted by the gradients the backward asks for
  tc::KBlocks kb;
  tc_kblocks(m, &kb);
  int ident = 0, gathered = 0;
  for (int k = 0; k < kb.nblk; ++k) {
    if (m->src[kb.src[k]].index) gathered = 1; else ++ident;
  }
  if (ident + gathered > 2) return false;
  return true;
}

// dx0_batch_sum (a function of the shape): src[0] is a plain 64-wide row source shared by the batch, and every CTA
// gets at least one whole row block (the samples of a block run back to back on one CTA)
bool ws_backward_bsum_ok(const p4c_mlp_t* m) {
  if (!ws_backward_shape_ok(m)) return false;
  const p4c_src_t& s0 = m->src[0];
  if (s0.index || s0.seg_rowptr || s0.width != P4C_HIDDEN || s0.batch_stride != 0) return false;
  return (m->rows + 127) / 128 >= kNumSMs;
}

static int wb_slots(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, const tc::KBlocks& kb, int* blk, int* kind) {
  int n = 0;
  for (int k = 0; k < kb.nblk; ++k)
    if (g->dx[kb.src[k]]) {
      if (n < 2) { blk[n] = k; kind[n] = 0; }
      ++n;
    }
  if (g->dx_group_out) {
    for (int k = 0; k < kb.nblk; ++k)
      if (kb.src[k] == g->dx_group_src) {
        if (n < 2) { blk[n] = k; kind[n] = 1; }
        ++n;
      }
  }
  return n;
}

bool ws_backward_supported(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g) {
  if (!ws_backward_shape_ok(m)) return false;
  if (g->dy_a_scale) return false;
  tc::KBlocks kb;
  tc_kblocks(m, &kb);
  for (int s = 0; s < m->nsrc; ++s)
    if (g->dx[s] && (m->src[s].index || g->dx_ld[s] % 8 != 0 || (reinterpret_cast<uintptr_t>(g->dx[s]) % 16) != 0))
      return false;
  if (g->dx_group_out) {
    const int gs = g->dx_group;
    const long long full = m->full_rows > 0 ? m->full_rows : m->rows;
    if (!(gs == 1 || gs == 2 || gs == 4 || gs == 8) || m->rows % gs != 0 || full % gs != 0) return false;
    if (g->dx_group_src < 0 || g->dx_group_src >= m->nsrc || m->src[g->dx_group_src].width != P4C_HIDDEN) return false;
    if (g->dx[g->dx_group_src] || (reinterpret_cast<uintptr_t>(g->dx_group_out) % 16) != 0) return false;
  }
  if (g->dx0_batch_sum) {  // 64-wide identity-indexed src[0] shared by the batch, its own gradient not requested
    if (!ws_backward_bsum_ok(m) || g->dx[0] || (reinterpret_cast<uintptr_t>(g->dx0_batch_sum) % 16) != 0) return false;
  }
  int blk[2], kind[2];
  if (wb_slots(m, g, kb, blk, kind) > 2) return false;
  if (g->dy_a_batch_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(g->dy_a) % 16) != 0) return false;
  if (g->dy_b && (reinterpret_cast<uintptr_t>(g->dy_b) % 16) != 0) return false;
  if (g->dh1_out && (reinterpret_cast<uintptr_t>(g->dh1_out) % 16) != 0) return false;
  if ((reinterpret_cast<uintptr_t>(g->wgrad_partials) % 16) != 0) return false;
  return true;
}

int ws_backward_grid(const p4c_mlp_t* m) {
  const long long ntiles = static_cast<long long>((m->rows + 127) / 128) * m->batch;
  return static_cast<int>(ntiles < kNumSMs ? ntiles : kNumSMs);
}

int launch_mlp2_ws_backward(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, int total_blocks, cudaStream_t st) {
  using namespace tc;
  WbArgs a;
  a.m = *m;
  a.g = *g;
  tc_kblocks(m, &a.kb);
  a.slot_blk[0] = a.slot_blk[1] = 0;
  a.slot_kind[0] = a.slot_kind[1] = 0;
  a.nslot = wb_slots(m, g, a.kb, a.slot_blk, a.slot_kind);
  a.tma_mask = 0;
  for (int k = 0; k < a.kb.nblk; ++k) {
    const p4c_src_t& src = m->src[a.kb.src[k]];
    a.tma_b[k] = (m->batch > 1 && src.batch_stride != 0) ? 1 : 0;
    if (tma_enabled() && !src.index && !src.seg_rowptr &&
        tma_rows_map(&a.tmap[k], src.ptr, src.ld, m->rows, src.ld, m->batch, src.batch_stride))
      a.tma_mask |= 1 << k;
  }
  a.zero_x = 0;
  for (int k = 0; k < a.kb.nblk; ++k) a.zero_x |= a.kb.nchunks[k] < 8 && !(a.tma_mask >> k & 1);
  a.bsum = g->dx0_batch_sum ? 1 : 0;
  a.nblocks = static_cast<uint32_t>((m->rows + WB_TM - 1) / WB_TM);
  a.dh1_tma = 0;
  if (g->dh1_out && tma_enabled()) {
    const long long fr = m->full_rows > 0 ? m->full_rows : m->rows;
    if (tma_rows_map(&a.tmap_dh1, g->dh1_out, P4C_HIDDEN, m->rows, P4C_HIDDEN, m->batch, fr * P4C_HIDDEN)) a.dh1_tma = 1;
  }
  const unsigned long long rows = static_cast<unsigned long long>(m->rows);
  const unsigned long long nb = static_cast<unsigned long long>(m->batch);
  a.div_magic = nb == 1 ? 0ull : (~0ull / nb) + 1ull;
  a.batch = static_cast<uint32_t>(m->batch);
  a.ntiles = static_cast<uint32_t>((rows + WB_TM - 1) / WB_TM * nb);
  a.full_rows = static_cast<uint32_t>(m->full_rows > 0 ? m->full_rows : m->rows);
  a.total_blocks = total_blocks;
  a.trace = tc_get_trace();
  {  // developer aid: P4C_WB_TRACE_NBLK=n restricts the phase trace to calls with n K-blocks
    static int want = -1;
    if (want < 0) {
      const char* e = getenv("P4C_WB_TRACE_NBLK");
      want = e ? atoi(e) : 0;
    }
    if (want > 0 && a.kb.nblk != want) a.trace = nullptr;
  }
  const int grid = ws_backward_grid(m);  // bsum: nblocks >= kNumSMs, i.e. kNumSMs CTAs
  const size_t smem = sizeof(WbSmem) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp2_ws_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mlp_backward(ws): cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    configured = true;
  }
  launch_pdl(mlp2_ws_bwd_kernel, dim3(grid), dim3(WB_THREADS), smem, st, a);
  P4C_CHECK_LAUNCH("mlp_backward(ws)");
  ++g_wb_launches;
  if (a.bsum) ++g_wb_bsum_launches;
  return P4C_OK;
}

}  // namespace p4c

extern "C" int32_t p4c_mlp_bwd_caps(const p4c_mlp_t* m) {
  if (!m || m->rows <= 0 || m->batch <= 0) return 0;
  return p4c::ws_backward_shape_ok(m) ? (1 | 4 | (p4c::ws_backward_bsum_ok(m) ? 2 : 0)) : 0;
}
extern "C" void p4c_set_ws_bwd_min_tiles(int64_t n) { p4c::g_wb_min_tiles = n > 0 ? n : -1; }
extern "C" int64_t p4c_get_ws_bwd_min_tiles(void) { return p4c::wb_min_tiles(); }
extern "C" int64_t p4c_ws_bwd_launch_count(void) { return p4c::g_wb_launches; }
extern "C" int64_t p4c_ws_bwd_bsum_launch_count(void) { return p4c::g_wb_bsum_launches; }
