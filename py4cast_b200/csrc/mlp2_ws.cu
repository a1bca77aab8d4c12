// Warp-specialised, software-pipelined FORWARD of the fused 2-layer MLP for the large edge / grid
// problems (bf16 storage): the same operator as mlp2_tc_fwd_kernel (mlp2_tc.cu), re-organised so
// that gather, tensor-core work and the two epilogues of DIFFERENT tiles overlap inside one CTA.
//
// One persistent CTA per SM, 21 warps:
//   warps 16-19 producer   cp.async gather of the up-to-three 64-column K-blocks of a 128-row tile into
//                          a 3-stage ring of 128-byte-swizzled smem tiles; row offsets are computed
//                          one per lane and handed out with shuffles so that every cp.async
//                          instruction fetches four complete 128-byte rows (coalesced requests);
//                          the gather indices of the next tile are loaded one iteration ahead.
//   warp  20    MMA        one lane issues tcgen05.mma: G1(t) = X(t) W1^T -> D1[t&1],
//                          G2(t-1) = A2(t-1) W2^T -> D2[(t-1)&1]; tcgen05.commit signals the
//                          mbarriers (X stage free, D1 / D2 ready).
//   warps 0-7 / 8-15       two epilogue groups, tile t belongs to group t & 1; inside a group warp w
//                          owns TMEM lane quadrant w % 4 (rows) and column half (w / 4) % 2: thread =
//                          (tile row, 32 columns):
//        epilogue 1: D1 + b1 -> bf16 h1, staged in the A2 tile; second pass (lanes re-mapped to
//                    16-byte chunks of eight consecutive rows): coalesced store of h1 (saved
//                    activation) and SiLU in place -> A2 operand of G2.
//        epilogue 2: D2 + b2, LayerNorm (partial sums exchanged between the two threads of a row
//                    through shared memory, one 64-thread named barrier), then
//                    - fused aggregation over groups of consecutive rows (shuffle butterfly), or
//                    - messages / residual output / edge update staged through the (free) A2 tile
//                      and stored with coalesced 16-byte chunks.  The residual / edge row is the
//                      tile's own K-block 0, read from shared memory instead of HBM.
//   While one group waits for G2 of its tile, the other runs the epilogues of the next one; the
//   producer runs two tiles ahead.  No thread spins hot: waits back off with nanosleep.
//
// TMEM: 256 columns (D1[2], D2[2], 64 fp32 columns each).  Shared memory: 3 x 48 KB X ring,
// 2 x 16 KB A2, 32 KB weights (bf16 image of p4c_pack_weights), barriers  = ~209 KB.
//
// Reference semantics: mfai/neural-lam `make_mlp` (Linear, SiLU, Linear, LayerNorm) on
// cat(edge, send[sender], rec[receiver]) and InteractionNet.aggregate.
#include "common.cuh"
#include "tc_common.cuh"
#include "tma.cuh"

namespace p4c {
namespace tc {

constexpr int WS_TM = 128;
constexpr int WS_HID = P4C_HIDDEN;
constexpr int WS_XBLK = WS_TM * 128;
constexpr int WS_WBLK = WS_HID * 128;
constexpr int WS_NS = 3;   // X ring stages
constexpr int WS_LAG = 2;  // tiles the producer keeps in flight before it signals the oldest
constexpr int WS_EPI_WARPS = 16, WS_PROD_WARPS = 4;
constexpr int WS_THREADS = (WS_EPI_WARPS + WS_PROD_WARPS + 1) * 32;
constexpr uint32_t WS_TMEM_COLS = 256;
constexpr uint32_t kWsNoRow = 0xFFFFFFFFu;

struct WsSmem {
  uint8_t X[WS_NS][P4C_MAX_SRC][WS_XBLK];
  uint8_t A2[2][WS_XBLK];
  uint8_t W1[P4C_MAX_SRC][WS_WBLK];
  uint8_t W2[WS_WBLK];
  uint8_t WBIAS[WS_WBLK];  // bias tile of the weight image (PK_BIAS): b1 in K-slice 0, b2 in K-slice 1
  uint8_t ONES[2048];      // 16 K-rows x 64 M of bf16 ones: the all-ones A operand of the bias products
  float vec[4][WS_HID];  // b1, b2, gamma, beta
  float2 ln_part[2][2][WS_TM];  // [group][column half][row]: LayerNorm partial (sum, sum of squares)
  uint64_t x_full[WS_NS], x_empty[WS_NS];
  uint64_t d1_full[2], a2_full[2], d2_full[2], d2_empty[2];
  uint32_t tmem_base;
};

static_assert(sizeof(WsSmem) + 1024 <= 232448, "pipelined forward: shared memory beyond the 227 KB a CTA can opt into");

struct WsArgs {
  p4c_mlp_t m;
  KBlocks kb;
  __nv_bfloat16* out;
  __nv_bfloat16* out2;
  __nv_bfloat16* save_h1;
  float* save_stats;
  __nv_bfloat16* aggr_out;
  int agg_group;
  float agg_scale;
  int residual_x0;  // out = y + (row of K-block 0)
  int out_width;    // 64, or the width of a narrow output without LayerNorm (output map: 64 -> 64 -> 21)
  int zero_x;
  // Tile order: tile T = (row block T / batch, sample T % batch), i.e. the samples of one row block are
  // processed back to back (by neighbouring CTAs at the same time), so that a source shared by the
  // batch (static edge / mesh embeddings: batch_stride 0) is fetched from HBM once and served to the
  // other samples from L2 instead of being re-read per sample (m2g edge embedding: 168 MB > L2).
  unsigned long long div_magic;  // ceil(2^64 / batch): row block = umul64hi(tile, magic)
  uint32_t batch, ntiles, full_rows;
  // identity-indexed K-blocks arrive by TMA (bit k of tma_mask; tma_b[k] = 0 for a source shared by the batch)
  int tma_mask;
  int tma_b[P4C_MAX_SRC];
  alignas(64) CUtensorMap tmap[P4C_MAX_SRC];
};

// NARROW: output narrower than 64 columns, no LayerNorm (its store path is kept out of the main instantiation:
// the epilogue warps are sensitive to every instruction in their tile loop, executed or not)
// AGGR_ONLY: the m2g edge call (LayerNorm, aggregation in the epilogue, nothing else stored or saved).
// NO_SAVE: h1 is not saved (calls whose backward recomputes it): epilogue 1 runs in one pass.
template <bool NARROW, bool AGGR_ONLY = false, bool NO_SAVE = AGGR_ONLY>
__global__ void __launch_bounds__(WS_THREADS, 1) mlp2_ws_fwd_kernel(const __grid_constant__ WsArgs a) {
  extern __shared__ unsigned char smem_raw[];
  WsSmem& S = *reinterpret_cast<WsSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const p4c_mlp_t& m = a.m;
  // warp index broadcast from lane 0: provably warp-uniform for the compiler (MMA warp: uniform registers)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int nblk = a.kb.nblk;
  const uint32_t nbatch = a.batch;
  const uint32_t rows = static_cast<uint32_t>(m.rows);
  const bool need_x0 = !AGGR_ONLY && (a.residual_x0 || a.out2 != nullptr);

  // ---- set-up: weight image, barriers, TMEM
  if (a.zero_x) {
    uint4* xz = reinterpret_cast<uint4*>(S.X);
    for (int i = tid; i < WS_NS * P4C_MAX_SRC * WS_XBLK / 16; i += WS_THREADS) xz[i] = make_uint4(0, 0, 0, 0);
  }
  {
    const uint8_t* img = reinterpret_cast<const uint8_t*>(m.packed);
    for (int i = tid; i < nblk * (WS_WBLK / 16); i += WS_THREADS)
      cp_async16(smem_u32(S.W1) + i * 16, img + PK_W1 + i * 16, 16u);
    for (int i = tid; i < WS_WBLK / 16; i += WS_THREADS) cp_async16(smem_u32(S.W2) + i * 16, img + PK_W2 + i * 16, 16u);
    for (int i = tid; i < WS_WBLK / 16; i += WS_THREADS) cp_async16(smem_u32(S.WBIAS) + i * 16, img + PK_BIAS + i * 16, 16u);
    if (tid < 64) cp_async16(smem_u32(S.vec) + tid * 16, img + PK_VEC + tid * 16, 16u);
    cp_async_commit();
  }
  if (tid < 2048 / 16)
    reinterpret_cast<uint4*>(S.ONES)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (tid == 0) {
    for (int i = 0; i < WS_NS; ++i) {
      mbar_init(&S.x_full[i], WS_PROD_WARPS);
      mbar_init(&S.x_empty[i], need_x0 ? 1 + 8 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.d1_full[i], 1);
      mbar_init(&S.a2_full[i], 8);
      mbar_init(&S.d2_full[i], 1);
      mbar_init(&S.d2_empty[i], 8);
    }
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == WS_EPI_WARPS + WS_PROD_WARPS) tmem_alloc(&S.tmem_base, WS_TMEM_COLS);
  cp_async_wait<0>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  // this CTA's tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
  const uint32_t ntl = (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
  pdl_launch_dependents();

  if (warp >= WS_EPI_WARPS && warp < WS_EPI_WARPS + WS_PROD_WARPS) {
    // =============================== producer ===============================
    const int pw = warp - WS_EPI_WARPS;
    const int sub = lane >> 3, j = lane & 7;  // lane -> (row within a group of 4, 16-byte chunk)
    int32_t idx_next[P4C_MAX_SRC];
    uint32_t b_next = 0, rr_next = 0;
    bool valid_next = false;
    auto load_indices = [&](uint32_t t) {  // gather indices of tile t (graph constants)
      const uint32_t tile = blockIdx.x + t * gridDim.x;
      const uint32_t blk = nbatch == 1 ? tile
                                       : static_cast<uint32_t>(__umul64hi(static_cast<unsigned long long>(tile), a.div_magic));
      b_next = tile - blk * nbatch;
      rr_next = blk * WS_TM + pw * 32 + lane;
      valid_next = t < ntl && rr_next < rows;
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k) {
        idx_next[k] = static_cast<int32_t>(rr_next);
        if (k < nblk && valid_next) {
          const int32_t* ix = m.src[a.kb.src[k]].index;
          if (ix) idx_next[k] = ix[rr_next];
        }
      }
    };
    load_indices(0);
    pdl_wait();  // everything above reads only constants of the step (weights image, graph indices)
    for (uint32_t t = 0; t < ntl; ++t) {
      const uint32_t s = t % WS_NS, n = t / WS_NS;
      uint32_t off[P4C_MAX_SRC];
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k) {
        off[k] = kWsNoRow;
        if (k < nblk && valid_next) {
          const p4c_src_t& src = m.src[a.kb.src[k]];
          off[k] = static_cast<uint32_t>(
              (static_cast<long long>(b_next) * src.batch_stride + static_cast<long long>(idx_next[k]) * src.ld +
               a.kb.col0[k]) >> 3);
        }
      }
      const uint32_t b_cur = b_next, row0 = rr_next - static_cast<uint32_t>(pw * 32 + lane);
      load_indices(t + 1);
      if (n > 0) mbar_wait_backoff<200, true>(&S.x_empty[s], (n - 1) & 1u);
      if (a.tma_mask && pw == 0 && lane == 0) {  // whole tiles by TMA: bytes counted on the stage's barrier
#pragma unroll
        for (int k = 0; k < P4C_MAX_SRC; ++k)
          if (k < nblk && (a.tma_mask >> k & 1)) {
            mbar_expect_tx(&S.x_full[s], WS_XBLK);
            tma_load_3d(smem_u32(S.X[s][k]), &a.tmap[k], &S.x_full[s], a.kb.col0[k], static_cast<int>(row0),
                        a.tma_b[k] ? static_cast<int>(b_cur) : 0);
          }
      }
#pragma unroll
      for (int k = 0; k < P4C_MAX_SRC; ++k) {
        if (k < nblk && !(a.tma_mask >> k & 1)) {
          const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(m.src[a.kb.src[k]].ptr);
          const uint32_t xs = smem_u32(S.X[s][k]);
          const bool fetch = j < a.kb.nchunks[k];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = 4 * i + sub;
            const uint32_t o = __shfl_sync(0xffffffffu, off[k], r);
            if (fetch) {
              const bool ok = o != kWsNoRow;
              cp_async16(xs + sw128_offset(pw * 32 + r, j), ok ? base + (static_cast<size_t>(o) << 3) + j * 8 : base,
                         ok ? 16u : 0u);
            }
          }
        }
      }
      cp_async_commit();
      if (t >= WS_LAG) {
        cp_async_wait<WS_LAG>();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.x_full[(t - WS_LAG) % WS_NS]);
      }
    }
    static_assert(WS_LAG == 2, "drain below is written for a lag of two tiles");
    if (ntl >= 2) {
      cp_async_wait<1>();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.x_full[(ntl - 2) % WS_NS]);
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&S.x_full[(ntl - 1) % WS_NS]);
  } else if (warp == WS_EPI_WARPS + WS_PROD_WARPS) {
    // ================================= MMA ==================================
    // warp-uniform control flow (every lane runs the loop, one elected lane issues): the operand descriptors
    // then live in uniform registers; under `if (lane == 0)` the compiler wrapped every UTCHMMA in an
    // ELECT / BRA.U.ANY loop of ~15 instructions.  One base descriptor per operand, a 64-bit add per MMA.
    {
      constexpr uint32_t IDESC = make_idesc_bf16(WS_TM, WS_HID, 0, 0);
      constexpr long long KSTEP = 2;  // 32 bytes = K 16 bf16 columns, in descriptor address units
      const uint64_t w2d = make_desc_sw128(smem_u32(S.W2), 16, 1024);
      // biases inside the products (see PK_BIAS): all-ones A (MN-major, both 64-wide M blocks at the ONES rows)
      // times K-slice 0 (b1 hi, lo) / 1 (b2 hi, lo) of the bias tile -- the epilogues neither load nor add them
      constexpr uint32_t IDESC_MK = make_idesc_bf16(WS_TM, WS_HID, 1, 0);
      const uint64_t ones_mn = make_desc_sw128(smem_u32(S.ONES), 0, 1024), wb_k = make_desc_sw128(smem_u32(S.WBIAS), 16, 1024);
      for (uint32_t t = 0; t <= ntl; ++t) {
        if (t < ntl) {
          const uint32_t s = t % WS_NS, n = t / WS_NS, buf = t & 1u;
          mbar_wait_backoff<20>(&S.x_full[s], n & 1u);
          tc_fence_after();
          uint32_t acc = 0u;
          for (int k = 0; k < nblk; ++k) {
            uint64_t xd = make_desc_sw128(smem_u32(S.X[s][k]), 16, 1024), wd = make_desc_sw128(smem_u32(S.W1[k]), 16, 1024);
            const int nk = (a.kb.kvalid[k] + 15) >> 4;
            for (int q = 0; q < nk; ++q, xd += KSTEP, wd += KSTEP, acc = 1u)
              mma_bf16_ss_elect(tmem + buf * WS_HID, xd, wd, IDESC, acc);
          }
          mma_commit_elect(&S.x_empty[s]);
          mma_bf16_ss_elect(tmem + buf * WS_HID, ones_mn, wb_k, IDESC_MK, acc);  // + b1
          mma_commit_elect(&S.d1_full[buf]);
        }
        if (t >= 1) {
          const uint32_t u = t - 1, buf = u & 1u, nn = u >> 1;
          mbar_wait_backoff<20>(&S.a2_full[buf], nn & 1u);
          if (nn > 0) mbar_wait_backoff<20>(&S.d2_empty[buf], (nn - 1) & 1u);
          tc_fence_after();
          const uint64_t ad = make_desc_sw128(smem_u32(S.A2[buf]), 16, 1024);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            mma_bf16_ss_elect(tmem + 2 * WS_HID + buf * WS_HID, ad + q * KSTEP, w2d + q * KSTEP, IDESC, q ? 1u : 0u);
          mma_bf16_ss_elect(tmem + 2 * WS_HID + buf * WS_HID, ones_mn, wb_k + KSTEP, IDESC_MK, 1u);  // + b2
          mma_commit_elect(&S.d2_full[buf]);
        }
      }
    }
  } else {
    // =============================== epilogues ==============================
    // warp w < 16: group = w / 8 (tile parity), TMEM lane quadrant = w % 4, column half = (w / 4) % 2
    const int grp = warp >> 3, quad = warp & 3, hsel = (warp >> 2) & 1;
    const int row = quad * 32 + lane;
    const int sub = lane >> 2, j = hsel * 4 + (lane & 3);  // pass-2 mapping: row 8 i + sub, chunk j
    const int col0 = hsel * 32;
    const uint32_t full = a.full_rows;
    const uint32_t a2 = smem_u32(S.A2[grp]);
    const uint32_t tm_d1 = tmem + (static_cast<uint32_t>(quad * 32) << 16) + grp * WS_HID + col0;
    const uint32_t tm_d2 = tm_d1 + 2 * WS_HID;
    const int pair_bar = 1 + grp * 4 + quad;  // named barrier of the two warps that share these rows
    __nv_bfloat16* const save_h1 = AGGR_ONLY ? nullptr : a.save_h1;
    __nv_bfloat16* const out = AGGR_ONLY ? nullptr : a.out;
    __nv_bfloat16* const out2 = AGGR_ONLY ? nullptr : a.out2;
    const float* const gamma = S.vec[2] + col0;
    const float* const beta = S.vec[3] + col0;
    pdl_wait();
    for (uint32_t t = grp; t < ntl; t += 2) {
      const uint32_t s = t % WS_NS, nn = t >> 1;
      const uint32_t tile = blockIdx.x + t * gridDim.x;
      const uint32_t blk = nbatch == 1 ? tile
                                       : static_cast<uint32_t>(__umul64hi(static_cast<unsigned long long>(tile), a.div_magic));
      const uint32_t rr = blk * WS_TM + row;
      int32_t orow = -1;  // row in the output-like arrays
      if (rr < rows) orow = static_cast<int32_t>((tile - blk * nbatch) * full + rr);
      mbar_wait_suspend(&S.d1_full[grp], nn & 1u);
      tc_fence_after();
      // the residual / edge row = this tile's K-block 0 (complete: G1 has consumed it)
      uint4 res[4];
      if (need_x0) {
        const uint32_t x0 = smem_u32(S.X[s][0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) res[c] = lds128(x0 + sw128_offset(row, hsel * 4 + c));
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.x_empty[s]);
      }
      if constexpr (NO_SAVE) {
        // ---- epilogue 1 in one pass (h1 is not saved: the backward of this call recomputes it): a1 = SiLU of the
        //      bf16-rounded h1 (what a saving call would have stored) -> A2 operand of G2, own row / column half
        float v[32];
        tmem_ld32(tm_d1, v);
        tmem_ld_wait();
        tc_fence_before();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float h[8];
          const uint4 hb = ws_pack8(v + c * 8);  // D1 holds X W1^T + b1
          ws_unpack8(hb, h);
          silu8_pairs(h);
          sts128(a2 + sw128_offset(row, hsel * 4 + c), ws_pack8(h));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a2_full[grp]);
      } else {
      // ---- epilogue 1, pass 1: h1 = D1 + b1 -> bf16 -> A2 (own row, own column half)
        {
          float v[32];
          tmem_ld32(tm_d1, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            sts128(a2 + sw128_offset(row, hsel * 4 + c), ws_pack8(v + c * 8));  // D1 holds X W1^T + b1
          }
        }
        tc_fence_before();
        __syncwarp();
        // ---- pass 2 (lane -> chunk j of row 8 i + sub of this warp): SiLU in place -> A2 operand of G2,
        //      hand the tile to the MMA warp, THEN the coalesced store of h1 (saved activation): the
        //      proxy fence (MEMBAR.ALL.CTA in SASS) would otherwise wait for the stores' acknowledgement
        //      on the critical path to G2
        uint4 hb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + sub;
          const uint32_t addr = a2 + sw128_offset(quad * 32 + r, j);
          hb[i] = lds128(addr);
          float h[8];
          ws_unpack8(hb[i], h);
          silu8_pairs(h);
          sts128(addr, ws_pack8(h));
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a2_full[grp]);
        if (save_h1) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int32_t orr = __shfl_sync(0xffffffffu, orow, 8 * i + sub);
            if (orr >= 0) *reinterpret_cast<uint4*>(save_h1 + static_cast<long long>(orr) * WS_HID + j * 8) = hb[i];
          }
        }

      }

      // ---- epilogue 2
      mbar_wait_suspend(&S.d2_full[grp], nn & 1u);
      tc_fence_after();
      float y[32];
      tmem_ld32(tm_d2, y);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.d2_empty[grp]);  // D2 is in registers: G2 of tile t + 2 may overwrite it
      // (D2 holds a1 W2^T + b2)  fp32 pairs (FADD2 / FFMA2): two columns per instruction
      if (!NARROW && (AGGR_ONLY || m.layer_norm)) {
        f32x2 ps2[2] = {0ull, 0ull}, pq2[2] = {0ull, 0ull};  // (ps[0], ps[1]), (ps[2], ps[3]) of the scalar form
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          const f32x2 a01 = f2_pack(y[q], y[q + 1]), a23 = f2_pack(y[q + 2], y[q + 3]);
          ps2[0] = f2_add(ps2[0], a01);
          ps2[1] = f2_add(ps2[1], a23);
          pq2[0] = f2_fma(a01, a01, pq2[0]);
          pq2[1] = f2_fma(a23, a23, pq2[1]);
        }
        float ps[4], pq[4];
        f2_unpack(ps2[0], ps[0], ps[1]);
        f2_unpack(ps2[1], ps[2], ps[3]);
        f2_unpack(pq2[0], pq[0], pq[1]);
        f2_unpack(pq2[1], pq[2], pq[3]);
        S.ln_part[grp][hsel][row] = make_float2((ps[0] + ps[1]) + (ps[2] + ps[3]), (pq[0] + pq[1]) + (pq[2] + pq[3]));
        asm volatile("bar.sync %0, 64;" : : "r"(pair_bar) : "memory");
        const float2 p0 = S.ln_part[grp][0][row], p1 = S.ln_part[grp][1][row];  // same order in both threads
        const float sx = p0.x + p1.x, sq = p0.y + p1.y;
        const float mean = sx * (1.0f / WS_HID);
        const float var = fmaxf(sq * (1.0f / WS_HID) - mean * mean, 0.f);
        const float rstd = rsqrtf(var + m.ln_eps);
        const float nm = -mean * rstd;
        const f32x2 rstd2 = f2_dup(rstd), nm2 = f2_dup(nm);
        if constexpr (AGGR_ONLY) {
          // only the sum over the rows of a group leaves this kernel: sum_rows (xhat gamma + beta) =
          // gamma sum_rows xhat + group beta -- the affine part is applied to the few columns a lane keeps after
          // the group sums (2 + 2 vector loads per thread instead of 8 + 8 broadcast ones, 8 instead of 32 FMAs)
#pragma unroll
          for (int c = 0; c < 16; ++c)
            f2_unpack(f2_fma(f2_pack(y[2 * c], y[2 * c + 1]), rstd2, nm2), y[2 * c], y[2 * c + 1]);
        } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 gv = *reinterpret_cast<const float4*>(gamma + c * 4);
          const float4 bv = *reinterpret_cast<const float4*>(beta + c * 4);
          f2_unpack(f2_fma(f2_fma(f2_pack(y[c * 4 + 0], y[c * 4 + 1]), rstd2, nm2), f2_pack(gv.x, gv.y), f2_pack(bv.x, bv.y)),
                    y[c * 4 + 0], y[c * 4 + 1]);
          f2_unpack(f2_fma(f2_fma(f2_pack(y[c * 4 + 2], y[c * 4 + 3]), rstd2, nm2), f2_pack(gv.z, gv.w), f2_pack(bv.z, bv.w)),
                    y[c * 4 + 2], y[c * 4 + 3]);
        }
        }
        if (!AGGR_ONLY && a.save_stats && orow >= 0 && hsel == 0)
          *reinterpret_cast<float2*>(a.save_stats + 2ll * orow) = make_float2(mean, rstd);
      }
      if constexpr (NARROW) {
        // Narrow output (no LayerNorm; columns >= out_width of y are zero): rows are out_width * 2 bytes and a
        // tile's rows are contiguous in `out`.  Compact [row][out_width] bf16 image of the row quad in the (free)
        // A2 tile, then the 64 threads that own the quad copy its bytes with 16-byte stores.  The image lives in
        // the 4 KB of the tile that hold the quad's OWN rows: the only other writers of those bytes are the same
        // two warps (epilogue 1 of their next tile), which the pair barriers order.  (Packed back to back, the
        // images of quads 1-3 sat in the rows of quads 0 / 1, whose warps may already be writing a1 of the next
        // tile: a race that corrupted the last chunks of a quad once epilogue 1 got faster.)
        const int OW = a.out_width;
        uint8_t* const ct = S.A2[grp] + quad * 4096;
        __nv_bfloat16* const crow = reinterpret_cast<__nv_bfloat16*>(ct) + lane * OW + col0;
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (col0 + c < OW) crow[c] = __float2bfloat16_rn(y[c]);
        asm volatile("bar.sync %0, 64;" : : "r"(pair_bar) : "memory");
        const uint32_t q0 = blk * WS_TM + static_cast<uint32_t>(quad * 32);  // first row of the quad
        const int nv = q0 < rows ? static_cast<int>(rows - q0 < 32u ? rows - q0 : 32u) : 0;
        const int nbytes = nv * OW * 2;
        const uint8_t* const src = ct;
        uint8_t* const dst = reinterpret_cast<uint8_t*>(out) +
                             (static_cast<long long>(tile - blk * nbatch) * full + q0) * OW * 2;
        for (int i = (hsel * 32 + lane) * 16; i < nbytes; i += 64 * 16) {
          if (i + 16 <= nbytes) {
            *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(src + i);
          } else {
            for (int b = i; b < nbytes; b += 2)
              *reinterpret_cast<uint16_t*>(dst + b) = *reinterpret_cast<const uint16_t*>(src + b);
          }
        }
        asm volatile("bar.sync %0, 64;" : : "r"(pair_bar) : "memory");  // the A2 tile is rewritten by the next tile
        continue;
      }
      // ---- outputs staged through the A2 tile (free: G2 of this tile is complete)
      if (out2) {  // edge update: src0 row + y
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float e[8];
          ws_unpack8(res[c], e);
#pragma unroll
          for (int q = 0; q < 8; q += 2)
            f2_unpack(f2_add(f2_pack(e[q], e[q + 1]), f2_pack(y[c * 8 + q], y[c * 8 + q + 1])), e[q], e[q + 1]);
          sts128(a2 + sw128_offset(row, hsel * 4 + c), ws_pack8(e));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + sub;
          const int32_t orr = __shfl_sync(0xffffffffu, orow, r);
          const uint4 u = lds128(a2 + sw128_offset(quad * 32 + r, j));
          if (orr >= 0) *reinterpret_cast<uint4*>(out2 + static_cast<long long>(orr) * WS_HID + j * 8) = u;
        }
        __syncwarp();
      }
      if (out) {  // messages, or y + residual
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float e[8];
          if (a.residual_x0) {
            ws_unpack8(res[c], e);
#pragma unroll
            for (int q = 0; q < 8; q += 2)
              f2_unpack(f2_add(f2_pack(e[q], e[q + 1]), f2_pack(y[c * 8 + q], y[c * 8 + q + 1])), e[q], e[q + 1]);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) e[q] = y[c * 8 + q];
          }
          sts128(a2 + sw128_offset(row, hsel * 4 + c), ws_pack8(e));
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + sub;
          const int32_t orr = __shfl_sync(0xffffffffu, orow, r);
          const uint4 u = lds128(a2 + sw128_offset(quad * 32 + r, j));
          if (orr >= 0) *reinterpret_cast<uint4*>(out + static_cast<long long>(orr) * WS_HID + j * 8) = u;
        }
        __syncwarp();
      }
      if (AGGR_ONLY || a.aggr_out) {
        // sum of the messages of `agg_group` consecutive rows (= lanes): transposing butterfly, at
        // every step a lane hands one half of its columns to its partner and adds the partner's
        // other half (fp32, fixed order; invalid rows add zero)
        const int gsz = a.agg_group;
        if (orow < 0) {
#pragma unroll
          for (int q = 0; q < 32; ++q) y[q] = 0.f;
        }
        int cbase = 0;
#define P4C_WS_AGG_STEP(OFF, N)                                              \
        if (gsz > OFF) {                                                     \
          const bool up = (lane & OFF) != 0;                                 \
          _Pragma("unroll") for (int i = 0; i < (N) / 2; ++i) {              \
            const float send = up ? y[i] : y[i + (N) / 2];                   \
            const float keep = up ? y[i + (N) / 2] : y[i];                   \
            y[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);           \
          }                                                                  \
          if (up) cbase += (N) / 2;                                          \
        }
        P4C_WS_AGG_STEP(1, 32)
        P4C_WS_AGG_STEP(2, 16)
        P4C_WS_AGG_STEP(4, 8)
        P4C_WS_AGG_STEP(8, 4)
#undef P4C_WS_AGG_STEP
        if (orow >= 0) {
          const int ncol = 32 / gsz;
          __nv_bfloat16* dst = a.aggr_out + static_cast<long long>(orow / gsz) * WS_HID + col0 + cbase;
          const float sc = a.agg_scale;
          const float gb = static_cast<float>(gsz);
          if (ncol >= 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (c < ncol / 8) {
                if constexpr (AGGR_ONLY) {  // affine part of the LayerNorm on the group sums (see above)
                  const float4 g0 = *reinterpret_cast<const float4*>(gamma + cbase + c * 8);
                  const float4 g1 = *reinterpret_cast<const float4*>(gamma + cbase + c * 8 + 4);
                  const float4 b0 = *reinterpret_cast<const float4*>(beta + cbase + c * 8);
                  const float4 b1 = *reinterpret_cast<const float4*>(beta + cbase + c * 8 + 4);
                  float* yy = y + c * 8;
                  yy[0] = fmaf(yy[0], g0.x, gb * b0.x); yy[1] = fmaf(yy[1], g0.y, gb * b0.y);
                  yy[2] = fmaf(yy[2], g0.z, gb * b0.z); yy[3] = fmaf(yy[3], g0.w, gb * b0.w);
                  yy[4] = fmaf(yy[4], g1.x, gb * b1.x); yy[5] = fmaf(yy[5], g1.y, gb * b1.y);
                  yy[6] = fmaf(yy[6], g1.z, gb * b1.z); yy[7] = fmaf(yy[7], g1.w, gb * b1.w);
                }
                float e[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) e[q] = y[c * 8 + q] * sc;
                *reinterpret_cast<uint4*>(dst + c * 8) = ws_pack8(e);
              }
            }
          } else if (ncol == 4) {
            if constexpr (AGGR_ONLY) {
#pragma unroll
              for (int q = 0; q < 4; ++q) y[q] = fmaf(y[q], gamma[cbase + q], gb * beta[cbase + q]);
            }
            uint2 o;
            o.x = pack_bf16x2(y[0] * sc, y[1] * sc);
            o.y = pack_bf16x2(y[2] * sc, y[3] * sc);
            *reinterpret_cast<uint2*>(dst) = o;
          } else {
            if constexpr (AGGR_ONLY) {
              y[0] = fmaf(y[0], gamma[cbase], gb * beta[cbase]);
              y[1] = fmaf(y[1], gamma[cbase + 1], gb * beta[cbase + 1]);
            }
            *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(y[0] * sc, y[1] * sc);
          }
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == WS_EPI_WARPS + WS_PROD_WARPS) tmem_dealloc(tmem, WS_TMEM_COLS);
}

}  // namespace tc

bool tc_kblocks(const p4c_mlp_t* m, tc::KBlocks* kb);

static long long g_ws_min_tiles = -1;
static long long g_ws_launches = 0;
static long long ws_min_tiles() {
  if (g_ws_min_tiles < 0) {
    const char* e = getenv("P4C_WS_MIN_TILES");
    g_ws_min_tiles = (e && atoll(e) > 0) ? atoll(e) : 2ll * kNumSMs;
  }
  return g_ws_min_tiles;
}

// The warp-specialised kernel covers the large calls; everything else stays on mlp2_tc_fwd_kernel.
bool ws_forward_supported(const p4c_mlp_t* m, const void* residual, const void* out2, const void* aggr_out,
                          int agg_group) {
  if (!m->packed) return false;
  if (m->out_features != P4C_HIDDEN) {
    // narrow output: no LayerNorm, plain store, 16-byte aligned row quads in every sample
    const long long fr = m->full_rows > 0 ? m->full_rows : m->rows;
    if (m->layer_norm || residual || out2 || aggr_out || m->out_features < 1 || m->out_features > P4C_HIDDEN) return false;
    if (m->batch > 1 && (fr * m->out_features * 2) % 16 != 0) return false;
  }
  if (m->tile_rows != 0 && m->tile_rows != 128) return false;
  for (int s = 0; s < m->nsrc; ++s)
    if (m->src[s].seg_rowptr) return false;
  const long long full = m->full_rows > 0 ? m->full_rows : m->rows;
  const long long total = static_cast<long long>(m->rows) * m->batch;
  if (total >= (1ll << 31) - 256 * 1024 || full * m->batch >= (1ll << 31)) return false;
  if ((total + 127) / 128 < ws_min_tiles()) return false;
  if (residual) {
    const p4c_src_t& s0 = m->src[0];
    if (residual != s0.ptr || s0.index || s0.ld != P4C_HIDDEN || s0.width != P4C_HIDDEN) return false;
    if (m->batch > 1 && s0.batch_stride != full * P4C_HIDDEN) return false;
  }
  if (aggr_out && (agg_group < 1 || agg_group > 16 || (agg_group & (agg_group - 1)))) return false;
  return true;
}

int launch_mlp2_ws_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2, void* save_h1,
                           float* save_stats, void* aggr_out, int agg_group, float agg_scale, cudaStream_t st) {
  using namespace tc;
  WsArgs a;
  a.m = *m;
  tc_kblocks(m, &a.kb);
  a.out = static_cast<__nv_bfloat16*>(out);
  a.out2 = static_cast<__nv_bfloat16*>(out2);
  a.save_h1 = static_cast<__nv_bfloat16*>(save_h1);
  a.save_stats = save_stats;
  a.aggr_out = static_cast<__nv_bfloat16*>(aggr_out);
  a.agg_group = aggr_out ? agg_group : 1;
  a.agg_scale = agg_scale;
  a.residual_x0 = residual != nullptr;
  a.out_width = m->out_features;
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
  const unsigned long long rows = static_cast<unsigned long long>(m->rows);
  const unsigned long long nb = static_cast<unsigned long long>(m->batch);
  a.div_magic = nb == 1 ? 0ull : (~0ull / nb) + 1ull;  // ceil(2^64 / batch) for batch >= 2 (batch 1: not used)
  a.batch = static_cast<uint32_t>(m->batch);
  a.ntiles = static_cast<uint32_t>((rows + WS_TM - 1) / WS_TM * nb);
  a.full_rows = static_cast<uint32_t>(m->full_rows > 0 ? m->full_rows : m->rows);
  int max_ctas = tc_max_ctas() / 2;
  if (max_ctas < 1) max_ctas = 1;
  const int grid = static_cast<int>(a.ntiles < static_cast<uint32_t>(max_ctas) ? a.ntiles : max_ctas);
  const size_t smem = sizeof(WsSmem) + 1024;
  const bool narrow = m->out_features != P4C_HIDDEN;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(mlp2_ws_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp2_ws_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp2_ws_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(mlp2_ws_fwd_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("mlp_forward(ws): cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    configured = true;
  }
  const bool aggr_only = !narrow && aggr_out && !out && !out2 && !save_h1 && !save_stats && !residual && m->layer_norm;
  if (narrow) launch_pdl(mlp2_ws_fwd_kernel<true>, dim3(grid), dim3(WS_THREADS), smem, st, a);
  else if (aggr_only) launch_pdl(mlp2_ws_fwd_kernel<false, true>, dim3(grid), dim3(WS_THREADS), smem, st, a);
  else if (!save_h1) launch_pdl(mlp2_ws_fwd_kernel<false, false, true>, dim3(grid), dim3(WS_THREADS), smem, st, a);
  else launch_pdl(mlp2_ws_fwd_kernel<false>, dim3(grid), dim3(WS_THREADS), smem, st, a);
  P4C_CHECK_LAUNCH("mlp_forward(ws)");
  ++g_ws_launches;
  return P4C_OK;
}

}  // namespace p4c

extern "C" void p4c_set_ws_min_tiles(int64_t n) { p4c::g_ws_min_tiles = n > 0 ? n : -1; }
extern "C" int64_t p4c_get_ws_min_tiles(void) { return p4c::ws_min_tiles(); }
extern "C" int64_t p4c_ws_launch_count(void) { return p4c::g_ws_launches; }
