// Weight images for the tensor-core engine, the (batched) weight-gradient reduction and the fused
// segment-sum + input-gradient GEMM of gathered sources.
//
//   p4c_pack_weights         fp32 nn.Linear / LayerNorm parameters -> bf16 images in the swizzled
//                            shared-memory layout of tcgen05.mma (ONE launch for a whole model; the
//                            forward / backward kernels then stage their weights with 16-byte copies
//                            instead of converting 16k values per CTA).
//   p4c_wgrad_reduce_batched per-CTA weight-gradient partials of many MLPs -> g_* in ONE launch,
//                            fixed summation order (bit-reproducible, no float atomics).
//   p4c_seg_linear           d_x(node) = (sum of dh1 over the node's edges) x W1 slice: the transpose
//                            of the gather of InteractionNet (K1/K3 backward), one kernel.
#include "common.cuh"
#include "tc_common.cuh"

namespace p4c {

constexpr int HIDP = P4C_HIDDEN;

// ------------------------------------------------------------------------------ weight images
// grid (n, 9): blockIdx.y = 8 KB block of the image (0-2 W1, 3 W2, 4 W2T, 5-7 W1T, 8 bias tile); y == 0 also
// writes the vectors.
__global__ void __launch_bounds__(256) pack_weights_kernel(const p4c_pack_t* __restrict__ table) {
  const p4c_pack_t e = table[blockIdx.x];
  const int y = blockIdx.y, tid = threadIdx.x;
  uint8_t* img = reinterpret_cast<uint8_t*>(e.dst);
  tc::KBlocks kb;
  tc::make_kblocks(e.nsrc, e.width, &kb);
  uint8_t* dst = y < 8 ? img + y * 8192 : img + tc::PK_BIAS;
  const int OUT = e.out_features;
  for (int i = tid; i < HIDP * HIDP; i += 256) {
    const int n = i >> 6, kk = i & 63;  // image element [n][kk]
    float v = 0.f;
    if (y == 8) {  // bias tile: hi / lo halves of b1 (k = 0, 1) and b2 (k = 16, 17)
      const int slice = kk >> 4, part = kk & 15;
      if (slice < 2 && part < 2 && (slice == 0 || n < OUT)) {
        const float b = slice == 0 ? e.b1[n] : e.b2[n];
        const float hi = __bfloat162float(__float2bfloat16_rn(b));
        v = part == 0 ? hi : b - hi;
      }
    } else if (y < 3) {  // W1 block y: [n = hidden unit][kk = input column of the block]
      if (y < kb.nblk && kk < kb.kvalid[y]) v = e.w1[static_cast<long long>(n) * e.k1 + kb.kstart[y] + kk];
    } else if (y == 3) {  // W2: [n = output unit][kk = hidden unit]
      if (n < OUT) v = e.w2[n * HIDP + kk];
    } else if (y == 4) {  // W2T: [n = hidden unit][kk = output unit]
      if (kk < OUT) v = e.w2[kk * HIDP + n];
    } else {  // W1T block y-5: [n = input column of the block][kk = hidden unit]
      const int b = y - 5;
      if (b < kb.nblk && n < kb.kvalid[b]) v = e.w1[static_cast<long long>(kk) * e.k1 + kb.kstart[b] + n];
    }
    *reinterpret_cast<__nv_bfloat16*>(dst + tc::sw128_offset(n, kk >> 3) + (kk & 7) * 2) = __float2bfloat16_rn(v);
  }
  if (y == 0 && tid < HIDP) {
    float* vec = reinterpret_cast<float*>(img + tc::PK_VEC);
    vec[tid] = e.b1[tid];
    vec[HIDP + tid] = tid < OUT ? e.b2[tid] : 0.f;
    vec[2 * HIDP + tid] = e.gamma ? e.gamma[tid] : 1.f;
    vec[3 * HIDP + tid] = e.beta ? e.beta[tid] : 0.f;
  }
}

// ---------------------------------------------------------------------- weight-gradient reduction
// Partial block layout (see mlp2.cu / mlp2_tc_bwd.cu): dW1[k1pad][64] (k-major), dW2[64 hidden][64 out],
// db1, db2, dgamma, dbeta [64] each.  Threads walk the partial layout (coalesced reads over all
// blocks); the transposition to the nn.Linear layout happens on the single write.
__host__ __device__ inline long long wg_partial_floats(int k1) {
  const int k1pad = (k1 + 31) / 32 * 32;
  return static_cast<long long>(k1pad) * HIDP + HIDP * HIDP + 4 * HIDP;
}

__device__ __forceinline__ void wgrad_reduce_entry(const p4c_wgrad_t& e, long long i0, long long stride) {
  const int k1pad = (e.k1 + 31) / 32 * 32;
  const long long pf = wg_partial_floats(e.k1);
  const long long n1 = static_cast<long long>(k1pad) * HIDP;
  for (long long i = i0; i < pf; i += stride) {
    float* dst = nullptr;
    if (i < n1) {  // partial dW1[k][c] -> g_w1[c][k]
      const int k = static_cast<int>(i >> 6), c = static_cast<int>(i & 63);
      if (e.g_w1 && k < e.k1) dst = e.g_w1 + static_cast<long long>(c) * e.k1 + k;
    } else if (i < n1 + HIDP * HIDP) {  // partial dW2[k][c] -> g_w2[c][k]
      const int j = static_cast<int>(i - n1);
      const int k = j >> 6, c = j & 63;
      if (e.g_w2 && c < e.out_features) dst = e.g_w2 + c * HIDP + k;
    } else {
      const int j = static_cast<int>(i - n1 - HIDP * HIDP);
      const int v = j >> 6, c = j & 63;
      float* base = v == 0 ? e.g_b1 : v == 1 ? e.g_b2 : v == 2 ? e.g_gamma : e.g_beta;
      if (base && (v != 1 || c < e.out_features)) dst = base + c;
    }
    if (!dst) continue;
    // fixed-shape reduction tree: 8 interleaved partial sums (8 independent loads in flight)
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float* p = e.partials + i;
    int b = 0;
    for (; b + 8 <= e.nblocks; b += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += p[static_cast<long long>(b + j) * pf];
    }
    for (int j = 0; b < e.nblocks; ++b, ++j) a[j] += p[static_cast<long long>(b) * pf];
    *dst += ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_batched_kernel(const p4c_wgrad_t* __restrict__ table) {
  const p4c_wgrad_t e = table[blockIdx.y];
  wgrad_reduce_entry(e, blockIdx.x * 256ll + threadIdx.x, static_cast<long long>(gridDim.x) * 256);
}

__global__ void __launch_bounds__(256) wgrad_reduce_single_kernel(const p4c_wgrad_t e) {
  wgrad_reduce_entry(e, blockIdx.x * 256ll + threadIdx.x, static_cast<long long>(gridDim.x) * 256);
}

int launch_wgrad_reduce_single(const p4c_wgrad_t* e, cudaStream_t st) {
  const long long pf = wg_partial_floats(e->k1);
  wgrad_reduce_single_kernel<<<static_cast<int>((pf + 255) / 256), 256, 0, st>>>(*e);
  P4C_CHECK_LAUNCH("mlp_wgrad_reduce");
  return P4C_OK;
}

// ------------------------------------------------------------------ segment sum + 64x64 GEMM (TC)
namespace tc {

// Set 0 first: launches that sum one set per tile (classic and split calls) allocate only SegLinSmem0.  The
// large gathers are bound by the requests the L1 can keep in flight, and every KB of shared memory is taken from
// it (measured: m2g by-sender pass 0.62 -> 0.84 ms with the two-set layout).
struct SegLinSmem0 {
  uint8_t A[128 * 128];  // summed rows, bf16, swizzled
  uint8_t W[64 * 128];   // [n = input column][k = hidden unit] = W1T block
  uint32_t lo[128], hi[128], sbase[128];
  float sc[128];
  int32_t orow[128];
  uint64_t mbar;
  uint32_t tmem_base;
  uint32_t pad[125];  // to a multiple of 1 KB (swizzled tiles follow)
};
struct SegLinSmem1 {  // the second set of a fused call
  uint8_t A[128 * 128];
  uint8_t W[64 * 128];
  uint32_t lo[128], hi[128], sbase[128];
  float sc[128];
};
static_assert(sizeof(SegLinSmem0) % 1024 == 0, "the second set's tiles must stay 1 KB aligned");

// One segment set of a launch: out[b, r, :] (+)= (scale[r] * sum over segment r of src rows) x W1 block
struct SegPart {
  const int32_t* rowptr;
  const int32_t* perm;
  const float* scale;
  const uint8_t* packed_block;  // PK_W1T block of the image, or NULL
  int col0;
  int num_seg, batch;
  long long src_bs;
  __nv_bfloat16* out;
  long long out_bs;
  int accumulate;
  int tile_rows;  // valid rows per tile (16..128)
};

// One or two segment sets over the SAME rows `src` (dh1 of an edge MLP):
//   nparts == 1                    the classic call;
//   nparts == 2, fused             both sets end in the same output rows (sender and receiver gradient of an
//                                  InteractionNet whose senders are its receivers): the two summed tiles are
//                                  the two K-blocks of ONE product, one read-modify-write of the output;
//   nparts == 2, not fused         independent problems, CTAs [0, grid0) run set 0 and the others set 1
//                                  (one launch instead of two on the launch-bound mesh levels).
struct SegLinArgs {
  const __nv_bfloat16* src;
  const float* w1;
  int k1;
  int nparts, fused, grid0;
  SegPart p[2];
};

// PAIR = false: one set per launch (everything about the second set folds away at compile time; the large
// gathers of the m2g / g2m backward run through this instantiation).
template <bool PAIR>
__global__ void __launch_bounds__(256, 2) seg_linear_tc_kernel(const __grid_constant__ SegLinArgs a) {
  extern __shared__ unsigned char smem_raw[];
  SegLinSmem0& S = *reinterpret_cast<SegLinSmem0*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  SegLinSmem1& S1 = *reinterpret_cast<SegLinSmem1*>(reinterpret_cast<unsigned char*>(&S) + sizeof(SegLinSmem0));
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;  // uniform for the compiler
  const int quad = warp & 3, hsel = warp >> 2;
  const int row = quad * 32 + lane, col0 = hsel * 32;
  const bool split = PAIR && a.nparts == 2 && !a.fused;
  const int first = (split && static_cast<int>(blockIdx.x) >= a.grid0) ? 1 : 0;  // this CTA's (first) set
  const int nsum = (PAIR && a.fused) ? 2 : 1;
  const int bid = static_cast<int>(blockIdx.x) - (first ? a.grid0 : 0);
  const int nct = split ? (first ? static_cast<int>(gridDim.x) - a.grid0 : a.grid0) : static_cast<int>(gridDim.x);
  const SegPart& P0 = PAIR ? a.p[first] : a.p[0];
  const long long total = static_cast<long long>(P0.num_seg) * P0.batch;
  const int TR = P0.tile_rows;
  const int32_t* const perm0 = P0.perm;
  const int32_t* const perm1 = a.p[1].perm;
  __nv_bfloat16* const outp = P0.out;
  const bool accum = P0.accumulate != 0;
  const long long ntiles = (total + TR - 1) / TR;

  for (int s = 0; s < nsum; ++s) {
    const SegPart& P = a.p[first + s];
    uint8_t* const Ws = s ? S1.W : S.W;
    if (P.packed_block) {
      for (int i = tid; i < 512; i += 256) cp_async16(smem_u32(Ws) + i * 16, P.packed_block + i * 16, 16u);
    } else {
      for (int i = tid; i < 64 * 64; i += 256) {
        const int c = i >> 6, k = i & 63;  // W[n=k][K=c] = w1[c][col0 + k]
        *reinterpret_cast<__nv_bfloat16*>(Ws + sw128_offset(k, c >> 3) + (c & 7) * 2) =
            __float2bfloat16_rn(a.w1[static_cast<long long>(c) * a.k1 + P.col0 + k]);
      }
    }
  }
  cp_async_commit();
  if (tid == 0) {
    mbar_init(&S.mbar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&S.tmem_base, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = S.tmem_base;
  const uint32_t tmem_row = tmem + (static_cast<uint32_t>(quad * 32) << 16) + col0;
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64, 0, 0);

  pdl_launch_dependents();
  pdl_wait();  // the previous kernel's outputs are complete and visible
  uint32_t it = 0;
  for (long long tile = bid; tile < ntiles; tile += nct, ++it) {
    {  // row tables: threads 0..127 fill set 0, threads 128..255 set 1 (fused calls)
      const int s = tid >> 7, t = tid & 127;
      if (s < nsum) {
        const SegPart& P = a.p[first + s];
        const long long g = tile * TR + t;
        uint32_t lo = 0xFFFFFFFFu, hi = 0, sb = 0;
        float sc = 1.f;
        int32_t o = -1;
        if (t < TR && g < total) {
          const long long b = g / P.num_seg;
          const int rr = static_cast<int>(g - b * P.num_seg);
          lo = static_cast<uint32_t>(P.rowptr[rr]);
          hi = static_cast<uint32_t>(P.rowptr[rr + 1]);
          sb = static_cast<uint32_t>((b * P.src_bs) >> 3);
          if (P.scale) sc = P.scale[rr];
          o = static_cast<int32_t>((b * P.out_bs) / 64 + rr);
        }
        if (s == 0) {
          S.lo[t] = lo; S.hi[t] = hi; S.sbase[t] = sb; S.sc[t] = sc; S.orow[t] = o;
        } else {
          S1.lo[t] = lo; S1.hi[t] = hi; S1.sbase[t] = sb; S1.sc[t] = sc;
        }
      }
    }
    __syncthreads();
    // (the branches on the null-ness of perm are taken here, outside the gather loops: inside them the compiler
    // would serialise the index loads)
    if (!PAIR) {
      // tiles of at most 16 rows (long segments: m2g by sender) leave half of a thread's four items idle: two items,
      // four rows each per iteration keep eight row loads in flight (same summation order)
      if (TR <= 16) seg_sum_tile<256, 2, 4, true>(a.src, a.p[0].perm, S.lo, S.hi, S.sbase, S.sc, S.orow, S.A, nullptr, tid, TR);
      else seg_sum_tile<256, 4, 2, true>(a.src, a.p[0].perm, S.lo, S.hi, S.sbase, S.sc, S.orow, S.A, nullptr, tid, TR);
    }
    else if (perm0) seg_sum_tile<256, 4, 2, true>(a.src, perm0, S.lo, S.hi, S.sbase, S.sc, S.orow, S.A, nullptr, tid, TR);
    else seg_sum_tile<256, 4, 2, true>(a.src, nullptr, S.lo, S.hi, S.sbase, S.sc, S.orow, S.A, nullptr, tid, TR);
    if (PAIR && nsum == 2) {
      if (perm1) seg_sum_tile<256, 4, 2, true>(a.src, perm1, S1.lo, S1.hi, S1.sbase, S1.sc, S.orow, S1.A, nullptr, tid, TR);
      else seg_sum_tile<256, 4, 2, true>(a.src, nullptr, S1.lo, S1.hi, S1.sbase, S1.sc, S.orow, S1.A, nullptr, tid, TR);
    }
    cp_async_wait<0>();  // the weight blocks (first tile)
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      for (int s = 0; s < nsum; ++s) {
        const uint32_t aa = smem_u32(s ? S1.A : S.A), wa = smem_u32(s ? S1.W : S.W);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          mma_bf16_ss_elect(tmem, make_desc_sw128(aa + j * 32, 16, 1024), make_desc_sw128(wa + j * 32, 16, 1024), IDESC,
                      (s | j) ? 1u : 0u);
      }
      mma_commit_elect(&S.mbar);
    }
    mbar_wait(&S.mbar, it & 1u);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem_row, v);
    tmem_ld_wait();
    const long long orow = S.orow[row];
    if (orow >= 0) {
      __nv_bfloat16* dst = outp + orow * 64 + col0;
      if (accum) {  // the four loads in flight together
        uint4 e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) e[j] = *reinterpret_cast<const uint4*>(dst + j * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* o = v + j * 8;
          float2 t;
          t = unpack_bf16x2(e[j].x); o[0] += t.x; o[1] += t.y;
          t = unpack_bf16x2(e[j].y); o[2] += t.x; o[3] += t.y;
          t = unpack_bf16x2(e[j].z); o[4] += t.x; o[5] += t.y;
          t = unpack_bf16x2(e[j].w); o[6] += t.x; o[7] += t.y;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* o = v + j * 8;
        uint4 u;
        u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
        u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
        *reinterpret_cast<uint4*>(dst + j * 8) = u;
      }
    }
    tc_fence_before();
    __syncthreads();  // A, row info and the accumulator may be overwritten
  }
  if (warp == 0) tmem_dealloc(tmem, 64);
}

}  // namespace tc
}  // namespace p4c

using namespace p4c;

extern "C" int p4c_pack_weights(const p4c_pack_t* table_dev, int32_t n, void* stream) {
  P4C_REQUIRE(n >= 0 && (n == 0 || table_dev), "pack_weights: bad arguments");
  if (n == 0) return P4C_OK;
  pack_weights_kernel<<<dim3(n, 9), 256, 0, (cudaStream_t)stream>>>(table_dev);
  P4C_CHECK_LAUNCH("pack_weights");
  pdl_block_next();  // the images are read before griddepcontrol.wait
  return P4C_OK;
}

extern "C" int p4c_wgrad_reduce_batched(const p4c_wgrad_t* table_dev, int32_t n, void* stream) {
  P4C_REQUIRE(n >= 0 && (n == 0 || table_dev), "wgrad_reduce_batched: bad arguments");
  if (n == 0) return P4C_OK;
  // 66 blocks of 256 threads cover the largest partial block (k1 = 192: 16 640 floats)
  wgrad_reduce_batched_kernel<<<dim3(66, n), 256, 0, (cudaStream_t)stream>>>(table_dev);
  P4C_CHECK_LAUNCH("wgrad_reduce_batched");
  return P4C_OK;
}

static int seg_check_part(const p4c_seg_part_t* q, int32_t k1, const void* packed, const char* what) {
  P4C_REQUIRE(q->rowptr && q->out, "%s: null pointer", what);
  P4C_REQUIRE(q->num_seg > 0 && q->batch > 0, "%s: empty problem", what);
  P4C_REQUIRE(q->col0 >= 0 && q->col0 + HIDP <= k1, "%s: column window [%d,%d) outside k1=%d", what, q->col0,
              q->col0 + HIDP, k1);
  P4C_REQUIRE(q->src_batch_stride % 8 == 0 && q->out_batch_stride % 64 == 0,
              "%s: batch strides must be multiples of 8 / 64 elements", what);
  P4C_REQUIRE(reinterpret_cast<uintptr_t>(q->out) % 16 == 0, "%s: out must be 16-byte aligned", what);
  P4C_REQUIRE(!packed || (q->kblock >= 0 && q->kblock < P4C_MAX_SRC), "%s: bad kblock %d", what, q->kblock);
  const int tr = q->tile_rows == 0 ? 128 : q->tile_rows;
  P4C_REQUIRE(tr == 16 || tr == 32 || tr == 64 || tr == 128, "%s: tile_rows=%d not in {16,32,64,128}", what, tr);
  return P4C_OK;
}

static void seg_fill_part(tc::SegPart* d, const p4c_seg_part_t* q, const void* packed) {
  d->rowptr = q->rowptr;
  d->perm = q->perm;
  d->scale = q->scale;
  d->packed_block = packed ? static_cast<const uint8_t*>(packed) + tc::PK_W1T + q->kblock * 8192 : nullptr;
  d->col0 = q->col0;
  d->num_seg = q->num_seg;
  d->batch = q->batch;
  d->src_bs = q->src_batch_stride;
  d->out = static_cast<__nv_bfloat16*>(q->out);
  d->out_bs = q->out_batch_stride;
  d->accumulate = q->accumulate;
  d->tile_rows = q->tile_rows == 0 ? 128 : q->tile_rows;
}

static long long seg_part_tiles(const tc::SegPart& d) {
  return (static_cast<long long>(d.num_seg) * d.batch + d.tile_rows - 1) / d.tile_rows;
}

extern "C" int p4c_seg_linear_pair(const void* src, const float* w1, int32_t k1, const void* packed,
                                   const p4c_seg_part_t* part0, const p4c_seg_part_t* part1, int dtype, void* stream) {
  P4C_REQUIRE(src && w1 && part0, "seg_linear_pair: null pointer");
  P4C_REQUIRE(dtype == P4C_BF16, "seg_linear_pair: tensor-core engine, bf16 storage only (dtype=%d)", dtype);
  P4C_REQUIRE(reinterpret_cast<uintptr_t>(src) % 16 == 0, "seg_linear_pair: src must be 16-byte aligned");
  int rc = seg_check_part(part0, k1, packed, "seg_linear_pair(part 0)");
  if (rc) return rc;
  if (part1 && (rc = seg_check_part(part1, k1, packed, "seg_linear_pair(part 1)"))) return rc;
  using namespace tc;
  SegLinArgs a{};
  a.src = static_cast<const __nv_bfloat16*>(src);
  a.w1 = w1;
  a.k1 = k1;
  a.nparts = part1 ? 2 : 1;
  seg_fill_part(&a.p[0], part0, packed);
  if (part1) seg_fill_part(&a.p[1], part1, packed);
  else a.p[1] = a.p[0];
  // both sets end in the same output rows: one product over two K-blocks (the second set's accumulate flag is
  // implied; the first set's decides whether the previous content of `out` is added)
  a.fused = part1 && part0->out == part1->out;
  if (a.fused) {
    P4C_REQUIRE(a.p[0].num_seg == a.p[1].num_seg && a.p[0].batch == a.p[1].batch && a.p[0].out_bs == a.p[1].out_bs &&
                    a.p[0].tile_rows == a.p[1].tile_rows,
                "seg_linear_pair: sets that share an output must have the same segments x batch shape and tile_rows");
    P4C_REQUIRE(part1->accumulate, "seg_linear_pair: the second set of a shared output accumulates by definition");
  }
  const long long t0 = seg_part_tiles(a.p[0]), t1 = part1 ? seg_part_tiles(a.p[1]) : 0;
  const long long cap = tc_max_ctas();
  if (part1 && !a.fused && t0 + t1 > cap) {
    // Independent sets that do not fit one wave of CTAs together gain nothing from sharing a launch (they are
    // bandwidth-bound, and a set of few long segments costs more per source row than one of many short ones:
    // measured, g2m by receiver + by sender, 0.93 ms as two launches, 1.7 ms as one with equal shares)
    int rc2 = p4c_seg_linear_pair(src, w1, k1, packed, part0, nullptr, dtype, stream);
    if (rc2) return rc2;
    return p4c_seg_linear_pair(src, w1, k1, packed, part1, nullptr, dtype, stream);
  }
  int grid;
  a.grid0 = 0;
  if (part1 && !a.fused) {  // every tile of both sets has its own CTA
    a.grid0 = static_cast<int>(t0);
    grid = static_cast<int>(t0 + t1);
  } else {
    grid = static_cast<int>(t0 < cap ? t0 : cap);
  }
  const size_t smem_max = sizeof(SegLinSmem0) + sizeof(SegLinSmem1) + 1024;
  const size_t smem = (a.fused ? sizeof(SegLinSmem0) + sizeof(SegLinSmem1) : sizeof(SegLinSmem0)) + 1024;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(seg_linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) {
      set_error("seg_linear: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    configured = true;
  }
  if (part1) launch_pdl(seg_linear_tc_kernel<true>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, a);
  else launch_pdl(seg_linear_tc_kernel<false>, dim3(grid), dim3(256), smem, (cudaStream_t)stream, a);
  P4C_CHECK_LAUNCH("seg_linear");
  return P4C_OK;
}

extern "C" int p4c_seg_linear(const void* src, int64_t src_bs, const int32_t* rowptr, const int32_t* perm,
                              const float* scale, int32_t num_seg, int32_t batch, const float* w1, int32_t k1,
                              int32_t col0, const void* packed, int32_t kblock, void* out, int64_t out_bs,
                              int accumulate, int32_t tile_rows, int dtype, void* stream) {
  P4C_REQUIRE(src && rowptr && out && w1, "seg_linear: null pointer");
  p4c_seg_part_t q;
  q.rowptr = rowptr; q.perm = perm; q.scale = scale;
  q.num_seg = num_seg; q.batch = batch; q.src_batch_stride = src_bs;
  q.col0 = col0; q.kblock = kblock;
  q.out = out; q.out_batch_stride = out_bs;
  q.accumulate = accumulate; q.tile_rows = tile_rows;
  return p4c_seg_linear_pair(src, w1, k1, packed, &q, nullptr, dtype, stream);
}
