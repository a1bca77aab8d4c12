// Hand-written sm_100a primitives for the tensor-core engine: tcgen05 MMA with TMEM
// accumulators, shared-memory matrix descriptors (128-byte swizzle), mbarriers, cp.async.
// Inline PTX only (no CUTLASS/CuTe dependency).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace p4c {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- shared-memory matrix descriptor (PTX "tcgen05 shared memory descriptor") ----------------
//   bits [ 0,14) start address >> 4          bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset >> 4  bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) swizzle: 0 none, 2 = 128 B, 4 = 64 B, 6 = 32 B
// 128-byte swizzle, rows of 128 bytes (64 bf16), 8-row atoms of 1024 bytes stacked every
// `sbo_bytes`.  The same physical tile serves as
//   * a K-major operand   (rows = M/N index, the 128-byte row = 64 consecutive K), LBO unused;
//   * an MN-major operand (the 128-byte row = 64 consecutive M/N, rows = K index): `lbo_bytes` is
//     then the distance between 64-wide M/N blocks, `sbo_bytes` the distance between 8-row K groups.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// ---- instruction descriptor for kind::f16 (bf16 x bf16 -> f32) --------------------------------
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Warp-uniform variants for a dedicated MMA warp: EVERY lane of the (converged) warp executes the
// call, one elected lane issues.  With the issuing code in warp-uniform control flow the descriptors
// live in uniform registers; under `if (lane == 0)` the compiler wraps every UTCHMMA in an
// ELECT / BRA.U.ANY "waterfall" loop (~15 instructions per MMA).
__device__ __forceinline__ void mma_bf16_ss_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, pe;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n"
      :
      : "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_arrive_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pe;\n\t"
      "elect.sync _|pe, 0xffffffff;\n\t"
      "@pe mbarrier.arrive.shared::cta.b64 _, [%0];\n\t"
      "}\n"
      :
      : "r"(smem_u32(bar))
      : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :
               : "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :
               : "r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" : : "r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes (st.shared / cp.async) -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base_lane + i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- programmatic dependent launch (see common.cuh) ---------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" : : "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Spins until the phase with the given parity completes.  A malformed MMA descriptor would
// otherwise hang the GPU: after ~2 s the kernel traps (reported as a launch failure) instead.
// The wait is RELAXED: every use waits for a tcgen05.commit and is followed by
// tcgen05.fence::after_thread_sync, which is what orders the TMEM reads / tile re-use after the
// MMAs.  The default (.acquire) form makes the thread drain all of its outstanding global loads,
// stores and cp.async copies first (ncu: 11 % of the backward kernel's stall samples were "membar"
// on this instruction), i.e. it serialises the next tile's prefetch with the tensor-core wait.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (done) return;
    // a failed try costs issue slots the co-resident CTA could use (ncu: 9 % of the backward
    // kernel's instructions were this loop): back off before the next one
    __nanosleep(64);
    if (spin > (1u << 24)) __trap();  // ~seconds: a protocol error, not a wait
  }
}

// ---- cp.async (16 bytes, zero-fill when src_bytes == 0) ------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" : : "r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" : : "n"(N) : "memory"); }

// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a 128-byte-swizzled tile
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
  return static_cast<uint32_t>(row) * 128u + (static_cast<uint32_t>(chunk ^ (row & 7)) << 4);
}

// bf16-path activation: sigmoid through ONE MUFU op (tanh.approx, rel. error ~2^-11, well below
// the bf16 rounding of the stored result); the fp32 engine keeps expf + IEEE division.
__device__ __forceinline__ float fast_sigmoid(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float fast_silu(float x) { return x * fast_sigmoid(x); }
__device__ __forceinline__ float fast_silu_grad(float x) {
  const float s = fast_sigmoid(x);
  return s * fmaf(x, 1.0f - s, 1.0f);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// ---- role-warp synchronisation (warp-specialised kernels) -----------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" : : "r"(smem_u32(bar)) : "memory");
}
// wait with back-off (role warps wait for microseconds; a hot spin would steal issue slots from
// the epilogue warps).  Traps after ~seconds instead of hanging the GPU on a protocol error.
// RELAXED = true: no acquire fence (which would first drain the thread's outstanding global stores /
// cp.async copies): for waits whose consumer is ordered by tcgen05.fence::after_thread_sync (TMEM
// reads after a tcgen05.commit) or that only guard the re-use of a buffer (producer: stage free).
template <int SLEEP_NS, bool RELAXED = false>
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    if (RELAXED) {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(addr), "r"(parity)
          : "memory");
    } else {
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t"
          "}\n"
          : "=r"(done)
          : "r"(addr), "r"(parity)
          : "memory");
    }
    if (done) return;
    __nanosleep(SLEEP_NS);
    if (spin > (1u << 24)) __trap();
  }
}
// Wait with the hardware's own suspension: try_wait with a suspend-time hint parks the warp until the phase
// completes (or the hint expires) without issuing instructions, and wakes it immediately -- instead of a
// poll / nanosleep loop, whose instructions were 37 % of everything the pipelined backward issued
// (profiles/r02v_ws_bwd_edge_lines.txt).  Relaxed: the consumer is ordered by tcgen05.fence::after_thread_sync.
__device__ __forceinline__ void mbar_wait_suspend(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(1000000u)
        : "memory");
    if (done) return;
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" : : "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 ws_pack8(const float* f) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
// bf16x2 -> two floats with integer ops (exact)
__device__ __forceinline__ void ws_unpack2(uint32_t u, float& lo, float& hi) {
  lo = __uint_as_float(u << 16);
  hi = __uint_as_float(u & 0xFFFF0000u);
}
__device__ __forceinline__ void ws_unpack8(const uint4& u, float* f) {
  ws_unpack2(u.x, f[0], f[1]); ws_unpack2(u.y, f[2], f[3]);
  ws_unpack2(u.z, f[4], f[5]); ws_unpack2(u.w, f[6], f[7]);
}

// TMEM store: 32 lanes x 16 columns of 32-bit (thread i of the warp writes lane base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- segment sum of 64-wide bf16 rows into one 128-row swizzled smem tile ------------------------
// Row r (< nrows) of the tile = sc[r] * sum_{e in [lo[r], hi[r])} src[sbase[r]*8 + (perm ? perm[e] : e)*64 ..]
// (lo[r] == ~0u: zero row).  Work items = (row, edge group g < G, 16-byte chunk j < 8) with
// G = 128 / nrows (1, 2 or 4): short tiles of long segments split every segment over G lanes.  A
// thread owns 1024 / NTHREADS items (one per pass) and walks them TOGETHER, two edges per item per
// iteration, so 2 * passes independent row loads are in flight per thread.  Each (row, g) sum is one
// chain in ascending edge order; the G partial sums of a row are combined with a fixed shuffle tree:
// bit-reproducible.  fp32 accumulation, one rounding to bf16.  Optionally also stores the rows to
// out_rows[orow[r]] (global).
__device__ __forceinline__ void seg_add8(float* s, const uint4& u) {
  float2 t;
  t = unpack_bf16x2(u.x); s[0] += t.x; s[1] += t.y;
  t = unpack_bf16x2(u.y); s[2] += t.x; s[3] += t.y;
  t = unpack_bf16x2(u.z); s[4] += t.x; s[5] += t.y;
  t = unpack_bf16x2(u.w); s[6] += t.x; s[7] += t.y;
}

// PV: items of a thread that can be valid (rows r = item >> (3 + lg) of the items p >= PV must lie beyond `nrows`:
// tiles of at most 16 rows with 256 threads use PV = 2); U: source rows per item and iteration.  The gather
// indices of iteration i + 1 are loaded while the rows of iteration i are in flight (one dependent latency per
// iteration instead of two); the rows of a segment are added in the same order whatever PV / U.
// PREFETCH = false keeps the index loads at the top of the iteration (the forward's aggregation prologue: short
// segments, 128 registers already in use -- the prefetch registers spilled there and cost 20 %).
template <int NTHREADS, int PV = 1024 / NTHREADS, int U = 2, bool PREFETCH = false>
__device__ __forceinline__ void seg_sum_tile(const __nv_bfloat16* __restrict__ src, const int32_t* __restrict__ perm,
                                             const uint32_t* lo, const uint32_t* hi, const uint32_t* sbase,
                                             const float* sc, const int32_t* orow, uint8_t* X,
                                             __nv_bfloat16* __restrict__ out_rows, int tid, int nrows = 128,
                                             long long* trace = nullptr) {
  constexpr int P = PV;  // items (passes) per thread
  if (trace) trace[10] = clock64();
  const int lg = nrows >= 128 ? 0 : (nrows >= 64 ? 1 : 2);  // log2(G); nrows < 32 still uses G = 4
  const uint32_t G = 1u << lg;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  float acc[P][8];
  uint32_t e[P], h[P];
  const __nv_bfloat16* sb[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int item = p * NTHREADS + tid;
    const int j = item & 7, g = (item >> 3) & (G - 1), r = item >> (3 + lg);
    const uint32_t l = r < nrows ? lo[r] : 0xFFFFFFFFu;
    const bool valid = l != 0xFFFFFFFFu;
    e[p] = valid ? l + g : 0u;
    h[p] = valid ? hi[r] : 0u;
    sb[p] = src + (valid ? (static_cast<long long>(sbase[r]) << 3) : 0ll) + j * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[p][q] = 0.f;
  }
  if (trace) trace[11] = clock64();
  uint32_t ix[P][U];
  bool act[P][U];
  auto fetch_indices = [&]() {
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t ee = e[p] + u * G;
        act[p][u] = ee < h[p];
        ix[p][u] = act[p][u] ? (perm ? static_cast<uint32_t>(perm[ee]) : ee) : 0u;
      }
  };
  if (PREFETCH) fetch_indices();
  int iters = 0;
  while (true) {
    ++iters;
    if (!PREFETCH) fetch_indices();
    uint4 v[P][U];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int u = 0; u < U; ++u)
        v[p][u] = act[p][u] ? *reinterpret_cast<const uint4*>(sb[p] + static_cast<long long>(ix[p][u]) * 64) : zero4;
    bool more = false;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      e[p] += U * G;
      more |= e[p] < h[p];
    }
    if (PREFETCH && more) fetch_indices();  // next iteration's indices, in flight together with this iteration's rows
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
      for (int u = 0; u < U; ++u) seg_add8(acc[p], v[p][u]);
    if (!more) break;
  }
  if (trace) { trace[12] = clock64(); trace[15] = iters; }
  __syncwarp();
  if (trace) trace[13] = clock64();
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int item = p * NTHREADS + tid;
    const int j = item & 7, g = (item >> 3) & (G - 1), r = item >> (3 + lg);
    const float scale = r < nrows ? sc[r] : 1.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v = acc[p][q];
      if (lg >= 1) v += __shfl_xor_sync(0xffffffffu, v, 8);
      if (lg >= 2) v += __shfl_xor_sync(0xffffffffu, v, 16);
      acc[p][q] = v * scale;
    }
    if (g == 0 && r < nrows) {
      uint4 o;
      o.x = pack_bf16x2(acc[p][0], acc[p][1]); o.y = pack_bf16x2(acc[p][2], acc[p][3]);
      o.z = pack_bf16x2(acc[p][4], acc[p][5]); o.w = pack_bf16x2(acc[p][6], acc[p][7]);
      *reinterpret_cast<uint4*>(X + sw128_offset(r, j)) = o;
      if (out_rows && lo[r] != 0xFFFFFFFFu)
        *reinterpret_cast<uint4*>(out_rows + static_cast<long long>(orow[r]) * 64 + j * 8) = o;
    }
  }
  if (trace) trace[14] = clock64();
}

// column sums over the 32 lanes of a warp (fixed order): a transposing butterfly that halves the
// columns a lane holds at every step.  N = 32: lane l ends with the sum of column l; N = 16: lanes
// 2c and 2c+1 both end with the sum of column c.
template <int N>
__device__ __forceinline__ float warp_colsum(float (&v)[N], int lane) {
  int n = N;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      const int h = n / 2;
      const bool hi = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < N / 2; ++i) {
        if (i < h) {
          const float send = hi ? v[i] : v[i + h];
          const float keep = hi ? v[i + h] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      n = h;
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
  return v[0];
}

// ---- packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2) ------------------------------------------------
// Two IEEE fp32 operations (round-to-nearest, the same results as two scalar instructions) per issued
// instruction: the element-wise phases of the pipelined kernels are bound by instruction issue, not by the
// fp32 pipe.  A pair lives in an aligned 64-bit register pair; packing / unpacking adjacent registers is free.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 f2_dup(float x) { return f2_pack(x, x); }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float f2_hsum(f32x2 v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return lo + hi;
}

// SiLU of eight values in place, two per instruction: the operations of fast_silu (0.5 x, tanh, fma, x s), so
// the results are bit-identical to the scalar form
__device__ __forceinline__ void silu8_pairs(float* h) {
  const f32x2 half2 = f2_dup(0.5f);
#pragma unroll
  for (int q = 0; q < 8; q += 2) {
    const f32x2 hp = f2_pack(h[q], h[q + 1]);
    float t0, t1;
    f2_unpack(f2_mul(half2, hp), t0, t1);
    asm("tanh.approx.f32 %0, %0;" : "+f"(t0));
    asm("tanh.approx.f32 %0, %0;" : "+f"(t1));
    f2_unpack(f2_mul(hp, f2_fma(half2, f2_pack(t0, t1), half2)), h[q], h[q + 1]);
  }
}

// ---- packed bf16x2 arithmetic (two elements per instruction) -----------------------------------------
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint32_t bf2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t bf2_tanh(uint32_t a) {
  uint32_t d;
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// ---- K-blocks of the layer-1 input ----------------------------------------------------------------
// A source of width w contributes ceil(w/64) blocks of 64 columns (the last one zero-padded): the
// unit of the gather (one 128x64 bf16 smem tile), of the W1 image and of the dW1 / dX products.
struct KBlocks {
  int nblk;
  int src[P4C_MAX_SRC];      // source the block belongs to
  int col0[P4C_MAX_SRC];     // first column inside the source (multiple of 64)
  int kstart[P4C_MAX_SRC];   // first column of w1 the block multiplies
  int kvalid[P4C_MAX_SRC];   // valid columns (1..64)
  int nchunks[P4C_MAX_SRC];  // 16-byte chunks to fetch per row: ceil(kvalid/8); reads up to ld
};
inline __host__ __device__ bool make_kblocks(int nsrc, const int* widths, KBlocks* kb) {
  int n = 0, k = 0;
  for (int s = 0; s < nsrc; ++s) {
    for (int c0 = 0; c0 < widths[s]; c0 += 64) {
      if (n >= P4C_MAX_SRC) return false;
      kb->src[n] = s;
      kb->col0[n] = c0;
      kb->kstart[n] = k + c0;
      kb->kvalid[n] = widths[s] - c0 < 64 ? widths[s] - c0 : 64;
      kb->nchunks[n] = (kb->kvalid[n] + 7) / 8;
      ++n;
    }
    k += widths[s];
  }
  kb->nblk = n;
  for (int i = n; i < P4C_MAX_SRC; ++i) kb->src[i] = kb->col0[i] = kb->kstart[i] = kb->kvalid[i] = kb->nchunks[i] = 0;
  return true;
}

// byte offsets inside a P4C_PACK_BYTES weight image
constexpr int PK_W1 = 0;                 // 3 x 8 KB  [n = hidden unit][k]   (forward GEMM 1)
constexpr int PK_W2 = 3 * 8192;          // [n = output unit][k = hidden]     (forward GEMM 2, bwd recompute)
constexpr int PK_W2T = 4 * 8192;         // [n = hidden][k = output unit]     (bwd: da1 = dh2 W2)
constexpr int PK_W1T = 5 * 8192;         // 3 x 8 KB  [n = input col][k = hidden unit]  (bwd: dX = dh1 W1)
constexpr int PK_VEC = 8 * 8192;         // b1, b2, gamma, beta (4 x 64 fp32)
// bias tile [n][k] (8 KB, same layout as a W1 block): b1[n] as bf16 hi + lo at k = 0, 1 (K-slice 0) and b2[n] at
// k = 16, 17 (K-slice 1), zeros elsewhere.  One extra MMA per GEMM with an all-ones A operand adds the bias inside
// the product (hi + lo = the fp32 bias to 2^-17 relative): the row-per-thread epilogues then neither load the bias
// vectors from shared memory (a broadcast 16-byte load costs four wavefronts of the L1 data pipe per warp) nor add them.
constexpr int PK_BIAS = 8 * 8192 + 4 * 64 * 4;
static_assert(PK_BIAS + 8192 == P4C_PACK_BYTES, "weight image layout");

}  // namespace tc
}  // namespace p4c
