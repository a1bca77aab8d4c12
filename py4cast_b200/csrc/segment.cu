// CSR construction, deterministic segment sums and small streaming utilities.
//
// Replaces PyG's COO scatter (`scatter_add` over edge_index[1]) used by mfai
// InteractionNet.aggregate and its autograd transpose: edges are grouped
// by receiver (or sender) once at set-up time, and every aggregation is a gather-free
// (receiver-sorted graphs) or permuted segment sum in a fixed order: no float atomics.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace p4c {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }

static int g_pdl = -1;
static bool g_pdl_block_next = false;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("P4C_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
void pdl_block_next() { g_pdl_block_next = true; }
bool pdl_take_attr() {
  const bool blocked = g_pdl_block_next;
  g_pdl_block_next = false;
  return pdl_enabled() && !blocked;
}

// ----------------------------------------------------------------------------- CSR
__global__ void csr_hist_kernel(const long long* __restrict__ keys, long long E, int num_keys,
                                int* __restrict__ counts, int* __restrict__ bad) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E;
       e += (long long)gridDim.x * blockDim.x) {
    const long long k = keys[e];
    if (k < 0 || k >= num_keys) {
      *bad = 1;
      continue;
    }
    atomicAdd(counts + k, 1);  // integer atomics: result independent of order
  }
}

// single-block exclusive scan: rowptr[0..n] from counts[0..n-1]
__global__ void csr_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ rowptr) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + tid;
    const int v = i < n ? counts[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int t = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;  // inclusive scan of warp totals
    }
    __syncthreads();
    const int carry = carry_s;
    const int prefix = carry + (wid > 0 ? warp_tot[wid - 1] : 0) + x - v;  // exclusive
    if (i < n) rowptr[i] = prefix;
    __syncthreads();
    if (tid == blockDim.x - 1) carry_s = prefix + v;
    __syncthreads();
  }
  if (tid == 0) rowptr[n] = carry_s;
}

// unordered placement of edge ids into their key's segment
__global__ void csr_fill_kernel(const long long* __restrict__ keys, long long E, int num_keys,
                                const int* __restrict__ rowptr, int* __restrict__ cursor,
                                int* __restrict__ tmp) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < E;
       e += (long long)gridDim.x * blockDim.x) {
    const long long kk = keys[e];
    if (kk < 0 || kk >= num_keys) continue;  // reported by csr_hist_kernel; never index with it
    const int k = (int)kk;
    const int slot = atomicAdd(cursor + k, 1);
    tmp[rowptr[k] + slot] = (int)e;
  }
}

// rank every edge id inside its segment -> ascending edge id (== stable sort by key)
__global__ void csr_rank_kernel(const long long* __restrict__ keys, int num_keys,
                                const int* __restrict__ rowptr, const int* __restrict__ tmp,
                                int* __restrict__ perm) {
  const long long placed = rowptr[num_keys];  // == E unless out-of-range keys were skipped
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < placed;
       j += (long long)gridDim.x * blockDim.x) {
    const int e = tmp[j];
    const int k = (int)keys[e];
    const int lo = rowptr[k], hi = rowptr[k + 1];
    int rank = 0;
    for (int q = lo; q < hi; ++q) rank += tmp[q] < e;
    perm[lo + rank] = e;
  }
}

__global__ void index_to_i32_kernel(const long long* __restrict__ src, int* __restrict__ dst,
                                    long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = (int)src[i];
}

// ------------------------------------------------------------------- segment sum (64 wide)
// One warp per (sample, segment).  Lanes 0-15 and 16-31 each walk every other edge of the
// segment with one float4 (4 columns) per lane; the two half-warp partial sums are combined
// with one shuffle.  Fixed association order => bit-reproducible.
template <typename T>
__global__ void __launch_bounds__(256)
segment_sum_kernel(const T* __restrict__ src, long long src_bs, const int* __restrict__ rowptr,
                   const int* __restrict__ perm, int num_seg, int batch, T* __restrict__ out,
                   long long out_bs, int mean, int accumulate) {
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, q = lane & 15;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long total = (long long)num_seg * batch;
  for (long long w = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); w < total;
       w += nwarps) {
    const int b = (int)(w / num_seg), s = (int)(w % num_seg);
    const int lo = rowptr[s], hi = rowptr[s + 1];
    const T* base = src + b * src_bs + 4 * q;
    float4 acc0 = make_float4(0, 0, 0, 0), acc1 = acc0;
    int j = lo + half;
    for (; j + 2 < hi; j += 4) {  // two independent loads in flight per lane
      const long long e0 = perm ? perm[j] : j;
      const long long e1 = perm ? perm[j + 2] : j + 2;
      const float4 v0 = load4<T>(base + e0 * HIDDEN_W);
      const float4 v1 = load4<T>(base + e1 * HIDDEN_W);
      acc0 = f4add(acc0, v0);
      acc1 = f4add(acc1, v1);
    }
    if (j < hi) {
      const long long e0 = perm ? perm[j] : j;
      acc0 = f4add(acc0, load4<T>(base + e0 * HIDDEN_W));
    }
    float4 acc = f4add(acc0, acc1);
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 16);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 16);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, 16);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, 16);
    if (half == 0) {
      if (mean) {
        const float sc = 1.0f / (float)max(1, hi - lo);
        acc.x *= sc; acc.y *= sc; acc.z *= sc; acc.w *= sc;
      }
      T* o = out + b * out_bs + (long long)s * HIDDEN_W + 4 * q;
      if (accumulate) acc = f4add(acc, load4<T>(o));
      store4<T>(o, acc);
    }
  }
}

template <typename T>
__global__ void batch_sum_kernel(const T* __restrict__ in, long long bs, int batch, long long n,
                                 T* __restrict__ out, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < batch; ++b) s += to_f32<T>(in[b * bs + i]);
    if (accumulate) s += to_f32<T>(out[i]);
    out[i] = from_f32<T>(s);
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<D>(to_f32<S>(src[i]));
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ dst,
                           long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f32<T>(to_f32<T>(a[i]) + to_f32<T>(b[i]));
}

static int grid_for(long long n, int block, int max_blocks = kNumSMs * 8) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  return (int)(g < max_blocks ? g : max_blocks);
}

}  // namespace p4c

using namespace p4c;

extern "C" const char* p4c_last_error(void) { return g_err; }
extern "C" int p4c_version(void) { return 100; }
extern "C" int64_t p4c_launch_count(void) { return g_launches.load(); }
extern "C" void p4c_reset_launch_count(void) { g_launches = 0; }
extern "C" int p4c_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(p4c_src_t);
    case 1: return (int)sizeof(p4c_mlp_t);
    case 2: return (int)sizeof(p4c_mlp_bwd_t);
    case 3: return (int)sizeof(p4c_ar_x_t);
    case 4: return (int)sizeof(p4c_adamw_state_t);
    case 5: return (int)sizeof(p4c_pack_t);
    case 6: return (int)sizeof(p4c_wgrad_t);
    case 7: return (int)sizeof(p4c_seg_part_t);
    default: return -1;
  }
}

extern "C" int p4c_csr_build(const int64_t* keys, int64_t E, int32_t num_keys, int32_t* rowptr,
                             int32_t* perm, int32_t* scratch, void* stream) {
  P4C_REQUIRE(num_keys >= 0 && E >= 0, "csr_build: negative size");
  P4C_REQUIRE(E < (1ll << 31) - 1, "csr_build: E=%lld exceeds int32 indexing", (long long)E);
  P4C_REQUIRE(rowptr && scratch && (E == 0 || (keys && perm)), "csr_build: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: [0, num_keys] counts/cursor (+1 slot = bad flag), then tmp[E]
  int* counts = scratch;
  int* bad = scratch + num_keys;
  int* tmp = scratch + num_keys + 1;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)(num_keys + 1), st);
  if (e != cudaSuccess) {
    set_error("csr_build: memset failed: %s", cudaGetErrorString(e));
    return P4C_ECUDA;
  }
  if (E > 0) {
    csr_hist_kernel<<<grid_for(E, 256), 256, 0, st>>>((const long long*)keys, E, num_keys, counts, bad);
    P4C_CHECK_LAUNCH("csr_hist");
  }
  if (E > 0) {
    // Set-up path (runs once per edge set, never under graph capture): read the out-of-range flag
    // back before any kernel indexes with the keys.  The one entry point that synchronises.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    if (cap == cudaStreamCaptureStatusNone) {
      int bad_h = 0;
      e = cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        set_error("csr_build: reading the key check failed: %s", cudaGetErrorString(e));
        return P4C_ECUDA;
      }
      P4C_REQUIRE(bad_h == 0, "csr_build: a key lies outside [0, %d)", num_keys);
    }
  }
  csr_scan_kernel<<<1, 1024, 0, st>>>(counts, num_keys, rowptr);
  P4C_CHECK_LAUNCH("csr_scan");
  if (E > 0) {
    e = cudaMemsetAsync(counts, 0, sizeof(int) * (size_t)num_keys, st);
    if (e != cudaSuccess) {
      set_error("csr_build: memset failed: %s", cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    csr_fill_kernel<<<grid_for(E, 256), 256, 0, st>>>((const long long*)keys, E, num_keys, rowptr, counts, tmp);
    P4C_CHECK_LAUNCH("csr_fill");
    csr_rank_kernel<<<grid_for(E, 256), 256, 0, st>>>((const long long*)keys, num_keys, rowptr, tmp, perm);
    P4C_CHECK_LAUNCH("csr_rank");
  }
  return P4C_OK;
}

extern "C" int p4c_index_to_i32(const int64_t* src, int32_t* dst, int64_t n, void* stream) {
  P4C_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "index_to_i32: bad arguments");
  if (n == 0) return P4C_OK;
  index_to_i32_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const long long*)src, dst, n);
  P4C_CHECK_LAUNCH("index_to_i32");
  return P4C_OK;
}

extern "C" int p4c_segment_sum(const void* src, int64_t src_bs, const int32_t* rowptr,
                               const int32_t* perm, int32_t num_seg, int32_t batch, void* out,
                               int64_t out_bs, int mean, int accumulate, int dtype, void* stream) {
  P4C_REQUIRE(src && rowptr && out, "segment_sum: null pointer");
  P4C_REQUIRE(num_seg > 0 && batch > 0, "segment_sum: empty problem");
  P4C_REQUIRE(src_bs % 4 == 0 && out_bs % 4 == 0, "segment_sum: batch strides must be multiples of 4");
  const long long total = (long long)num_seg * batch;
  const int grid = grid_for(total, 8, kNumSMs * 16);  // 8 warps per block
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == P4C_F32)
    segment_sum_kernel<float><<<grid, 256, 0, st>>>((const float*)src, src_bs, rowptr, perm, num_seg,
                                                    batch, (float*)out, out_bs, mean, accumulate);
  else if (dtype == P4C_BF16)
    segment_sum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, src_bs, rowptr,
                                                            perm, num_seg, batch, (__nv_bfloat16*)out,
                                                            out_bs, mean, accumulate);
  else
    P4C_REQUIRE(false, "segment_sum: bad dtype %d", dtype);
  P4C_CHECK_LAUNCH("segment_sum");
  return P4C_OK;
}

extern "C" int p4c_batch_sum(const void* in, int64_t bs, int32_t batch, int64_t n, void* out,
                             int accumulate, int dtype, void* stream) {
  P4C_REQUIRE(in && out && batch > 0 && n > 0, "batch_sum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == P4C_F32)
    batch_sum_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)in, bs, batch, n, (float*)out, accumulate);
  else if (dtype == P4C_BF16)
    batch_sum_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, st>>>((const __nv_bfloat16*)in, bs, batch, n,
                                                                      (__nv_bfloat16*)out, accumulate);
  else
    P4C_REQUIRE(false, "batch_sum: bad dtype %d", dtype);
  P4C_CHECK_LAUNCH("batch_sum");
  return P4C_OK;
}

extern "C" int p4c_cast(const void* src, int sd, void* dst, int dd, int64_t n, void* stream) {
  P4C_REQUIRE(src && dst && n > 0, "cast: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(n, 256);
  if (sd == P4C_F32 && dd == P4C_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (sd == P4C_BF16 && dd == P4C_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (sd == P4C_F32 && dd == P4C_F32)
    cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (sd == P4C_BF16 && dd == P4C_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else
    P4C_REQUIRE(false, "cast: bad dtypes %d -> %d", sd, dd);
  P4C_CHECK_LAUNCH("cast");
  return P4C_OK;
}

extern "C" int p4c_add(const void* a, const void* b, void* dst, int64_t n, int dtype, void* stream) {
  P4C_REQUIRE(a && b && dst && n > 0, "add: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == P4C_F32)
    add_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)a, (const float*)b, (float*)dst, n);
  else if (dtype == P4C_BF16)
    add_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,
                                                                (__nv_bfloat16*)dst, n);
  else
    P4C_REQUIRE(false, "add: bad dtype %d", dtype);
  P4C_CHECK_LAUNCH("add");
  return P4C_OK;
}
