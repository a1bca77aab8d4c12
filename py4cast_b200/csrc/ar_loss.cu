// AR-step state update / input concatenation, losses and the flat AdamW step.
//
// Streaming (HBM-bound) kernels: flat, fully coalesced indexing over the innermost
// (node, feature) plane; reductions use fixed-order trees (no float atomics).
//
// Reference lines replaced (relative to the reference tree):
//   p4c_ar_concat         py4cast/lightning.py:711-767  (_next_x: select x T_in + cat)
//   p4c_ar_step_forward   py4cast/lightning.py:598-662  (clone, scaled/diff update, border blend,
//                                                        prev-state roll, torch.stack slot)
//   p4c_loss_*            py4cast/losses.py:130-169 (WeightedLoss), :186-210 (ScaledLoss)
//   p4c_adamw_*           py4cast/lightning.py:442-467  (torch.optim.AdamW + cosine/min-lr schedule)
#include <math.h>

#include "common.cuh"

namespace p4c {

struct ArX {
  p4c_ar_x_t a;
};

template <typename TX>
__device__ __forceinline__ float ar_x_value(const p4c_ar_x_t& a, int b, int n, int k,
                                            bool* is_last_slot, int* f_out) {
  const int F = a.f_state;
  const int ks = a.num_prev * F;
  *is_last_slot = false;
  if (k < ks) {
    const int t = k / F, f = k - t * F;
    if (t == a.num_prev - 1) {
      *is_last_slot = true;
      *f_out = f;
      return 0.f;
    }
    return a.prev[t][(long long)b * a.prev_batch_stride[t] + (long long)n * F + f];
  }
  k -= ks;
  if (k < a.f_static) return a.stat[(long long)b * a.stat_batch_stride + (long long)n * a.f_static + k];
  k -= a.f_static;
  return a.forcing[(long long)b * a.forcing_batch_stride + (long long)n * a.f_forcing + k];
}

template <typename TX>
__global__ void __launch_bounds__(256)
ar_concat_kernel(const p4c_ar_x_t a, TX* __restrict__ x) {
  const int F = a.f_state;
  const int f_in = a.num_prev * F + a.f_static + a.f_forcing;
  const int ld = a.x_ld > 0 ? a.x_ld : f_in;  // padded row stride: padding columns are zeroed
  const long long per_b = (long long)a.nodes * ld;
  const long long total = per_b * a.batch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const long long rem = i - (long long)b * per_b;
    const int n = (int)(rem / ld), k = (int)(rem - (long long)n * ld);
    if (k >= f_in) {
      x[i] = from_f32<TX>(0.f);
      continue;
    }
    bool last;
    int f;
    float v = ar_x_value<TX>(a, b, n, k, &last, &f);
    if (last) {
      const int t = a.num_prev - 1;
      v = a.prev[t][(long long)b * a.prev_batch_stride[t] + (long long)n * F + f];
    }
    x[i] = from_f32<TX>(v);
  }
}

struct ArStep {
  const void* y;
  const float* prev_last;
  long long prev_last_bs;
  const float* truth;
  long long truth_bs;
  const float* border;
  const float* std_;
  const float* mean_;
  int batch, nodes, F;
  float* new_state;
  long long new_state_bs;
};

template <typename TY>
__device__ __forceinline__ float ar_new_state(const ArStep& s, int b, int n, int f) {
  const long long o = (long long)n * s.F + f;
  const float yv = to_f32<TY>(reinterpret_cast<const TY*>(s.y)[(long long)b * s.nodes * s.F + o]);
  const float pl = s.prev_last[(long long)b * s.prev_last_bs + o];
  float pred = s.std_ ? pl + yv * s.std_[f] + s.mean_[f] : pl + yv;
  if (s.border) {
    const float bm = s.border[n];
    // same expression as the reference: border*truth + interior*pred, interior = 1 - border
    pred = bm * s.truth[(long long)b * s.truth_bs + o] + (1.0f - bm) * pred;
  }
  return pred;
}

// update only (last AR step: no next x)
template <typename TY>
__global__ void __launch_bounds__(256) ar_update_kernel(const ArStep s) {
  const long long per_b = (long long)s.nodes * s.F;
  const long long total = per_b * s.batch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const long long rem = i - (long long)b * per_b;
    const int n = (int)(rem / s.F), f = (int)(rem - (long long)n * s.F);
    s.new_state[(long long)b * s.new_state_bs + rem] = ar_new_state<TY>(s, b, n, f);
  }
}

// update + concat of the next model input in one pass over x_next
template <typename TY, typename TX>
__global__ void __launch_bounds__(256)
ar_step_kernel(const ArStep s, const p4c_ar_x_t a, TX* __restrict__ x) {
  const int F = a.f_state;
  const int f_in = a.num_prev * F + a.f_static + a.f_forcing;
  const int ld = a.x_ld > 0 ? a.x_ld : f_in;  // padded row stride: padding columns are zeroed
  const long long per_b = (long long)a.nodes * ld;
  const long long total = per_b * a.batch;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const long long rem = i - (long long)b * per_b;
    const int n = (int)(rem / ld), k = (int)(rem - (long long)n * ld);
    if (k >= f_in) {
      x[i] = from_f32<TX>(0.f);
      continue;
    }
    bool last;
    int f;
    float v = ar_x_value<TX>(a, b, n, k, &last, &f);
    if (last) {
      v = ar_new_state<TY>(s, b, n, f);
      s.new_state[(long long)b * s.new_state_bs + (long long)n * F + f] = v;
    }
    x[i] = from_f32<TX>(v);
  }
}

// Tiled variant of ar_concat / ar_step (the one used whenever the x tile fits shared memory):
// one CTA = AR_TN consecutive nodes of one sample.  Every input array is contiguous over a node
// range, so each is read with fully coalesced loads into an fp32 staging tile laid out exactly like
// the x rows ([older states | new state | statics | forcing | zero padding]); the new state is
// computed into its slot (and written to the prediction slab), then the tile is converted and
// streamed out as one contiguous block.  No 64-bit divisions, no strided stores.
constexpr int AR_TN = 64;

template <typename T>
__device__ __forceinline__ void ar_store_pair(T* p, float a, float b);
template <>
__device__ __forceinline__ void ar_store_pair<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <>
__device__ __forceinline__ void ar_store_pair<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

template <typename TY, typename TX>
__global__ void __launch_bounds__(256)
ar_tile_kernel(const ArStep s, const p4c_ar_x_t a, TX* __restrict__ x, int has_update) {
  extern __shared__ float S[];  // [AR_TN][ld]
  const int F = a.f_state, T = a.num_prev, Fs = a.f_static, Ff = a.f_forcing;
  const int f_in = T * F + Fs + Ff;
  const int ld = a.x_ld > 0 ? a.x_ld : f_in;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * AR_TN;
  const int nn = min(AR_TN, a.nodes - n0);
  const int tid = threadIdx.x;
  const float invF = 1.0f / (float)F;
  // older states (and, without update, the newest one too)
  const int ncopy = has_update ? T - 1 : T;
  for (int t = 0; t < ncopy; ++t) {
    const float* src = a.prev[t] + (long long)b * a.prev_batch_stride[t] + (long long)n0 * F;
    for (int i = tid; i < nn * F; i += 256) {
      const int n = (int)(((float)i + 0.5f) * invF), f = i - n * F;
      S[n * ld + t * F + f] = src[i];
    }
  }
  if (Fs > 0) {
    const float* src = a.stat + (long long)b * a.stat_batch_stride + (long long)n0 * Fs;
    const float inv = 1.0f / (float)Fs;
    for (int i = tid; i < nn * Fs; i += 256) {
      const int n = (int)(((float)i + 0.5f) * inv), f = i - n * Fs;
      S[n * ld + T * F + f] = src[i];
    }
  }
  if (Ff > 0) {
    const float* src = a.forcing + (long long)b * a.forcing_batch_stride + (long long)n0 * Ff;
    const float inv = 1.0f / (float)Ff;
    for (int i = tid; i < nn * Ff; i += 256) {
      const int n = (int)(((float)i + 0.5f) * inv), f = i - n * Ff;
      S[n * ld + T * F + Fs + f] = src[i];
    }
  }
  const int npad = ld - f_in;
  for (int i = tid; i < nn * npad; i += 256) {
    const int n = i / npad, j = i - n * npad;
    S[n * ld + f_in + j] = 0.f;
  }
  if (has_update) {
    for (int i = tid; i < nn * F; i += 256) {
      const int n = (int)(((float)i + 0.5f) * invF), f = i - n * F;
      const float v = ar_new_state<TY>(s, b, n0 + n, f);
      s.new_state[(long long)b * s.new_state_bs + (long long)n0 * F + i] = v;
      S[n * ld + (T - 1) * F + f] = v;
    }
  }
  __syncthreads();
  TX* dst = x + ((long long)b * a.nodes + n0) * ld;
  const int total = nn * ld;
  if ((ld & 1) == 0) {
    for (int i = tid; i < total / 2; i += 256) ar_store_pair<TX>(dst + 2 * i, S[2 * i], S[2 * i + 1]);
  } else {
    for (int i = tid; i < total; i += 256) dst[i] = from_f32<TX>(S[i]);
  }
}

// Staged variant of the tile kernel (nodes a multiple of AR_TN, 16-byte aligned arrays, even row stride of x):
// every input array of the CTA's AR_TN nodes is one contiguous span, copied global -> shared with 16-byte
// cp.async (no registers, every copy of the CTA in flight together: the scalar-load tile kernel above keeps one
// 4-byte load per thread in flight and reaches 0.47 of the HBM peak); the new state is then computed from the
// staged spans four elements at a time (same expressions as ar_new_state) and the x rows are assembled from the
// spans through a per-column (offset, stride) table and streamed out as pairs.
__device__ __forceinline__ void ar_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" : : "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float ar_blend(float pl, float yv, const float* std_, const float* mean_, int f,
                                          const float* border, int n, float truth) {
  float pred = std_ ? pl + yv * std_[f] + mean_[f] : pl + yv;
  if (border) {
    const float bm = border[n];
    pred = bm * truth + (1.0f - bm) * pred;  // same expression as the reference (see ar_new_state)
  }
  return pred;
}

template <typename TY, typename TX>
__global__ void __launch_bounds__(256)
ar_stage_kernel(const ArStep s, const p4c_ar_x_t a, TX* __restrict__ x, int has_update) {
  extern __shared__ __align__(16) float R[];
  const int F = a.f_state, T = a.num_prev, Fs = a.f_static, Ff = a.f_forcing;
  const int f_in = T * F + Fs + Ff;
  const int ld = a.x_ld > 0 ? a.x_ld : f_in;
  const int b = blockIdx.y;
  const int n0 = blockIdx.x * AR_TN;
  const int tid = threadIdx.x;
  const int span = AR_TN * F;  // floats of one state span
  // layout (floats; every offset a multiple of 4): T state spans | statics | forcing | zeros(4) | prev_last |
  // truth | y | column table
  const int offS = T * span, offF = offS + AR_TN * Fs, offZ = offF + AR_TN * Ff;
  const int offPL = offZ + 4, offTR = offPL + span, offY = offTR + span;
  const int ybytes = span * (int)sizeof(TY);
  const int offTab = offY + (ybytes + 15) / 16 * 4;
  const uint32_t r0 = (uint32_t)__cvta_generic_to_shared(R);
  int2* tab = reinterpret_cast<int2*>(R + offTab);

  const int ncopy = has_update ? T - 1 : T;
  for (int t = 0; t < ncopy; ++t) {
    const float* src = a.prev[t] + (long long)b * a.prev_batch_stride[t] + (long long)n0 * F;
    for (int i = tid; i < span / 4; i += 256) ar_cp16(r0 + (t * span + 4 * i) * 4, src + 4 * i);
  }
  if (Fs > 0) {
    const float* src = a.stat + (long long)b * a.stat_batch_stride + (long long)n0 * Fs;
    for (int i = tid; i < AR_TN * Fs / 4; i += 256) ar_cp16(r0 + (offS + 4 * i) * 4, src + 4 * i);
  }
  if (Ff > 0) {
    const float* src = a.forcing + (long long)b * a.forcing_batch_stride + (long long)n0 * Ff;
    for (int i = tid; i < AR_TN * Ff / 4; i += 256) ar_cp16(r0 + (offF + 4 * i) * 4, src + 4 * i);
  }
  if (has_update) {
    const float* pl = s.prev_last + (long long)b * s.prev_last_bs + (long long)n0 * F;
    for (int i = tid; i < span / 4; i += 256) ar_cp16(r0 + (offPL + 4 * i) * 4, pl + 4 * i);
    if (s.border) {
      const float* tr = s.truth + (long long)b * s.truth_bs + (long long)n0 * F;
      for (int i = tid; i < span / 4; i += 256) ar_cp16(r0 + (offTR + 4 * i) * 4, tr + 4 * i);
    }
    const char* yb = reinterpret_cast<const char*>(s.y) + ((long long)b * s.nodes * F + (long long)n0 * F) * (long long)sizeof(TY);
    for (int i = tid; i < ybytes / 16; i += 256) ar_cp16(r0 + offY * 4 + 16 * i, yb + 16 * i);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (tid < 4) R[offZ + tid] = 0.f;
  for (int k = tid; k < ld; k += 256) {  // column k of an x row -> (offset of node 0's value, stride per node)
    int2 e;
    if (k < T * F) {
      const int t = k / F;
      e = make_int2(t * span + (k - t * F), F);
    } else if (k < T * F + Fs) {
      e = make_int2(offS + (k - T * F), Fs);
    } else if (k < f_in) {
      e = make_int2(offF + (k - T * F - Fs), Ff);
    } else {
      e = make_int2(offZ, 0);
    }
    tab[k] = e;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  if (has_update) {
    const float invF = 1.0f / (float)F;
    float* ns = s.new_state + (long long)b * s.new_state_bs + (long long)n0 * F;
    const TY* ys = reinterpret_cast<const TY*>(R + offY);
    for (int i = tid; i < span / 4; i += 256) {
      const int e0 = 4 * i;
      int n = (int)(((float)e0 + 0.5f) * invF), f = e0 - n * F;
      const float4 pl = *reinterpret_cast<const float4*>(R + offPL + e0);
      float4 tr = make_float4(0.f, 0.f, 0.f, 0.f);
      if (s.border) tr = *reinterpret_cast<const float4*>(R + offTR + e0);
      const float plv[4] = {pl.x, pl.y, pl.z, pl.w}, trv[4] = {tr.x, tr.y, tr.z, tr.w};
      float v[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[q] = ar_blend(plv[q], to_f32<TY>(ys[e0 + q]), s.std_, s.mean_, f, s.border, n0 + n, trv[q]);
        if (++f == F) {
          f = 0;
          ++n;
        }
      }
      const float4 o = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(R + (T - 1) * span + e0) = o;
      *reinterpret_cast<float4*>(ns + e0) = o;
    }
    __syncthreads();
  }

  TX* dst = x + ((long long)b * a.nodes + n0) * ld;
  const int half = ld >> 1;
  const float inv_half = 1.0f / (float)half;
  for (int p = tid; p < AR_TN * half; p += 256) {
    const int n = (int)(((float)p + 0.5f) * inv_half), k = 2 * (p - n * half);
    const int2 t0 = tab[k], t1 = tab[k + 1];
    ar_store_pair<TX>(dst + 2 * p, R[t0.x + n * t0.y], R[t1.x + n * t1.y]);
  }
}

static size_t ar_stage_smem(const p4c_ar_x_t* a, int y_bytes_per_elem) {
  const int F = a->f_state, T = a->num_prev;
  const int f_in = T * F + a->f_static + a->f_forcing;
  const int ld = a->x_ld > 0 ? a->x_ld : f_in;
  const size_t span = (size_t)AR_TN * F;
  size_t fl = T * span + (size_t)AR_TN * a->f_static + (size_t)AR_TN * a->f_forcing + 4 + 2 * span;
  fl += (span * y_bytes_per_elem + 15) / 16 * 4;
  return fl * 4 + (size_t)ld * 8;
}

static bool ar_al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static int g_ar_vector_paths = 1;  // p4c_set_ar_vector_paths (tests compare the staged / vectorised kernels with the scalar ones)

// the staged kernel's requirements (else the scalar tile kernel / flat kernels serve the call)
static bool ar_stage_ok(const p4c_ar_x_t* a, const void* x, int x_dtype, const ArStep* s, int y_dtype, size_t* smem) {
  const int F = a->f_state, T = a->num_prev;
  const int f_in = T * F + a->f_static + a->f_forcing;
  const int ld = a->x_ld > 0 ? a->x_ld : f_in;
  if (!g_ar_vector_paths || a->nodes % AR_TN != 0 || (ld & 1) || a->batch > 65535 || F > 512) return false;
  if (!ar_al16(x) && x_dtype == P4C_F32) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 3) != 0) return false;
  const int ncopy = s ? T - 1 : T;
  for (int t = 0; t < ncopy; ++t)
    if (!ar_al16(a->prev[t]) || a->prev_batch_stride[t] % 4 != 0) return false;
  if (a->f_static > 0 && (!ar_al16(a->stat) || a->stat_batch_stride % 4 != 0)) return false;
  if (a->f_forcing > 0 && (!ar_al16(a->forcing) || a->forcing_batch_stride % 4 != 0)) return false;
  int ybytes = 4;
  if (s) {
    ybytes = y_dtype == P4C_BF16 ? 2 : 4;
    if (!ar_al16(s->prev_last) || s->prev_last_bs % 4 != 0) return false;
    if (s->border && (!ar_al16(s->truth) || s->truth_bs % 4 != 0)) return false;
    if (!ar_al16(s->y) || ((long long)a->nodes * F * ybytes) % 16 != 0) return false;
    if (!ar_al16(s->new_state) || s->new_state_bs % 4 != 0) return false;
  }
  *smem = ar_stage_smem(a, ybytes);
  return *smem <= 48 * 1024;
}

struct ArBwd {
  const float* g_new;
  long long g_new_bs;
  const float* g_new2;
  long long g_new2_bs;
  const void* g_x;
  const float* border;
  const float* std_;
  int batch, nodes, F, num_prev, f_in;
  void* dy;
  float* d_prev[P4C_MAX_PREV];
};

template <typename TY, typename TX>
__global__ void __launch_bounds__(256) ar_step_bwd_kernel(const ArBwd s) {
  const long long per_b = (long long)s.nodes * s.F;
  const long long total = per_b * s.batch;
  const TX* gx = reinterpret_cast<const TX*>(s.g_x);
  TY* dy = reinterpret_cast<TY*>(s.dy);
  const int T = s.num_prev;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_b);
    const long long rem = i - (long long)b * per_b;
    const int n = (int)(rem / s.F), f = (int)(rem - (long long)n * s.F);
    const long long xrow = ((long long)b * s.nodes + n) * s.f_in;
    float G = s.g_new ? s.g_new[(long long)b * s.g_new_bs + rem] : 0.f;
    if (s.g_new2) G += s.g_new2[(long long)b * s.g_new2_bs + rem];
    if (gx) G += to_f32<TX>(gx[xrow + (long long)(T - 1) * s.F + f]);
    const float interior = s.border ? 1.0f - s.border[n] : 1.0f;
    const float gi = G * interior;
    dy[i] = from_f32<TY>(s.std_ ? gi * s.std_[f] : gi);
    // states copied into x_next slots 0..T-2 are the step inputs prev[1..T-1]
    float dlast = gi;
    if (gx && T >= 2) dlast += to_f32<TX>(gx[xrow + (long long)(T - 2) * s.F + f]);
    if (s.d_prev[T - 1]) s.d_prev[T - 1][i] = dlast;
    if (gx) {
      for (int t = 1; t < T - 1; ++t)
        if (s.d_prev[t]) s.d_prev[t][i] = to_f32<TX>(gx[xrow + (long long)(t - 1) * s.F + f]);
    }
  }
}

// Vectorised form of ar_step_bwd_kernel: blockIdx.y = sample, a thread handles four consecutive elements of the
// (nodes x F) plane with 16-byte loads / stores of the state-shaped arrays and ONE exact 32-bit division (the scalar
// kernel pays two 64-bit divisions per element).  Same arithmetic per element: bit-identical results.  Needs
// nodes * F a multiple of 4 and 16-byte aligned arrays (p4c_ar_step_backward checks).
template <typename TY, typename TX>
__global__ void __launch_bounds__(256) ar_step_bwd_vec_kernel(const ArBwd s, unsigned long long magic_f) {
  const uint32_t per_b = (uint32_t)s.nodes * (uint32_t)s.F;
  const uint32_t e0 = 4u * (blockIdx.x * 256u + threadIdx.x);
  if (e0 >= per_b) return;
  const int b = blockIdx.y;
  const int T = s.num_prev, F = s.F;
  const TX* gx = reinterpret_cast<const TX*>(s.g_x);
  int n = (int)__umul64hi((unsigned long long)e0, magic_f);
  int f = (int)e0 - n * F;
  const long long i0 = (long long)b * per_b + e0;
  float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.g_new) g4 = *reinterpret_cast<const float4*>(s.g_new + (long long)b * s.g_new_bs + e0);
  float G[4] = {g4.x, g4.y, g4.z, g4.w};
  if (s.g_new2) {
    const float4 h = *reinterpret_cast<const float4*>(s.g_new2 + (long long)b * s.g_new2_bs + e0);
    G[0] += h.x; G[1] += h.y; G[2] += h.z; G[3] += h.w;
  }
  float dyv[4], dl[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const long long xrow = ((long long)b * s.nodes + n) * s.f_in;
    float Gq = G[q];
    if (gx) Gq += to_f32<TX>(gx[xrow + (long long)(T - 1) * F + f]);
    const float interior = s.border ? 1.0f - s.border[n] : 1.0f;
    const float gi = Gq * interior;
    dyv[q] = s.std_ ? gi * s.std_[f] : gi;
    float dlast = gi;
    if (gx && T >= 2) dlast += to_f32<TX>(gx[xrow + (long long)(T - 2) * F + f]);
    dl[q] = dlast;
    if (gx) {
      for (int t = 1; t < T - 1; ++t)
        if (s.d_prev[t]) s.d_prev[t][i0 + q] = to_f32<TX>(gx[xrow + (long long)(t - 1) * F + f]);  // T_in >= 3 only
    }
    if (++f == F) {
      f = 0;
      ++n;
    }
  }
  if (sizeof(TY) == 2) {
    uint2 o;
    __nv_bfloat162 lo = __floats2bfloat162_rn(dyv[0], dyv[1]), hi = __floats2bfloat162_rn(dyv[2], dyv[3]);
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(s.dy) + i0) = o;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(s.dy) + i0) = make_float4(dyv[0], dyv[1], dyv[2], dyv[3]);
  }
  if (s.d_prev[T - 1]) *reinterpret_cast<float4*>(s.d_prev[T - 1] + i0) = make_float4(dl[0], dl[1], dl[2], dl[3]);
}

// ------------------------------------------------------------------------------ losses
constexpr int LOSS_NT = 256;
constexpr int LOSS_CHUNK_ITERS = 64;  // node-iterations per CTA

__device__ __forceinline__ float loss_elem(float p, float t, float m, int kind) {
  const float d = p * m - t * m;
  return kind == 0 ? d * d : fabsf(d);
}

__host__ __device__ inline int loss_nodes_per_iter(int F) { return LOSS_NT / F; }
__host__ __device__ inline int loss_chunk_nodes(int F) { return loss_nodes_per_iter(F) * LOSS_CHUNK_ITERS; }

// partial[bt][chunk][F]  (per-feature interior-weighted sums, un-normalised)
__global__ void __launch_bounds__(LOSS_NT)
loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                    const float* __restrict__ mask, const float* __restrict__ interior, int nodes,
                    int F, int kind, float* __restrict__ partials) {
  __shared__ float red[LOSS_NT];
  const int npi = loss_nodes_per_iter(F);
  const int active = npi * F;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, nchunks = gridDim.x;
  const long long bt = blockIdx.y;
  const int n0 = chunk * loss_chunk_nodes(F);
  const int n1 = min(nodes, n0 + loss_chunk_nodes(F));
  float acc = 0.f;
  if (tid < active) {
    const int f = tid % F, no = tid / F;
    const long long base = bt * (long long)nodes * F;
    // four independent chains (fixed association order): four row loads in flight per thread
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    int n = n0 + no;
    for (; n + 3 * npi < n1; n += 4 * npi) {
      float pv[4], tv[4], mv[4], iv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long idx = base + (long long)(n + u * npi) * F + f;
        pv[u] = pred[idx];
        tv[u] = target[idx];
        mv[u] = mask ? mask[idx] : 1.0f;
        iv[u] = interior[n + u * npi];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) a4[u] += iv[u] * loss_elem(pv[u], tv[u], mv[u], kind);
    }
    for (; n < n1; n += npi) {
      const long long idx = base + (long long)n * F + f;
      const float m = mask ? mask[idx] : 1.0f;
      a4[0] += interior[n] * loss_elem(pred[idx], target[idx], m, kind);
    }
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  red[tid] = acc;
  __syncthreads();
  if (tid < F) {
    float s = 0.f;
    for (int j = 0; j < npi; ++j) s += red[tid + j * F];
    partials[(bt * nchunks + chunk) * F + tid] = s;
  }
}

// form 0: out[bt] = sum_f w_f * S_f / denom ; form 1: out[bt][f] = w_f * sqrt?(S_f/denom)
__global__ void loss_finalize_kernel(const float* __restrict__ partials, int nchunks, int F,
                                     const float* __restrict__ w, const float* __restrict__ denom,
                                     int kind, int form, float* __restrict__ out) {
  __shared__ float sf[LOSS_NT];
  const long long bt = blockIdx.x;
  const int tid = threadIdx.x;
  float s = 0.f;
  if (tid < F) {
    for (int c = 0; c < nchunks; ++c) s += partials[(bt * nchunks + c) * F + tid];
    s /= *denom;
    if (form == 1) {
      if (kind == 0) s = sqrtf(s);
      out[bt * F + tid] = s * w[tid];
    } else {
      sf[tid] = s * w[tid];
    }
  }
  if (form == 0) {
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int f = 0; f < F; ++f) t += sf[f];
      out[bt] = t;
    }
  }
}

// form 2: out[bt][n] = sum_f w_f * l(p,t)
__global__ void loss_map_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                const float* __restrict__ mask, const float* __restrict__ w,
                                long long rows, int F, int kind, float* __restrict__ out) {
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < rows;
       r += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int f = 0; f < F; ++f) {
      const long long idx = r * F + f;
      s += w[f] * loss_elem(pred[idx], target[idx], mask ? mask[idx] : 1.0f, kind);
    }
    out[r] = s;
  }
}

__global__ void loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                const float* __restrict__ mask, const float* __restrict__ interior,
                                const float* __restrict__ w, const float* __restrict__ denom,
                                const float* __restrict__ gout, const float* __restrict__ out,
                                long long bt_total, int nodes, int F, int kind, int form,
                                float* __restrict__ dpred, float* __restrict__ dtarget) {
  const long long per_bt = (long long)nodes * F;
  const long long total = per_bt * bt_total;
  const float inv_denom = 1.0f / *denom;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long bt = i / per_bt;
    const long long rem = i - bt * per_bt;
    const int n = (int)(rem / F), f = (int)(rem - (long long)n * F);
    const float m = mask ? mask[i] : 1.0f;
    const float d = pred[i] * m - target[i] * m;
    float coef;
    if (form == 0) {
      coef = gout[bt] * w[f] * interior[n] * inv_denom;
    } else if (form == 1) {
      if (kind == 0) {
        const float o = out[bt * F + f];  // = w_f * sqrt(S_f/denom)
        coef = o > 0.f ? gout[bt * F + f] * w[f] * w[f] / o * interior[n] * inv_denom * 0.5f : 0.f;
      } else {
        coef = gout[bt * F + f] * w[f] * interior[n] * inv_denom;
      }
    } else {
      coef = gout[bt * (long long)nodes + n] * w[f];
    }
    const float dl = kind == 0 ? 2.0f * d * m : (d > 0.f ? m : (d < 0.f ? -m : 0.f));
    const float gval = coef * dl;
    dpred[i] = gval;
    if (dtarget) dtarget[i] = -gval;
  }
}

// Vectorised form of loss_bwd_kernel: blockIdx.y = (sample, step), a thread handles four consecutive
// elements of the (nodes x F) slab per iteration with 16-byte loads / stores and ONE 32-bit division
// (the scalar kernel pays two 64-bit divisions per element and reaches 2 TB/s).  Needs F >= 4 and
// nodes * F a multiple of 4; same arithmetic per element, so results are bit-identical.
__global__ void loss_bwd_vec_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                    const float* __restrict__ mask, const float* __restrict__ interior,
                                    const float* __restrict__ w, const float* __restrict__ denom,
                                    const float* __restrict__ gout, const float* __restrict__ out,
                                    int nodes, int F, int kind, int form,
                                    float* __restrict__ dpred, float* __restrict__ dtarget) {
  const long long bt = blockIdx.y;
  const unsigned per_bt = (unsigned)nodes * (unsigned)F;
  const long long base = bt * (long long)per_bt;
  const float inv_denom = 1.0f / *denom;
  const float gbt = form == 0 ? gout[bt] : 0.f;
  for (unsigned e = (blockIdx.x * blockDim.x + threadIdx.x) * 4u; e < per_bt; e += gridDim.x * blockDim.x * 4u) {
    unsigned n = e / (unsigned)F, f = e - n * (unsigned)F;
    const float4 p4 = *reinterpret_cast<const float4*>(pred + base + e);
    const float4 t4 = *reinterpret_cast<const float4*>(target + base + e);
    float4 m4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (mask) m4 = *reinterpret_cast<const float4*>(mask + base + e);
    const float pv[4] = {p4.x, p4.y, p4.z, p4.w}, tv[4] = {t4.x, t4.y, t4.z, t4.w}, mv[4] = {m4.x, m4.y, m4.z, m4.w};
    float gv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float m = mv[q];
      const float d = pv[q] * m - tv[q] * m;
      float coef;
      if (form == 0) {
        coef = gbt * w[f] * interior[n] * inv_denom;
      } else if (form == 1) {
        if (kind == 0) {
          const float o = out[bt * F + f];  // = w_f * sqrt(S_f/denom)
          coef = o > 0.f ? gout[bt * F + f] * w[f] * w[f] / o * interior[n] * inv_denom * 0.5f : 0.f;
        } else {
          coef = gout[bt * F + f] * w[f] * interior[n] * inv_denom;
        }
      } else {
        coef = gout[bt * (long long)nodes + n] * w[f];
      }
      const float dl = kind == 0 ? 2.0f * d * m : (d > 0.f ? m : (d < 0.f ? -m : 0.f));
      gv[q] = coef * dl;
      if (++f == (unsigned)F) { f = 0; ++n; }
    }
    *reinterpret_cast<float4*>(dpred + base + e) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    if (dtarget) *reinterpret_cast<float4*>(dtarget + base + e) = make_float4(-gv[0], -gv[1], -gv[2], -gv[3]);
  }
}

// pixels whose mask is zero for every (sample, step, feature)
__global__ void loss_never_kernel(const float* __restrict__ mask, long long bt, int nodes, int F,
                                  int* __restrict__ count) {
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nodes; n += gridDim.x * blockDim.x) {
    bool any = false;
    for (long long q = 0; q < bt && !any; ++q)
      for (int f = 0; f < F; ++f)
        if (mask[(q * nodes + n) * F + f] != 0.f) { any = true; break; }
    if (!any) atomicAdd(count, 1);  // integer atomic: order-independent
  }
}
__global__ void loss_denom_kernel(float num_interior, const int* __restrict__ count,
                                  float* __restrict__ denom) {
  *denom = num_interior - (count ? (float)*count : 0.f);
}

// ------------------------------------------------------------------------------- ACC metric
// Anomaly correlation coefficient of reference py4cast/metrics.py:407-420 (MetricACC.update):
//   acc[bt][f] = mean_s((p - c_f)(t - c_f) m) / sqrt(mean_s(((p - c_f) m)^2) * mean_s(((t - c_f) m)^2))
// = S0 / sqrt(S1 * S2) with the three spatial sums below (the 1/N of the means cancels).  Same
// partial / finalize structure as the loss (fixed summation order, no atomics): ONE pass over
// prediction and target instead of the reference's ~10 temporaries.
// partial[bt][chunk][3][F]
__global__ void __launch_bounds__(LOSS_NT)
acc_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                   const float* __restrict__ mask, const float* __restrict__ clim, int nodes, int F,
                   float* __restrict__ partials) {
  __shared__ float red[3][LOSS_NT];
  const int npi = loss_nodes_per_iter(F);
  const int active = npi * F;
  const int tid = threadIdx.x;
  const int chunk = blockIdx.x, nchunks = gridDim.x;
  const long long bt = blockIdx.y;
  const int n0 = chunk * loss_chunk_nodes(F);
  const int n1 = min(nodes, n0 + loss_chunk_nodes(F));
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  if (tid < active) {
    const int f = tid % F, no = tid / F;
    const float c = clim[f];
    const long long base = bt * (long long)nodes * F;
    for (int n = n0 + no; n < n1; n += npi) {
      const long long idx = base + (long long)n * F + f;
      const float m = mask ? mask[idx] : 1.0f;
      const float a = (pred[idx] - c) * m, b = (target[idx] - c) * m;
      s0 = fmaf((pred[idx] - c) * (target[idx] - c), m, s0);
      s1 = fmaf(a, a, s1);
      s2 = fmaf(b, b, s2);
    }
  }
  red[0][tid] = s0;
  red[1][tid] = s1;
  red[2][tid] = s2;
  __syncthreads();
  if (tid < F) {
    for (int k = 0; k < 3; ++k) {
      float s = 0.f;
      for (int j = 0; j < npi; ++j) s += red[k][tid + j * F];
      partials[((bt * nchunks + chunk) * 3 + k) * F + tid] = s;
    }
  }
}

__global__ void acc_finalize_kernel(const float* __restrict__ partials, int nchunks, int F,
                                    float* __restrict__ out) {
  const long long bt = blockIdx.x;
  const int f = threadIdx.x;
  if (f >= F) return;
  float s[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < nchunks; ++c)
    for (int k = 0; k < 3; ++k) s[k] += partials[((bt * nchunks + c) * 3 + k) * F + f];
  out[bt * F + f] = s[0] / sqrtf(s[1] * s[2]);
}

// --------------------------------------------------------------------- external forcings
// The five generated forcing channels of one sample window (reference py4cast/forcingutils.py:61-132,
// assembled by datasets/base.py:233-274 `generate_forcings` and broadcast over the grid at :509-511):
// [cos_hour, sin_hour, cos_doy, sin_doy] (per time slot, the same for every pixel) and the
// top-of-atmosphere solar irradiation max(0, E0 cos(zenith)) with
//   cos(zenith) = sin(phi) sin(dec) + cos(phi) cos(dec) cos(omega),  omega = 15 (hour + lon/15 - 12) deg,
// written straight into channels ch0.. of the (slots, nodes, ld) forcing tensor (no per-channel
// tensors, no expand + cat).  fp32 throughout, like the reference.
// one association order for both forcing kernels (the compiler is free to contract a*b + c*d*e either way)
__device__ __forceinline__ float toa_radiation(float sin_phi, float cos_phi, float sd, float cd, float omega_rad) {
  const float cos_sza = __fmaf_rn(sin_phi, sd, __fmul_rn(__fmul_rn(cos_phi, cd), cosf(omega_rad)));
  return fmaxf(0.f, 1366.f * cos_sza);
}

__global__ void forcing_fill_kernel(const float* __restrict__ lat, const float* __restrict__ lon,
                                    const float* __restrict__ hour, const float* __restrict__ sin_dec,
                                    const float* __restrict__ cos_dec, const float* __restrict__ dt4,
                                    int slots, int nodes, float* __restrict__ out, int ld, int ch0) {
  const float kRad = 0.017453292519943295f;  // float32(pi / 180), what np.radians multiplies by
  const int t = blockIdx.y;
  const float hr = hour[t], sd = sin_dec[t], cd = cos_dec[t];
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
  if (dt4) { d0 = dt4[4 * t]; d1 = dt4[4 * t + 1]; d2 = dt4[4 * t + 2]; d3 = dt4[4 * t + 3]; }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < nodes; n += gridDim.x * blockDim.x) {
    const float hours_lcl = hr + lon[n] / 15.f;
    const float omega_rad = (15.f * (hours_lcl - 12.f)) * kRad;
    const float phi_rad = lat[n] * kRad;
    float* o = out + ((long long)t * nodes + n) * ld + ch0;
    if (dt4) { o[0] = d0; o[1] = d1; o[2] = d2; o[3] = d3; o += 4; }
    o[0] = toa_radiation(sinf(phi_rad), cosf(phi_rad), sd, cd, omega_rad);
  }
}

// The input path's forcing tensor in ONE pass (datasets/base.py:503-515: NWP forcing channels, then the five
// generated ones): rows of the (slots, nodes, ld) fp32 tensor are written whole and in order -- channels
// [0, n_ship) converted from the shipped array (bf16 under `bf16-true`, else fp32), then cos/sin hour, cos/sin
// day of year, TOA radiation -- instead of a strided channel copy plus forcing_fill_kernel's 20-byte writes into
// 84-byte rows (two partial-sector passes over the tensor).  A CTA owns FA_TN nodes: sin / cos of their
// latitude are computed once and kept in shared memory, then for each of the CTA's time slots the radiation of
// the nodes (same expressions as forcing_fill_kernel: bit-identical) and a linear, coalesced sweep over the
// FA_TN * ld output floats.
constexpr int FA_TN = 256;
template <typename TS>
__global__ void __launch_bounds__(256)
forcing_assemble_kernel(const TS* __restrict__ src, int n_ship, const float* __restrict__ lat,
                        const float* __restrict__ lon, const float* __restrict__ hour,
                        const float* __restrict__ sin_dec, const float* __restrict__ cos_dec,
                        const float* __restrict__ dt4, int slots, int nodes, float* __restrict__ out, int ld,
                        int slots_per_cta) {
  __shared__ float s_sp[FA_TN], s_cp[FA_TN], s_lon[FA_TN], s_toa[FA_TN];
  const float kRad = 0.017453292519943295f;
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * FA_TN;
  const int nn = min(FA_TN, nodes - n0);
  if (tid < nn) {
    const float phi_rad = lat[n0 + tid] * kRad;
    s_sp[tid] = sinf(phi_rad);
    s_cp[tid] = cosf(phi_rad);
    s_lon[tid] = lon[n0 + tid];
  }
  const int t_lo = blockIdx.y * slots_per_cta, t_hi = min(slots, t_lo + slots_per_cta);
  const int total = nn * ld;
  const int dn = 256 / ld, dc = 256 - dn * ld;  // i += 256  ->  (n += dn, c += dc) with one carry
  for (int t = t_lo; t < t_hi; ++t) {
    const float hr = hour[t], sd = sin_dec[t], cd = cos_dec[t];
    const float d0 = dt4[4 * t], d1 = dt4[4 * t + 1], d2 = dt4[4 * t + 2], d3 = dt4[4 * t + 3];
    __syncthreads();  // s_sp.. written (first slot); s_toa of the previous slot consumed
    if (tid < nn) {
      const float hours_lcl = hr + s_lon[tid] / 15.f;
      const float omega_rad = (15.f * (hours_lcl - 12.f)) * kRad;
      s_toa[tid] = toa_radiation(s_sp[tid], s_cp[tid], sd, cd, omega_rad);
    }
    __syncthreads();
    const long long row0 = (long long)t * nodes + n0;
    float* o = out + row0 * ld;
    const TS* sr = src + row0 * n_ship;
    int n = tid / ld, c = tid - n * ld;
    for (int i = tid; i < total; i += 256) {
      float v;
      const int e = c - n_ship;
      if (e < 0) v = to_f32<TS>(sr[n * n_ship + c]);
      else if (e == 0) v = d0;
      else if (e == 1) v = d1;
      else if (e == 2) v = d2;
      else if (e == 3) v = d3;
      else if (e == 4) v = s_toa[n];
      else v = 0.f;
      o[i] = v;
      n += dn;
      c += dc;
      if (c >= ld) { c -= ld; ++n; }
    }
  }
}

// Staged variant (nodes a multiple of FA_TN, 16-byte aligned shipped array, two slot buffers within 48 KB of
// shared memory): the shipped channels of the CTA's nodes for one slot are ONE contiguous span, fetched with
// 16-byte cp.async one slot ahead of the sweep that writes the rows (the scalar kernel above keeps a single
// 2-byte load per thread in flight: 0.20 of the HBM peak).
template <typename TS>
__global__ void __launch_bounds__(256)
forcing_assemble_staged_kernel(const TS* __restrict__ src, int n_ship, const float* __restrict__ lat,
                               const float* __restrict__ lon, const float* __restrict__ hour,
                               const float* __restrict__ sin_dec, const float* __restrict__ cos_dec,
                               const float* __restrict__ dt4, int slots, int nodes, float* __restrict__ out, int ld,
                               int slots_per_cta) {
  extern __shared__ __align__(16) unsigned char fa_raw[];  // 2 x [FA_TN][n_ship] of TS
  __shared__ float s_sp[FA_TN], s_cp[FA_TN], s_lon[FA_TN], s_toa[FA_TN];
  const float kRad = 0.017453292519943295f;
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * FA_TN;
  {
    const float phi_rad = lat[n0 + tid] * kRad;
    s_sp[tid] = sinf(phi_rad);
    s_cp[tid] = cosf(phi_rad);
    s_lon[tid] = lon[n0 + tid];
  }
  const int t_lo = blockIdx.y * slots_per_cta, t_hi = min(slots, t_lo + slots_per_cta);
  const int span_bytes = FA_TN * n_ship * (int)sizeof(TS);  // a multiple of 16
  const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(fa_raw);
  auto fetch = [&](int t, int buf) {
    const char* g = reinterpret_cast<const char*>(src + ((long long)t * nodes + n0) * n_ship);
    for (int i = tid; i < span_bytes / 16; i += 256) ar_cp16(b0 + buf * span_bytes + 16 * i, g + 16 * i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (t_lo < t_hi) fetch(t_lo, 0);
  const int total = FA_TN * ld;
  const int dn = 256 / ld, dc = 256 - dn * ld;
  for (int t = t_lo; t < t_hi; ++t) {
    const int buf = (t - t_lo) & 1;
    const float hr = hour[t], sd = sin_dec[t], cd = cos_dec[t];
    const float d0 = dt4[4 * t], d1 = dt4[4 * t + 1], d2 = dt4[4 * t + 2], d3 = dt4[4 * t + 3];
    __syncthreads();  // the sweep of slot t - 1 (buffer buf ^ 1, s_toa) is over
    if (t + 1 < t_hi) {
      fetch(t + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    {
      const float hours_lcl = hr + s_lon[tid] / 15.f;
      const float omega_rad = (15.f * (hours_lcl - 12.f)) * kRad;
      s_toa[tid] = toa_radiation(s_sp[tid], s_cp[tid], sd, cd, omega_rad);
    }
    __syncthreads();
    const TS* sr = reinterpret_cast<const TS*>(fa_raw + buf * span_bytes);
    float* o = out + ((long long)t * nodes + n0) * ld;
    int n = tid / ld, c = tid - n * ld;
#pragma unroll 4
    for (int i = tid; i < total; i += 256) {
      float v;
      const int e = c - n_ship;
      if (e < 0) v = to_f32<TS>(sr[n * n_ship + c]);
      else if (e == 4) v = s_toa[n];
      else if (e == 0) v = d0;
      else if (e == 1) v = d1;
      else if (e == 2) v = d2;
      else if (e == 3) v = d3;
      else v = 0.f;
      __stcs(o + i, v);
      n += dn;
      c += dc;
      if (c >= ld) { c -= ld; ++n; }
    }
  }
}

// Row-block variant of the staged kernel (additionally: n_ship a multiple of one 16-byte vector of TS, ld a
// multiple of 4 floats per FA_TN rows -- always -- and FA_TN * ld floats of shared memory): the fp32 rows of the
// CTA's nodes are BUILT in shared memory (16-byte reads of the staged span, one conversion per element, the five
// generated values per node) and leave as 16-byte vectors.  The staged kernel above decides per output element
// which of seven sources it comes from (~30 instructions per element: issue-bound at 0.40 of the HBM peak).
template <typename TS>
__global__ void __launch_bounds__(256)
forcing_assemble_rows_kernel(const TS* __restrict__ src, int n_ship, const float* __restrict__ lat,
                             const float* __restrict__ lon, const float* __restrict__ hour,
                             const float* __restrict__ sin_dec, const float* __restrict__ cos_dec,
                             const float* __restrict__ dt4, int slots, int nodes, float* __restrict__ out, int ld,
                             int slots_per_cta) {
  constexpr int VEC = 16 / (int)sizeof(TS);
  extern __shared__ __align__(16) unsigned char fa_raw[];  // 2 x [FA_TN][n_ship] of TS | [FA_TN][ld] fp32
  __shared__ float s_sp[FA_TN], s_cp[FA_TN], s_lon[FA_TN];
  const float kRad = 0.017453292519943295f;
  const int tid = threadIdx.x;
  const int n0 = blockIdx.x * FA_TN;
  {
    const float phi_rad = lat[n0 + tid] * kRad;
    s_sp[tid] = sinf(phi_rad);
    s_cp[tid] = cosf(phi_rad);
    s_lon[tid] = lon[n0 + tid];
  }
  const int t_lo = blockIdx.y * slots_per_cta, t_hi = min(slots, t_lo + slots_per_cta);
  const int span_bytes = FA_TN * n_ship * (int)sizeof(TS);
  float* rows = reinterpret_cast<float*>(fa_raw + 2 * span_bytes);
  const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(fa_raw);
  auto fetch = [&](int t, int buf) {
    const char* g = reinterpret_cast<const char*>(src + ((long long)t * nodes + n0) * n_ship);
    for (int i = tid; i < span_bytes / 16; i += 256) ar_cp16(b0 + buf * span_bytes + 16 * i, g + 16 * i);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (t_lo < t_hi) fetch(t_lo, 0);
  const int vpr = n_ship / VEC;  // 16-byte vectors per row of the span
  const int nvec = FA_TN * vpr;
  const int nout4 = FA_TN * ld / 4;
  for (int t = t_lo; t < t_hi; ++t) {
    const int buf = (t - t_lo) & 1;
    const float hr = hour[t], sd = sin_dec[t], cd = cos_dec[t];
    const float d0 = dt4[4 * t], d1 = dt4[4 * t + 1], d2 = dt4[4 * t + 2], d3 = dt4[4 * t + 3];
    __syncthreads();  // the rows of slot t - 1 have left; buffer buf ^ 1 was consumed
    if (t + 1 < t_hi) {
      fetch(t + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    {  // the five generated values of node tid
      const float hours_lcl = hr + s_lon[tid] / 15.f;
      const float omega_rad = (15.f * (hours_lcl - 12.f)) * kRad;
      float* r = rows + tid * ld + n_ship;
      r[0] = d0; r[1] = d1; r[2] = d2; r[3] = d3;
      r[4] = toa_radiation(s_sp[tid], s_cp[tid], sd, cd, omega_rad);
      for (int c = n_ship + 5; c < ld; ++c) rows[tid * ld + c] = 0.f;
    }
    __syncthreads();  // everybody's part of the span has landed
    const uint4* sv = reinterpret_cast<const uint4*>(fa_raw + buf * span_bytes);
    int n = tid / vpr, q = tid - n * vpr;
    const int dn = 256 / vpr, dq = 256 - dn * vpr;
    for (int j = tid; j < nvec; j += 256) {
      const uint4 w = sv[j];
      float* r = rows + n * ld + q * VEC;
      if constexpr (sizeof(TS) == 2) {
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          r[2 * k] = __uint_as_float(ww[k] << 16);
          r[2 * k + 1] = __uint_as_float(ww[k] & 0xFFFF0000u);
        }
      } else {
        r[0] = __uint_as_float(w.x); r[1] = __uint_as_float(w.y); r[2] = __uint_as_float(w.z); r[3] = __uint_as_float(w.w);
      }
      n += dn;
      q += dq;
      if (q >= vpr) { q -= vpr; ++n; }
    }
    __syncthreads();
    float4* o4 = reinterpret_cast<float4*>(out + ((long long)t * nodes + n0) * ld);
    const float4* r4 = reinterpret_cast<const float4*>(rows);
    for (int i = tid; i < nout4; i += 256) __stcs(o4 + i, r4[i]);
  }
}

// ------------------------------------------------------------------------------- AdamW
__global__ void adamw_advance_kernel(p4c_adamw_state_t* st, float base_lr, float min_lr,
                                     long long warmup, long long total_steps, float beta1,
                                     float beta2) {
  const long long cur = st->step;  // scheduler step index for this optimiser step
  double lam;
  if (cur < warmup) {
    lam = (double)cur / (double)(warmup > 1 ? warmup : 1);
  } else {
    const long long span = total_steps - warmup;
    const double progress = (double)(cur - warmup) / (double)(span > 1 ? span : 1);
    double factor = 0.5 * (1.0 + cos(3.14159265358979323846 * progress));
    const double rate = (double)min_lr / (double)base_lr;
    factor = factor * (1.0 - rate) + rate;
    lam = factor > 0.0 ? factor : 0.0;
  }
  const long long t = cur + 1;
  st->step = t;
  st->lr = (float)((double)base_lr * lam);
  st->bias1 = (float)(1.0 - pow((double)beta1, (double)t));
  st->bias2 = (float)(1.0 - pow((double)beta2, (double)t));
}

__global__ void __launch_bounds__(256)
adamw_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, long long n, const p4c_adamw_state_t* __restrict__ st,
                  float beta1, float beta2, float eps, float wd, float grad_scale) {
  const float lr = st->lr, bias1 = st->bias1, bias2 = st->bias2;
  const float step_size = lr / bias1;
  const float inv_sqrt_b2 = 1.0f / sqrtf(bias2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    const float den = sqrtf(vi) * inv_sqrt_b2 + eps;
    pi -= step_size * (mi / den);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}

static int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)kNumSMs * 16;
  if (g < 1) g = 1;
  return (int)(g < cap ? g : cap);
}

// min_prev = 0 for the plain concatenation: the downscaling strategy feeds statics and forcings only
// (reference lightning.py:759-765, `inputs * (1 - ds)`)
static int check_ar_x(const p4c_ar_x_t* a, bool last_may_be_null, int min_prev = 1) {
  P4C_REQUIRE(a != nullptr, "ar: null descriptor");
  P4C_REQUIRE(a->batch > 0 && a->nodes > 0 && a->f_state > 0, "ar: empty problem");
  P4C_REQUIRE(a->num_prev >= min_prev && a->num_prev <= P4C_MAX_PREV, "ar: num_prev=%d not in [%d,%d]",
              a->num_prev, min_prev, P4C_MAX_PREV);
  P4C_REQUIRE(a->num_prev * a->f_state + a->f_static + a->f_forcing > 0, "ar: x has no columns");
  for (int t = 0; t < a->num_prev; ++t)
    P4C_REQUIRE(a->prev[t] != nullptr || (last_may_be_null && t == a->num_prev - 1),
                "ar: prev[%d] is null", t);
  P4C_REQUIRE(a->f_static == 0 || a->stat, "ar: static features missing");
  P4C_REQUIRE(a->f_forcing == 0 || a->forcing, "ar: forcing missing");
  P4C_REQUIRE(a->x_ld == 0 || a->x_ld >= a->num_prev * a->f_state + a->f_static + a->f_forcing,
              "ar: x_ld=%d smaller than the row width", a->x_ld);
  return P4C_OK;
}

}  // namespace p4c

using namespace p4c;

extern "C" void p4c_set_ar_vector_paths(int enabled) { p4c::g_ar_vector_paths = enabled ? 1 : 0; }

extern "C" int p4c_ar_concat(const p4c_ar_x_t* a, void* x, int x_dtype, void* stream) {
  int rc = check_ar_x(a, false, 0);
  if (rc) return rc;
  P4C_REQUIRE(x != nullptr, "ar_concat: x is null");
  const int f_in = a->num_prev * a->f_state + a->f_static + a->f_forcing;
  const long long total = (long long)a->batch * a->nodes * (a->x_ld > 0 ? a->x_ld : f_in);
  cudaStream_t st = (cudaStream_t)stream;
  P4C_REQUIRE(x_dtype == P4C_F32 || x_dtype == P4C_BF16, "ar_concat: bad dtype %d", x_dtype);
  const int ld = a->x_ld > 0 ? a->x_ld : f_in;
  const size_t tile_smem = (size_t)AR_TN * ld * sizeof(float);
  size_t stage_smem = 0;
  if (ar_stage_ok(a, x, x_dtype, nullptr, 0, &stage_smem)) {
    const dim3 grid(a->nodes / AR_TN, a->batch);
    ArStep none{};
    if (x_dtype == P4C_F32) ar_stage_kernel<float, float><<<grid, 256, stage_smem, st>>>(none, *a, (float*)x, 0);
    else ar_stage_kernel<float, __nv_bfloat16><<<grid, 256, stage_smem, st>>>(none, *a, (__nv_bfloat16*)x, 0);
    P4C_CHECK_LAUNCH("ar_concat(staged)");
    return P4C_OK;
  }
  if (tile_smem <= 48 * 1024 && a->batch <= 65535) {
    const dim3 grid((a->nodes + AR_TN - 1) / AR_TN, a->batch);
    ArStep none{};
    if (x_dtype == P4C_F32) ar_tile_kernel<float, float><<<grid, 256, tile_smem, st>>>(none, *a, (float*)x, 0);
    else ar_tile_kernel<float, __nv_bfloat16><<<grid, 256, tile_smem, st>>>(none, *a, (__nv_bfloat16*)x, 0);
    P4C_CHECK_LAUNCH("ar_concat(tile)");
    return P4C_OK;
  }
  if (x_dtype == P4C_F32)
    ar_concat_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(*a, (float*)x);
  else
    ar_concat_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(*a, (__nv_bfloat16*)x);
  P4C_CHECK_LAUNCH("ar_concat");
  return P4C_OK;
}

extern "C" int p4c_ar_step_forward(const void* y, int y_dtype, const float* prev_last,
                                   int64_t prev_last_bs, const float* truth, int64_t truth_bs,
                                   const float* border, const float* diff_std,
                                   const float* diff_mean, int32_t batch, int32_t nodes,
                                   int32_t f_state, float* new_state, int64_t new_state_bs,
                                   const p4c_ar_x_t* next, void* x_next, int x_dtype,
                                   void* stream) {
  P4C_REQUIRE(y && prev_last && new_state, "ar_step_forward: null pointer");
  P4C_REQUIRE(batch > 0 && nodes > 0 && f_state > 0, "ar_step_forward: empty problem");
  P4C_REQUIRE((border == nullptr) == (truth == nullptr), "ar_step_forward: border and truth go together");
  P4C_REQUIRE((diff_std == nullptr) == (diff_mean == nullptr), "ar_step_forward: std and mean go together");
  P4C_REQUIRE(y_dtype == P4C_F32 || y_dtype == P4C_BF16, "ar_step_forward: bad y dtype");
  ArStep s{y, prev_last, prev_last_bs, truth, truth_bs, border, diff_std, diff_mean,
           batch, nodes, f_state, new_state, new_state_bs};
  cudaStream_t st = (cudaStream_t)stream;
  if (!next) {
    const long long total = (long long)batch * nodes * f_state;
    if (y_dtype == P4C_F32) ar_update_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(s);
    else ar_update_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(s);
    P4C_CHECK_LAUNCH("ar_update");
    return P4C_OK;
  }
  int rc = check_ar_x(next, true);
  if (rc) return rc;
  P4C_REQUIRE(x_next != nullptr, "ar_step_forward: x_next is null");
  P4C_REQUIRE(next->batch == batch && next->nodes == nodes && next->f_state == f_state,
              "ar_step_forward: next-x descriptor does not match the state shape");
  P4C_REQUIRE(x_dtype == P4C_F32 || x_dtype == P4C_BF16, "ar_step_forward: bad x dtype");
  const int f_in = next->num_prev * f_state + next->f_static + next->f_forcing;
  const long long total = (long long)batch * nodes * (next->x_ld > 0 ? next->x_ld : f_in);
  const int grid = grid_for(total, 256);
  const int ld = next->x_ld > 0 ? next->x_ld : f_in;
  const size_t tile_smem = (size_t)AR_TN * ld * sizeof(float);
  size_t stage_smem = 0;
  if (ar_stage_ok(next, x_next, x_dtype, &s, y_dtype, &stage_smem)) {
    const dim3 tg(nodes / AR_TN, batch);
    if (y_dtype == P4C_F32 && x_dtype == P4C_F32)
      ar_stage_kernel<float, float><<<tg, 256, stage_smem, st>>>(s, *next, (float*)x_next, 1);
    else if (y_dtype == P4C_F32 && x_dtype == P4C_BF16)
      ar_stage_kernel<float, __nv_bfloat16><<<tg, 256, stage_smem, st>>>(s, *next, (__nv_bfloat16*)x_next, 1);
    else if (y_dtype == P4C_BF16 && x_dtype == P4C_F32)
      ar_stage_kernel<__nv_bfloat16, float><<<tg, 256, stage_smem, st>>>(s, *next, (float*)x_next, 1);
    else
      ar_stage_kernel<__nv_bfloat16, __nv_bfloat16><<<tg, 256, stage_smem, st>>>(s, *next, (__nv_bfloat16*)x_next, 1);
    P4C_CHECK_LAUNCH("ar_step(staged)");
    return P4C_OK;
  }
  if (tile_smem <= 48 * 1024 && batch <= 65535) {
    const dim3 tg((nodes + AR_TN - 1) / AR_TN, batch);
    if (y_dtype == P4C_F32 && x_dtype == P4C_F32)
      ar_tile_kernel<float, float><<<tg, 256, tile_smem, st>>>(s, *next, (float*)x_next, 1);
    else if (y_dtype == P4C_F32 && x_dtype == P4C_BF16)
      ar_tile_kernel<float, __nv_bfloat16><<<tg, 256, tile_smem, st>>>(s, *next, (__nv_bfloat16*)x_next, 1);
    else if (y_dtype == P4C_BF16 && x_dtype == P4C_F32)
      ar_tile_kernel<__nv_bfloat16, float><<<tg, 256, tile_smem, st>>>(s, *next, (float*)x_next, 1);
    else
      ar_tile_kernel<__nv_bfloat16, __nv_bfloat16><<<tg, 256, tile_smem, st>>>(s, *next, (__nv_bfloat16*)x_next, 1);
    P4C_CHECK_LAUNCH("ar_step(tile)");
    return P4C_OK;
  }
  if (y_dtype == P4C_F32 && x_dtype == P4C_F32)
    ar_step_kernel<float, float><<<grid, 256, 0, st>>>(s, *next, (float*)x_next);
  else if (y_dtype == P4C_F32 && x_dtype == P4C_BF16)
    ar_step_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(s, *next, (__nv_bfloat16*)x_next);
  else if (y_dtype == P4C_BF16 && x_dtype == P4C_F32)
    ar_step_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(s, *next, (float*)x_next);
  else
    ar_step_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(s, *next, (__nv_bfloat16*)x_next);
  P4C_CHECK_LAUNCH("ar_step");
  return P4C_OK;
}

extern "C" int p4c_ar_step_backward(const float* g_new, int64_t g_new_bs, const float* g_new2, int64_t g_new2_bs,
                                    const void* g_xnext, int x_dtype, const float* border, const float* diff_std,
                                    int32_t batch, int32_t nodes, int32_t f_state,
                                    int32_t num_prev, int32_t f_in, void* dy, int y_dtype,
                                    float* const* d_prev, void* stream) {
  P4C_REQUIRE(dy != nullptr, "ar_step_backward: dy is null");
  P4C_REQUIRE(g_new || g_new2 || g_xnext, "ar_step_backward: no incoming gradient");
  P4C_REQUIRE(batch > 0 && nodes > 0 && f_state > 0, "ar_step_backward: empty problem");
  P4C_REQUIRE(num_prev >= 1 && num_prev <= P4C_MAX_PREV, "ar_step_backward: bad num_prev %d", num_prev);
  P4C_REQUIRE(!g_xnext || f_in >= num_prev * f_state, "ar_step_backward: f_in too small");
  ArBwd s{};
  s.g_new = g_new; s.g_new_bs = g_new_bs; s.g_new2 = g_new2; s.g_new2_bs = g_new2_bs; s.g_x = g_xnext; s.border = border; s.std_ = diff_std;
  s.batch = batch; s.nodes = nodes; s.F = f_state; s.num_prev = num_prev; s.f_in = f_in; s.dy = dy;
  for (int t = 0; t < P4C_MAX_PREV; ++t) s.d_prev[t] = (d_prev && t < num_prev) ? d_prev[t] : nullptr;
  const long long total = (long long)batch * nodes * f_state;
  const int grid = grid_for(total, 256);
  cudaStream_t st = (cudaStream_t)stream;
  const bool xb = g_xnext && x_dtype == P4C_BF16;
  {
    const long long per_b = (long long)nodes * f_state;
    bool vec = g_ar_vector_paths && per_b % 4 == 0 && per_b < (1ll << 31) && batch <= 65535 && ar_al16(dy) && (!g_new || (ar_al16(g_new) && g_new_bs % 4 == 0)) &&
               (!g_new2 || (ar_al16(g_new2) && g_new2_bs % 4 == 0)) && (y_dtype == P4C_F32 || y_dtype == P4C_BF16);
    for (int t = 0; t < num_prev; ++t) vec = vec && (!s.d_prev[t] || ar_al16(s.d_prev[t]));
    if (vec) {
      const unsigned long long magic = (~0ull / (unsigned long long)f_state) + 1ull;  // ceil(2^64 / F): exact for 32-bit dividends
      const dim3 vg((unsigned)((per_b / 4 + 255) / 256), (unsigned)batch);
      if (y_dtype == P4C_F32 && !xb) ar_step_bwd_vec_kernel<float, float><<<vg, 256, 0, st>>>(s, magic);
      else if (y_dtype == P4C_F32 && xb) ar_step_bwd_vec_kernel<float, __nv_bfloat16><<<vg, 256, 0, st>>>(s, magic);
      else if (y_dtype == P4C_BF16 && !xb) ar_step_bwd_vec_kernel<__nv_bfloat16, float><<<vg, 256, 0, st>>>(s, magic);
      else ar_step_bwd_vec_kernel<__nv_bfloat16, __nv_bfloat16><<<vg, 256, 0, st>>>(s, magic);
      P4C_CHECK_LAUNCH("ar_step_bwd(vec)");
      return P4C_OK;
    }
  }
  if (y_dtype == P4C_F32 && !xb) ar_step_bwd_kernel<float, float><<<grid, 256, 0, st>>>(s);
  else if (y_dtype == P4C_F32 && xb) ar_step_bwd_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(s);
  else if (y_dtype == P4C_BF16 && !xb) ar_step_bwd_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(s);
  else if (y_dtype == P4C_BF16 && xb) ar_step_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(s);
  else P4C_REQUIRE(false, "ar_step_backward: bad dtype");
  P4C_CHECK_LAUNCH("ar_step_bwd");
  return P4C_OK;
}

extern "C" int p4c_loss_denominator(const float* mask, int64_t bt, int32_t nodes, int32_t f,
                                    float num_interior, float* denom_out, int32_t* scratch,
                                    void* stream) {
  P4C_REQUIRE(denom_out != nullptr, "loss_denominator: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (mask) {
    P4C_REQUIRE(scratch != nullptr && bt > 0 && nodes > 0 && f > 0, "loss_denominator: bad arguments");
    cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(int), st);
    if (e != cudaSuccess) {
      set_error("loss_denominator: memset failed: %s", cudaGetErrorString(e));
      return P4C_ECUDA;
    }
    loss_never_kernel<<<grid_for(nodes, 128), 128, 0, st>>>(mask, bt, nodes, f, scratch);
    P4C_CHECK_LAUNCH("loss_never");
  }
  loss_denom_kernel<<<1, 1, 0, st>>>(num_interior, mask ? scratch : nullptr, denom_out);
  P4C_CHECK_LAUNCH("loss_denom");
  return P4C_OK;
}

static int loss_nchunks(int nodes, int F) {
  const int cn = loss_chunk_nodes(F);
  return (nodes + cn - 1) / cn;
}

extern "C" int64_t p4c_loss_partial_floats(int64_t bt, int32_t nodes, int32_t f, int form) {
  if (form == 2 || f <= 0 || f > LOSS_NT) return 0;
  return bt * (int64_t)loss_nchunks(nodes, f) * f;
}

extern "C" int p4c_loss_forward(const float* pred, const float* target, const float* mask,
                                const float* interior, const float* w, const float* denom,
                                int64_t bt, int32_t nodes, int32_t f, int kind, int form,
                                float* partials, float* out, void* stream) {
  P4C_REQUIRE(pred && target && w && out, "loss_forward: null pointer");
  P4C_REQUIRE(bt > 0 && nodes > 0 && f > 0 && f <= LOSS_NT, "loss_forward: bad sizes (f must be <= 256)");
  P4C_REQUIRE(kind == 0 || kind == 1, "loss_forward: kind must be 0 (MSE) or 1 (L1)");
  P4C_REQUIRE(form >= 0 && form <= 2, "loss_forward: bad form %d", form);
  P4C_REQUIRE(bt <= 65535, "loss_forward: batch*steps=%lld exceeds 65535", (long long)bt);
  cudaStream_t st = (cudaStream_t)stream;
  if (form == 2) {
    const long long rows = bt * nodes;
    loss_map_kernel<<<grid_for(rows, 256), 256, 0, st>>>(pred, target, mask, w, rows, f, kind, out);
    P4C_CHECK_LAUNCH("loss_map");
    return P4C_OK;
  }
  P4C_REQUIRE(interior && denom && partials, "loss_forward: interior / denom / partials missing");
  const int nchunks = loss_nchunks(nodes, f);
  dim3 grid(nchunks, (unsigned)bt);
  loss_partial_kernel<<<grid, LOSS_NT, 0, st>>>(pred, target, mask, interior, nodes, f, kind, partials);
  P4C_CHECK_LAUNCH("loss_partial");
  loss_finalize_kernel<<<(unsigned)bt, LOSS_NT, 0, st>>>(partials, nchunks, f, w, denom, kind, form, out);
  P4C_CHECK_LAUNCH("loss_finalize");
  return P4C_OK;
}

extern "C" int p4c_loss_backward(const float* pred, const float* target, const float* mask,
                                 const float* interior, const float* w, const float* denom,
                                 const float* gout, const float* out, int64_t bt, int32_t nodes,
                                 int32_t f, int kind, int form, float* dpred, float* dtarget,
                                 void* stream) {
  P4C_REQUIRE(pred && target && w && gout && dpred, "loss_backward: null pointer");
  P4C_REQUIRE(bt > 0 && nodes > 0 && f > 0, "loss_backward: bad sizes");
  P4C_REQUIRE(form == 2 || (interior && denom), "loss_backward: interior / denom missing");
  P4C_REQUIRE(!(form == 1 && kind == 0) || out, "loss_backward: forward output needed for ScaledLoss(MSE)");
  const long long total = bt * (long long)nodes * f;
  const float* dn = denom;
  P4C_REQUIRE(dn != nullptr || form == 2, "loss_backward: denom missing");
  if (!dn) dn = w;  // form 2 never uses inv_denom's value; any valid address
  const long long per_bt = (long long)nodes * f;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  if (f >= 4 && per_bt % 4 == 0 && per_bt < (1ll << 31) && bt <= 65535 && al16(pred) && al16(target) && al16(mask) &&
      al16(dpred) && al16(dtarget)) {
    long long bx = (per_bt / 4 + 255) / 256;
    const long long cap = (8ll * kNumSMs + bt - 1) / bt;  // ~8 CTAs per SM over the whole grid
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)bx, (unsigned)bt);
    loss_bwd_vec_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pred, target, mask, interior, w, dn, gout, out,
                                                                nodes, f, kind, form, dpred, dtarget);
  } else {
    loss_bwd_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        pred, target, mask, interior, w, dn, gout, out, bt, nodes, f, kind, form, dpred, dtarget);
  }
  P4C_CHECK_LAUNCH("loss_bwd");
  return P4C_OK;
}

extern "C" int p4c_adamw_advance(p4c_adamw_state_t* state, float base_lr, float min_lr,
                                 int64_t warmup_steps, int64_t total_steps, float beta1,
                                 float beta2, void* stream) {
  P4C_REQUIRE(state != nullptr && base_lr > 0.f, "adamw_advance: bad arguments");
  adamw_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state, base_lr, min_lr, warmup_steps,
                                                          total_steps, beta1, beta2);
  P4C_CHECK_LAUNCH("adamw_advance");
  return P4C_OK;
}

extern "C" int p4c_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                              int64_t n, const p4c_adamw_state_t* state, float beta1, float beta2,
                              float eps, float weight_decay, float grad_scale, void* stream) {
  P4C_REQUIRE(param && grad && exp_avg && exp_avg_sq && state && n > 0, "adamw_step: bad arguments");
  adamw_step_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, state, beta1, beta2, eps, weight_decay, grad_scale);
  P4C_CHECK_LAUNCH("adamw_step");
  return P4C_OK;
}

extern "C" int p4c_forcing_assemble(const void* src, int src_dtype, int32_t n_ship, const float* lat, const float* lon,
                                    const float* hour, const float* sin_dec, const float* cos_dec,
                                    const float* datetime4, int32_t slots, int32_t nodes, float* out, int32_t ld,
                                    void* stream) {
  P4C_REQUIRE(lat && lon && hour && sin_dec && cos_dec && datetime4 && out, "forcing_assemble: null pointer");
  P4C_REQUIRE(n_ship >= 0 && (n_ship == 0 || src), "forcing_assemble: shipped channels missing");
  P4C_REQUIRE(slots > 0 && nodes > 0, "forcing_assemble: bad sizes (slots=%d nodes=%d)", slots, nodes);
  P4C_REQUIRE(ld >= n_ship + 5 && ld <= 256, "forcing_assemble: ld=%d does not hold %d + 5 channels (or > 256)", ld,
              n_ship);
  P4C_REQUIRE(src_dtype == P4C_F32 || src_dtype == P4C_BF16, "forcing_assemble: bad dtype %d", src_dtype);
  const int bx = (nodes + FA_TN - 1) / FA_TN;
  // ~32 CTAs per SM over the launch (several waves of the 8 resident ones: a small tail); a CTA keeps its nodes'
  // sin / cos latitude over its slots
  int by = (32 * kNumSMs + bx - 1) / bx;
  if (by > slots) by = slots;
  if (by > 65535) by = 65535;
  if (by < 1) by = 1;
  const int per = (slots + by - 1) / by;
  by = (slots + per - 1) / per;
  const dim3 grid(bx, by);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t esize = src_dtype == P4C_BF16 ? 2 : 4;
  const size_t stage = 2 * (size_t)FA_TN * n_ship * esize;
  if (g_ar_vector_paths && n_ship > 0 && nodes % FA_TN == 0 && stage <= 44 * 1024 &&
      (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int vec = (int)(16 / esize);
    const size_t rows_smem = stage + (size_t)FA_TN * ld * sizeof(float);
    if (n_ship % vec == 0 && rows_smem <= 44 * 1024 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      if (src_dtype == P4C_BF16)
        forcing_assemble_rows_kernel<__nv_bfloat16><<<grid, 256, rows_smem, st>>>(
            (const __nv_bfloat16*)src, n_ship, lat, lon, hour, sin_dec, cos_dec, datetime4, slots, nodes, out, ld, per);
      else
        forcing_assemble_rows_kernel<float><<<grid, 256, rows_smem, st>>>((const float*)src, n_ship, lat, lon, hour, sin_dec,
                                                                         cos_dec, datetime4, slots, nodes, out, ld, per);
      P4C_CHECK_LAUNCH("forcing_assemble(rows)");
      return P4C_OK;
    }
    if (src_dtype == P4C_BF16)
      forcing_assemble_staged_kernel<__nv_bfloat16><<<grid, 256, stage, st>>>(
          (const __nv_bfloat16*)src, n_ship, lat, lon, hour, sin_dec, cos_dec, datetime4, slots, nodes, out, ld, per);
    else
      forcing_assemble_staged_kernel<float><<<grid, 256, stage, st>>>((const float*)src, n_ship, lat, lon, hour, sin_dec,
                                                                     cos_dec, datetime4, slots, nodes, out, ld, per);
    P4C_CHECK_LAUNCH("forcing_assemble(staged)");
    return P4C_OK;
  }
  if (src_dtype == P4C_BF16)
    forcing_assemble_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, n_ship, lat, lon, hour, sin_dec,
                                                                cos_dec, datetime4, slots, nodes, out, ld, per);
  else
    forcing_assemble_kernel<float><<<grid, 256, 0, st>>>((const float*)src, n_ship, lat, lon, hour, sin_dec, cos_dec,
                                                        datetime4, slots, nodes, out, ld, per);
  P4C_CHECK_LAUNCH("forcing_assemble");
  return P4C_OK;
}

extern "C" int p4c_forcing_fill(const float* lat, const float* lon, const float* hour, const float* sin_dec,
                                const float* cos_dec, const float* datetime4, int32_t slots, int32_t nodes,
                                float* out, int32_t ld, int32_t ch0, void* stream) {
  P4C_REQUIRE(lat && lon && hour && sin_dec && cos_dec && out, "forcing_fill: null pointer");
  P4C_REQUIRE(slots > 0 && slots <= 65535 && nodes > 0, "forcing_fill: bad sizes (slots=%d nodes=%d)", slots, nodes);
  P4C_REQUIRE(ch0 >= 0 && ch0 + (datetime4 ? 5 : 1) <= ld, "forcing_fill: channels %d.. do not fit ld=%d", ch0, ld);
  int bx = (nodes + 255) / 256;
  const int cap = (16 * kNumSMs + slots - 1) / slots;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  forcing_fill_kernel<<<dim3(bx, slots), 256, 0, (cudaStream_t)stream>>>(lat, lon, hour, sin_dec, cos_dec, datetime4,
                                                                         slots, nodes, out, ld, ch0);
  P4C_CHECK_LAUNCH("forcing_fill");
  return P4C_OK;
}


extern "C" int64_t p4c_acc_partial_floats(int64_t bt, int32_t nodes, int32_t f) {
  if (f <= 0 || f > LOSS_NT || bt <= 0 || nodes <= 0) return 0;
  return bt * (int64_t)loss_nchunks(nodes, f) * 3 * f;
}

extern "C" int p4c_acc_forward(const float* pred, const float* target, const float* mask,
                               const float* climate_means, int64_t bt, int32_t nodes, int32_t f,
                               float* partials, float* out, void* stream) {
  P4C_REQUIRE(pred && target && climate_means && partials && out, "acc_forward: null pointer");
  P4C_REQUIRE(bt > 0 && bt <= 65535 && nodes > 0 && f > 0 && f <= LOSS_NT, "acc_forward: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunks = loss_nchunks(nodes, f);
  acc_partial_kernel<<<dim3(nchunks, (unsigned)bt), LOSS_NT, 0, st>>>(pred, target, mask, climate_means, nodes, f,
                                                                      partials);
  P4C_CHECK_LAUNCH("acc_partial");
  acc_finalize_kernel<<<(unsigned)bt, LOSS_NT, 0, st>>>(partials, nchunks, f, out);
  P4C_CHECK_LAUNCH("acc_finalize");
  return P4C_OK;
}
