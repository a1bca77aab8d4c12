// Fused 2-layer MLP (+LayerNorm, +residual, +gather prologue) forward / backward.
//
// "Exact" engine: fp32 FFMA on CUDA cores with fp32 or bf16 activation storage.  This is
// the engine used for the fp32 configuration (1e-5 parity rules out bf16/TF32 MMA) and the
// numerical reference of the tcgen05 engine (mlp2_tc.cu).
//
// Replaces, per call, the reference-side ATen chain of mfai make_mlp / InteractionNet:
// index_select x2 + cat + Linear + SiLU + Linear + LayerNorm (+ residual add)
// (reference call site py4cast/lightning.py:593-596).
//
// Tiling: one CTA = 256 threads processes tiles of TM=128 rows, persistent over tiles.
//   thread (tx = tid&15, ty = tid>>4) owns rows {ty + 16 i, i<8} x cols {4 tx .. 4 tx+3}.
//   Layer-1 input is streamed through shared memory in K-chunks of 32 columns
//   (register-prefetched, so global loads of chunk c+1 overlap the FFMAs of chunk c).
#include <stdlib.h>

#include "common.cuh"

namespace p4c {

constexpr int TM = 128;
constexpr int HID = P4C_HIDDEN;  // 64
constexpr int KC = 32;
constexpr int NT = 256;
constexpr int XS_LD = KC + 4;   // 36
constexpr int HS_LD = HID + 4;  // 68

struct RowInfo {
  // element offset of the row in each source (-1 = row out of range)
  long long off[P4C_MAX_SRC];
};

// ------------------------------------------------------------------------------------
// Chunk loader: X chunk (TM x KC) + optional W1 chunk, via registers.
// VEC : every source is 64 wide -> chunk c = source c/2, columns (c&1)*32..+32, float4 loads
// !VEC: single dense source of width k1 -> scalar loads with column guard
// ------------------------------------------------------------------------------------
template <typename T, bool VEC>
struct XPrefetch {
  float4 v[4];  // VEC: 4 quads ; dense: 16 scalars
  __device__ __forceinline__ void load(const p4c_mlp_t& m, const long long* row_off /*[3][TM]*/,
                                       int chunk, int tid) {
    if (VEC) {
      const int s = chunk >> 1;
      const int col = (chunk & 1) * KC + (tid & 7) * 4;
      const T* base = reinterpret_cast<const T*>(m.src[s].ptr);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (tid >> 3) + 32 * j;
        const long long off = row_off[s * TM + r];
        v[j] = off >= 0 ? load4<T>(base + off + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const int col = chunk * KC + (tid & 31);
      const bool colok = col < m.k1;
      const T* base = reinterpret_cast<const T*>(m.src[0].ptr);
      float* f = reinterpret_cast<float*>(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = (tid >> 5) + 8 * j;
        const long long off = row_off[r];
        f[j] = (colok && off >= 0) ? to_f32<T>(base[off + col]) : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(float* Xs, int tid) const {
    if (VEC) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (tid >> 3) + 32 * j;
        *reinterpret_cast<float4*>(Xs + r * XS_LD + (tid & 7) * 4) = v[j];
      }
    } else {
      const float* f = reinterpret_cast<const float*>(v);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = (tid >> 5) + 8 * j;
        Xs[r * XS_LD + (tid & 31)] = f[j];
      }
    }
  }
};

// W1 chunk in "k-major" layout Ws[k][c] (forward):  Ws[k*HS_LD + c] = w1[c*k1 + kbase + k]
struct WkPrefetch {
  float v[8];
  __device__ __forceinline__ void load(const float* w1, int k1, int chunk, int tid) {
    const int k = chunk * KC + (tid & 31);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (tid >> 5) + 8 * j;
      v[j] = k < k1 ? __ldg(w1 + (long long)c * k1 + k) : 0.f;
    }
  }
  __device__ __forceinline__ void store(float* Ws, int tid) const {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = (tid >> 5) + 8 * j;
      Ws[(tid & 31) * HS_LD + c] = v[j];
    }
  }
};

__device__ __forceinline__ void compute_row_info(const p4c_mlp_t& m, long long tile_row0,
                                                 long long total_rows, long long* row_off,
                                                 long long* orow, int tid) {
  // orow[r]: row index into the output-like arrays (sample stride = full_rows), -1 = invalid
  if (tid < TM) {
    const long long g = tile_row0 + tid;
    long long o = -1;
    if (g < total_rows) {
      const long long b = g / m.rows;
      o = b * (m.full_rows > 0 ? m.full_rows : m.rows) + (g - b * m.rows);
    }
    orow[tid] = o;
  }
  // row_off[s*TM + r]
  for (int i = tid; i < m.nsrc * TM; i += NT) {
    const int s = i / TM, r = i % TM;
    const long long g = tile_row0 + r;
    long long off = -1;
    if (g < total_rows) {
      const long long b = g / m.rows;
      const int rr = (int)(g - b * m.rows);
      const long long idx = m.src[s].index ? (long long)m.src[s].index[rr] : (long long)rr;
      off = b * m.src[s].batch_stride + idx * m.src[s].ld;
    }
    row_off[i] = off;
  }
}

// acc[i] (8 rows x 4 cols) += A[rows][k0..k0+nk) * B[k][cols]   with A in smem (ld = lda),
// B in smem k-major (ld = HS_LD)
__device__ __forceinline__ void tile_fma(float4 (&acc)[8], const float* As, int lda,
                                         const float* Bs, int nk, int tx, int ty) {
#pragma unroll 2
  for (int kk = 0; kk < nk; kk += 4) {
    const float4 w0 = *reinterpret_cast<const float4*>(Bs + (kk + 0) * HS_LD + 4 * tx);
    const float4 w1 = *reinterpret_cast<const float4*>(Bs + (kk + 1) * HS_LD + 4 * tx);
    const float4 w2 = *reinterpret_cast<const float4*>(Bs + (kk + 2) * HS_LD + 4 * tx);
    const float4 w3 = *reinterpret_cast<const float4*>(Bs + (kk + 3) * HS_LD + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 a = *reinterpret_cast<const float4*>(As + (ty + 16 * i) * lda + kk);
      acc[i].x = fmaf(a.x, w0.x, acc[i].x); acc[i].y = fmaf(a.x, w0.y, acc[i].y);
      acc[i].z = fmaf(a.x, w0.z, acc[i].z); acc[i].w = fmaf(a.x, w0.w, acc[i].w);
      acc[i].x = fmaf(a.y, w1.x, acc[i].x); acc[i].y = fmaf(a.y, w1.y, acc[i].y);
      acc[i].z = fmaf(a.y, w1.z, acc[i].z); acc[i].w = fmaf(a.y, w1.w, acc[i].w);
      acc[i].x = fmaf(a.z, w2.x, acc[i].x); acc[i].y = fmaf(a.z, w2.y, acc[i].y);
      acc[i].z = fmaf(a.z, w2.z, acc[i].z); acc[i].w = fmaf(a.z, w2.w, acc[i].w);
      acc[i].x = fmaf(a.w, w3.x, acc[i].x); acc[i].y = fmaf(a.w, w3.y, acc[i].y);
      acc[i].z = fmaf(a.w, w3.z, acc[i].z); acc[i].w = fmaf(a.w, w3.w, acc[i].w);
    }
  }
}

// sum over the 16 lanes (tx) that share a row
__device__ __forceinline__ float row_sum16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  return v;
}

struct FwdSmem {
  float Xs[TM * XS_LD];
  float Ws[KC * HS_LD];
  float W2s[HID * HS_LD];  // k-major: W2s[k*HS_LD + c] = w2[c*64 + k]
  float Hs[TM * HS_LD];
  long long row_off[P4C_MAX_SRC * TM];
  long long orow[TM];
  float b1[HID], b2[HID], gamma[HID], beta[HID];
};

template <typename T, bool VEC>
__global__ void __launch_bounds__(NT, 2)
mlp2_fwd_kernel(const p4c_mlp_t m, const T* __restrict__ residual, T* __restrict__ out,
                T* __restrict__ out2, T* __restrict__ save_h1, float* __restrict__ save_stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long total_rows = (long long)m.rows * m.batch;
  const long long ntiles = (total_rows + TM - 1) / TM;
  const int nchunks = (m.k1 + KC - 1) / KC;
  const int OUT = m.out_features;

  // one-time staging of layer-2 weights and vectors
  for (int i = tid; i < HID * HID; i += NT) {
    const int c = i / HID, k = i % HID;  // coalesced over k in global
    S.W2s[k * HS_LD + c] = c < OUT ? m.w2[c * HID + k] : 0.f;
  }
  if (tid < HID) {
    S.b1[tid] = m.b1[tid];
    S.b2[tid] = tid < OUT ? m.b2[tid] : 0.f;
    S.gamma[tid] = m.gamma ? m.gamma[tid] : 1.f;
    S.beta[tid] = m.beta ? m.beta[tid] : 0.f;
  }

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    __syncthreads();  // previous tile done with row_off / Hs
    compute_row_info(m, row0, total_rows, S.row_off, S.orow, tid);
    __syncthreads();

    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    XPrefetch<T, VEC> xp;
    WkPrefetch wp;
    xp.load(m, S.row_off, 0, tid);
    wp.load(m.w1, m.k1, 0, tid);
    for (int c = 0; c < nchunks; ++c) {
      __syncthreads();  // everyone finished reading Xs/Ws of chunk c-1
      xp.store(S.Xs, tid);
      wp.store(S.Ws, tid);
      __syncthreads();
      if (c + 1 < nchunks) {
        xp.load(m, S.row_off, c + 1, tid);
        wp.load(m.w1, m.k1, c + 1, tid);
      }
      tile_fma(acc, S.Xs, XS_LD, S.Ws, KC, tx, ty);
    }

    // layer-1 epilogue: + b1, save, SiLU -> Hs
    const float4 bb1 = *reinterpret_cast<const float4*>(S.b1 + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      float4 h = f4add(acc[i], bb1);
      const long long g = S.orow[r];
      if (save_h1 && g >= 0) store4<T>(save_h1 + g * HID + 4 * tx, h);
      if (save_h1 && sizeof(T) == 2) {
        // backward recomputes from the STORED value: round-trip through T so that
        // forward and backward see the same activation
        h.x = to_f32<T>(from_f32<T>(h.x)); h.y = to_f32<T>(from_f32<T>(h.y));
        h.z = to_f32<T>(from_f32<T>(h.z)); h.w = to_f32<T>(from_f32<T>(h.w));
      }
      float4 a = make_float4(silu(h.x), silu(h.y), silu(h.z), silu(h.w));
      *reinterpret_cast<float4*>(S.Hs + r * HS_LD + 4 * tx) = a;
    }
    __syncthreads();

    // layer 2
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    tile_fma(acc, S.Hs, HS_LD, S.W2s, HID, tx, ty);

    const float4 bb2 = *reinterpret_cast<const float4*>(S.b2 + 4 * tx);
    const float4 gm = *reinterpret_cast<const float4*>(S.gamma + 4 * tx);
    const float4 bt = *reinterpret_cast<const float4*>(S.beta + 4 * tx);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      const long long g = S.orow[r];
      float4 y = f4add(acc[i], bb2);
      if (m.layer_norm) {
        const float mean = row_sum16(y.x + y.y + y.z + y.w) * (1.0f / HID);
        const float4 d = make_float4(y.x - mean, y.y - mean, y.z - mean, y.w - mean);
        const float var = row_sum16(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w) * (1.0f / HID);
        const float rstd = 1.0f / sqrtf(var + m.ln_eps);
        y = make_float4(d.x * rstd * gm.x + bt.x, d.y * rstd * gm.y + bt.y,
                        d.z * rstd * gm.z + bt.z, d.w * rstd * gm.w + bt.w);
        if (save_stats && tx == 0 && g >= 0) {
          save_stats[2 * g] = mean;
          save_stats[2 * g + 1] = rstd;
        }
      }
      if (g < 0) continue;
      if (OUT == HID) {
        if (out2) {
          const float4 e = load4<T>(reinterpret_cast<const T*>(m.src[0].ptr) + S.row_off[r] + 4 * tx);
          store4<T>(out2 + g * HID + 4 * tx, f4add(e, y));
        }
        if (residual) y = f4add(y, load4<T>(residual + g * HID + 4 * tx));
        store4<T>(out + g * HID + 4 * tx, y);
      } else {
        const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 4 * tx + j;
          if (c < OUT) {
            float v = yv[j];
            if (residual) v += to_f32<T>(residual[g * OUT + c]);
            out[g * OUT + c] = from_f32<T>(v);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------
struct BwdSmemFixed {
  float Xs[TM * XS_LD];
  float W1Ts[HID * XS_LD];  // c-major chunk: W1Ts[c*XS_LD + k] = w1[c*k1 + kbase + k]
  float W2s[HID * HS_LD];   // k-major  W2s[k*HS_LD + c] = w2[c*64+k]
  float W2Ts[HID * HS_LD];  // c-major  W2Ts[c*HS_LD + k] = w2[c*64+k]
  float Hs[TM * HS_LD];     // a1 = silu(h1)
  float Ds[TM * HS_LD];     // dh2, then dh1
  long long row_off[P4C_MAX_SRC * TM];
  long long orow[TM];
  float b2[HID], gamma[HID];
  float red[16 * HID];      // cross-thread reduction scratch for the vector grads
  // followed by dW1 accumulators: [k1pad][HID] floats
};

// per-CTA partial layout (floats): dW1 [k1pad*64] (k-major), dW2 [64*64] (k-major),
// db1[64], db2[64], dgamma[64], dbeta[64]
__host__ __device__ inline long long partial_floats(int k1) {
  const int k1pad = (k1 + KC - 1) / KC * KC;
  return (long long)k1pad * HID + HID * HID + 4 * HID;
}

template <typename T, bool VEC>
__global__ void __launch_bounds__(NT, 1)
mlp2_bwd_kernel(const p4c_mlp_t m, const p4c_mlp_bwd_t g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmemFixed& S = *reinterpret_cast<BwdSmemFixed*>(smem_raw);
  float* dW1s = reinterpret_cast<float*>(smem_raw + sizeof(BwdSmemFixed));
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const long long total_rows = (long long)m.rows * m.batch;
  const long long ntiles = (total_rows + TM - 1) / TM;
  const int nchunks = (m.k1 + KC - 1) / KC;
  const int k1pad = nchunks * KC;
  const int OUT = m.out_features;
  const bool need_dw1 = g.g_w1 != nullptr;

  for (int i = tid; i < HID * HID; i += NT) {
    const int c = i / HID, k = i % HID;
    const float w = c < OUT ? m.w2[c * HID + k] : 0.f;
    S.W2s[k * HS_LD + c] = w;
    S.W2Ts[c * HS_LD + k] = w;
  }
  if (tid < HID) {
    S.b2[tid] = tid < OUT ? m.b2[tid] : 0.f;
    S.gamma[tid] = m.gamma ? m.gamma[tid] : 1.f;
  }
  for (int i = tid; i < k1pad * HID; i += NT) dW1s[i] = 0.f;

  // register accumulators that persist over tiles
  float dW2r[16];  // dW2[k = 4*ty + a][c = 4*tx + b] -> dW2r[a*4+b]
#pragma unroll
  for (int i = 0; i < 16; ++i) dW2r[i] = 0.f;
  float4 db1p = make_float4(0, 0, 0, 0), db2p = db1p, dgp = db1p, dbp = db1p;

  const T* h1p = reinterpret_cast<const T*>(g.h1);
  const T* dya = reinterpret_cast<const T*>(g.dy_a);
  const T* dyb = reinterpret_cast<const T*>(g.dy_b);
  T* dh1_out = reinterpret_cast<T*>(g.dh1_out);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    __syncthreads();
    compute_row_info(m, row0, total_rows, S.row_off, S.orow, tid);
    __syncthreads();

    // ---- 1. h1 -> a1 (Hs)
    float4 h1r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      const long long gr = S.orow[r];
      h1r[i] = gr >= 0 ? load4<T>(h1p + gr * HID + 4 * tx) : make_float4(0, 0, 0, 0);
      *reinterpret_cast<float4*>(S.Hs + r * HS_LD + 4 * tx) =
          make_float4(silu(h1r[i].x), silu(h1r[i].y), silu(h1r[i].z), silu(h1r[i].w));
    }
    __syncthreads();

    // ---- 2. dy tile (+ LayerNorm backward) -> dh2
    float4 dh2[8];
    if (m.layer_norm) {
      // recompute h2 = a1 W2 + b2
      float4 acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = make_float4(0, 0, 0, 0);
      tile_fma(acc, S.Hs, HS_LD, S.W2s, HID, tx, ty);
      const float4 bb2 = *reinterpret_cast<const float4*>(S.b2 + 4 * tx);
      const float4 gm = *reinterpret_cast<const float4*>(S.gamma + 4 * tx);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty + 16 * i;
        const long long gr = S.orow[r];
        const bool ok = gr >= 0;
        float4 dy = make_float4(0, 0, 0, 0);
        float mean = 0.f, rstd = 0.f;
        if (ok) {
          const long long gl = row0 + r;
          const long long b = gl / m.rows;
          const int rr = (int)(gl - b * m.rows);
          const long long ia = g.dy_a_index ? (long long)g.dy_a_index[rr] : (long long)rr;
          dy = load4<T>(dya + b * g.dy_a_batch_stride + ia * HID + 4 * tx);
          if (g.dy_a_scale) {
            const float sc = g.dy_a_scale[ia];
            dy.x *= sc; dy.y *= sc; dy.z *= sc; dy.w *= sc;
          }
          if (dyb) dy = f4add(dy, load4<T>(dyb + gr * HID + 4 * tx));
          mean = g.stats[2 * gr];
          rstd = g.stats[2 * gr + 1];
        }
        const float4 y = f4add(acc[i], bb2);
        const float4 xh = make_float4((y.x - mean) * rstd, (y.y - mean) * rstd,
                                      (y.z - mean) * rstd, (y.w - mean) * rstd);
        dgp.x += dy.x * xh.x; dgp.y += dy.y * xh.y; dgp.z += dy.z * xh.z; dgp.w += dy.w * xh.w;
        dbp = f4add(dbp, dy);
        const float4 dxh = make_float4(dy.x * gm.x, dy.y * gm.y, dy.z * gm.z, dy.w * gm.w);
        const float s1 = row_sum16(dxh.x + dxh.y + dxh.z + dxh.w) * (1.0f / HID);
        const float s2 =
            row_sum16(dxh.x * xh.x + dxh.y * xh.y + dxh.z * xh.z + dxh.w * xh.w) * (1.0f / HID);
        dh2[i] = make_float4(rstd * (dxh.x - s1 - xh.x * s2), rstd * (dxh.y - s1 - xh.y * s2),
                             rstd * (dxh.z - s1 - xh.z * s2), rstd * (dxh.w - s1 - xh.w * s2));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = ty + 16 * i;
        const long long gr = S.orow[r];
        float dv[4] = {0.f, 0.f, 0.f, 0.f};
        if (gr >= 0) {
          const long long gl = row0 + r;
          const long long b = gl / m.rows;
          const int rr = (int)(gl - b * m.rows);
          const long long ia = g.dy_a_index ? (long long)g.dy_a_index[rr] : (long long)rr;
          const float sc = g.dy_a_scale ? g.dy_a_scale[ia] : 1.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = 4 * tx + j;
            if (c < OUT) {
              float v = to_f32<T>(dya[b * g.dy_a_batch_stride + ia * OUT + c]) * sc;
              if (dyb) v += to_f32<T>(dyb[gr * OUT + c]);
              dv[j] = v;
            }
          }
        }
        dh2[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      db2p = f4add(db2p, dh2[i]);
      *reinterpret_cast<float4*>(S.Ds + (ty + 16 * i) * HS_LD + 4 * tx) = dh2[i];
    }
    __syncthreads();

    // ---- 3. dW2[k][c] += sum_r a1[r][k] dh2[r][c]   (thread: k = 4ty.., c = 4tx..)
    if (g.g_w2) {
#pragma unroll 4
      for (int r = 0; r < TM; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(S.Hs + r * HS_LD + 4 * ty);
        const float4 d = *reinterpret_cast<const float4*>(S.Ds + r * HS_LD + 4 * tx);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          dW2r[p * 4 + 0] = fmaf(av[p], d.x, dW2r[p * 4 + 0]);
          dW2r[p * 4 + 1] = fmaf(av[p], d.y, dW2r[p * 4 + 1]);
          dW2r[p * 4 + 2] = fmaf(av[p], d.z, dW2r[p * 4 + 2]);
          dW2r[p * 4 + 3] = fmaf(av[p], d.w, dW2r[p * 4 + 3]);
        }
      }
    }

    // ---- 4. da1 = dh2 W2  (A = Ds, B[c][k] = W2Ts) ; dh1 = da1 * silu'(h1)
    float4 dh1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dh1[i] = make_float4(0, 0, 0, 0);
    tile_fma(dh1, S.Ds, HS_LD, S.W2Ts, HID, tx, ty);
    __syncthreads();  // all reads of Ds (dh2) done
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      const long long gr = S.orow[r];
      dh1[i].x *= silu_grad(h1r[i].x); dh1[i].y *= silu_grad(h1r[i].y);
      dh1[i].z *= silu_grad(h1r[i].z); dh1[i].w *= silu_grad(h1r[i].w);
      if (gr < 0) dh1[i] = make_float4(0, 0, 0, 0);
      db1p = f4add(db1p, dh1[i]);
      *reinterpret_cast<float4*>(S.Ds + r * HS_LD + 4 * tx) = dh1[i];
      if (dh1_out && gr >= 0) store4<T>(dh1_out + gr * HID + 4 * tx, dh1[i]);
    }
    // (Ds visibility is ensured by the __syncthreads at the top of the chunk loop)

    // ---- 5. per K-chunk: dW1 += X^T dh1 ; dX = dh1 W1
    XPrefetch<T, VEC> xp;
    if (need_dw1) xp.load(m, S.row_off, 0, tid);
    for (int c = 0; c < nchunks; ++c) {
      const int s = VEC ? (c >> 1) : 0;
      const bool need_dx = g.dx[s] != nullptr;
      __syncthreads();
      if (need_dw1) xp.store(S.Xs, tid);
      if (need_dx) {
        // W1Ts[c'][k] = w1[c'*k1 + kbase + k]  : 64 x 32, coalesced over k
        for (int i = tid; i < HID * KC; i += NT) {
          const int cc = i / KC, k = i % KC;
          const int kg = c * KC + k;
          S.W1Ts[cc * XS_LD + k] = kg < m.k1 ? __ldg(m.w1 + (long long)cc * m.k1 + kg) : 0.f;
        }
      }
      __syncthreads();
      if (need_dw1 && c + 1 < nchunks) xp.load(m, S.row_off, c + 1, tid);

      if (need_dw1) {
        // halves of the tile rows, summed in a fixed order (half 0 then half 1)
        const int half = tid >> 7, t = tid & 127;
        const int cq = t & 15, kq = t >> 4;  // c = 4cq.., k = 4kq.. (within chunk)
        float a16[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a16[i] = 0.f;
#pragma unroll 4
        for (int r = half * 64; r < half * 64 + 64; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(S.Xs + r * XS_LD + 4 * kq);
          const float4 d = *reinterpret_cast<const float4*>(S.Ds + r * HS_LD + 4 * cq);
          const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            a16[p * 4 + 0] = fmaf(xv[p], d.x, a16[p * 4 + 0]);
            a16[p * 4 + 1] = fmaf(xv[p], d.y, a16[p * 4 + 1]);
            a16[p * 4 + 2] = fmaf(xv[p], d.z, a16[p * 4 + 2]);
            a16[p * 4 + 3] = fmaf(xv[p], d.w, a16[p * 4 + 3]);
          }
        }
        float* dst = dW1s + (c * KC + 4 * kq) * HID + 4 * cq;
        if (half == 0) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float4* q = reinterpret_cast<float4*>(dst + p * HID);
            float4 v = *q;
            v.x += a16[p * 4 + 0]; v.y += a16[p * 4 + 1]; v.z += a16[p * 4 + 2]; v.w += a16[p * 4 + 3];
            *q = v;
          }
        }
        __syncthreads();
        if (half == 1) {
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float4* q = reinterpret_cast<float4*>(dst + p * HID);
            float4 v = *q;
            v.x += a16[p * 4 + 0]; v.y += a16[p * 4 + 1]; v.z += a16[p * 4 + 2]; v.w += a16[p * 4 + 3];
            *q = v;
          }
        }
      }

      if (need_dx) {
        // dX[r][k] = sum_c dh1[r][c] W1[c][k] ; thread: rows ty+16i, k = 2tx, 2tx+1
        float2 ax[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ax[i] = make_float2(0.f, 0.f);
#pragma unroll 2
        for (int cc = 0; cc < HID; cc += 4) {
          const float2 w0 = *reinterpret_cast<const float2*>(S.W1Ts + (cc + 0) * XS_LD + 2 * tx);
          const float2 w1 = *reinterpret_cast<const float2*>(S.W1Ts + (cc + 1) * XS_LD + 2 * tx);
          const float2 w2 = *reinterpret_cast<const float2*>(S.W1Ts + (cc + 2) * XS_LD + 2 * tx);
          const float2 w3 = *reinterpret_cast<const float2*>(S.W1Ts + (cc + 3) * XS_LD + 2 * tx);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 d = *reinterpret_cast<const float4*>(S.Ds + (ty + 16 * i) * HS_LD + cc);
            ax[i].x = fmaf(d.x, w0.x, ax[i].x); ax[i].y = fmaf(d.x, w0.y, ax[i].y);
            ax[i].x = fmaf(d.y, w1.x, ax[i].x); ax[i].y = fmaf(d.y, w1.y, ax[i].y);
            ax[i].x = fmaf(d.z, w2.x, ax[i].x); ax[i].y = fmaf(d.z, w2.y, ax[i].y);
            ax[i].x = fmaf(d.w, w3.x, ax[i].x); ax[i].y = fmaf(d.w, w3.y, ax[i].y);
          }
        }
        T* dxp = reinterpret_cast<T*>(g.dx[s]);
        const int ld = g.dx_ld[s];
        const int colbase = VEC ? (c & 1) * KC : c * KC;
        const int width = m.src[s].width;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long long gr = S.orow[ty + 16 * i];
          if (gr < 0) continue;
          const float av[2] = {ax[i].x, ax[i].y};
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int col = colbase + 2 * tx + j;
            if (col < width) {
              T* p = dxp + gr * ld + col;
              float v = av[j];
              if (g.dx_accumulate[s]) v += to_f32<T>(*p);
              if (s == 0 && g.dx0_add_dy) {
                // residual / edge-update pass-through: d src0 += dy (dy_b only when == 1)
                if (g.dx0_add_dy == 2) {
                  const long long gl = row0 + ty + 16 * i;
                  const long long b = gl / m.rows;
                  const int rr = (int)(gl - b * m.rows);
                  const long long ia = g.dy_a_index ? (long long)g.dy_a_index[rr] : (long long)rr;
                  float da = to_f32<T>(dya[b * g.dy_a_batch_stride + ia * HID + col]);
                  if (g.dy_a_scale) da *= g.dy_a_scale[ia];
                  v += da;
                }
                if (dyb) v += to_f32<T>(dyb[gr * HID + col]);
              }
              *p = from_f32<T>(v);
            }
          }
        }
      }
    }
  }

  // ---- write per-CTA partials
  __syncthreads();
  float* part = g.wgrad_partials + (long long)blockIdx.x * partial_floats(m.k1);
  for (int i = tid; i < k1pad * HID; i += NT) part[i] = dW1s[i];
  float* p2 = part + (long long)k1pad * HID;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) p2[(4 * ty + a) * HID + 4 * tx + b] = dW2r[a * 4 + b];
  float* pv = p2 + HID * HID;
  // vector grads: reduce the 16 ty-groups in a fixed order through smem
  const float4 vecs[4] = {db1p, db2p, dgp, dbp};
  for (int v = 0; v < 4; ++v) {
    __syncthreads();
    *reinterpret_cast<float4*>(S.red + ty * HID + 4 * tx) = vecs[v];
    __syncthreads();
    if (tid < HID) {
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 16; ++q) s += S.red[q * HID + tid];
      pv[v * HID + tid] = s;
    }
  }
}

// out[b,r,:] (+)= in[b,r,:] x w1[:, col0:col0+64]      (64x64 slice of the layer-1 weight)
struct RmSmem {
  float Xs[TM * HS_LD];
  float Ws[HID * HS_LD];  // Ws[c][j] = w1[c*k1 + col0 + j]
};
template <typename T>
__global__ void __launch_bounds__(NT, 2)
rows_matmul_kernel(const T* __restrict__ in, long long in_bs, const float* __restrict__ w, int k1,
                   int col0, int rows, int batch, T* __restrict__ out, long long out_bs,
                   int accumulate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RmSmem& S = *reinterpret_cast<RmSmem*>(smem_raw);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  for (int i = tid; i < HID * HID; i += NT) {
    const int c = i / HID, j = i % HID;
    S.Ws[c * HS_LD + j] = w[(long long)c * k1 + col0 + j];
  }
  const long long total = (long long)rows * batch;
  const long long ntiles = (total + TM - 1) / TM;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row0 = tile * TM;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = ty + 16 * i;
      const long long gr = row0 + r;
      float4 v = make_float4(0, 0, 0, 0);
      if (gr < total) {
        const long long b = gr / rows, rr = gr - b * rows;
        v = load4<T>(in + b * in_bs + rr * HID + 4 * tx);
      }
      *reinterpret_cast<float4*>(S.Xs + r * HS_LD + 4 * tx) = v;
    }
    __syncthreads();
    float4 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float4(0, 0, 0, 0);
    tile_fma(acc, S.Xs, HS_LD, S.Ws, HID, tx, ty);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long gr = row0 + ty + 16 * i;
      if (gr >= total) continue;
      const long long b = gr / rows, rr = gr - b * rows;
      T* p = out + b * out_bs + rr * HID + 4 * tx;
      float4 v = acc[i];
      if (accumulate) v = f4add(v, load4<T>(p));
      store4<T>(p, v);
    }
  }
}

static int validate(const p4c_mlp_t* m) {
  P4C_REQUIRE(m != nullptr, "mlp: null descriptor");
  P4C_REQUIRE(m->rows > 0 && m->batch > 0, "mlp: empty problem (rows=%d batch=%d)", m->rows, m->batch);
  P4C_REQUIRE(m->nsrc >= 1 && m->nsrc <= P4C_MAX_SRC, "mlp: nsrc=%d out of range", m->nsrc);
  P4C_REQUIRE(m->dtype == P4C_F32 || m->dtype == P4C_BF16, "mlp: bad dtype %d", m->dtype);
  P4C_REQUIRE(m->full_rows == 0 || m->full_rows >= m->rows, "mlp: full_rows=%d < rows=%d", m->full_rows, m->rows);
  P4C_REQUIRE(m->out_features >= 1 && m->out_features <= HID, "mlp: out_features=%d not in [1,64]",
              m->out_features);
  P4C_REQUIRE(!m->layer_norm || m->out_features == HID, "mlp: layer_norm requires out_features == 64");
  int k = 0;
  for (int s = 0; s < m->nsrc; ++s) {
    P4C_REQUIRE(m->src[s].ptr != nullptr, "mlp: src[%d] null", s);
    P4C_REQUIRE(m->src[s].width >= 1 && m->src[s].ld >= m->src[s].width, "mlp: src[%d] bad width/ld", s);
    if (m->nsrc > 1 || m->src[s].width == HID)
      P4C_REQUIRE(m->src[s].width == HID && m->src[s].ld % 4 == 0 && m->src[s].batch_stride % 4 == 0,
                  "mlp: multi-source / vector inputs must be 64 wide with 4-element aligned strides");
    k += m->src[s].width;
  }
  P4C_REQUIRE(k == m->k1, "mlp: k1=%d does not match sum of widths %d", m->k1, k);
  P4C_REQUIRE(m->nsrc > 1 || m->src[0].index == nullptr || m->src[0].width == HID,
              "mlp: dense (non-64) source cannot be gathered");
  P4C_REQUIRE(m->w1 && m->b1 && m->w2 && m->b2, "mlp: null weights");
  P4C_REQUIRE(m->tile_rows == 0 || m->tile_rows == 16 || m->tile_rows == 32 || m->tile_rows == 64 ||
                  m->tile_rows == 128, "mlp: tile_rows=%d not in {0,16,32,64,128}", m->tile_rows);
  for (int s = 0; s < m->nsrc; ++s)
    P4C_REQUIRE(!m->src[s].seg_rowptr || (m->dtype == P4C_BF16 && tc_engine_enabled()),
                "mlp: segment-sum sources need the tensor-core engine (bf16 storage)");
  return P4C_OK;
}

static bool is_vec(const p4c_mlp_t* m) { return m->src[0].width == HID; }

static int g_tc_enabled = -1;
bool tc_engine_enabled() {
  if (g_tc_enabled < 0) {
    const char* e = getenv("P4C_ENGINE");
    g_tc_enabled = (e && (e[0] == 's' || e[0] == 'S')) ? 0 : 1;
  }
  return g_tc_enabled == 1;
}

// Opt in to > 48 KB of dynamic shared memory, once per kernel instantiation (the attribute is
// sticky), so that nothing but launches happens while a CUDA graph is being captured.
template <auto kKernel>
static int set_smem(size_t bytes, const char* name) {
  static size_t configured = 0;  // one static per kernel (non-type template parameter)
  if (bytes <= configured) return P4C_OK;
  cudaError_t e = cudaFuncSetAttribute(kKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu) failed: %s", name, bytes, cudaGetErrorString(e));
    return P4C_ECUDA;
  }
  configured = bytes;
  return P4C_OK;
}

}  // namespace p4c

using namespace p4c;

static int g_cta_reserve = 0;
namespace p4c {
int tc_max_ctas() { return 2 * kNumSMs - g_cta_reserve; }
}  // namespace p4c
extern "C" void p4c_set_cta_reserve(int n) { g_cta_reserve = n < 0 ? 0 : (n > 2 * kNumSMs - 16 ? 2 * kNumSMs - 16 : n); }
extern "C" int p4c_get_cta_reserve(void) { return g_cta_reserve; }
extern "C" void p4c_set_trace_buffer(void* device_buffer) { tc_set_trace(static_cast<long long*>(device_buffer)); }
extern "C" void p4c_set_engine(int tensor_cores) { g_tc_enabled = tensor_cores ? 1 : 0; }
extern "C" int p4c_get_engine(void) { return tc_engine_enabled() ? 1 : 0; }

extern "C" int p4c_mlp_forward_aggregate(const p4c_mlp_t* m, void* out, void* out2, void* save_h1,
                                         float* save_stats, void* aggr_out, int32_t group, float scale,
                                         void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  P4C_REQUIRE(aggr_out != nullptr, "mlp_forward_aggregate: aggr_out is null");
  P4C_REQUIRE(group >= 1 && group <= 8 && (group & (group - 1)) == 0,
              "mlp_forward_aggregate: group=%d must be 1, 2, 4 or 8", group);
  const int full = m->full_rows > 0 ? m->full_rows : m->rows;
  P4C_REQUIRE(m->rows % group == 0 && full % group == 0, "mlp_forward_aggregate: rows=%d / full_rows=%d not a multiple of group=%d",
              m->rows, full, group);
  P4C_REQUIRE(m->out_features == HID, "mlp_forward_aggregate: out_features must be 64");
  P4C_REQUIRE(!out2 || (m->src[0].width == HID && m->src[0].index == nullptr),
              "mlp_forward: out2 needs a 64-wide identity-indexed src[0]");
  P4C_REQUIRE(tc_engine_enabled() && tc_forward_supported(m) && reinterpret_cast<uintptr_t>(aggr_out) % 16 == 0,
              "mlp_forward_aggregate: needs the tensor-core engine (bf16 storage, aligned 64-wide sources)");
  return launch_mlp2_tc_forward(m, nullptr, out, out2, save_h1, save_stats, aggr_out, group, scale,
                                (cudaStream_t)stream);
}

extern "C" int p4c_mlp_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2,
                               void* save_h1, float* save_stats, void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  P4C_REQUIRE(out != nullptr, "mlp_forward: out is null");
  P4C_REQUIRE(!out2 || (m->out_features == HID && m->src[0].width == HID && m->src[0].index == nullptr),
              "mlp_forward: out2 needs a 64-wide identity-indexed src[0]");
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_engine_enabled() && tc_forward_supported(m))
    return launch_mlp2_tc_forward(m, residual, out, out2, save_h1, save_stats, nullptr, 0, 0.f, st);
  for (int s = 0; s < m->nsrc; ++s)
    P4C_REQUIRE(!m->src[s].seg_rowptr, "mlp_forward: segment-sum source not supported for this shape / alignment");
  const long long total = (long long)m->rows * m->batch;
  const long long ntiles = (total + TM - 1) / TM;
  const int grid = (int)(ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs);
  const size_t smem = sizeof(FwdSmem);
#define LAUNCH_FWD(T, VEC)                                                                   \
  do {                                                                                       \
    rc = set_smem<mlp2_fwd_kernel<T, VEC>>(smem, "mlp_forward");                             \
    if (rc) return rc;                                                                       \
    mlp2_fwd_kernel<T, VEC><<<grid, NT, smem, st>>>(*m, (const T*)residual, (T*)out, (T*)out2, \
                                                    (T*)save_h1, save_stats);                \
  } while (0)
  if (m->dtype == P4C_F32) {
    if (is_vec(m)) LAUNCH_FWD(float, true); else LAUNCH_FWD(float, false);
  } else {
    if (is_vec(m)) LAUNCH_FWD(__nv_bfloat16, true); else LAUNCH_FWD(__nv_bfloat16, false);
  }
#undef LAUNCH_FWD
  P4C_CHECK_LAUNCH("mlp_forward");
  return P4C_OK;
}

static int bwd_grid(const p4c_mlp_t* m) {
  const long long total = (long long)m->rows * m->batch;
  const long long ntiles = (total + TM - 1) / TM;
  return (int)(ntiles < kNumSMs ? ntiles : kNumSMs);
}

extern "C" int32_t p4c_mlp_wgrad_blocks(const p4c_mlp_t* m) {
  if (!m || m->rows <= 0 || m->batch <= 0) return 0;
  const int g_simt = bwd_grid(m), g_tc = tc_backward_grid(m);
  return g_simt > g_tc ? g_simt : g_tc;
}

extern "C" int64_t p4c_mlp_wgrad_partial_floats(const p4c_mlp_t* m) {
  if (!m || m->rows <= 0 || m->batch <= 0) return 0;
  const int g_simt = bwd_grid(m), g_tc = tc_backward_grid(m);  // either engine may serve the call
  return (int64_t)(g_simt > g_tc ? g_simt : g_tc) * partial_floats(m->k1);
}

extern "C" int p4c_mlp_backward(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, void* stream) {
  int rc = validate(m);
  if (rc) return rc;
  P4C_REQUIRE(g != nullptr && g->dy_a != nullptr, "mlp_backward: dy_a null");
  // the pipelined engine recomputes h1 and the LayerNorm statistics; every other engine reads them
  const bool ws = tc_engine_enabled() && ws_backward_supported(m, g);
  P4C_REQUIRE(ws || g->h1 != nullptr,
              "mlp_backward: h1 null (only calls with p4c_mlp_bwd_caps(m) & 1 and no dy_a_scale may omit it)");
  P4C_REQUIRE(ws || !m->layer_norm || g->stats != nullptr, "mlp_backward: LayerNorm stats missing");
  P4C_REQUIRE(ws || (!g->dx_group_out && !g->dx0_batch_sum),
              "mlp_backward: dx_group_out / dx0_batch_sum need the pipelined engine (p4c_mlp_bwd_caps)");
  P4C_REQUIRE(g->wgrad_partials != nullptr &&
                  g->wgrad_partials_floats >= p4c_mlp_wgrad_partial_floats(m),
              "mlp_backward: wgrad workspace too small");
  for (int s = 0; s < m->nsrc; ++s)
    P4C_REQUIRE(!g->dx[s] || m->src[s].index == nullptr,
                "mlp_backward: dx[%d] requested for a gathered source (use dh1 + segment_sum)", s);
  P4C_REQUIRE(g->partials_mode >= 0 && g->partials_mode <= 2, "mlp_backward: bad partials_mode %d", g->partials_mode);
  cudaStream_t st = (cudaStream_t)stream;
  int grid;
  if (ws) {
    grid = ws_backward_grid(m);
    rc = launch_mlp2_ws_backward(m, g, p4c_mlp_wgrad_blocks(m), st);
    if (rc) return rc;
    if (g->partials_mode != 0) return P4C_OK;
  } else if (tc_engine_enabled() && tc_backward_supported(m, g)) {
    grid = tc_backward_grid(m);
    rc = launch_mlp2_tc_backward(m, g, st);
    if (rc) return rc;
    if (g->partials_mode != 0) return P4C_OK;  // reduction deferred to p4c_wgrad_reduce_batched
  } else {
  P4C_REQUIRE(g->partials_mode == 0, "mlp_backward: deferred weight-gradient reduction needs the tensor-core engine");
  for (int s = 0; s < m->nsrc; ++s)
    P4C_REQUIRE(!m->src[s].seg_rowptr, "mlp_backward: segment-sum sources are forward-only (pass the saved sums)");
  const int k1pad = (m->k1 + KC - 1) / KC * KC;
  const size_t smem = sizeof(BwdSmemFixed) + (size_t)k1pad * HID * sizeof(float);
  P4C_REQUIRE(smem <= 227 * 1024, "mlp_backward: k1=%d needs %zu B of shared memory", m->k1, smem);
  grid = bwd_grid(m);
#define LAUNCH_BWD(T, VEC)                                               \
  do {                                                                   \
    rc = set_smem<mlp2_bwd_kernel<T, VEC>>(smem, "mlp_backward");        \
    if (rc) return rc;                                                   \
    mlp2_bwd_kernel<T, VEC><<<grid, NT, smem, st>>>(*m, *g);             \
  } while (0)
  if (m->dtype == P4C_F32) {
    if (is_vec(m)) LAUNCH_BWD(float, true); else LAUNCH_BWD(float, false);
  } else {
    if (is_vec(m)) LAUNCH_BWD(__nv_bfloat16, true); else LAUNCH_BWD(__nv_bfloat16, false);
  }
#undef LAUNCH_BWD
  P4C_CHECK_LAUNCH("mlp_backward");
  }
  p4c_wgrad_t e;
  e.partials = g->wgrad_partials;
  e.g_w1 = g->g_w1; e.g_b1 = g->g_b1; e.g_w2 = g->g_w2; e.g_b2 = g->g_b2;
  e.g_gamma = g->g_gamma; e.g_beta = g->g_beta;
  e.nblocks = grid; e.k1 = m->k1; e.out_features = m->out_features; e._pad = 0;
  return launch_wgrad_reduce_single(&e, st);
}

extern "C" int p4c_rows_matmul(const void* in, int64_t in_bs, const float* w, int32_t k1,
                               int32_t col0, int32_t rows, int32_t batch, void* out, int64_t out_bs,
                               int accumulate, int dtype, void* stream) {
  P4C_REQUIRE(in && w && out, "rows_matmul: null pointer");
  P4C_REQUIRE(rows > 0 && batch > 0, "rows_matmul: empty problem");
  P4C_REQUIRE(col0 >= 0 && col0 + HID <= k1, "rows_matmul: column window [%d,%d) outside k1=%d", col0,
              col0 + HID, k1);
  const long long total = (long long)rows * batch;
  const long long ntiles = (total + TM - 1) / TM;
  const int grid = (int)(ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs);
  const size_t smem = sizeof(RmSmem);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (dtype == P4C_F32) {
    rc = set_smem<rows_matmul_kernel<float>>(smem, "rows_matmul");
    if (rc) return rc;
    rows_matmul_kernel<float><<<grid, NT, smem, st>>>((const float*)in, in_bs, w, k1, col0, rows, batch,
                                                      (float*)out, out_bs, accumulate);
  } else if (dtype == P4C_BF16) {
    rc = set_smem<rows_matmul_kernel<__nv_bfloat16>>(smem, "rows_matmul");
    if (rc) return rc;
    rows_matmul_kernel<__nv_bfloat16><<<grid, NT, smem, st>>>((const __nv_bfloat16*)in, in_bs, w, k1,
                                                              col0, rows, batch, (__nv_bfloat16*)out,
                                                              out_bs, accumulate);
  } else {
    P4C_REQUIRE(false, "rows_matmul: bad dtype %d", dtype);
  }
  P4C_CHECK_LAUNCH("rows_matmul");
  return P4C_OK;
}
