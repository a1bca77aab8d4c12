// SURVEY 8(f) rows 3 / 4, the data either side of the training path, as HBM-bound one-pass kernels:
//
//  * power spectral density (reference py4cast/metrics.py:253-352, used by MetricPSDK / MetricPSDVar): the
//    reference takes an orthonormal 2-D DCT-II of every (sample, feature) field on the host (scipy), squares it,
//    averages over the batch and bins "radially" -- but its binning indexes the FLATTENED spectrum with 2 r
//    (metrics.py:311-315), so only the first 2 Rmax + 1 entries of the spectrum's first row and its very last
//    entry are ever read.  The first row of a separable DCT is the 1-D DCT of the column sums, the last entry a
//    doubly-weighted sum: ONE pass over the fields (two column reductions) + a tiny DCT of the sums.
//  * per-feature statistics (datasets/compute_dataset_stats.py:11-127): NaN-ignoring mean, mean of squares, min,
//    max per (sample, feature) of a batch tensor, or of the time difference of two tensors (diff stats).
//  * un-normalisation of predictions (lightning.py:1162-1169): x = x * std_f + mean_f in place.
//
// All reductions have a fixed order (no atomics): bit-reproducible.
#include "common.cuh"

namespace p4c {

constexpr int SM_NT = 256;

// ---------------------------------------------------------------------------------- PSD: column sums
// x [bt][H][W*F] (features last), optional mask of the same shape; s1 = sum_y x m, s2 = sum_y wy[y] x m.
// grid (ceil(WF / 256), bt, ysplit): partial[(ys * bt_n + bt) * WF + e][2]
__global__ void __launch_bounds__(SM_NT) psd_colsum_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                           const float* __restrict__ wy, int H, long long WF,
                                                           int rows_per_split, float2* __restrict__ partial) {
  const long long e = static_cast<long long>(blockIdx.x) * SM_NT + threadIdx.x;
  if (e >= WF) return;
  const long long bt = blockIdx.y, ys = blockIdx.z;
  const int y0 = static_cast<int>(ys) * rows_per_split;
  const int y1 = min(H, y0 + rows_per_split);
  const float* p = x + (bt * H + y0) * WF + e;
  const float* m = mask ? mask + (bt * H + y0) * WF + e : nullptr;
  float a = 0.f, b = 0.f;
  for (int y = y0; y < y1; ++y, p += WF) {
    float v = *p;
    if (m) {
      v *= *m;
      m += WF;
    }
    a += v;
    b = fmaf(wy[y], v, b);
  }
  partial[(ys * gridDim.y + bt) * WF + e] = make_float2(a, b);
}

// s[bt][e] = sum over the row splits of partial (fixed order), in place in split 0
__global__ void __launch_bounds__(SM_NT) psd_split_reduce_kernel(float2* __restrict__ partial, int nsplit, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * SM_NT + threadIdx.x;
  if (i >= n) return;
  float2 a = partial[i];
  for (int ys = 1; ys < nsplit; ++ys) {
    const float2 v = partial[ys * n + i];
    a.x += v.x;
    a.y += v.y;
  }
  partial[i] = a;
}

// spec[bt][f][j], j < J: ((alpha_0(H) alpha_j(W) sum_x s1[x] cos(pi (2x+1) j / 2W))^2) / W^2   (first spectrum row)
// spec[bt][f][J]       : ((alpha_{H-1}(H) alpha_{W-1}(W) sum_x s2[x] cos(pi (2x+1)(W-1) / 2W))^2) / W^2   (last entry)
// ctab [W][J + 1]: the cosines (column J = those of k = W - 1), computed once per grid shape in fp64 on the host.
// fp64 accumulation: the sums cancel strongly for smooth fields.
// block = 32 consecutive j (lanes: coalesced table rows, the column sum is a broadcast) x 8 ranges of x (warps), four
// loads in flight per thread; the eight partial sums are added in warp order (fixed association).  One thread per
// output with a serial loop over x -- the first version -- left this tiny kernel at twice the time of the pass over
// the fields (198 us against 92 us for 16 x 512 x 640 x 21: every iteration waited for an L2 round trip).
constexpr int PSD_XS = 8;
__global__ void __launch_bounds__(32 * PSD_XS) psd_finalize_kernel(const float2* __restrict__ s,
                                                                   const float* __restrict__ ctab, long long bt_n,
                                                                   int H, int W, int F, int J,
                                                                   float* __restrict__ spec) {
  __shared__ double red[PSD_XS][32];
  const int lane = threadIdx.x & 31, xs = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int f = blockIdx.y;
  const long long bt = blockIdx.z;
  const bool valid = j <= J;
  const bool last = j == J;
  const float2* sp = s + bt * W * F + f;
  const float* cp = ctab + (valid ? j : 0);
  const int per = (W + PSD_XS - 1) / PSD_XS;
  const int x0 = xs * per, x1 = min(W, x0 + per);
  double acc = 0.0;
  int x = x0;
  for (; x + 4 <= x1; x += 4) {
    float2 v[4];
    float c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u] = sp[static_cast<long long>(x + u) * F];
      c[u] = cp[static_cast<long long>(x + u) * (J + 1)];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      acc = fma(static_cast<double>(last ? v[u].y : v[u].x), static_cast<double>(c[u]), acc);
  }
  for (; x < x1; ++x) {
    const float2 v = sp[static_cast<long long>(x) * F];
    acc = fma(static_cast<double>(last ? v.y : v.x), static_cast<double>(cp[static_cast<long long>(x) * (J + 1)]), acc);
  }
  red[xs][lane] = acc;
  __syncthreads();
  if (xs != 0 || !valid) return;
  for (int k = 1; k < PSD_XS; ++k) acc += red[k][lane];
  const int k = last ? W - 1 : j;
  const double ax = k == 0 ? sqrt(1.0 / W) : sqrt(2.0 / W);
  const double ay = last ? (H == 1 ? 1.0 : sqrt(2.0 / H)) : sqrt(1.0 / H);
  const double d = ax * ay * acc;
  spec[(bt * F + f) * (J + 1) + j] = static_cast<float>(d * d / (static_cast<double>(W) * W));
}

// prof[f][r] = mean_b spec[b][f][2r] + 0.5 mean_b spec[b][f][left(r)] + 0.5 mean_b spec[b][f][2r + 1], left(0) = J (the
// reference's flat index -1 wraps to the last spectrum entry), left(r) = 2r - 1; NaN where the reference's bin is
// empty (metrics.py:311-321, 343).  Replaces a dozen latency-bound torch launches after the spectrum kernel.
__global__ void __launch_bounds__(SM_NT) psd_profile_kernel(const float* __restrict__ spec,
                                                            const unsigned char* __restrict__ present, int B, int F, int R,
                                                            float* __restrict__ prof) {
  const int i = blockIdx.x * SM_NT + threadIdx.x;
  if (i >= F * R) return;
  const int f = i / R, r = i - f * R;
  const int J = 2 * R;
  const int il = r == 0 ? J : 2 * r - 1;
  float c = 0.f, l = 0.f, g = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* sp = spec + (static_cast<long long>(b) * F + f) * (J + 1);
    c += sp[2 * r];
    l += sp[il];
    g += sp[2 * r + 1];
  }
  const float inv = static_cast<float>(B);
  c = c / inv; l = l / inv; g = g / inv;
  const float v = __fadd_rn(__fadd_rn(c, __fmul_rn(0.5f, l)), __fmul_rn(0.5f, g));
  prof[i] = present[r] ? v : __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------- per-feature statistics
// x [B][X][F] (optionally minus prev, same layout): per (b, f) NaN-ignoring count, sum, sum of squares, min, max.
// block = R rows x F features (R = SM_NT / F), grid (chunks, B); partial[(b * chunks + c) * F + f] = 5 doubles
struct StatPart {
  double n, s, q, mn, mx;
};
__global__ void __launch_bounds__(SM_NT) feature_stats_kernel(const float* __restrict__ x, const float* __restrict__ prev,
                                                              long long X, int F, int R, long long rows_per_chunk,
                                                              StatPart* __restrict__ partial) {
  __shared__ StatPart red[SM_NT];
  const int tid = threadIdx.x;
  const int r = tid / F, f = tid - r * F;
  const long long b = blockIdx.y, c = blockIdx.x;
  StatPart a = {0.0, 0.0, 0.0, INFINITY, -INFINITY};
  if (r < R) {
    const long long x0 = c * rows_per_chunk, x1 = min(X, x0 + rows_per_chunk);
    // groups of 8 rows: loads issued together, group sums in fp32 (8 terms), accumulation over groups in fp64
    float mn = INFINITY, mx = -INFINITY;
    // pointers advance by a constant stride (no 64-bit index products per load); whole groups of 8 rows take the
    // unchecked path, the tail of the chunk the checked one -- same values, same order of additions
    const long long step = static_cast<long long>(R) * F;
    const float* px = x + (b * X + x0 + r) * F + f;
    const float* pp = prev ? prev + (b * X + x0 + r) * F + f : nullptr;
    long long i0 = x0 + r;
    auto fold = [&](const float (&v)[8]) {
      float n8 = 0.f, s8 = 0.f, q8 = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = v[u] == v[u];  // not NaN (fminf / fmaxf ignore a NaN operand)
        const float w = ok ? v[u] : 0.f;
        n8 += ok ? 1.f : 0.f;
        s8 += w;
        q8 = fmaf(w, w, q8);
        mn = fminf(mn, v[u]);
        mx = fmaxf(mx, v[u]);
      }
      a.n += n8;
      a.s += s8;
      a.q += q8;
    };
    for (; i0 + 7ll * R < x1; i0 += 8ll * R, px += 8 * step) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = px[u * step];
      if (pp) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] -= pp[u * step];
        pp += 8 * step;
      }
      fold(v);
    }
    if (i0 < x1) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool in = i0 + static_cast<long long>(u) * R < x1;
        v[u] = in ? (pp ? px[u * step] - pp[u * step] : px[u * step]) : __int_as_float(0x7fc00000);
      }
      fold(v);
    }
    a.mn = mn;
    a.mx = mx;
  }
  red[tid] = a;
  __syncthreads();
  if (tid < F) {  // rows in ascending order: fixed association
    StatPart t = red[tid];
    for (int rr = 1; rr < R; ++rr) {
      const StatPart u = red[rr * F + tid];
      t.n += u.n; t.s += u.s; t.q += u.q;
      t.mn = fmin(t.mn, u.mn); t.mx = fmax(t.mx, u.mx);
    }
    partial[(b * gridDim.x + c) * F + tid] = t;
  }
}
// out[b][f] = (nanmean, nanmean of squares, min (NaN -> +inf), max (NaN -> -inf))
// one warp per (b, f): lane l folds chunks l, l + 32, ... in ascending order, then the 32 lane results are folded by a
// fixed butterfly (bit-reproducible).  One thread per (b, f) looping over up to 1024 chunks of 40-byte records -- the
// first version -- took 157 us next to a 472 us pass over the data.
__global__ void __launch_bounds__(128) feature_stats_finalize_kernel(const StatPart* __restrict__ partial, int chunks, int F,
                                                                     long long BF, float4* __restrict__ out) {
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= BF) return;  // whole warps leave together (BF is counted in warps)
  const long long b = i / F;
  const int f = static_cast<int>(i - b * F);
  StatPart t = {0.0, 0.0, 0.0, INFINITY, -INFINITY};
  for (int c = lane; c < chunks; c += 32) {
    const StatPart u = partial[(b * chunks + c) * F + f];
    t.n += u.n; t.s += u.s; t.q += u.q;
    t.mn = fmin(t.mn, u.mn); t.mx = fmax(t.mx, u.mx);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    t.n += __shfl_xor_sync(0xffffffffu, t.n, o);
    t.s += __shfl_xor_sync(0xffffffffu, t.s, o);
    t.q += __shfl_xor_sync(0xffffffffu, t.q, o);
    t.mn = fmin(t.mn, __shfl_xor_sync(0xffffffffu, t.mn, o));
    t.mx = fmax(t.mx, __shfl_xor_sync(0xffffffffu, t.mx, o));
  }
  if (lane != 0) return;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  out[i] = make_float4(static_cast<float>(t.n > 0 ? t.s / t.n : nan), static_cast<float>(t.n > 0 ? t.q / t.n : nan),
                       static_cast<float>(t.mn), static_cast<float>(t.mx));
}

// ------------------------------------------------------------------------------- un-normalisation
// x[i][f] = x[i][f] * scale[f] + shift[f], two roundings as the reference's `*=` then `+=`
// thread = (row within a group of R = 256 / F rows, feature): consecutive threads touch consecutive addresses,
// the feature index is computed once; a block walks `rows_per_block` rows
__global__ void __launch_bounds__(SM_NT) scale_shift_kernel(float* __restrict__ x, long long rows, int F, int R,
                                                            long long rows_per_block, const float* __restrict__ scale,
                                                            const float* __restrict__ shift) {
  const int r = threadIdx.x / F, f = threadIdx.x - r * F;
  if (r >= R) return;
  const float sc = scale[f], sh = shift[f];
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  for (long long i0 = r0 + r; i0 < r1; i0 += 4ll * R) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + static_cast<long long>(u) * R;
      if (i < r1) v[u] = x[i * F + f];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + static_cast<long long>(u) * R;
      if (i < r1) x[i * F + f] = __fadd_rn(__fmul_rn(v[u], sc), sh);
    }
  }
}

// 16-byte variant (rows * F a multiple of 4, x 16-byte aligned): the feature of element 4 i + k repeats with period
// P = F / gcd(F, 4) in i; blocks of a multiple of P threads striding by a multiple of P keep every thread on ONE
// phase, so its four (scale, shift) pairs are loaded once.  Same two roundings per element: bit-identical.
__global__ void __launch_bounds__(SM_NT) scale_shift_vec_kernel(float4* __restrict__ x4, long long n4, int F, int P,
                                                                const float* __restrict__ scale,
                                                                const float* __restrict__ shift) {
  const int p = threadIdx.x % P;
  float sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int f = (4 * p + k) % F;
    sc[k] = scale[f];
    sh[k] = shift[f];
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = x4[i + u * stride];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      v[u].x = __fadd_rn(__fmul_rn(v[u].x, sc[0]), sh[0]);
      v[u].y = __fadd_rn(__fmul_rn(v[u].y, sc[1]), sh[1]);
      v[u].z = __fadd_rn(__fmul_rn(v[u].z, sc[2]), sh[2]);
      v[u].w = __fadd_rn(__fmul_rn(v[u].w, sc[3]), sh[3]);
      x4[i + u * stride] = v[u];
    }
  }
  for (; i < n4; i += stride) {
    float4 v = x4[i];
    v.x = __fadd_rn(__fmul_rn(v.x, sc[0]), sh[0]);
    v.y = __fadd_rn(__fmul_rn(v.y, sc[1]), sh[1]);
    v.z = __fadd_rn(__fmul_rn(v.z, sc[2]), sh[2]);
    v.w = __fadd_rn(__fmul_rn(v.w, sc[3]), sh[3]);
    x4[i] = v;
  }
}

}  // namespace p4c

using namespace p4c;

static int psd_splits(int H) {
  int s = (H + 63) / 64;
  return s < 1 ? 1 : (s > 16 ? 16 : s);
}

extern "C" int64_t p4c_psd_partial_floats(int64_t bt, int32_t H, int32_t W, int32_t F) {
  if (bt <= 0 || H <= 0 || W <= 0 || F <= 0) return 0;
  return 2ll * psd_splits(H) * bt * W * F;
}

extern "C" int p4c_psd_first_row(const float* x, const float* mask, const float* wy, const float* ctab, int64_t bt,
                                 int32_t H, int32_t W, int32_t F, int32_t J, float* partials, float* spec, void* stream) {
  P4C_REQUIRE(x && wy && ctab && partials && spec, "psd_first_row: null pointer");
  P4C_REQUIRE(bt > 0 && H > 0 && W > 0 && F > 0, "psd_first_row: empty problem");
  P4C_REQUIRE(J >= 1 && J <= W, "psd_first_row: J=%d spectrum entries requested from a row of %d", J, W);
  P4C_REQUIRE(bt <= 65535, "psd_first_row: more than 65535 fields per call");
  cudaStream_t st = (cudaStream_t)stream;
  const int ns = psd_splits(H);
  const int rps = (H + ns - 1) / ns;
  const long long WF = static_cast<long long>(W) * F;
  dim3 grid(static_cast<unsigned>((WF + SM_NT - 1) / SM_NT), static_cast<unsigned>(bt), static_cast<unsigned>(ns));
  psd_colsum_kernel<<<grid, SM_NT, 0, st>>>(x, mask, wy, H, WF, rps, reinterpret_cast<float2*>(partials));
  P4C_CHECK_LAUNCH("psd_colsum");
  const long long n = bt * WF;
  if (ns > 1) {
    psd_split_reduce_kernel<<<static_cast<unsigned>((n + SM_NT - 1) / SM_NT), SM_NT, 0, st>>>(
        reinterpret_cast<float2*>(partials), ns, n);
    P4C_CHECK_LAUNCH("psd_split_reduce");
  }
  P4C_REQUIRE(F <= 65535, "psd_first_row: more than 65535 features");
  const dim3 fgrid(static_cast<unsigned>((J + 1 + 31) / 32), static_cast<unsigned>(F), static_cast<unsigned>(bt));
  psd_finalize_kernel<<<fgrid, 32 * PSD_XS, 0, st>>>(reinterpret_cast<const float2*>(partials), ctab, bt, H, W, F, J, spec);
  P4C_CHECK_LAUNCH("psd_finalize");
  return P4C_OK;
}

extern "C" int p4c_psd_profile(const float* spec, const unsigned char* present, int32_t B, int32_t F, int32_t R,
                               float* prof, void* stream) {
  P4C_REQUIRE(spec && present && prof, "psd_profile: null pointer");
  P4C_REQUIRE(B > 0 && F > 0 && R > 0, "psd_profile: empty problem");
  const int n = F * R;
  psd_profile_kernel<<<(n + SM_NT - 1) / SM_NT, SM_NT, 0, (cudaStream_t)stream>>>(spec, present, B, F, R, prof);
  P4C_CHECK_LAUNCH("psd_profile");
  return P4C_OK;
}

static int stat_chunks(long long X) {
  long long c = (X + 4095) / 4096;
  return static_cast<int>(c < 1 ? 1 : (c > 1024 ? 1024 : c));
}

extern "C" int64_t p4c_feature_stats_partial_bytes(int32_t B, int64_t X, int32_t F) {
  if (B <= 0 || X <= 0 || F <= 0) return 0;
  return static_cast<int64_t>(sizeof(StatPart)) * B * stat_chunks(X) * F;
}

extern "C" int p4c_feature_stats(const float* x, const float* prev, int32_t B, int64_t X, int32_t F, void* partials,
                                 float* out, void* stream) {
  P4C_REQUIRE(x && partials && out, "feature_stats: null pointer");
  P4C_REQUIRE(B > 0 && X > 0 && F > 0 && F <= SM_NT, "feature_stats: B=%d X=%lld F=%d (F <= 256)", B, (long long)X, F);
  P4C_REQUIRE(B <= 65535, "feature_stats: more than 65535 samples per call");
  cudaStream_t st = (cudaStream_t)stream;
  const int chunks = stat_chunks(X);
  const long long rpc = (X + chunks - 1) / chunks;
  dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(B));
  feature_stats_kernel<<<grid, SM_NT, 0, st>>>(x, prev, X, F, SM_NT / F, rpc, static_cast<StatPart*>(partials));
  P4C_CHECK_LAUNCH("feature_stats");
  const long long BF = static_cast<long long>(B) * F;
  feature_stats_finalize_kernel<<<static_cast<unsigned>((BF * 32 + 127) / 128), 128, 0, st>>>(
      static_cast<const StatPart*>(partials), chunks, F, BF, reinterpret_cast<float4*>(out));
  P4C_CHECK_LAUNCH("feature_stats_finalize");
  return P4C_OK;
}

extern "C" int p4c_scale_shift(float* x, int64_t rows, int32_t F, const float* scale, const float* shift, void* stream) {
  P4C_REQUIRE(x && scale && shift, "scale_shift: null pointer");
  P4C_REQUIRE(rows > 0 && F > 0, "scale_shift: empty problem");
  P4C_REQUIRE(F <= SM_NT, "scale_shift: F=%d > 256", F);
  const long long total = rows * F;
  const int g4 = (F % 4 == 0) ? 4 : (F % 2 == 0 ? 2 : 1);
  const int P = F / g4;  // period of the feature pattern in units of 4 elements
  if (total % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && P <= SM_NT) {
    const int nt = (SM_NT / P) * P;
    const long long n4 = total / 4;
    long long blocks = (n4 + 4ll * nt - 1) / (4ll * nt);  // four vectors per thread and iteration
    const long long cap = 32ll * kNumSMs;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    scale_shift_vec_kernel<<<static_cast<unsigned>(blocks), nt, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4*>(x), n4, F,
                                                                                       P, scale, shift);
    P4C_CHECK_LAUNCH("scale_shift(vec)");
    return P4C_OK;
  }
  const int R = SM_NT / F;
  const long long rpb = 64ll * R;  // 64 iterations of R rows per block
  scale_shift_kernel<<<static_cast<unsigned>((rows + rpb - 1) / rpb), SM_NT, 0, (cudaStream_t)stream>>>(x, rows, F, R, rpb,
                                                                                                       scale, shift);
  P4C_CHECK_LAUNCH("scale_shift");
  return P4C_OK;
}
