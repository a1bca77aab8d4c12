// Shared helpers for the py4cast_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/py4cast_b200.h"

namespace p4c {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define P4C_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::p4c::set_error(__VA_ARGS__);      \
      return P4C_EINVAL;                  \
    }                                     \
  } while (0)

#define P4C_CHECK_LAUNCH(name)                                                       \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      ::p4c::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(_e));  \
      return P4C_ECUDA;                                                              \
    }                                                                                \
    ::p4c::count_launch();                                                           \
  } while (0)

// tensor-core engine (mlp2_tc.cu)
bool tc_forward_supported(const p4c_mlp_t* m);
int launch_mlp2_tc_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2, void* save_h1,
                           float* save_stats, void* aggr_out, int agg_group, float agg_scale, cudaStream_t st);
void tc_set_trace(long long* buf);
long long* tc_get_trace();
// warp-specialised pipelined forward for the large calls (mlp2_ws.cu)
bool ws_forward_supported(const p4c_mlp_t* m, const void* residual, const void* out2, const void* aggr_out,
                          int agg_group);
int launch_mlp2_ws_forward(const p4c_mlp_t* m, const void* residual, void* out, void* out2, void* save_h1,
                           float* save_stats, void* aggr_out, int agg_group, float agg_scale, cudaStream_t st);
// weight-gradient partial reduction of one MLP (pack.cu)
int launch_wgrad_reduce_single(const p4c_wgrad_t* e, cudaStream_t st);
bool tc_backward_supported(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g);
int tc_backward_grid(const p4c_mlp_t* m);
int launch_mlp2_tc_backward(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, cudaStream_t st);
// warp-specialised pipelined backward that recomputes h1 / LayerNorm statistics (mlp2_ws_bwd.cu)
bool ws_backward_shape_ok(const p4c_mlp_t* m);
bool ws_backward_supported(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g);
int ws_backward_grid(const p4c_mlp_t* m);
int launch_mlp2_ws_backward(const p4c_mlp_t* m, const p4c_mlp_bwd_t* g, int total_blocks, cudaStream_t st);
// P4C_ENGINE=simt forces the CUDA-core engine for every call (A/B testing); default: tensor cores
// wherever the shape allows (bf16 storage, 64-wide sources)
bool tc_engine_enabled();

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The tensor-core kernels are launched with cudaLaunchAttributeProgrammaticStreamSerialization:
// kernel N+1 may start once every CTA of kernel N has executed griddepcontrol.launch_dependents
// (placed after N's own set-up: TMEM allocated, barriers initialised), run ITS set-up (weight image
// -> shared memory, TMEM allocation, graph-index loads: constants of the step) and then blocks in
// griddepcontrol.wait until N has completed and its writes are visible.  Launch gaps and set-up
// latency of the ~1200 launches of a HiLAM step overlap the previous kernel's tail.
// P4C_PDL=0 disables the attribute (the device instructions are then no-ops).
bool pdl_enabled();
// the next PDL-capable launch must serialise fully (called after kernels whose output is read
// BEFORE griddepcontrol.wait, i.e. p4c_pack_weights)
void pdl_block_next();
bool pdl_take_attr();  // true: launch with the attribute

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_take_attr() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kNumSMs = 148;  // B200
// persistent tensor-core kernels use at most 2 * kNumSMs - reserve CTAs (p4c_set_cta_reserve)
int tc_max_ctas();
constexpr int HIDDEN_W = P4C_HIDDEN;

// ---- storage <-> fp32 conversion -------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 4 consecutive elements (16-byte aligned for float, 8-byte for bf16)
template <typename T>
__device__ __forceinline__ float4 load4(const T* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  uint2 r = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&r.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&r.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, float4 v);
template <>
__device__ __forceinline__ void store4<float>(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&a);
  r.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }
// d/dx silu(x) = s + x*s*(1-s), s = sigmoid(x)
__device__ __forceinline__ float silu_grad(float x) {
  float s = 1.0f / (1.0f + expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float4 f4add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

}  // namespace p4c
