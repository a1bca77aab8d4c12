// Micro-benchmark: cost of one tcgen05.mma (cta_group::1, kind::f16, bf16 x bf16 -> f32, M = 128, K = 16) as a
// function of N and of where / how the operands are read: A, B K-major from shared memory (SS), MN-major
// ("transposed") operands, A from tensor memory (TS).  One CTA per SM, one warp; REPS MMAs are issued back to
// back into the same accumulator, one commit, clock64 around issue and around completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_cost scripts/mma_cost.cu && ./mma_cost
// Feeds DESIGN.md 3.1 (what bounds the fused d = 64 MLP kernels).
#include <cstdio>
#include <cuda_runtime.h>
#include "../py4cast_b200/csrc/common.cuh"
#include "../py4cast_b200/csrc/tc_common.cuh"

using namespace p4c::tc;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: SS K-major A, K-major B; 1: SS MN-major A and B; 2: TS (A in TMEM), K-major B; 3: SS K-major A, MN-major B
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) mma_cost_kernel(long long* out, int reps) {
  extern __shared__ unsigned char smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  // A: 128 x 64 bf16 tile (16 KB), B: N x 64 bf16 (N / 64 tiles of 8 KB .. as 128-row blocks of 16 KB)
  for (int i = threadIdx.x; i < (16384 + 4 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (threadIdx.x < 32) tmem_alloc(&tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  const uint32_t a = smem_u32(base), b = smem_u32(base + 16384);
  constexpr uint32_t idesc = make_idesc_bf16(128, N, MODE == 1, MODE == 1 || MODE == 3);
  if (threadIdx.x == 0) {
    for (int rep = 0; rep < 2; ++rep) {  // second pass is the measured one
      const long long t0 = clock64();
      for (int i = 0; i < reps; ++i) {
        const int q = i & 3;
        uint64_t da, db;
        if (MODE == 1) {
          da = make_desc_sw128(a + q * 2048, 16384, 1024);
          db = make_desc_sw128(b + q * 2048, 16384, 1024);
        } else {
          da = make_desc_sw128(a + q * 32, 16, 1024);
          db = MODE == 3 ? make_desc_sw128(b + q * 2048, 16384, 1024) : make_desc_sw128(b + q * 32, 16, 1024);
        }
        if (MODE == 2) mma_ts(tmem, tmem + 256 + q * 8, db, idesc, i ? 1u : 0u);
        else mma_bf16_ss(tmem, da, db, idesc, i ? 1u : 0u);
      }
      const long long t1 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      const long long t2 = clock64();
      if (rep == 1 && blockIdx.x == 0) {
        out[0] = t1 - t0;
        out[1] = t2 - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(const char* name, long long* d_out, int ctas) {
  const int reps = 256;
  const size_t smem = 16384 + 4 * 16384 + 1024;
  cudaFuncSetAttribute(mma_cost_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  mma_cost_kernel<N, MODE><<<ctas, 128, smem>>>(d_out, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2] = {0, 0};
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-44s N=%3d  issue %7.1f cycles/MMA   complete %7.1f cycles/MMA   (%s)\n", name, N, double(h[0]) / reps,
         double(h[1]) / reps, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int ctas : {1, 148}) {
    printf("---- %d CTA(s), 256 MMAs (M = 128, K = 16) back to back into one accumulator\n", ctas);
    run<64, 0>("SS  A K-major, B K-major", d_out, ctas);
    run<128, 0>("SS  A K-major, B K-major", d_out, ctas);
    run<256, 0>("SS  A K-major, B K-major", d_out, ctas);
    run<64, 3>("SS  A K-major, B MN-major", d_out, ctas);
    run<64, 1>("SS  A MN-major, B MN-major", d_out, ctas);
    run<128, 1>("SS  A MN-major, B MN-major", d_out, ctas);
    run<64, 2>("TS  A in TMEM, B K-major", d_out, ctas);
    run<128, 2>("TS  A in TMEM, B K-major", d_out, ctas);
  }
  return 0;
}
