// Microbenchmark: how long does one tcgen05.mma (kind::f16, bf16 operands in shared memory, cta_group::1, M = 128)
// take as a function of N, issued back to back by ONE thread (and by two threads of different warps on separate
// accumulators)?  No loads: the operands are whatever shared memory holds.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dcr_b200/csrc -o /tmp/umma_rate tools/microbench/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace dcr;

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int issuers, int mode, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint64_t ring[8];    // mode 1/3: a commit per 4 MMAs lands here; mode 4: a helper thread bounces it back
  __shared__ uint32_t tmem_slot;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); for (int i = 0; i < 8; ++i) mbar_init(&ring[i], 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<1>(&tmem_slot, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (lane == 0 && warp < static_cast<uint32_t>(issuers)) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t da = umma_desc_sw128(smem_u32(smem));
    const uint64_t db = umma_desc_sw128(smem_u32(smem + 16384));
    const uint32_t tmem_d = tmem_base + warp * 256 % 512;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (mode == 2 || mode == 3) tc_fence_after();
      if (mode == 5) {   // wait for the commit of 8 groups ago (what a stage ring does), then fence
        if (i >= 8) mbar_wait(&ring[i & 7], ((i >> 3) - 1) & 1);
        tc_fence_after();
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16<1>(tmem_d, da + 2 * k, db + 2 * k, idesc, 1);
      if (mode == 1 || mode == 3 || mode == 5) umma_commit<1>(&ring[i & 7]);
    }
    umma_commit<1>(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[warp] = static_cast<unsigned long long>(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_base, 512);
}

// A operand from tensor memory (".ts" form): D[tmem] += A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts_cg1(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  const uint32_t z = 0, one = 1;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(one), "r"(z)
      : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) rate_ts_kernel(int iters, unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f803f80u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc<1>(&tmem_slot, 512); tmem_relinquish<1>(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t db = umma_desc_sw128(smem_u32(smem));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts_cg1(tmem_base + 256, tmem_base + k * 8, db + 2 * k, idesc);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit<1>(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[0] = static_cast<unsigned long long>(t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmem_base, 512);
}

template <int N>
void run_ts() {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int iters = 2000;
  const size_t smem = 1024 + N * 128;
  cudaFuncSetAttribute(rate_ts_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_ts_kernel<N><<<148, 128, smem>>>(iters, d);
  cudaEventRecord(e0);
  rate_ts_kernel<N><<<148, 128, smem>>>(iters, d);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  unsigned long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = 4.0 * iters;
  printf("A-in-TMEM N=%3d: %s  cycles/MMA = %.1f  kernel %.3f ms  %.0f TFLOP/s\n", N, cudaGetErrorString(err), h[0] / mmas, ms,
         2.0 * 128 * N * 16 * mmas * 148 / (ms * 1e-3) / 1e12);
  cudaFree(d);
}

template <int N>
void run(int issuers, int mode = 0) {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  const int iters = 2000;
  const size_t smem = 1024 + 16384 + N * 128;
  cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  rate_kernel<N><<<148, 128, smem>>>(iters, issuers, mode, d);
  cudaEventRecord(e0);
  rate_kernel<N><<<148, 128, smem>>>(iters, issuers, mode, d);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  unsigned long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double mmas = 4.0 * iters;
  const double flops = 2.0 * 128 * N * 16 * mmas * issuers * 148;
  printf("mode=%d N=%3d issuers=%d: %s  cycles/MMA (thread 0) = %.1f  (thread 1) = %.1f  kernel %.3f ms  %.0f TFLOP/s  SM clock ~%.0f MHz\n", mode, N, issuers,
         cudaGetErrorString(err), h[0] / mmas, h[1] / mmas, ms, flops / (ms * 1e-3) / 1e12, h[0] / (ms * 1e3));
  cudaFree(d);
}

int main() {
  // mode 0: MMAs only; 1: + tcgen05.commit per 4 MMAs; 2: + tcgen05.fence::after_thread_sync per 4 MMAs; 3: both;
  // 5: commit per 4 MMAs and wait (mbarrier try_wait) for the commit of 8 groups earlier before each group
  for (int mode : {0, 1, 2, 3, 5}) { run<64>(1, mode); run<256>(1, mode); }
  run<64>(2); run<256>(2);
  run_ts<64>(); run_ts<128>(); run_ts<256>();
  return 0;
}
