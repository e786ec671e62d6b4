// Probe: can a tcgen05 shared-memory descriptor (K-major, SWIZZLE_NONE) describe an OVERLAPPING-window (Toeplitz) A
// operand?   A[m][k] = X[(m + k/8) * 8 + k % 8]   -- row m starts 16 bytes after row m-1, K-chunk j starts 16 bytes
// after chunk j-1: leading-dimension byte offset = 16, stride-dimension byte offset = 128 (8 rows x 16 B).
// This is the access pattern of a stride-1 convolution along one image row when the input is stored as 16-byte pixels:
// used by the fused ResNet stem (csrc/stem_fused.cu).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -I dcr_b200/csrc
// tools/microbench/toeplitz_probe.cu -o tools/microbench/toeplitz_probe ; prints the max abs error vs the CPU.
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx.cuh"
using namespace dcr;

__device__ uint64_t desc_none(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;   // layout type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* x, const __nv_bfloat16* b, float* out, int ksteps, int shift_units) {
  __shared__ __align__(1024) uint8_t sx[8192];   // X units (16 B each)
  __shared__ __align__(1024) uint8_t sbm[4096 * 4];  // B: [kstep][chunk 2][64 rows][16 B]
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 8192 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(sx)[i] = x[i];
  for (int i = threadIdx.x; i < 4096 * 4 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(sbm)[i] = b[i];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc<1>(&slot, 64); tmem_relinquish<1>(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int ks = 0; ks < ksteps; ++ks) {
      // A: K-step ks covers X units (m + shift + 2*ks) and (m + shift + 2*ks + 1)
      const uint64_t da = desc_none(smem_u32(sx) + (shift_units + 2 * ks) * 16, 16, 128);
      const uint64_t db = desc_none(smem_u32(sbm) + ks * 2048, 1024, 128);
      umma_f16<1>(tmem, da, db, idesc, ks != 0);
    }
    umma_commit<1>(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((warp * 32u) << 16) + c * 32, r);
    tmem_ld_wait_regs(r);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<1>(tmem, 64);
}

int main() {
  const int ksteps = 4, shift = 5;
  std::vector<__nv_bfloat16> hx(4096), hb(4096 * 4 / 2);
  std::vector<float> fx(4096), fb(64 * 16 * ksteps);
  srand(1);
  for (int i = 0; i < 4096; ++i) { float v = (rand() % 17 - 8) / 8.f; hx[i] = __float2bfloat16(v); fx[i] = __bfloat162float(hx[i]); }
  for (auto& v : hb) v = __float2bfloat16(0.f);
  for (int ks = 0; ks < ksteps; ++ks)
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 16; ++k) {
        float v = (rand() % 13 - 6) / 4.f;
        fb[(ks * 64 + n) * 16 + k] = v;
        // canonical K-major no-swizzle: chunk kc = k/8 at kc*1024 bytes, row n at n*16 bytes, element (k%8)*2
        hb[(ks * 2048 + (k / 8) * 1024 + n * 16) / 2 + (k % 8)] = __float2bfloat16(v);
      }
  __nv_bfloat16 *dx, *db; float* dout;
  cudaMalloc(&dx, 8192); cudaMalloc(&db, 4096 * 4); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dx, hx.data(), 8192, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), 4096 * 4, cudaMemcpyHostToDevice);
  probe<<<1, 128>>>(dx, db, dout, ksteps, shift);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> out(128 * 64);
  cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double acc = 0;
      for (int ks = 0; ks < ksteps; ++ks)
        for (int k = 0; k < 16; ++k) acc += fx[(m + shift + 2 * ks + k / 8) * 8 + k % 8] * fb[(ks * 64 + n) * 16 + k];
      maxerr = fmax(maxerr, fabs(acc - out[m * 64 + n]));
    }
  printf("toeplitz probe: max abs err %.3e (%s)\n", maxerr, maxerr < 1e-3 ? "OVERLAPPING WINDOWS WORK" : "MISMATCH");
  return 0;
}
