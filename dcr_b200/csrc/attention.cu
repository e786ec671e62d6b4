// Multi-head self-attention of the DINO ViT blocks:  softmax(q k^T * scale) v   per (image, head).
// Reference: dino_vits.py:117-128 (Attention.forward): qkv Linear output reshaped to [B, N, 3, heads, dh];
// attn = (q @ k^T) * scale; softmax(dim=-1); x = attn @ v; heads concatenated along the channel dim.
//
// Input  qkv : bf16 planes [B*T, 3*heads*dh] (columns [q | k | v], each heads x dh) -- the fused qkv GEMM output.
// Output out : bf16 planes [B*T, heads*dh].
//
// attention_fp32_kernel: exact-fp32 path (used by parity mode, and by fast mode until the tcgen05 kernel below
// is enabled): K and V of one (image, head) live in shared memory as fp32, one warp per query row.
#include <cuda_bf16.h>

#include "dcr_internal.cuh"
#include "host_util.cuh"

namespace dcr {
namespace {

constexpr uint32_t kFull = 0xffffffffu;

struct AttnParams {
  const __nv_bfloat16* qkv;
  long long qkv_plane_stride;
  __nv_bfloat16* out;
  long long out_plane_stride;
  int planes, B, T, heads, dh;
  float scale;
};

__device__ __forceinline__ float load1(const __nv_bfloat16* base, long long ps, int planes, size_t idx) {
  float v = 0.f;
  for (int p = 0; p < planes; ++p) v += __bfloat162float(base[p * ps + idx]);
  return v;
}

// dh == 64 only (ViT-S/B).  grid = B*heads, block = 256 (8 warps).
__global__ void __launch_bounds__(256) attention_fp32_kernel(const AttnParams p) {
  extern __shared__ __align__(16) float sm[];
  const int T = p.T;
  float* ks = sm;                       // [T][65]
  float* vs = ks + T * 65;              // [T][64]
  float* qs = vs + T * 64;              // [8 warps][64]
  float* ps = qs + 8 * 64;              // [8 warps][Tpad]
  const int Tpad = (T + 31) / 32 * 32;
  const int b = blockIdx.x / p.heads, h = blockIdx.x % p.heads;
  const int ld = 3 * p.heads * 64;
  const size_t row0 = static_cast<size_t>(b) * T;
  for (int i = threadIdx.x; i < T * 64; i += blockDim.x) {
    const int t = i >> 6, d = i & 63;
    ks[t * 65 + d] = load1(p.qkv, p.qkv_plane_stride, p.planes, (row0 + t) * ld + p.heads * 64 + h * 64 + d);
    vs[t * 64 + d] = load1(p.qkv, p.qkv_plane_stride, p.planes, (row0 + t) * ld + 2 * p.heads * 64 + h * 64 + d);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = qs + warp * 64;
  float* pr = ps + warp * Tpad;
  for (int t = warp; t < T; t += 8) {
    q[lane] = load1(p.qkv, p.qkv_plane_stride, p.planes, (row0 + t) * ld + h * 64 + lane);
    q[lane + 32] = load1(p.qkv, p.qkv_plane_stride, p.planes, (row0 + t) * ld + h * 64 + lane + 32);
    __syncwarp();
    float mx = -INFINITY;
    for (int j = lane; j < T; j += 32) {
      float s = 0.f;
#pragma unroll 16
      for (int d = 0; d < 64; ++d) s = fmaf(q[d], ks[j * 65 + d], s);
      s *= p.scale;
      pr[j] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, off));
    float sum = 0.f;
    for (int j = lane; j < T; j += 32) {
      const float e = expf(pr[j] - mx);
      pr[j] = e;
      sum += e;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(kFull, sum, off);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < T; ++j) {
      const float w = pr[j];
      o0 = fmaf(w, vs[j * 64 + lane], o0);
      o1 = fmaf(w, vs[j * 64 + lane + 32], o1);
    }
    o0 /= sum;
    o1 /= sum;
    const size_t oidx = (row0 + t) * (p.heads * 64) + h * 64;
    for (int pl = 0; pl < p.planes; ++pl) {
      const __nv_bfloat16 a = __float2bfloat16_rn(o0), c = __float2bfloat16_rn(o1);
      p.out[pl * p.out_plane_stride + oidx + lane] = a;
      p.out[pl * p.out_plane_stride + oidx + lane + 32] = c;
      o0 -= __bfloat162float(a);
      o1 -= __bfloat162float(c);
    }
    __syncwarp();
  }
}

}  // namespace

int attention(const __nv_bfloat16* qkv, long long qkv_plane_stride, __nv_bfloat16* out, long long out_plane_stride,
              int planes, int B, int T, int heads, int dh, float scale, cudaStream_t stream) {
  DCR_REQUIRE(dh == 64, "attention: head dim %d not supported (64 only)", dh);
  DCR_REQUIRE(T >= 1 && T <= 1024, "attention: sequence length %d out of range", T);
  if (B == 0) return 0;
  AttnParams p;
  p.qkv = qkv; p.qkv_plane_stride = qkv_plane_stride; p.out = out; p.out_plane_stride = out_plane_stride;
  p.planes = planes; p.B = B; p.T = T; p.heads = heads; p.dh = dh; p.scale = scale;
  const int Tpad = (T + 31) / 32 * 32;
  const size_t smem = (static_cast<size_t>(T) * 65 + static_cast<size_t>(T) * 64 + 8 * 64 + 8 * Tpad) * 4;
  DCR_CUDA_CHECK(cudaFuncSetAttribute(attention_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  attention_fp32_kernel<<<B * heads, 256, smem, stream>>>(p);
  count_launch();
  DCR_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace dcr
