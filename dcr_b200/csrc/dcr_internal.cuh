// Internal C++ declarations shared between the kernel translation units and the C-ABI layer (api.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace dcr {

struct SimStats {
  int cta_group;
  int grid;
  int smem_bytes;
  int stages;
  int kp;
  int cap;
  int n_flagged;
  int d_pad;
};

size_t sim_topk_workspace_size(int nq, int ng, int d, int k);
int sim_topk(const float* q, int nq, const float* g, int ng, int d, int k, long long g_index_base,
             long long g_index_stride, float* out_scores, long long* out_idx, void* ws, size_t ws_bytes,
             cudaStream_t stream, SimStats* stats);

int l2_normalize(float* x, int n, int d, float eps, cudaStream_t stream);
int topk_merge(const float* scores, const long long* idx, int nq, int nlists, int k_in, int k_out, float* out_scores,
               long long* out_idx, cudaStream_t stream);


// ---- tcgen05 GEMM / implicit-GEMM convolution (conv_gemm.cu) ---------------------------------------------------
constexpr int kMaxGemmTerms = 6;

// Y = act(scale * conv(X, W) + bias (+ R)).  X: NHWC bf16 planes [B,H,W,C] (row stride ld_in for the 1x1/linear
// case), W: prepared weights [planes][N][kh*kw*ceil64(C)] bf16 (tap-major, channels zero-padded to 64).
struct ConvGemmDesc {
  const __nv_bfloat16* in = nullptr;
  long long in_plane_stride = 0;
  int B = 0, H = 1, W = 1, C = 0, ld_in = 0;
  const __nv_bfloat16* weight = nullptr;
  long long w_plane_stride = 0;
  int N = 0, kh = 1, kw = 1, stride = 1, pad_h = 0, pad_w = 0;
  int n_terms = 1;
  int term_a[kMaxGemmTerms] = {0, 0, 0, 0, 0, 0};
  int term_w[kMaxGemmTerms] = {0, 0, 0, 0, 0, 0};
  const float* scale = nullptr;
  const float* bias = nullptr;
  const __nv_bfloat16* res = nullptr;
  int ld_res = 0, res_planes = 1;
  long long res_plane_stride = 0;
  __nv_bfloat16* out = nullptr;
  int ld_out = 0, out_col_off = 0, out_planes = 1;
  long long out_plane_stride = 0;
  float* out_f32 = nullptr;
  int ld_out_f32 = 0;
  int act = 0;        // 0 none, 1 relu, 2 gelu(erf)
  int force_bn = 0;   // 0 = auto tile width
};
int conv_gemm(const ConvGemmDesc& d, cudaStream_t stream);

}  // namespace dcr
