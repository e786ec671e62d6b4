// Internal C++ declarations shared between the kernel translation units and the C-ABI layer (api.cu).
#pragma once
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

}  // namespace dcr
