// Streaming FID activation statistics on the device (float64).
// Replaces   pred_arr[...] = pred  (float64 [N,2048] on the host, metrics/fid.py:118,135) followed by
//            mu = np.mean(act, axis=0); sigma = np.cov(act, rowvar=False)      metrics/fid.py:219-220
// The [N, d] activation matrix is never kept: every batch is folded into  s = sum(x - c)  and  S = (x-c)^T (x-c)
// (c = mean of the first batch, a fixed shift that removes the cancellation of the one-pass formula), and
//   mu = c + s/n,   sigma = (S - s s^T / n) / (n - 1)        (unbiased, as np.cov)
// Roofline: 2*n*d^2 fp64 FLOP (0.42 TFLOP per 50k x 2048 set) on the fp64 pipe; the upper triangle only is computed.
#include <vector>

#include "dcr_internal.cuh"
#include "host_util.cuh"

namespace dcr {

struct FidState {
  int d = 0;
  long long n = 0;
  double* shift = nullptr;   // [d]
  double* sum = nullptr;     // [d]
  double* xtx = nullptr;     // [d, d], upper-triangular tiles valid
};

namespace {
constexpr int kTile = 64;
constexpr int kRows = 32;

__global__ void fid_shift_kernel(const float* __restrict__ x, int n, int d, double* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  double s = 0.0;
  for (int r = 0; r < n; ++r) s += static_cast<double>(x[static_cast<size_t>(r) * d + c]);
  shift[c] = s / n;
}

// grid (d/64, d/64) upper triangle, 256 threads, each thread a 4x4 block of the 64x64 tile
__global__ void __launch_bounds__(256)
    fid_accumulate_kernel(const float* __restrict__ x, int n, int d, const double* __restrict__ shift,
                          double* __restrict__ sum, double* __restrict__ xtx) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj < ti) return;
  __shared__ double a[kRows][kTile + 1];
  __shared__ double b[kRows][kTile + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4] = {};
  double colsum = 0.0;   // diagonal tiles, threads 0..63: column sums
  for (int r0 = 0; r0 < n; r0 += kRows) {
    for (int i = threadIdx.x; i < kRows * kTile; i += 256) {
      const int r = i / kTile, c = i % kTile;
      const bool ok = r0 + r < n;
      const int ca = ti * kTile + c, cb = tj * kTile + c;
      a[r][c] = (ok && ca < d) ? static_cast<double>(x[static_cast<size_t>(r0 + r) * d + ca]) - shift[ca] : 0.0;
      b[r][c] = (ok && cb < d) ? static_cast<double>(x[static_cast<size_t>(r0 + r) * d + cb]) - shift[cb] : 0.0;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < kRows; ++r) {
      double av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = a[r][ty * 4 + i];
        bv[i] = b[r][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    if (ti == tj && threadIdx.x < kTile)
      for (int r = 0; r < kRows; ++r) colsum += a[r][threadIdx.x];
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = ti * kTile + ty * 4 + i, col = tj * kTile + tx * 4 + j;
      if (row < d && col < d) xtx[static_cast<size_t>(row) * d + col] += acc[i][j];   // one block owns the tile
    }
  if (ti == tj && threadIdx.x < kTile && ti * kTile + threadIdx.x < d) sum[ti * kTile + threadIdx.x] += colsum;
}
}  // namespace

int fid_create(int d, FidState** out) {
  DCR_REQUIRE(d >= 1 && d <= 16384, "fid_create: bad dim %d", d);
  if (!device_info()) return -2;
  FidState* s = new FidState();
  s->d = d;
  DCR_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->shift), sizeof(double) * d));
  DCR_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->sum), sizeof(double) * d));
  DCR_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->xtx), sizeof(double) * d * d));
  DCR_CUDA_CHECK(cudaMemset(s->sum, 0, sizeof(double) * d));
  DCR_CUDA_CHECK(cudaMemset(s->xtx, 0, sizeof(double) * d * d));
  *out = s;
  return 0;
}

void fid_destroy(FidState* s) {
  if (!s) return;
  cudaFree(s->shift);
  cudaFree(s->sum);
  cudaFree(s->xtx);
  delete s;
}

int fid_accumulate(FidState* s, const float* act, int n, cudaStream_t stream) {
  DCR_REQUIRE(s && act, "fid_accumulate: null argument");
  if (n <= 0) return 0;
  if (s->n == 0) {
    fid_shift_kernel<<<(s->d + 127) / 128, 128, 0, stream>>>(act, n, s->d, s->shift);
    count_launch();
  }
  const int t = (s->d + kTile - 1) / kTile;
  fid_accumulate_kernel<<<dim3(t, t), 256, 0, stream>>>(act, n, s->d, s->shift, s->sum, s->xtx);
  count_launch();
  DCR_CUDA_CHECK(cudaGetLastError());
  s->n += n;
  return 0;
}

int fid_finalize(FidState* s, double* mu_host, double* sigma_host, long long* n_out, cudaStream_t stream) {
  DCR_REQUIRE(s && mu_host && sigma_host, "fid_finalize: null argument");
  DCR_REQUIRE(s->n >= 2, "fid_finalize: need at least 2 samples (have %lld)", s->n);
  const int d = s->d;
  std::vector<double> shift(d), sum(d);
  DCR_CUDA_CHECK(cudaMemcpyAsync(shift.data(), s->shift, sizeof(double) * d, cudaMemcpyDeviceToHost, stream));
  DCR_CUDA_CHECK(cudaMemcpyAsync(sum.data(), s->sum, sizeof(double) * d, cudaMemcpyDeviceToHost, stream));
  DCR_CUDA_CHECK(cudaMemcpyAsync(sigma_host, s->xtx, sizeof(double) * d * d, cudaMemcpyDeviceToHost, stream));
  DCR_CUDA_CHECK(cudaStreamSynchronize(stream));
  const double n = static_cast<double>(s->n);
  for (int i = 0; i < d; ++i) mu_host[i] = shift[i] + sum[i] / n;
  for (int i = 0; i < d; ++i)
    for (int j = i; j < d; ++j) {
      const double v = (sigma_host[static_cast<size_t>(i) * d + j] - sum[i] * sum[j] / n) / (n - 1.0);
      sigma_host[static_cast<size_t>(i) * d + j] = v;
      sigma_host[static_cast<size_t>(j) * d + i] = v;
    }
  if (n_out) *n_out = s->n;
  return 0;
}

}  // namespace dcr
