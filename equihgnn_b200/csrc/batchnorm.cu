// BatchNorm1d over atom rows in training mode (mhnn.py:182,206: `x = self.batch_norms[i](x)` after every MHNNConv of
// mhnnm / egnn_equihnnm), forward and backward, as two kernels each instead of ~8 ATen launches + their autograd.
//
// Statistics are taken over the first n_stat rows only: a ghost-padded batch (graphed.pad_batch) appends its padding atoms
// after the real ones, and they must not enter the batch statistics (the real rows then get exactly the unpadded
// result); all `rows` rows are normalised.
//
//   pass 1  k_bn_partial   : P CTAs, each a contiguous chunk of rows; thread = column (coalesced rows); per-column Welford
//                            (count, mean, M2) of the chunk -> partial[P][3][C]
//   pass 2  k_bn_apply     : every CTA combines the P partials of its columns in chunk order (Chan's formula: fixed order,
//                            deterministic, every CTA gets the same bits), normalises its rows; CTA 0 also stores
//                            mean / rstd for the backward and updates the running statistics (momentum, UNBIASED variance,
//                            as nn.BatchNorm1d).
// Backward: dx = gamma * rstd * (dy - mean_r(dy) - xhat * mean_r(dy * xhat)) with the means over the n_stat rows;
//   k_bn_bwd_partial: per-chunk column sums of dy and dy * xhat; k_bn_bwd_apply: combines them in chunk order, writes dx and
//   (CTA 0) dgamma / dbeta, optionally accumulating into existing gradient buffers.
#include "common.cuh"

namespace eqh {
namespace bn {

constexpr int kChunks = 64;     // partial rows (chunks of the statistic rows)

// n_stat may live in device memory (n_stat_dev: the number of real atoms of a ghost-padded batch = ptr[real graphs], which
// changes from replay to replay of a captured graph); clamped to [0, rows]
__device__ __forceinline__ int64_t stat_rows(int64_t n_stat, const long long* n_stat_dev, int64_t rows) {
  int64_t n = n_stat_dev ? (int64_t)__ldg(n_stat_dev) : n_stat;
  return n < 0 ? 0 : (n > rows ? rows : n);
}

__global__ void __launch_bounds__(256) k_bn_partial(const float* __restrict__ x, int64_t n_stat_in, int C, int64_t ld,
                                                     float* __restrict__ partial, const long long* __restrict__ n_stat_dev,
                                                     int64_t rows) {
  pdl_prologue();
  const int64_t n_stat = stat_rows(n_stat_in, n_stat_dev, rows);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int64_t per = (n_stat + kChunks - 1) / kChunks;
  const int64_t r0 = blockIdx.x * per, r1 = min(n_stat, r0 + per);
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float v = __ldg(x + r * ld + c);
    cnt += 1.f;
    const float d = v - mean;
    mean += d / cnt;
    m2 += d * (v - mean);
  }
  float* p = partial + (int64_t)blockIdx.x * 3 * C;
  p[c] = cnt; p[C + c] = mean; p[2 * C + c] = m2;
}

__device__ __forceinline__ void combine(const float* __restrict__ partial, int C, int c, float& cnt, float& mean, float& m2) {
  cnt = 0.f; mean = 0.f; m2 = 0.f;
  for (int k = 0; k < kChunks; ++k) {
    const float* p = partial + (int64_t)k * 3 * C;
    const float nb = __ldg(p + c);
    if (nb == 0.f) continue;
    const float mb = __ldg(p + C + c), qb = __ldg(p + 2 * C + c);
    const float n = cnt + nb, d = mb - mean;
    mean += d * (nb / n);
    m2 += qb + d * d * (cnt * nb / n);
    cnt = n;
  }
}

__global__ void __launch_bounds__(256) k_bn_apply(const float* __restrict__ x, const float* __restrict__ partial,
                                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                   float momentum, int64_t rows, int C, int64_t ld, float* __restrict__ y,
                                                   float* __restrict__ save_mean, float* __restrict__ save_rstd,
                                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                                   long long* __restrict__ num_batches_tracked) {
  pdl_prologue();
  if (num_batches_tracked && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *num_batches_tracked += 1;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float cnt, mean, m2;
  combine(partial, C, c, cnt, mean, m2);
  const float var = cnt > 0.f ? m2 / cnt : 0.f;                // biased: what BatchNorm normalises with
  const float rstd = rsqrtf(var + eps);
  const float g = gamma ? __ldg(gamma + c) : 1.f, b = beta ? __ldg(beta + c) : 0.f;
  if (blockIdx.x == 0) {
    save_mean[c] = mean;
    save_rstd[c] = rstd;
    if (running_mean) {
      const float unb = cnt > 1.f ? m2 / (cnt - 1.f) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unb;
    }
  }
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  for (int64_t r = r0; r < r1; ++r) y[r * ld + c] = (__ldg(x + r * ld + c) - mean) * rstd * g + b;
}

__global__ void __launch_bounds__(256) k_bn_bwd_partial(const float* __restrict__ x, const float* __restrict__ dy,
                                                         const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                         int64_t n_stat_in, int C, int64_t ld, float* __restrict__ partial,
                                                         const long long* __restrict__ n_stat_dev, int64_t rows) {
  pdl_prologue();
  const int64_t n_stat = stat_rows(n_stat_in, n_stat_dev, rows);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c);
  const int64_t per = (n_stat + kChunks - 1) / kChunks;
  const int64_t r0 = blockIdx.x * per, r1 = min(n_stat, r0 + per);
  float s1 = 0.f, s2 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const float g = __ldg(dy + r * ld + c);
    s1 += g;
    s2 += g * ((__ldg(x + r * ld + c) - mean) * rstd);
  }
  float* p = partial + (int64_t)blockIdx.x * 2 * C;
  p[c] = s1; p[C + c] = s2;
}

__global__ void __launch_bounds__(256) k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy,
                                                       const float* __restrict__ partial, const float* __restrict__ gamma,
                                                       const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                                                       int64_t rows, int64_t n_stat_in, int C, int64_t ld, float* __restrict__ dx,
                                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate,
                                                       const long long* __restrict__ n_stat_dev) {
  pdl_prologue();
  const int64_t n_stat = stat_rows(n_stat_in, n_stat_dev, rows);
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s1 = 0.f, s2 = 0.f;
  for (int k = 0; k < kChunks; ++k) {                          // chunk order: deterministic, same bits in every CTA
    s1 += __ldg(partial + (int64_t)k * 2 * C + c);
    s2 += __ldg(partial + (int64_t)k * 2 * C + C + c);
  }
  if (blockIdx.x == 0) {
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + s2 : s2;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + s1 : s1;
  }
  const float mean = __ldg(save_mean + c), rstd = __ldg(save_rstd + c), g = gamma ? __ldg(gamma + c) : 1.f;
  const float inv = n_stat > 0 ? 1.f / (float)n_stat : 0.f;
  const float m1 = s1 * inv, m2 = s2 * inv;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  for (int64_t r = r0; r < r1; ++r) {
    const float xh = (__ldg(x + r * ld + c) - mean) * rstd;
    const float gy = __ldg(dy + r * ld + c);
    // rows past n_stat are not part of the statistics: they only see the affine map
    dx[r * ld + c] = r < n_stat ? g * rstd * (gy - m1 - xh * m2) : g * rstd * gy;
  }
}

}  // namespace bn
}  // namespace eqh

extern "C" {

int eqh_batchnorm_partial_floats(int64_t C) { return (int)(eqh::bn::kChunks * 3 * C); }

int eqh_batchnorm_fwd(const float* x, int64_t rows, int64_t n_stat, int64_t C, const float* gamma, const float* beta, float eps,
                      float momentum, float* y, float* save_mean, float* save_rstd, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, const int64_t* n_stat_dev, float* partial, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::bn;
  EQH_CHECK_ARG(rows >= 0 && n_stat >= 0 && n_stat <= rows && C > 0 && save_mean && save_rstd && partial &&
                (rows == 0 || (x && y)));
  EQH_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned cb = (unsigned)((C + 255) / 256);
  const int threads = C >= 256 ? 256 : (int)((C + 31) / 32 * 32);
  launch_kernel(k_bn_partial, dim3(kChunks, cb), dim3(threads), 0, st, x, n_stat, (int)C, C, partial,
                reinterpret_cast<const long long*>(n_stat_dev), rows);
  int64_t rb = (rows + 63) / 64;
  if (rb > 2 * kSMs) rb = 2 * kSMs;
  if (rb < 1) rb = 1;
  launch_kernel(k_bn_apply, dim3((unsigned)rb, cb), dim3(threads), 0, st, x, partial, gamma, beta, eps, momentum, rows, (int)C, C,
                y, save_mean, save_rstd, running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked));
  EQH_LAUNCH_RC();
}

int eqh_batchnorm_bwd(const float* x, const float* dy, int64_t rows, int64_t n_stat, int64_t C, const float* gamma,
                      const float* save_mean, const float* save_rstd, float* dx, float* dgamma, float* dbeta, int accumulate,
                      const int64_t* n_stat_dev, float* partial, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::bn;
  EQH_CHECK_ARG(rows >= 0 && n_stat >= 0 && n_stat <= rows && C > 0 && save_mean && save_rstd && partial &&
                (rows == 0 || (x && dy && dx)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned cb = (unsigned)((C + 255) / 256);
  const int threads = C >= 256 ? 256 : (int)((C + 31) / 32 * 32);
  launch_kernel(k_bn_bwd_partial, dim3(kChunks, cb), dim3(threads), 0, st, x, dy, save_mean, save_rstd, n_stat, (int)C, C,
                partial, reinterpret_cast<const long long*>(n_stat_dev), rows);
  int64_t rb = (rows + 63) / 64;
  if (rb > 2 * kSMs) rb = 2 * kSMs;
  if (rb < 1) rb = 1;
  launch_kernel(k_bn_bwd_apply, dim3((unsigned)rb, cb), dim3(threads), 0, st, x, dy, partial, gamma, save_mean, save_rstd, rows,
                n_stat, (int)C, C, dx, dgamma, dbeta, accumulate, reinterpret_cast<const long long*>(n_stat_dev));
  EQH_LAUNCH_RC();
}

}  // extern "C"
