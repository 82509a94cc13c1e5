// Readout head + loss of the training step as two launches: the last Linear of the readout MLP (out_features = 1:
// mhnn.py:216-218 / equihnn_egnn.py:167-169 with num_target = 1), the MSE loss over the first B rows (main.py:49-51; the
// ghost molecule of a padded batch is row B) and their backward.  In the captured step this tail was 14 ATen / cuBLAS
// launches of ~2 us each on the critical path (gemv, mse, mean, fills, mse_backward, gemv, two reductions, gradient
// adds); SURVEY.md 8(f) row 2.
//
//   forward : pred[r] = z[r,:] . w + b ;  resid[r] = 2 (pred[r] - y[r]) / B for r < B, else 0 ;  loss = mean_{r<B} (pred - y)^2
//   backward: dz[r,:] = g resid[r] w ;  dw (+)= g sum_r resid[r] z[r,:] ;  db (+)= g sum_r resid[r]     (g = d loss, device scalar)
// One CTA each, fixed summation orders (deterministic).  R rows, K <= 1024 columns.
#include "common.cuh"

namespace eqh {

constexpr int kHeadThreads = 1024;

__global__ void __launch_bounds__(kHeadThreads) k_head_mse_fwd(const float* __restrict__ z, const float* __restrict__ w,
                                                              const float* __restrict__ b, const float* __restrict__ y, int R,
                                                              int K, int B, float* __restrict__ pred, float* __restrict__ resid,
                                                              float* __restrict__ loss) {
  pdl_prologue();
  __shared__ float part[kHeadThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float bias = b ? __ldg(b) : 0.f;
  const float invB = 1.f / (float)max(B, 1);
  float sq = 0.f;                                      // this warp's rows, ascending
  for (int r = warp; r < R; r += kHeadThreads / 32) {
    float d = 0.f;
    for (int k = lane; k < K; k += 32) d += __ldg(z + (int64_t)r * K + k) * __ldg(w + k);
    d = warp_sum(d) + bias;
    float e = 0.f;
    if (r < B) e = d - __ldg(y + r);
    if (lane == 0) {
      pred[r] = d;
      resid[r] = 2.f * e * invB;
    }
    sq += e * e;
  }
  if (lane == 0) part[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kHeadThreads / 32; ++i) t += part[i];
    *loss = t * invB;
  }
}

__global__ void __launch_bounds__(kHeadThreads) k_head_mse_bwd(const float* __restrict__ z, const float* __restrict__ w,
                                                              const float* __restrict__ resid, const float* __restrict__ g,
                                                              int R, int K, float* __restrict__ dz, float* __restrict__ dw,
                                                              float* __restrict__ db, int accumulate) {
  pdl_prologue();
  __shared__ float part[kHeadThreads];
  const float gs = g ? __ldg(g) : 1.f;
  const int groups = kHeadThreads / K;                 // row groups: thread (rg, k) sums rows rg, rg + groups, ...
  const int k = threadIdx.x % K, rg = threadIdx.x / K;
  float acc = 0.f, accb = 0.f;
  if (rg < groups) {
    const float wk = __ldg(w + k);
#pragma unroll 8
    for (int r = rg; r < R; r += groups) {
      const float c = gs * __ldg(resid + r);
      dz[(int64_t)r * K + k] = c * wk;
      acc += c * __ldg(z + (int64_t)r * K + k);
      accb += c;
    }
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (rg == 0) {
    float t = 0.f;
    for (int i = 0; i < groups; ++i) t += part[i * K + k];
    dw[k] = accumulate ? dw[k] + t : t;
  }
  __syncthreads();
  if (k == 0 && rg < groups) part[rg] = accb;
  __syncthreads();
  if (threadIdx.x == 0 && db) {
    float t = 0.f;
    for (int i = 0; i < groups; ++i) t += part[i];
    *db = accumulate ? *db + t : t;
  }
}

}  // namespace eqh

extern "C" {

int eqh_head_mse_fwd(const float* z, const float* w, const float* b, const float* y, int64_t R, int64_t K, int64_t B,
                     float* pred, float* resid, float* loss, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(z && w && y && pred && resid && loss && R >= 1 && B >= 1 && B <= R);
  if (K < 1 || K > kHeadThreads || R > (1 << 24)) return EQH_ERR_UNSUPPORTED;
  launch_kernel(k_head_mse_fwd, dim3(1), dim3(kHeadThreads), 0, static_cast<cudaStream_t>(stream), z, w, b, y, (int)R, (int)K, (int)B,
                pred, resid, loss);
  EQH_LAUNCH_RC();
}

int eqh_head_mse_bwd(const float* z, const float* w, const float* resid, const float* g, int64_t R, int64_t K, float* dz,
                     float* dw, float* db, int accumulate, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(z && w && resid && dz && dw && R >= 1);
  if (K < 1 || K > kHeadThreads || R > (1 << 24)) return EQH_ERR_UNSUPPORTED;
  launch_kernel(k_head_mse_bwd, dim3(1), dim3(kHeadThreads), 0, static_cast<cudaStream_t>(stream), z, w, resid, g, (int)R, (int)K, dz,
                dw, db, accumulate);
  EQH_LAUNCH_RC();
}

}  // extern "C"
