// Flat fused Adam over one contiguous fp32 buffer (torch.optim.Adam semantics, main.py:137-140).
// Pure streaming kernel: 4 loads + 3 stores of 4 B per element, HBM-bound.
#include "common.cuh"

namespace eqh {

// one element of the update (the same expression order in the scalar and the float4 loops: bit-identical results)
__device__ __forceinline__ void adam_elem(float& pi, float g, float& mi, float& vi, float lr, float b1, float b2, float eps, float wd,
                                          float bc1, float bc2_sqrt, float gscale) {
  float gi = g * gscale;
  if (wd != 0.f) gi += wd * pi;
  mi = b1 * mi + (1.f - b1) * gi;
  vi = b2 * vi + (1.f - b2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi = pi - (lr / bc1) * (mi / denom);
}

// the whole buffer as float4 lanes when it is 16-byte aligned (flat parameter buffers are 128-byte aligned and every
// bucket boundary is a multiple of 32 floats): 4 loads + 3 stores of 16 B per 4 elements, two groups in flight per thread
__device__ __forceinline__ void adam_body(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                          float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                                          float bc1, float bc2_sqrt, float gscale) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = tid; i < n4; i += 2 * stride) {
      const int64_t j = i + stride;
      const bool two = j < n4;
      float4 pa = p4[i], ga = g4[i], ma = m4[i], va = v4[i];
      float4 pb = pa, gb = ga, mb = ma, vb = va;
      if (two) { pb = p4[j]; gb = g4[j]; mb = m4[j]; vb = v4[j]; }
      adam_elem(pa.x, ga.x, ma.x, va.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adam_elem(pa.y, ga.y, ma.y, va.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adam_elem(pa.z, ga.z, ma.z, va.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adam_elem(pa.w, ga.w, ma.w, va.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      m4[i] = ma; v4[i] = va; p4[i] = pa;
      if (two) {
        adam_elem(pb.x, gb.x, mb.x, vb.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
        adam_elem(pb.y, gb.y, mb.y, vb.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
        adam_elem(pb.z, gb.z, mb.z, vb.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
        adam_elem(pb.w, gb.w, mb.w, vb.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
        m4[j] = mb; v4[j] = vb; p4[j] = pb;
      }
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_elem(pi, g[i], mi, vi, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
    m[i] = mi; v[i] = vi; p[i] = pi;
  }
}

__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                               float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                               float wd, float bc1, float bc2_sqrt, float gscale) {
  pdl_prologue();
  adam_body(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
}

// Same update with the step count in DEVICE memory, so that the optimizer can live inside a captured CUDA graph
// (a host-side `step` would be frozen into the graph's kernel parameters).  Every block reads t = *step + 1 and derives
// the bias corrections in double, as torch.optim.Adam does on the host; with `advance` the block that finishes last
// (integer arrival counter, left zero) stores t back.  Several launches of one optimizer step (one per gradient bucket)
// pass advance = 0 except the last one.
__global__ void __launch_bounds__(256) k_adam_dev(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                                                   float wd, long long* __restrict__ step, int advance,
                                                   int32_t* __restrict__ counter, float gscale) {
  pdl_prologue();
  __shared__ float s_bc[2];
  const long long t = *reinterpret_cast<volatile long long*>(step) + 1;
  if (threadIdx.x == 0) {
    s_bc[0] = (float)(1.0 - pow((double)b1, (double)t));
    s_bc[1] = (float)sqrt(1.0 - pow((double)b2, (double)t));
  }
  __syncthreads();
  const float bc1 = s_bc[0], bc2_sqrt = s_bc[1];
  adam_body(p, g, m, v, n, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
  if (advance) {
    __syncthreads();                       // every thread of this block has read *step
    if (threadIdx.x == 0) {
      __threadfence();
      const int prev = atomicAdd(counter, 1);
      if (prev == (int)gridDim.x - 1) {    // all blocks have read the old value: publish the new one
        *counter = 0;
        *step = t;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_rank1(float* __restrict__ out, int64_t ld, const float* __restrict__ u,
                                                const float* __restrict__ v, int64_t M, int64_t N, float scale) {
  pdl_prologue();
  const int64_t total = M * N;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = t / N, n = t - m * N;
    out[m * ld + n] += scale * __ldg(u + m) * __ldg(v + n);
  }
}

}  // namespace eqh

extern "C" {

int eqh_rank1_update(float* out, int64_t ld, const float* u, const float* v, int64_t M, int64_t N, float scale,
                     eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(M >= 0 && N >= 0 && ld >= N && (M * N == 0 || (out && u && v)));
  if (M * N == 0) return EQH_OK;
  int64_t blocks = (M * N + 255) / 256;
  if (blocks > (int64_t)kSMs * 8) blocks = (int64_t)kSMs * 8;
  launch_kernel(k_rank1, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), out, ld, u, v, M, N, scale);
  EQH_LAUNCH_RC();
}

int eqh_version(void) { return 100; /* 0.1.0 */ }

const char* eqh_error_string(int rc) {
  switch (rc) {
    case EQH_OK: return "ok";
    case EQH_ERR_ARG: return "invalid argument (null pointer, negative size or misaligned buffer)";
    case EQH_ERR_UNSUPPORTED: return "unsupported configuration (feature width / hidden width / K outside the implemented set)";
    case EQH_ERR_WORKSPACE: return "workspace too small";
    default: return rc > 0 ? cudaGetErrorString(static_cast<cudaError_t>(rc)) : "unknown error";
  }
}

int eqh_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int64_t step, float grad_scale,
                  eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(n >= 0 && step >= 1 && (n == 0 || (param && grad && exp_avg && exp_avg_sq)));
  if (n == 0) return EQH_OK;
  // bias corrections in double on the host, as torch.optim.Adam computes them (1 - 0.999f in fp32 carries ~1e-5)
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)beta2, (double)step));
  int64_t blocks = (n / 4 + 255) / 256;          // one float4 group per thread and pass
  if (blocks > (int64_t)kSMs * 8) blocks = (int64_t)kSMs * 8;
  if (blocks < 1) blocks = 1;
  launch_kernel(k_adam, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                         beta2, eps, weight_decay, bc1, sqrtf(bc2),
                                                                         grad_scale);
  EQH_LAUNCH_RC();
}

int eqh_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, int64_t* step_dev, int advance,
                      int32_t* counter, float grad_scale, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(n >= 0 && step_dev && (!advance || counter) && (n == 0 || (param && grad && exp_avg && exp_avg_sq)));
  if (n == 0 && !advance) return EQH_OK;
  int64_t blocks = (n / 4 + 255) / 256;          // one float4 group per thread and pass
  if (blocks > (int64_t)kSMs * 8) blocks = (int64_t)kSMs * 8;
  if (blocks < 1) blocks = 1;
  launch_kernel(k_adam_dev, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), param, grad, exp_avg,
                exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, reinterpret_cast<long long*>(step_dev), advance, counter,
                grad_scale);
  EQH_LAUNCH_RC();
}

}  // extern "C"
