// EGNN coordinate branch phi_x (egnn_layer.py:325-340, CoorsNorm :71-81), forward and backward, as kernels:
//     w_ij   = W4 . SiLU(W3 . m_ij + b3) + b4                        (16 -> 64 -> 1 per kNN edge)
//     r^_ij  = (x_i - x_j) / max(|x_i - x_j|, 1e-8) * scale
//     x'_i   = x_i + sum_k w_ij r^_ij
// The reference runs this as ~10 ATen / cuBLAS launches on materialised [1,N,16,*] tensors; every wrapper of the
// reference then discards x' (equihnn_egnn.py:77,158,242), so this is the drop-in path's cost, not the training step's.
//
// One thread per kNN edge (the 16 edges of an atom are 16 consecutive lanes: the sum over k is four shuffles); the
// 16 x 64 coordinate MLP lives in shared memory and is read by warp-wide broadcasts.  fp32 with accurate expf / IEEE
// division (the bar is 1e-5 against the reference).
//
// Backward (only reached when a caller actually uses x'): the per-edge kernel recomputes the hidden layer and writes
// what the parameter gradients are contractions of -- d pre3 [E,64], t [E,64], dw [E], the per-edge scale term -- plus
// d m_ij [E,16] and the gradient g_e of each edge's relative vector; the contractions themselves (dW3 = dpre3^T . m,
// dW4 = t^T . dw, column sums) run on the library's deterministic GEMM / column-sum kernels, and eqh_egnn_coors_bwd_pos
// folds g into d x (own edges minus incoming edges over the transposed kNN incidence: no atomics).  The self edge
// (j == i, rel = 0) contributes +g and -g to the same atom: it is skipped, where the reference's fp32 autograd cancels
// two 1e8-scaled terms.
#include "common.cuh"

namespace eqh {
namespace coors {

constexpr int kM = EQH_M_DIM;      // 16
constexpr int kNbr = EQH_KNN_K;    // 16
constexpr int kHid = 4 * kM;       // 64
constexpr float kEps = 1e-8f;

struct Mlp {
  float w3[kHid * kM];   // [64][16]
  float b3[kHid];
  float w4[kHid];
  float b4;
};

__device__ __forceinline__ void load_mlp(Mlp* s, const float* W3, const float* b3, const float* W4, const float* b4) {
  for (int t = threadIdx.x; t < kHid * kM; t += blockDim.x) s->w3[t] = __ldg(W3 + t);
  for (int t = threadIdx.x; t < kHid; t += blockDim.x) {
    s->b3[t] = __ldg(b3 + t);
    s->w4[t] = __ldg(W4 + t);
  }
  if (threadIdx.x == 0) s->b4 = __ldg(b4);
  __syncthreads();
}

__device__ __forceinline__ void load_m(const float* Mij, int64_t e, float* m) {
  const float4* p = reinterpret_cast<const float4*>(Mij + e * kM);
#pragma unroll
  for (int q = 0; q < kM / 4; ++q) {
    const float4 t = __ldg(p + q);
    m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w;
  }
}

__global__ void __launch_bounds__(256) k_coors_fwd(const float* __restrict__ Mij, const float* __restrict__ pos,
                                                    const int32_t* __restrict__ idx, const float* __restrict__ W3,
                                                    const float* __restrict__ b3, const float* __restrict__ W4,
                                                    const float* __restrict__ b4, const float* __restrict__ scale, int64_t N,
                                                    float* __restrict__ out, float* __restrict__ wij) {
  pdl_prologue();
  __shared__ Mlp s;
  load_mlp(&s, W3, b3, W4, b4);
  const float sc = scale ? __ldg(scale) : 1.f;
  const int64_t n_edges = N * kNbr;
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x; e0 < n_edges; e0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = e0 + threadIdx.x;
    const bool ok = e < n_edges;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (ok) {
      float m[kM];
      load_m(Mij, e, m);
      float w = s.b4;
#pragma unroll 4
      for (int u = 0; u < kHid; ++u) {
        float pre = s.b3[u];
#pragma unroll
        for (int o = 0; o < kM; ++o) pre += s.w3[u * kM + o] * m[o];
        w += s.w4[u] * silu_f(pre);
      }
      if (wij) wij[e] = w;
      const int64_t i = e >> 4;
      const int64_t j = __ldg(idx + e);
      const float rx = __ldg(pos + 3 * i) - __ldg(pos + 3 * j), ry = __ldg(pos + 3 * i + 1) - __ldg(pos + 3 * j + 1),
                  rz = __ldg(pos + 3 * i + 2) - __ldg(pos + 3 * j + 2);
      float f = w;
      if (scale) f = w * sc / fmaxf(sqrtf(rx * rx + ry * ry + rz * rz), kEps);       // CoorsNorm
      cx = f * rx; cy = f * ry; cz = f * rz;
    }
    // sum over the 16 neighbours of an atom: 16 consecutive lanes (256 threads = 16 atoms per block pass)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      cx += __shfl_xor_sync(kFull, cx, o);
      cy += __shfl_xor_sync(kFull, cy, o);
      cz += __shfl_xor_sync(kFull, cz, o);
    }
    if (ok && (e & 15) == 0) {
      const int64_t i = e >> 4;
      out[3 * i] = __ldg(pos + 3 * i) + cx;
      out[3 * i + 1] = __ldg(pos + 3 * i + 1) + cy;
      out[3 * i + 2] = __ldg(pos + 3 * i + 2) + cz;
    }
  }
}

__global__ void __launch_bounds__(256) k_coors_bwd(const float* __restrict__ Mij, const float* __restrict__ pos,
                                                    const int32_t* __restrict__ idx, const float* __restrict__ W3,
                                                    const float* __restrict__ b3, const float* __restrict__ W4,
                                                    const float* __restrict__ b4, const float* __restrict__ scale,
                                                    const float* __restrict__ dout, int64_t N, float* __restrict__ dMij,
                                                    float* __restrict__ dpre3, float* __restrict__ thid, float* __restrict__ dw,
                                                    float* __restrict__ dscale_e, float* __restrict__ grel) {
  pdl_prologue();
  __shared__ Mlp s;
  load_mlp(&s, W3, b3, W4, b4);
  const float sc = scale ? __ldg(scale) : 1.f;
  const int64_t n_edges = N * kNbr;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e >> 4;
    const int64_t j = __ldg(idx + e);
    const float gx = __ldg(dout + 3 * i), gy = __ldg(dout + 3 * i + 1), gz = __ldg(dout + 3 * i + 2);
    const float rx = __ldg(pos + 3 * i) - __ldg(pos + 3 * j), ry = __ldg(pos + 3 * i + 1) - __ldg(pos + 3 * j + 1),
                rz = __ldg(pos + 3 * i + 2) - __ldg(pos + 3 * j + 2);
    const float nrm = sqrtf(rx * rx + ry * ry + rz * rz);
    const float c = fmaxf(nrm, kEps);
    const float f = scale ? sc / c : 1.f;                       // r^ = rel * f
    const float gr = gx * rx + gy * ry + gz * rz;               // dout_i . rel
    const float dwe = gr * f;                                   // d w_e = dout_i . r^_e
    float m[kM], dm[kM];
    load_m(Mij, e, m);
#pragma unroll
    for (int o = 0; o < kM; ++o) dm[o] = 0.f;
    float w = s.b4;
    for (int u = 0; u < kHid; ++u) {
      float pre = s.b3[u];
#pragma unroll
      for (int o = 0; o < kM; ++o) pre += s.w3[u * kM + o] * m[o];
      const float sg = 1.f / (1.f + expf(-pre));
      const float t = pre * sg;
      w += s.w4[u] * t;
      const float dpre = dwe * s.w4[u] * (sg * (1.f + pre * (1.f - sg)));
      thid[e * kHid + u] = t;
      dpre3[e * kHid + u] = dpre;
#pragma unroll
      for (int o = 0; o < kM; ++o) dm[o] += dpre * s.w3[u * kM + o];
    }
    float4* dmp = reinterpret_cast<float4*>(dMij + e * kM);
#pragma unroll
    for (int q = 0; q < kM / 4; ++q) dmp[q] = make_float4(dm[4 * q], dm[4 * q + 1], dm[4 * q + 2], dm[4 * q + 3]);
    dw[e] = dwe;
    dscale_e[e] = scale ? w * gr / c : 0.f;                     // d scale = sum_e w_e dout_i . rel / c
    // gradient of the relative vector: G = w * dout_i through r^ = rel * scale / max(|rel|, eps)
    float hx = 0.f, hy = 0.f, hz = 0.f;
    if (j != i) {
      const float Gx = w * gx, Gy = w * gy, Gz = w * gz;
      if (!scale) {
        hx = Gx; hy = Gy; hz = Gz;
      } else if (nrm > kEps) {                                  // (I - u u^T) G * scale / |rel|, u = rel / |rel|
        const float ug = (Gx * rx + Gy * ry + Gz * rz) / (nrm * nrm);
        hx = (Gx - ug * rx) * f; hy = (Gy - ug * ry) * f; hz = (Gz - ug * rz) * f;
      } else {                                                  // clamped norm: no gradient through it
        hx = Gx * f; hy = Gy * f; hz = Gz * f;
      }
    }
    grel[3 * e] = hx; grel[3 * e + 1] = hy; grel[3 * e + 2] = hz;
  }
}

// d x_a = dout_a + sum over a's own edges of g_e - sum over the edges that point AT a (transposed kNN incidence), in
// ascending edge order: deterministic.
__global__ void __launch_bounds__(256) k_coors_bwd_pos(const float* __restrict__ dout, const float* __restrict__ grel,
                                                        const int32_t* __restrict__ rowptr_j, const int32_t* __restrict__ perm_j,
                                                        int64_t N, float* __restrict__ dpos) {
  pdl_prologue();
  for (int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; a < N; a += (int64_t)gridDim.x * blockDim.x) {
    float x = __ldg(dout + 3 * a), y = __ldg(dout + 3 * a + 1), z = __ldg(dout + 3 * a + 2);
    for (int k = 0; k < kNbr; ++k) {
      const int64_t e = a * kNbr + k;
      x += __ldg(grel + 3 * e); y += __ldg(grel + 3 * e + 1); z += __ldg(grel + 3 * e + 2);
    }
    const int32_t lo = __ldg(rowptr_j + a), hi = __ldg(rowptr_j + a + 1);
    for (int32_t q = lo; q < hi; ++q) {
      const int64_t e = __ldg(perm_j + q);
      x -= __ldg(grel + 3 * e); y -= __ldg(grel + 3 * e + 1); z -= __ldg(grel + 3 * e + 2);
    }
    dpos[3 * a] = x; dpos[3 * a + 1] = y; dpos[3 * a + 2] = z;
  }
}

static unsigned blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  if (b > (int64_t)kSMs * 8) b = (int64_t)kSMs * 8;
  return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace coors
}  // namespace eqh

extern "C" {

int eqh_egnn_coors_fwd(const float* Mij, const float* pos, const int32_t* idx, const float* W3, const float* b3,
                       const float* W4, const float* b4, const float* scale, int64_t N, float* coors_out, float* wij,
                       eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::coors;
  EQH_CHECK_ARG(N >= 0 && W3 && b3 && W4 && b4 && (N == 0 || (Mij && pos && idx && coors_out)));
  EQH_CHECK_ARG(aligned16(Mij));
  if (N == 0) return EQH_OK;
  launch_kernel(k_coors_fwd, dim3(blocks_for(N * kNbr)), dim3(256), 0, static_cast<cudaStream_t>(stream), Mij, pos, idx, W3, b3,
                W4, b4, scale, N, coors_out, wij);
  EQH_LAUNCH_RC();
}

int eqh_egnn_coors_bwd(const float* Mij, const float* pos, const int32_t* idx, const float* W3, const float* b3,
                       const float* W4, const float* b4, const float* scale, const float* dcoors_out, int64_t N, float* dMij,
                       float* dpre3, float* thid, float* dw, float* dscale_e, float* grel, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::coors;
  EQH_CHECK_ARG(N >= 0 && W3 && b3 && W4 && b4 &&
                (N == 0 || (Mij && pos && idx && dcoors_out && dMij && dpre3 && thid && dw && dscale_e && grel)));
  EQH_CHECK_ARG(aligned16(Mij) && aligned16(dMij));
  if (N == 0) return EQH_OK;
  launch_kernel(k_coors_bwd, dim3(blocks_for(N * kNbr)), dim3(256), 0, static_cast<cudaStream_t>(stream), Mij, pos, idx, W3, b3,
                W4, b4, scale, dcoors_out, N, dMij, dpre3, thid, dw, dscale_e, grel);
  EQH_LAUNCH_RC();
}

int eqh_egnn_coors_bwd_pos(const float* dcoors_out, const float* grel, const int32_t* rowptr_j, const int32_t* perm_j,
                           int64_t N, float* dpos, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::coors;
  EQH_CHECK_ARG(N >= 0 && (N == 0 || (dcoors_out && grel && rowptr_j && perm_j && dpos)));
  if (N == 0) return EQH_OK;
  launch_kernel(k_coors_bwd_pos, dim3(blocks_for(N)), dim3(256), 0, static_cast<cudaStream_t>(stream), dcoors_out, grel,
                rowptr_j, perm_j, N, dpos);
  EQH_LAUNCH_RC();
}

}  // extern "C"
