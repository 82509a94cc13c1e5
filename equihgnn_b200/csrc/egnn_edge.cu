// EGNN edge model phi_e + neighbour pooling, fp32 CUDA-core path (parity path, 1e-5).
//
// The reference materialises cat(h_i, h_j, d2) [1,N,16,2C+1] and runs two Linear layers on
// N*16 rows (egnn_layer.py:298-310).  Here the first Linear is split over the concat:
//     W1 [h_i | h_j | d2] + b1 = (h W1a^T + b1)[i] + (h W1b^T)[j] + d2 * wc = P[i] + Q[j] + d2 * wc
// so the per-edge work is a gathered add + SiLU + a 16-wide contraction; P and Q are N-row GEMMs
// done by the host module.  H (the hidden width 2(2C+1)) is zero-padded by the host to a multiple
// of 64; padded columns of W2 / wc are zero, so they contribute nothing.
//
// forward : one warp per atom; W2^T and wc live in shared memory (bank-conflict-free pitch 20);
//           the 16 outputs of an edge are combined over the warp with a 16-shuffle
//           reduce-scatter.
// backward: one CTA per atom, each thread owns HC columns of H for the whole launch, so dW2 /
//           dwc / dP accumulate in registers in a fixed edge order (deterministic, no atomics);
//           per-CTA partials are finished by eqh_colsum.
#include "common.cuh"

namespace eqh {

constexpr int kM = EQH_M_DIM;     // 16
constexpr int kNbr = EQH_KNN_K;   // 16
constexpr int kPitch = 20;        // floats per W2^T row in smem (16 + 4 pad): conflict-free LDS.128

__global__ void __launch_bounds__(256) k_edge_fwd(const float* __restrict__ P, const float* __restrict__ Q,
                                                   const float* __restrict__ wc, const float* __restrict__ W2,
                                                   const float* __restrict__ b2, const int32_t* __restrict__ idx,
                                                   const float* __restrict__ d2, int64_t N, int H,
                                                   float* __restrict__ S2, float* __restrict__ Mij, float* __restrict__ Mi) {
  pdl_prologue();
  extern __shared__ __align__(16) float smem[];
  float* w2t = smem;                 // [H][kPitch]
  float* wcs = smem + (size_t)H * kPitch;  // [H]
  for (int t = threadIdx.x; t < H * kM; t += 256) {
    const int o = t / H, h = t - o * H;  // coalesced read of W2 [16,H]
    w2t[h * kPitch + o] = __ldg(W2 + t);
  }
  for (int t = threadIdx.x; t < H; t += 256) wcs[t] = __ldg(wc + t);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int o_mine = lane >> 1;  // output owned by this lane after the reduce-scatter
  const float bias = __ldg(b2 + o_mine);

  for (int64_t i = warp; i < N; i += nwarps) {
    const float* Pi = P + i * H;
    float mi = 0.f;
    for (int k = 0; k < kNbr; ++k) {
      const int64_t e = i * kNbr + k;
      const int32_t j = __ldg(idx + e);
      const float dd = __ldg(d2 + e);
      const float* Qj = Q + (int64_t)j * H;
      float acc[kM];
#pragma unroll
      for (int o = 0; o < kM; ++o) acc[o] = 0.f;
      for (int h0 = 0; h0 < H; h0 += 64) {  // H % 64 == 0
        float pre[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int h = h0 + u * 32 + lane;
          pre[u] = (__ldg(Pi + h) + __ldg(Qj + h)) + dd * wcs[h];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int h = h0 + u * 32 + lane;
          const float hid = silu_f(pre[u]);
          const float4* w = reinterpret_cast<const float4*>(w2t + h * kPitch);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 ww = w[q4];
            acc[q4 * 4 + 0] += ww.x * hid;
            acc[q4 * 4 + 1] += ww.y * hid;
            acc[q4 * 4 + 2] += ww.z * hid;
            acc[q4 * 4 + 3] += ww.w * hid;
          }
        }
      }
      // reduce-scatter 16 values over 32 lanes: after the 4 halving steps lane l holds the
      // partial of output l>>1 for its lane-pair; one more xor-1 step completes it.
      float r8[8], r4[4], r2[2], r1;
      {
        const bool up = lane & 16;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float keep = up ? acc[t + 8] : acc[t];
          const float send = up ? acc[t] : acc[t + 8];
          r8[t] = keep + __shfl_xor_sync(kFull, send, 16);
        }
      }
      {
        const bool up = lane & 8;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float keep = up ? r8[t + 4] : r8[t];
          const float send = up ? r8[t] : r8[t + 4];
          r4[t] = keep + __shfl_xor_sync(kFull, send, 8);
        }
      }
      {
        const bool up = lane & 4;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float keep = up ? r4[t + 2] : r4[t];
          const float send = up ? r4[t] : r4[t + 2];
          r2[t] = keep + __shfl_xor_sync(kFull, send, 4);
        }
      }
      {
        const bool up = lane & 2;
        const float keep = up ? r2[1] : r2[0];
        const float send = up ? r2[0] : r2[1];
        r1 = keep + __shfl_xor_sync(kFull, send, 2);
      }
      r1 += __shfl_xor_sync(kFull, r1, 1);
      const float s2 = r1 + bias;
      const float m = silu_f(s2);
      if (!(lane & 1)) {
        S2[e * kM + o_mine] = s2;
        if (Mij) Mij[e * kM + o_mine] = m;
      }
      mi += m;
    }
    if (!(lane & 1)) Mi[i * kM + o_mine] = mi;
  }
}

// threads own columns (HC each); fewer columns per thread = fewer registers = more resident warps
constexpr int kBwdMaxThreads = 384;   // (576 threads x 2 columns was measured: 96 registers with spills, 20 % slower)

template <int HC>
__global__ void __launch_bounds__(kBwdMaxThreads) k_edge_bwd(const float* __restrict__ P, const float* __restrict__ Q,
                                                   const float* __restrict__ wc, const float* __restrict__ W2,
                                                   const int32_t* __restrict__ idx, const float* __restrict__ d2,
                                                   const float* __restrict__ S2, const float* __restrict__ dMi,
                                                   const float* __restrict__ dMij, int64_t N, int H,
                                                   float* __restrict__ dP, float* __restrict__ dPre,
                                                   float* __restrict__ dd2, float* __restrict__ part,
                                                   const int32_t* __restrict__ order) {
  pdl_prologue();
  __shared__ __align__(16) float ds2s[kNbr][kM];
  __shared__ float dd2s[kBwdMaxThreads / 32][kNbr];
  __shared__ float dds[kNbr];
  // The 16 neighbour rows Q[j] of the NEXT atom are copied into shared memory with cp.async while the current atom
  // is processed (double buffer [2][16][H]): the gather latency -- 42 % of this kernel's stall samples when each row
  // was prefetched only one neighbour ahead through registers -- is hidden behind a whole atom of arithmetic.
  extern __shared__ __align__(16) float qs[];
  const int T = blockDim.x;
  const int tid = threadIdx.x;
  const int lane = tid & 31, wib = tid >> 5;
  const int H4 = H >> 2;
  // per buffer: 16 neighbour rows of Q, the atom's own row of P, its 16x16 block of S2, dMi [16], d2 [16]
  const size_t kBufFloats = (size_t)(kNbr + 1) * H + kNbr * kM + kM + kNbr;
  auto cp16 = [](float* dst, const float* src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  };
  auto prefetch = [&](int64_t atom, int buf) {
    float* dst = qs + (size_t)buf * kBufFloats;
    for (int t = tid; t < kNbr * H4; t += T) {
      const int k = t / H4, c4 = t - k * H4;
      cp16(dst + (size_t)k * H + c4 * 4, Q + (int64_t)__ldg(idx + atom * kNbr + k) * H + c4 * 4);
    }
    float* pd = dst + (size_t)kNbr * H;
    for (int t = tid; t < H4; t += T) cp16(pd + t * 4, P + atom * H + t * 4);
    float* sd = pd + H;
    if (tid < kNbr * kM / 4) cp16(sd + tid * 4, S2 + atom * kNbr * kM + tid * 4);
    else if (tid < kNbr * kM / 4 + kNbr / 4) cp16(sd + kNbr * kM + kM + (tid - kNbr * kM / 4) * 4, d2 + atom * kNbr + (tid - kNbr * kM / 4) * 4);
    else if (dMi && tid < kNbr * kM / 4 + kNbr / 4 + kM / 4)
      cp16(sd + kNbr * kM + (tid - kNbr * kM / 4 - kNbr / 4) * 4, dMi + atom * kM + (tid - kNbr * kM / 4 - kNbr / 4) * 4);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float w2[kM][HC], dw2[kM][HC], wcr[HC], dwc[HC];
  bool act[HC];
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    const int h = tid + c * T;
    act[c] = h < H;
    wcr[c] = act[c] ? __ldg(wc + h) : 0.f;
    dwc[c] = 0.f;
#pragma unroll
    for (int o = 0; o < kM; ++o) {
      w2[o][c] = act[c] ? __ldg(W2 + (int64_t)o * H + h) : 0.f;
      dw2[o][c] = 0.f;
    }
  }
  float db2 = 0.f;  // threads < 16

  if ((int64_t)blockIdx.x < N) prefetch(order ? (int64_t)__ldg(order + blockIdx.x) : (int64_t)blockIdx.x, 0);
  int buf = 0;
  for (int64_t slot = blockIdx.x; slot < N; slot += gridDim.x, buf ^= 1) {
    // atoms in `order` (the cell order of the kNN grid) when given: the CTAs running side by side work on spatially
    // close atoms, whose neighbour rows Q[j] overlap and hit in L2
    const int64_t i = order ? (int64_t)__ldg(order + slot) : slot;
    __syncthreads();  // previous atom's readers of ds2s/js/dds and of the other Q buffer are done
    const int64_t nslot = slot + gridDim.x;
    if (nslot < N) {
      prefetch(order ? (int64_t)__ldg(order + nslot) : nslot, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // this atom's rows have landed (the next may be in flight)
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // every thread's cp.async chunks of this atom are visible
    const float* cur = qs + (size_t)buf * kBufFloats;
    const float* pcur = cur + (size_t)kNbr * H;
    const float* s2cur = pcur + H;
    for (int t = tid; t < kNbr * kM; t += T) {
      const int k = t >> 4, o = t & 15;
      const int64_t e = i * kNbr + k;
      float g = dMi ? s2cur[kNbr * kM + o] : 0.f;
      if (dMij) g += __ldg(dMij + e * kM + o);
      ds2s[k][o] = g * dsilu_f(s2cur[t]);
    }
    if (tid < kNbr) dds[tid] = s2cur[kNbr * kM + kM + tid];
    __syncthreads();
    if (tid < kM) {
#pragma unroll
      for (int k = 0; k < kNbr; ++k) db2 += ds2s[k][tid];
    }
    float pi[HC], dpa[HC];
#pragma unroll
    for (int c = 0; c < HC; ++c) {
      pi[c] = act[c] ? pcur[tid + c * T] : 0.f;
      dpa[c] = 0.f;
    }
    const float* qbuf = cur;
    for (int k = 0; k < kNbr; ++k) {
      const float dd = dds[k];
      float q[HC];
#pragma unroll
      for (int c = 0; c < HC; ++c) q[c] = act[c] ? qbuf[(size_t)k * H + tid + c * T] : 0.f;
      float g[kM];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const float4 t4 = *reinterpret_cast<const float4*>(&ds2s[k][q4 * 4]);
        g[q4 * 4 + 0] = t4.x; g[q4 * 4 + 1] = t4.y; g[q4 * 4 + 2] = t4.z; g[q4 * 4 + 3] = t4.w;
      }
      float dd2_part = 0.f;
#pragma unroll
      for (int c = 0; c < HC; ++c) {
        const float pre = (pi[c] + q[c]) + dd * wcr[c];
        float ex, sg;                                   // SFU approximations (~1 ulp each), see egnn_edge_tc.cu
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(pre * -1.4426950408889634f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + ex));
        const float hid = pre * sg;
        const float dsl = sg * (1.f + pre * (1.f - sg));
        float dh = 0.f;
#pragma unroll
        for (int o = 0; o < kM; ++o) {
          dh += w2[o][c] * g[o];
          dw2[o][c] += g[o] * hid;
        }
        const float dpre = dh * dsl;
        dpa[c] += dpre;
        dwc[c] += dd * dpre;
        dd2_part += dpre * wcr[c];
        if (act[c]) dPre[(i * kNbr + k) * H + tid + c * T] = dpre;
      }
      if (dd2) {
        dd2_part = warp_sum(dd2_part);
        if (lane == 0) dd2s[wib][k] = dd2_part;
      }
    }
#pragma unroll
    for (int c = 0; c < HC; ++c)
      if (act[c]) dP[i * H + tid + c * T] = dpa[c];
    if (dd2) {
      __syncthreads();
      if (tid < kNbr) {
        float t = 0.f;
        for (int w = 0; w < (T >> 5); ++w) t += dd2s[w][tid];
        dd2[i * kNbr + tid] = t;
      }
    }
  }
  float* row = part + (int64_t)blockIdx.x * ((int64_t)(kM + 1) * H + kM);
#pragma unroll
  for (int c = 0; c < HC; ++c) {
    if (act[c]) {
      const int h = tid + c * T;
#pragma unroll
      for (int o = 0; o < kM; ++o) row[(int64_t)o * H + h] = dw2[o][c];
      row[(int64_t)kM * H + h] = dwc[c];
    }
  }
  if (tid < kM) row[(int64_t)(kM + 1) * H + tid] = db2;
}

// phi_e's parameters re-laid for the kernels: W1 [H0, 2C+1] -> W1a | W1b [Hp, C] (16-byte row strides for TMA) and the
// d2 column wc [Hp]; b1 -> [Hp]; W2 [16, H0] -> [16, Hp].  Pad rows / columns are never written (they stay zero).
__global__ void __launch_bounds__(256) k_egnn_pack(const float* __restrict__ W1, const float* __restrict__ b1,
                                                    const float* __restrict__ W2, int H0, int C, int Hp, float* __restrict__ W1a,
                                                    float* __restrict__ W1b, float* __restrict__ wc, float* __restrict__ b1p,
                                                    float* __restrict__ W2p) {
  pdl_prologue();
  // one block per row of W1 (coalesced reads and writes, no integer division); the last kM blocks copy b1 and W2
  const int r = blockIdx.x;
  if (r < H0) {
    const float* src = W1 + (int64_t)r * (2 * C + 1);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      W1a[(int64_t)r * C + c] = __ldg(src + c);
      W1b[(int64_t)r * C + c] = __ldg(src + C + c);
    }
    if (threadIdx.x == 0) {
      wc[r] = __ldg(src + 2 * C);
      b1p[r] = __ldg(b1 + r);
    }
  } else {
    const int o = r - H0;
    for (int h = threadIdx.x; h < H0; h += blockDim.x) W2p[(int64_t)o * Hp + h] = __ldg(W2 + (int64_t)o * H0 + h);
  }
}

// The inverse re-layout for the gradients, accumulating: T [2 Hp, C] = [dP | dQ]^T . h (rows 0..Hp-1: d W1a, rows Hp..: d W1b)
// and red = (dW2 [16, Hp] | dwc [Hp] | db2 [16]) are added into the parameters' own gradient tensors
// gW1 [H0, 2C+1], gW2 [16, H0], gb2 [16] (one launch instead of four strided ATen adds).
__global__ void __launch_bounds__(256) k_egnn_unpack_grads(const float* __restrict__ T, const float* __restrict__ red, int H0,
                                                            int C, int Hp, float* __restrict__ gW1, float* __restrict__ gW2,
                                                            float* __restrict__ gb2) {
  pdl_prologue();
  const int64_t ld = 2 * (int64_t)C + 1;
  const int64_t n1 = (int64_t)H0 * ld, n2 = n1 + (int64_t)kM * H0, n3 = n2 + kM;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n3; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < n1) {
      const int64_t r = t / ld, c = t - r * ld;
      float v;
      if (c < C) v = __ldg(T + r * C + c);
      else if (c < 2 * C) v = __ldg(T + ((int64_t)Hp + r) * C + (c - C));
      else v = __ldg(red + (int64_t)kM * Hp + r);
      gW1[t] += v;
    } else if (t < n2) {
      const int64_t u = t - n1, o = u / H0, h = u - o * H0;
      gW2[u] += __ldg(red + o * Hp + h);
    } else {
      gb2[t - n2] += __ldg(red + (int64_t)(kM + 1) * Hp + (t - n2));
    }
  }
}

}  // namespace eqh

extern "C" {

int eqh_egnn_unpack_grads(const float* T, const float* red, int64_t H0, int64_t C, int64_t Hp, float* gW1, float* gW2,
                          float* gb2, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(T && red && gW1 && gW2 && gb2 && H0 > 0 && C > 0 && Hp >= H0);
  const int64_t n = H0 * (2 * C + 1) + (int64_t)kM * H0 + kM;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)kSMs * 8) blocks = (int64_t)kSMs * 8;
  launch_kernel(k_egnn_unpack_grads, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), T, red, (int)H0,
                (int)C, (int)Hp, gW1, gW2, gb2);
  EQH_LAUNCH_RC();
}

int eqh_egnn_pack(const float* W1, const float* b1, const float* W2, int64_t H0, int64_t C, int64_t Hp, float* W1a, float* W1b,
                  float* wc, float* b1p, float* W2p, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(W1 && b1 && W2 && W1a && W1b && wc && b1p && W2p && H0 > 0 && C > 0 && Hp >= H0);
  launch_kernel(k_egnn_pack, dim3((unsigned)(H0 + kM)), dim3(256), 0, static_cast<cudaStream_t>(stream), W1, b1, W2, (int)H0, (int)C,
                (int)Hp, W1a, W1b, wc, b1p, W2p);
  EQH_LAUNCH_RC();
}

size_t eqh_egnn_edge_fwd_smem_bytes(int64_t H) {
  return H > 0 ? sizeof(float) * (size_t)H * (eqh::kPitch + 1) : 0;
}

int eqh_egnn_edge_fwd(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                      const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij, float* Mi,
                      eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(P && Q && wc && W2 && b2 && idx && d2 && S2 && Mi && N >= 0);
  if (H <= 0 || H % 64 != 0) return EQH_ERR_UNSUPPORTED;
  const size_t smem = eqh_egnn_edge_fwd_smem_bytes(H);
  if (smem > 227 * 1024) return EQH_ERR_UNSUPPORTED;
  if (N == 0) return EQH_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaFuncSetAttribute(k_edge_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  int64_t blocks = (N + 7) / 8;
  const int64_t cap = smem > 110 * 1024 ? kSMs : 2 * kSMs;  // persistent: W2^T staging is amortised over many atoms
  if (blocks > cap) blocks = cap;
  launch_kernel(k_edge_fwd, dim3((unsigned)blocks), dim3(256), smem, st, P, Q, wc, W2, b2, idx, d2, N, (int)H, S2, Mij, Mi);
  EQH_LAUNCH_RC();
}

int eqh_egnn_edge_bwd(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                      const int32_t* idx, const float* d2, const float* S2, const float* dMi, const float* dMij,
                      int64_t N, int64_t H, float* dP, float* dPre, float* dd2, float* part, eqh_stream_t stream) {
  return eqh_egnn_edge_bwd_ordered(P, Q, wc, W2, b2, idx, d2, S2, dMi, dMij, N, H, dP, dPre, dd2, part, nullptr, stream);
}

int eqh_egnn_edge_bwd_ordered(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                              const int32_t* idx, const float* d2, const float* S2, const float* dMi, const float* dMij,
                              int64_t N, int64_t H, float* dP, float* dPre, float* dd2, float* part, const int32_t* order,
                              eqh_stream_t stream) {
  using namespace eqh;
  (void)b2;
  EQH_CHECK_ARG(P && Q && wc && W2 && idx && d2 && S2 && dP && dPre && part && N >= 0);
  EQH_CHECK_ARG(aligned16(Q) && aligned16(P) && aligned16(S2) && aligned16(d2) && aligned16(dMi));   // 16-byte cp.async
  if (H <= 0 || H % 64 != 0) return EQH_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // threads own columns: HC = ceil(H / 576) columns per thread, T = ceil(H / HC) rounded up to a warp
  const int HC = (int)((H + kBwdMaxThreads - 1) / kBwdMaxThreads);
  int T = (int)((H + HC - 1) / HC);
  T = (T + 31) / 32 * 32;
  if (T < 96) T = 96;   // the prefetch of S2 / d2 / dMi uses threads 0..71 (threads past H own no column)
  const unsigned grid = kPartialRows;  // every CTA writes its partial row (zeros when it owns no atom)
  // double-buffered: 16 neighbour rows of Q + the atom's P row + its S2 block, dMi and d2
  const size_t qsmem = sizeof(float) * 2 * ((size_t)(kNbr + 1) * H + kNbr * kM + kM + kNbr);
  if (qsmem > 200 * 1024) return EQH_ERR_UNSUPPORTED;
#define EQH_BWD(HCV)                                                                                              \
  do {                                                                                                            \
    cudaError_t e_ = cudaFuncSetAttribute(k_edge_bwd<HCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsmem); \
    if (e_ != cudaSuccess) return (int)e_;                                                                        \
    launch_kernel(k_edge_bwd<HCV>, dim3(grid), dim3(T), qsmem, st, P, Q, wc, W2, idx, d2, S2, dMi, dMij, N, (int)H, dP, dPre, \
                  dd2, part, order);                                                                              \
  } while (0)
  switch (HC) {
    case 1: EQH_BWD(1); break;
    case 2: EQH_BWD(2); break;
    case 3: EQH_BWD(3); break;
    case 4: EQH_BWD(4); break;
    case 5: EQH_BWD(5); break;
    case 6: EQH_BWD(6); break;
    default: return EQH_ERR_UNSUPPORTED;
  }
#undef EQH_BWD
  EQH_LAUNCH_RC();
}

}  // extern "C"
