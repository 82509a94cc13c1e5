// Batch-wide exact 16-nearest-neighbour search on squared distance, O(N*K) memory.
// Replaces the reference's all-pairs [1,N,N,3] / [1,N,N] temporaries + topk
// (egnn_layer.py:253-259,281-283).  CUDA-core fp32 work: candidates are staged through shared
// memory in tiles; each query is scanned by S cooperating lanes, each keeping a sorted top-16
// list in registers, merged at the end by a 16-round tournament over warp shuffles.
//
// d2 is evaluated exactly as the reference's (rel ** 2).sum(-1): separate IEEE multiplies
// and adds, (dx*dx + dy*dy) + dz*dz, no FMA contraction -> bit-identical ranking keys.
// Ties (equal d2) are broken by the smaller index; torch.topk leaves tie order unspecified.
#include <float.h>

#include "common.cuh"

namespace eqh {

constexpr int kK = EQH_KNN_K;
constexpr int kTile = 2048;

struct TopK {
  float d[kK];
  int32_t j[kK];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int p = 0; p < kK; ++p) { d[p] = FLT_MAX; j[p] = 0x7fffffff; }
  }
  // candidates arrive in ascending index order per lane: strict '<' keeps equal keys in index order
  __device__ __forceinline__ void push(float nd, int32_t nj) {
    if (nd < d[kK - 1]) {
      d[kK - 1] = nd; j[kK - 1] = nj;
#pragma unroll
      for (int p = kK - 1; p > 0; --p) {
        if (d[p] < d[p - 1]) {
          float td = d[p]; d[p] = d[p - 1]; d[p - 1] = td;
          int32_t tj = j[p]; j[p] = j[p - 1]; j[p - 1] = tj;
        }
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int p = 0; p < kK - 1; ++p) { d[p] = d[p + 1]; j[p] = j[p + 1]; }
    d[kK - 1] = FLT_MAX; j[kK - 1] = 0x7fffffff;
  }
};

template <int S>  // lanes per query (8 or 32); 256/S queries per CTA
__global__ void __launch_bounds__(256) k_knn16(const float* __restrict__ pos, int64_t N, int32_t* __restrict__ idx,
                                                float* __restrict__ d2out) {
  pdl_prologue();
  __shared__ float sp[kTile * 3];
  constexpr int QPB = 256 / S;
  const int sub = threadIdx.x % S;
  const int64_t q = (int64_t)blockIdx.x * QPB + threadIdx.x / S;
  const int64_t qc = q < N ? q : N - 1;
  const float xi = __ldg(pos + qc * 3 + 0), yi = __ldg(pos + qc * 3 + 1), zi = __ldg(pos + qc * 3 + 2);
  TopK best;
  best.init();
  for (int64_t base = 0; base < N; base += kTile) {
    const int n = (int)min((int64_t)kTile, N - base);
    __syncthreads();
    for (int t = threadIdx.x; t < n * 3; t += 256) sp[t] = __ldg(pos + base * 3 + t);
    __syncthreads();
    for (int c = sub; c < n; c += S) {
      const float dx = __fsub_rn(xi, sp[3 * c + 0]);
      const float dy = __fsub_rn(yi, sp[3 * c + 1]);
      const float dz = __fsub_rn(zi, sp[3 * c + 2]);
      const float dd = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      best.push(dd, (int32_t)(base + c));
    }
  }
  // merge the S sorted lists of this query: 16 rounds, lexicographic (d2, index) minimum
#pragma unroll 1
  for (int r = 0; r < kK; ++r) {
    float md = best.d[0];
    int32_t mj = best.j[0];
#pragma unroll
    for (int o = S / 2; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(kFull, md, o);
      const int32_t oj = __shfl_xor_sync(kFull, mj, o);
      if (od < md || (od == md && oj < mj)) { md = od; mj = oj; }
    }
    if (best.j[0] == mj && best.d[0] == md) best.pop();
    if (sub == (r % S) && q < N) {
      idx[q * kK + r] = mj;
      d2out[q * kK + r] = md;
    }
  }
}

// Small clouds (N < 16384): ONE WARP PER QUERY with the top-16 list distributed over lanes 0-15 (sorted, one entry per
// lane).  Each lane evaluates one candidate per round; a ballot finds the few that beat the current 16th distance and they
// are inserted one by one (ascending lane = ascending index, so equal keys stay in index order) by a shuffle-up of the
// list's tail.  ~16 (1 + ln(N/16)) insertions per query instead of a 16-deep insertion sort in every lane on (almost)
// every candidate: 3-4x fewer instructions than k_knn16<32> at N = 2560, same keys, same tie rule, bit-identical result.
constexpr int kWarpsPerCta = 8;
// small candidate tile (6 KB): two or three of these CTAs per SM leave room for a 194 KB GEMM CTA next to them (the P|Q
// GEMM of phi_e runs beside the search)
constexpr int kTileW = 512;
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_knn16_warp(const float* __restrict__ pos, int64_t N,
                                                                  int32_t* __restrict__ idx, float* __restrict__ d2out) {
  pdl_prologue();
  __shared__ float sp[kTileW * 3];
  static_assert(kK == 16, "one list entry per lane of a half warp");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * kWarpsPerCta + warp;
  const int64_t qc = q < N ? q : N - 1;
  const float xi = __ldg(pos + qc * 3 + 0), yi = __ldg(pos + qc * 3 + 1), zi = __ldg(pos + qc * 3 + 2);
  float ld = FLT_MAX;            // lanes 0-15: entry `lane` of the sorted list (lanes 16-31 mirror nothing: +inf keys)
  int32_t lj = 0x7fffffff;
  float thr = FLT_MAX;           // the list's 16th key, on all lanes
  for (int64_t base = 0; base < N; base += kTileW) {
    const int n = (int)min((int64_t)kTileW, N - base);
    __syncthreads();
    for (int t = threadIdx.x; t < n * 3; t += kWarpsPerCta * 32) sp[t] = __ldg(pos + base * 3 + t);
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 32) {
      const int c = c0 + lane;
      float dd = FLT_MAX;
      if (c < n) {
        const float dx = __fsub_rn(xi, sp[3 * c + 0]);
        const float dy = __fsub_rn(yi, sp[3 * c + 1]);
        const float dz = __fsub_rn(zi, sp[3 * c + 2]);
        dd = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      }
      unsigned m = __ballot_sync(kFull, dd < thr);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float nd = __shfl_sync(kFull, dd, src);
        if (nd < thr) {                                   // (thr may have dropped since the ballot; warp-uniform)
          const int32_t nj = (int32_t)(base + c0 + src);
          // insertion position = number of entries with key <= nd (equal keys have smaller indices: they stay in front)
          const int p = __popc(__ballot_sync(kFull, lane < kK && ld <= nd));
          const float ud = __shfl_up_sync(kFull, ld, 1);
          const int32_t uj = __shfl_up_sync(kFull, lj, 1);
          if (lane > p) { ld = ud; lj = uj; }
          else if (lane == p) { ld = nd; lj = nj; }
          thr = __shfl_sync(kFull, ld, kK - 1);
        }
      }
    }
  }
  if (lane < kK && q < N) {
    idx[q * kK + lane] = lj;
    d2out[q * kK + lane] = ld;
  }
}

}  // namespace eqh

extern "C" int eqh_knn16(const float* pos, int64_t N, int32_t* idx, float* d2, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(pos && idx && d2);
  EQH_CHECK_ARG(N >= kK && N < ((int64_t)1 << 31));  // the reference's topk raises for k > N
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (N >= 16384) {
    launch_kernel(k_knn16<8>, dim3((unsigned)((N + 31) / 32)), dim3(256), 0, st, pos, N, idx, d2);
  } else {
    launch_kernel(k_knn16_warp, dim3((unsigned)((N + kWarpsPerCta - 1) / kWarpsPerCta)), dim3(kWarpsPerCta * 32), 0, st, pos, N,
                  idx, d2);
  }
  EQH_LAUNCH_RC();
}
