// Shared neighbour-search primitive for the geometric encoders next to the hot path:
//   eqh_knn    -- exact k-nearest neighbours on squared distance for k in [1, 32], either over the whole cloud (EGNN,
//                 egnn_layer.py:281-283) or per molecule of a collated batch, self optionally excluded (FAFormer's dense
//                 batch search, fa_former_layer.py:650-667; Equiformer / SE(3)-Transformer topk, equiformer_layer.py:1326,
//                 se3_transformer_layer.py:1577);
//   eqh_radius -- torch_cluster.radius_graph semantics (ViSNet, visnet_layer.py:336): per molecule, the first
//                 max_num_neighbors atoms j (ascending index) with d2 < r*r.
// Both are O(N * k) memory: nothing of shape [N, N] exists.  CUDA-core fp32 work; d2 = (dx*dx + dy*dy) + dz*dz with
// explicit round-to-nearest multiplies / adds (no FMA contraction), the association of `(rel ** 2).sum(-1)`, so that
// the ranking keys are bit-identical to the PyTorch expression; ties are broken by the smaller index.
//
// Atoms of one molecule are contiguous in a collated batch (`batch` is sorted, PyG Batch.ptr gives the ranges), so a
// query scans only its own range [seg_ptr[s], seg_ptr[s+1]).
#include <float.h>

#include "common.cuh"

namespace eqh {
namespace knng {

template <int KMAX>
struct TopList {
  float d[KMAX];
  int32_t j[KMAX];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int p = 0; p < KMAX; ++p) { d[p] = FLT_MAX; j[p] = 0x7fffffff; }
  }
  // candidates arrive in ascending index order per lane: strict '<' keeps equal keys in index order
  __device__ __forceinline__ void push(float nd, int32_t nj) {
    if (nd < d[KMAX - 1]) {
      d[KMAX - 1] = nd; j[KMAX - 1] = nj;
#pragma unroll
      for (int p = KMAX - 1; p > 0; --p) {
        if (d[p] < d[p - 1]) {
          float td = d[p]; d[p] = d[p - 1]; d[p - 1] = td;
          int32_t tj = j[p]; j[p] = j[p - 1]; j[p - 1] = tj;
        }
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int p = 0; p < KMAX - 1; ++p) { d[p] = d[p + 1]; j[p] = j[p + 1]; }
    d[KMAX - 1] = FLT_MAX; j[KMAX - 1] = 0x7fffffff;
  }
};

__device__ __forceinline__ float dist2(float xi, float yi, float zi, float xj, float yj, float zj) {
  const float dx = __fsub_rn(xi, xj), dy = __fsub_rn(yi, yj), dz = __fsub_rn(zi, zj);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

constexpr int kTile = 1024;

// S lanes cooperate on one query (256 / S queries per CTA).  SEG: per-molecule ranges (candidates read through the
// read-only cache: a molecule's coordinates are a few hundred bytes); otherwise the whole cloud, staged through shared
// memory in tiles shared by the CTA's queries.
template <int KMAX, int S, bool SEG>
__global__ void __launch_bounds__(256) k_knn(const float* __restrict__ pos, int64_t N, int k,
                                             const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_of,
                                             int exclude_self, int32_t* __restrict__ idx, float* __restrict__ d2out) {
  pdl_prologue();
  __shared__ float sp[SEG ? 1 : kTile * 3];
  constexpr int QPB = 256 / S;
  const int sub = threadIdx.x % S;
  const int64_t q = (int64_t)blockIdx.x * QPB + threadIdx.x / S;
  const int64_t qc = q < N ? q : N - 1;
  const float xi = __ldg(pos + qc * 3 + 0), yi = __ldg(pos + qc * 3 + 1), zi = __ldg(pos + qc * 3 + 2);
  TopList<KMAX> best;
  best.init();
  if constexpr (SEG) {
    const int32_t s = __ldg(seg_of + qc);
    const int32_t lo = __ldg(seg_ptr + s), hi = __ldg(seg_ptr + s + 1);
    for (int32_t c = lo + sub; c < hi; c += S) {
      if (exclude_self && c == (int32_t)qc) continue;
      const float dd = dist2(xi, yi, zi, __ldg(pos + 3 * (int64_t)c), __ldg(pos + 3 * (int64_t)c + 1), __ldg(pos + 3 * (int64_t)c + 2));
      best.push(dd, c);
    }
  } else {
    for (int64_t base = 0; base < N; base += kTile) {
      const int n = (int)min((int64_t)kTile, N - base);
      __syncthreads();
      for (int t = threadIdx.x; t < n * 3; t += 256) sp[t] = __ldg(pos + base * 3 + t);
      __syncthreads();
      for (int c = sub; c < n; c += S) {
        if (exclude_self && base + c == qc) continue;
        best.push(dist2(xi, yi, zi, sp[3 * c], sp[3 * c + 1], sp[3 * c + 2]), (int32_t)(base + c));
      }
    }
  }
  // merge the S sorted lists of this query: k rounds, lexicographic (d2, index) minimum
#pragma unroll 1
  for (int r = 0; r < k; ++r) {
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
      const bool found = mj != 0x7fffffff;                 // fewer than k candidates in the range: slot stays empty
      idx[q * k + r] = found ? mj : -1;
      d2out[q * k + r] = found ? md : __int_as_float(0x7f800000);
    }
  }
}

// One warp per query: candidates of the query's range in ascending index order, 32 at a time; the hits of a chunk are
// compacted in order with a ballot, so the output is the first `max_nb` neighbours by index (torch_cluster semantics).
__global__ void __launch_bounds__(256) k_radius(const float* __restrict__ pos, int64_t N, float r2,
                                                 const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_of,
                                                 int loop, int max_nb, int32_t* __restrict__ idx, int32_t* __restrict__ count,
                                                 float* __restrict__ d2out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= N) return;
  const float xi = __ldg(pos + q * 3 + 0), yi = __ldg(pos + q * 3 + 1), zi = __ldg(pos + q * 3 + 2);
  int32_t lo = 0, hi = (int32_t)N;
  if (seg_ptr) {
    const int32_t s = __ldg(seg_of + q);
    lo = __ldg(seg_ptr + s);
    hi = __ldg(seg_ptr + s + 1);
  }
  int n = 0;
  for (int32_t base = lo; base < hi && n < max_nb; base += 32) {
    const int32_t c = base + lane;
    bool hit = false;
    float dd = 0.f;
    if (c < hi && (loop || c != (int32_t)q)) {
      dd = dist2(xi, yi, zi, __ldg(pos + 3 * (int64_t)c), __ldg(pos + 3 * (int64_t)c + 1), __ldg(pos + 3 * (int64_t)c + 2));
      hit = dd < r2;
    }
    const unsigned m = __ballot_sync(kFull, hit);
    const int at = n + __popc(m & ((1u << lane) - 1u));
    if (hit && at < max_nb) {
      idx[q * max_nb + at] = c;
      if (d2out) d2out[q * max_nb + at] = dd;
    }
    n += __popc(m);
  }
  n = min(n, max_nb);
  for (int t = n + lane; t < max_nb; t += 32) {
    idx[q * max_nb + t] = -1;
    if (d2out) d2out[q * max_nb + t] = __int_as_float(0x7f800000);
  }
  if (lane == 0) count[q] = n;
}

}  // namespace knng
}  // namespace eqh

extern "C" {

int eqh_knn(const float* pos, int64_t N, int k, const int32_t* seg_ptr, const int32_t* seg_of, int exclude_self,
            int32_t* idx, float* d2, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::knng;
  EQH_CHECK_ARG(N >= 0 && N < ((int64_t)1 << 31) && (N == 0 || (pos && idx && d2)));
  EQH_CHECK_ARG((seg_ptr == nullptr) == (seg_of == nullptr));
  if (k < 1 || k > 32) return EQH_ERR_UNSUPPORTED;
  if (N == 0) return EQH_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool seg = seg_ptr != nullptr;
#define EQH_KNN_LAUNCH(KMAX, S, SEGV)                                                                            \
  launch_kernel(k_knn<KMAX, S, SEGV>, dim3((unsigned)((N + 256 / S - 1) / (256 / S))), dim3(256), 0, st, pos, N, k, \
                seg_ptr, seg_of, exclude_self, idx, d2)
  if (seg) {                                 // a molecule has tens to a few hundred atoms: 8 lanes per query
    if (k <= 16) EQH_KNN_LAUNCH(16, 8, true);
    else EQH_KNN_LAUNCH(32, 8, true);
  } else if (N >= 16384) {
    if (k <= 16) EQH_KNN_LAUNCH(16, 8, false);
    else EQH_KNN_LAUNCH(32, 8, false);
  } else {
    if (k <= 16) EQH_KNN_LAUNCH(16, 32, false);
    else EQH_KNN_LAUNCH(32, 32, false);
  }
#undef EQH_KNN_LAUNCH
  EQH_LAUNCH_RC();
}

int eqh_radius(const float* pos, int64_t N, float r, const int32_t* seg_ptr, const int32_t* seg_of, int loop,
               int max_num_neighbors, int32_t* idx, int32_t* count, float* d2, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::knng;
  EQH_CHECK_ARG(N >= 0 && N < ((int64_t)1 << 31) && r >= 0.f && max_num_neighbors >= 1 && (N == 0 || (pos && idx && count)));
  EQH_CHECK_ARG((seg_ptr == nullptr) == (seg_of == nullptr));
  if (N == 0) return EQH_OK;
  const float r2 = r * r;
  launch_kernel(k_radius, dim3((unsigned)((N + 7) / 8)), dim3(256), 0, static_cast<cudaStream_t>(stream), pos, N, r2, seg_ptr,
                seg_of, loop, max_num_neighbors, idx, count, d2);
  EQH_LAUNCH_RC();
}

}  // extern "C"
