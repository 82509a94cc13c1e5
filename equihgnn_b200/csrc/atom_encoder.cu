// AtomEncoder (ogb 1.3.6, third-party): h[n,:] = sum_f table_f[x[n,f], :] over the 9 categorical
// atom features (reference call sites mhnn.py:201, equihnn_egnn.py:157), forward and backward, as
// two kernels instead of 17 ATen launches forward and 9 atomic embedding_backward launches.
//
// forward : one warp per atom, tables summed in feature order (same association as the
//           reference's running `x_embedding + emb(x[:,i])`, so fp32 results are bit-identical).
// backward: dtable_f[v,:] = sum over atoms n with x[n,f]==v of dh[n,:].  Segments are long and
//           skewed (a binary feature has N/2 rows per value), so the atoms are cut into S slices;
//           one warp owns (feature, 32-column block, slice), walks its atoms in ascending order and
//           accumulates into a private shared-memory table acc[v][lane] (no atomics, fixed order);
//           slice partials are combined in slice order by eqh_colsum.  Deterministic.
#include "common.cuh"

namespace eqh {

constexpr int kMaxFeat = 16;
struct TablePtrs { const float* p[kMaxFeat]; };
struct FeatMeta { int32_t rows[kMaxFeat]; int32_t off[kMaxFeat]; };   // rows per table, row offset in the concatenation

template <int V>
__global__ void __launch_bounds__(256) k_atom_fwd(const int64_t* __restrict__ x, TablePtrs tabs, FeatMeta meta, int F,
                                                   int64_t N, float* __restrict__ out, int32_t* __restrict__ status) {
  pdl_prologue();
  constexpr int C = 32 * V;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < N; n += nwarps) {
    Row<V> acc;
#pragma unroll
    for (int f = 0; f < kMaxFeat; ++f) {   // unrolled so the by-value pointer table stays in the constant bank
      if (f >= F) break;
      int64_t v = __ldg(x + n * F + f);
      if (v < 0 || v >= meta.rows[f]) {
        if (status && lane == 0) atomicExch(status, 1);
        v = 0;
      }
      Row<V> t;
      t.load(tabs.p[f] + v * C, lane);
      if (f == 0) acc = t;
      else {
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] += t.v[i];
      }
    }
    acc.store(out + n * C, lane);
  }
}

// grid: (slices, C/32, F); block: 32 threads (one warp); dynamic smem: max_rows * 32 floats
__global__ void __launch_bounds__(32) k_atom_bwd(const int64_t* __restrict__ x, const float* __restrict__ dh, FeatMeta meta,
                                                  int F, int64_t N, int C, int64_t slice, int total_rows,
                                                  float* __restrict__ partial) {
  pdl_prologue();
  extern __shared__ float acc[];   // [rows_f][32]
  const int lane = threadIdx.x;
  const int f = blockIdx.z, cb = blockIdx.y;
  const int rows = meta.rows[f];
  for (int r = 0; r < rows; ++r) acc[r * 32 + lane] = 0.f;
  __syncwarp();
  const int64_t lo = (int64_t)blockIdx.x * slice, hi = min(N, lo + slice);
  const float* g = dh + cb * 32 + lane;
  int64_t n = lo;
  for (; n + 16 <= hi; n += 16) {   // 16 rows in flight: the kernel sits on the tail of the backward, latency is what counts
    int64_t v[16];
    float d[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = __ldg(x + (n + u) * F + f);
#pragma unroll
    for (int u = 0; u < 16; ++u) d[u] = __ldg(g + (n + u) * C);
#pragma unroll
    for (int u = 0; u < 16; ++u)
      if (v[u] >= 0 && v[u] < rows) acc[(int)v[u] * 32 + lane] += d[u];   // ascending n: fixed order
  }
  for (; n + 4 <= hi; n += 4) {
    int64_t v[4];
    float d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(x + (n + u) * F + f);
#pragma unroll
    for (int u = 0; u < 4; ++u) d[u] = __ldg(g + (n + u) * C);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (v[u] >= 0 && v[u] < rows) acc[(int)v[u] * 32 + lane] += d[u];
  }
  for (; n < hi; ++n) {
    const int64_t v = __ldg(x + n * F + f);
    if (v >= 0 && v < rows) acc[(int)v * 32 + lane] += __ldg(g + n * C);
  }
  __syncwarp();
  float* dst = partial + ((int64_t)blockIdx.x * total_rows + meta.off[f]) * C + cb * 32 + lane;
  for (int r = 0; r < rows; ++r) dst[(int64_t)r * C] = acc[r * 32 + lane];
}

}  // namespace eqh

extern "C" {

int eqh_atom_encoder_slices(int64_t N) {
  if (N <= 0) return 1;
  int64_t slice = N / 64 > 64 ? N / 64 : 64;
  return (int)((N + slice - 1) / slice);
}

int eqh_atom_encoder_fwd(const int64_t* x, const float* const* tables, const int32_t* rows_per_table, int F, int64_t N,
                         int64_t C, float* out, int32_t* status, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(tables && rows_per_table && F >= 1 && F <= kMaxFeat && N >= 0 && (N == 0 || (x && out)));
  if (!supported_C(C)) return EQH_ERR_UNSUPPORTED;
  if (N == 0) return EQH_OK;
  TablePtrs tp;
  FeatMeta meta;
  int off = 0;
  for (int f = 0; f < F; ++f) {
    EQH_CHECK_ARG(tables[f] && aligned16(tables[f]) && rows_per_table[f] > 0);
    tp.p[f] = tables[f];
    meta.rows[f] = rows_per_table[f];
    meta.off[f] = off;
    off += rows_per_table[f];
  }
  int64_t blocks = (N + 7) / 8;
  if (blocks > (int64_t)kSMs * 16) blocks = (int64_t)kSMs * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EQH_DISPATCH_V(C, { launch_kernel(k_atom_fwd<V>, dim3((unsigned)blocks), dim3(256), 0, st, x, tp, meta, F, N, out, status); });
  EQH_LAUNCH_RC();
}

int eqh_atom_encoder_bwd(const int64_t* x, const float* dh, const int32_t* rows_per_table, int F, int64_t N, int64_t C,
                         float* partial, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(rows_per_table && F >= 1 && F <= kMaxFeat && N >= 0 && partial && (N == 0 || (x && dh)));
  if (C <= 0 || C % 32 != 0) return EQH_ERR_UNSUPPORTED;
  FeatMeta meta;
  int off = 0, max_rows = 0;
  for (int f = 0; f < F; ++f) {
    EQH_CHECK_ARG(rows_per_table[f] > 0);
    meta.rows[f] = rows_per_table[f];
    meta.off[f] = off;
    off += rows_per_table[f];
    if (rows_per_table[f] > max_rows) max_rows = rows_per_table[f];
  }
  const size_t smem = sizeof(float) * 32 * (size_t)max_rows;
  if (smem > 48 * 1024) return EQH_ERR_UNSUPPORTED;   // tables with more than 384 rows are not implemented
  const int S = eqh_atom_encoder_slices(N);
  const int64_t slice = (N + S - 1) / S > 0 ? (N + S - 1) / S : 1;
  dim3 grid((unsigned)S, (unsigned)(C / 32), (unsigned)F);
  launch_kernel(k_atom_bwd, dim3(grid), dim3(32), smem, static_cast<cudaStream_t>(stream), x, dh, meta, F, N, (int)C, slice, off, partial);
  EQH_LAUNCH_RC();
}

}  // extern "C"
