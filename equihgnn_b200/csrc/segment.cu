// Deterministic segmented reductions over CSR-grouped rows (no atomics) and the matching row
// gather.  HBM-bound: one warp streams one segment with 128-bit coalesced loads (a 512 B
// contiguous run per warp instruction for C >= 128), four rows in flight per warp.
//
// Algorithmic bytes per call (figure P): 4C(Z + n_out) + 4(Z + n_out + 1).
#include "common.cuh"

namespace eqh {

template <int V, bool WEIGHTED>
__global__ void __launch_bounds__(256) k_gather_reduce(const float* __restrict__ table, const int32_t* __restrict__ idx,
                                                        const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ src_rowptr, int64_t n_out, int mode,
                                                        float* __restrict__ out, int64_t ld, int64_t ld_out) {
  pdl_prologue();
  table += (int64_t)blockIdx.y * (32 * V);  // column block of a wider row (row stride ld)
  out += (int64_t)blockIdx.y * (32 * V);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < n_out; s += nwarps) {
    const int32_t lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
    Row<V> acc;
    acc.zero();
    int32_t k = lo;
    for (; k + 4 <= hi; k += 4) {
      int32_t r[4];
      float w[4];
      Row<V> x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] = idx ? __ldg(idx + k + u) : (k + u);
      // a negative row id (an out-of-range index the incidence builder marked with -1) contributes nothing
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r[u] >= 0) x[u].load(table + (int64_t)r[u] * ld, lane);
        else x[u].zero();
      }
      if constexpr (WEIGHTED) {
#pragma unroll
        for (int u = 0; u < 4; ++u) w[u] = r[u] >= 0 ? inv_count(src_rowptr, r[u]) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {  // strictly ascending k: same association as a sequential scatter_add_
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] += WEIGHTED ? w[u] * x[u].v[i] : x[u].v[i];
      }
    }
    for (; k < hi; ++k) {
      const int32_t r = idx ? __ldg(idx + k) : k;
      if (r < 0) continue;
      Row<V> x;
      x.load(table + (int64_t)r * ld, lane);
      const float w = WEIGHTED ? inv_count(src_rowptr, r) : 1.f;
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] += WEIGHTED ? w * x.v[i] : x.v[i];
    }
    if (mode == EQH_REDUCE_MEAN) {
      const float cnt = (float)max(hi - lo, 1);
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = acc.v[i] / cnt;  // true division, as torch_scatter's true_divide_
    }
    acc.store(out + s * ld_out, lane);
  }
}

// Wide rows whose width is a multiple of 4 but not of 128 (the phi_e hidden width 1088 = 8.5 x 128): column blocks of 128
// floats, one float4 per lane (a 512-byte run per warp instruction instead of the 256 bytes of the 64-column blocks the
// divisor rule would pick), lanes past the row's end idle.  Unweighted sum only (the dQ reduction of the EGNN backward).
__global__ void __launch_bounds__(256) k_gather_reduce_wide(const float* __restrict__ table, const int32_t* __restrict__ idx,
                                                            const int32_t* __restrict__ rowptr, int64_t n_out, int64_t C,
                                                            float* __restrict__ out, int64_t ld, int64_t ld_out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t col = (int64_t)blockIdx.y * 128 + lane * 4;
  const bool live = col < C;                    // (C % 4 == 0: a float4 is entirely inside or outside)
  table += col;
  out += col;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < n_out; s += nwarps) {
    const int32_t lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int32_t k = lo;
    for (; k + 8 <= hi; k += 8) {
      int32_t r[8];
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = idx ? __ldg(idx + k + u) : (k + u);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        x[u] = (live && r[u] >= 0) ? __ldg(reinterpret_cast<const float4*>(table + (int64_t)r[u] * ld)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 8; ++u) {             // strictly ascending k: same association as a sequential scatter_add_
        acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w;
      }
    }
    for (; k < hi; ++k) {
      const int32_t r = idx ? __ldg(idx + k) : k;
      if (!live || r < 0) continue;
      const float4 x = __ldg(reinterpret_cast<const float4*>(table + (int64_t)r * ld));
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    if (live) *reinterpret_cast<float4*>(out + s * ld_out) = acc;
  }
}

template <int V>
__global__ void __launch_bounds__(256) k_gather_rows(const float* __restrict__ table, const int32_t* __restrict__ idx,
                                                      const int32_t* __restrict__ scale_rowptr, int64_t n_rows,
                                                      float* __restrict__ out, int64_t ld) {
  pdl_prologue();
  table += (int64_t)blockIdx.y * (32 * V);
  out += (int64_t)blockIdx.y * (32 * V);
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp * 2; r0 < n_rows; r0 += nwarps * 2) {
    const bool two = r0 + 1 < n_rows;
    const int32_t s0 = idx ? __ldg(idx + r0) : (int32_t)r0;
    const int32_t s1 = two ? (idx ? __ldg(idx + r0 + 1) : (int32_t)(r0 + 1)) : s0;
    Row<V> a, b;
    // a negative source id (out-of-range index dropped by the incidence builder) yields a zero row, as the forward
    // left that row out of every segment
    if (s0 >= 0) a.load(table + (int64_t)s0 * ld, lane); else a.zero();
    if (s1 >= 0) b.load(table + (int64_t)s1 * ld, lane); else b.zero();
    const float w0 = s0 >= 0 ? inv_count(scale_rowptr, s0) : 0.f, w1 = s1 >= 0 ? inv_count(scale_rowptr, s1) : 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { a.v[i] *= w0; b.v[i] *= w1; }
    a.store(out + r0 * ld, lane);
    if (two) b.store(out + (r0 + 1) * ld, lane);
  }
}

// widest supported column block (32*V, V in {16,8,4,2,1}) dividing C; 0 if C % 32 != 0
static int64_t col_block(int64_t C) {
  for (int64_t cb = 512; cb >= 32; cb >>= 1)
    if (C % cb == 0) return cb;
  return 0;
}

static unsigned grid_for_warps(int64_t warps_needed, int threads) {
  int64_t per = threads / 32;
  int64_t blocks = (warps_needed + per - 1) / per;
  const int64_t cap = (int64_t)kSMs * 8 * 4;  // 8 CTAs of 256 threads per SM resident; allow 4 waves before striding
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace eqh

extern "C" {

int eqh_gather_reduce(const float* table, const int32_t* idx, const int32_t* rowptr, const int32_t* src_rowptr,
                      int64_t n_out, int64_t C, int mode, float* out, eqh_stream_t stream) {
  return eqh_gather_reduce_ld(table, C, idx, rowptr, src_rowptr, n_out, C, mode, out, C, stream);
}

int eqh_gather_reduce_ld(const float* table, int64_t ld_table, const int32_t* idx, const int32_t* rowptr,
                         const int32_t* src_rowptr, int64_t n_out, int64_t C, int mode, float* out, int64_t ld_out,
                         eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(ld_table >= C && ld_out >= C && ld_table % 4 == 0 && ld_out % 4 == 0);
  EQH_CHECK_ARG(n_out >= 0 && rowptr && (n_out == 0 || out));  // table may be NULL when every segment is empty
  EQH_CHECK_ARG(mode == EQH_REDUCE_SUM || mode == EQH_REDUCE_MEAN);
  EQH_CHECK_ARG(aligned16(table) && aligned16(out));
  const int64_t CB = col_block(C);
  if (C <= 0 || CB == 0 || C / CB > 65535) return EQH_ERR_UNSUPPORTED;
  if (n_out == 0) return EQH_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (CB < 128 && C > 128 && C % 4 == 0 && !src_rowptr && mode == EQH_REDUCE_SUM) {
    const dim3 gridw(grid_for_warps(n_out, 256), (unsigned)((C + 127) / 128));
    launch_kernel(k_gather_reduce_wide, dim3(gridw), dim3(256), 0, st, table, idx, rowptr, n_out, C, out, ld_table, ld_out);
    EQH_LAUNCH_RC();
  }
  const dim3 grid(grid_for_warps(n_out, 256), (unsigned)(C / CB));
  EQH_DISPATCH_V(CB, {
    if (src_rowptr)
      launch_kernel(k_gather_reduce<V, true>, dim3(grid), dim3(256), 0, st, table, idx, rowptr, src_rowptr, n_out, mode, out, ld_table, ld_out);
    else
      launch_kernel(k_gather_reduce<V, false>, dim3(grid), dim3(256), 0, st, table, idx, rowptr, nullptr, n_out, mode, out, ld_table, ld_out);
  });
  EQH_LAUNCH_RC();
}

int eqh_gather_rows(const float* table, const int32_t* idx, const int32_t* scale_rowptr, int64_t n_rows, int64_t C,
                    float* out, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(n_rows >= 0 && (n_rows == 0 || (out && table)));
  EQH_CHECK_ARG(aligned16(table) && aligned16(out));
  const int64_t CB = col_block(C);
  if (C <= 0 || CB == 0 || C / CB > 65535) return EQH_ERR_UNSUPPORTED;
  if (n_rows == 0) return EQH_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(grid_for_warps((n_rows + 1) / 2, 256), (unsigned)(C / CB));
  EQH_DISPATCH_V(CB, { launch_kernel(k_gather_rows<V>, dim3(grid), dim3(256), 0, st, table, idx, scale_rowptr, n_rows, out, C); });
  EQH_LAUNCH_RC();
}

}  // extern "C"
