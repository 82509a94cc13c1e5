// Incidence CSR build: stable counting sort by destination, on device, int32 outputs.
// HBM-bound integer work: every pass is a coalesced stream over the Z incidences plus
// L2-resident histogram traffic.  The result (rowptr, perm) is unique, so the integer atomics
// used for counting / slot claiming do not affect determinism: each segment is put into
// ascending incidence order afterwards (== torch.sort(stable=True) by destination).
#include "common.cuh"

namespace eqh {

template <typename IdxT>
__global__ void k_hist(const IdxT* __restrict__ dst, int64_t Z, int64_t n_dst, int32_t* __restrict__ cnt,
                       int32_t* __restrict__ dst32, int32_t* __restrict__ status) {
  pdl_prologue();
  int64_t z = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (z >= Z) return;
  int64_t d = (int64_t)dst[z];
  if (d < 0 || d >= n_dst) {
    if (status) atomicExch(status, 1);
    if (dst32) dst32[z] = -1;
    return;
  }
  if (dst32) dst32[z] = (int32_t)d;
  atomicAdd(cnt + d, 1);
}

// exclusive scan of cnt[0..n) -> rowptr[0..n], cursor[0..n) ; one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_scan(const int32_t* cnt, int64_t n, int32_t* rowptr, int32_t* cursor /* may alias cnt */) {
  pdl_prologue();
  __shared__ int32_t part[1024];
  const int t = threadIdx.x;
  const int64_t per = (n + 1023) / 1024;
  const int64_t lo = t * per, hi = min(n, lo + per);
  int32_t s = 0;
  for (int64_t i = lo; i < hi; ++i) s += cnt[i];
  part[t] = s;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int o = 1; o < 1024; o <<= 1) {
    int32_t v = (t >= o) ? part[t - o] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int32_t run = (t == 0) ? 0 : part[t - 1];
  for (int64_t i = lo; i < hi; ++i) {
    int32_t c = cnt[i];
    rowptr[i] = run;
    cursor[i] = run;
    run += c;
  }
  if (t == 1023) rowptr[n] = part[1023];
}

// Multi-CTA exclusive scan for long count arrays (the single CTA above walks n / 1024 counts per thread serially: 98 us for
// the transposed kNN incidence of a Molecule3D-1024 batch).  Two launches: every CTA sums its contiguous chunk, then every
// CTA adds up the (<= kScanBlocks) chunk sums before it and scans its own chunk with a block-wide scan.  Integer sums:
// the result does not depend on the order.
constexpr int kScanBlocks = 2 * kSMs;

__global__ void __launch_bounds__(256) k_scan_partial(const int32_t* __restrict__ cnt, int64_t n, int64_t per,
                                                       int32_t* __restrict__ block_sum) {
  pdl_prologue();
  __shared__ int32_t red[8];
  const int64_t lo = blockIdx.x * per, hi = min(n, lo + per);
  int32_t s = 0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) s += cnt[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int w = 0; w < 8; ++w) t += red[w];
    block_sum[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) k_scan_apply(const int32_t* cnt, int64_t n, int64_t per,
                                                     const int32_t* __restrict__ block_sum, int nblocks, int32_t* rowptr,
                                                     int32_t* cursor /* may alias cnt */) {
  pdl_prologue();
  __shared__ int32_t warp_tot[8];
  __shared__ int32_t base_s;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (wib == 0) {                                    // offset of this chunk: sum of the chunk sums before it
    int32_t b = 0;
    for (int k = lane; k < (int)blockIdx.x; k += 32) b += __ldg(block_sum + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(kFull, b, o);
    if (lane == 0) base_s = b;
  }
  __syncthreads();
  int32_t run = base_s;
  const int64_t lo = blockIdx.x * per, hi = min(n, lo + per);
  for (int64_t i0 = lo; i0 < hi; i0 += 256) {
    const int64_t i = i0 + threadIdx.x;
    const int32_t c = i < hi ? cnt[i] : 0;
    int32_t inc = c;                                 // inclusive scan inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t v = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) warp_tot[wib] = inc;
    __syncthreads();
    int32_t wbase = 0, tile = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      const int32_t t = warp_tot[w];
      if (w < wib) wbase += t;
      tile += t;
    }
    if (i < hi) {
      const int32_t ex = run + wbase + inc - c;
      rowptr[i] = ex;
      cursor[i] = ex;
    }
    run += tile;
    __syncthreads();
  }
  if (blockIdx.x == nblocks - 1 && threadIdx.x == 0) rowptr[n] = run;
}

template <typename IdxT>
__global__ void k_place(const IdxT* __restrict__ dst, int64_t Z, int64_t n_dst, int32_t* __restrict__ cursor,
                        int32_t* __restrict__ slot_z) {
  pdl_prologue();
  int64_t z = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (z >= Z) return;
  int64_t d = (int64_t)dst[z];
  if (d < 0 || d >= n_dst) return;
  int32_t p = atomicAdd(cursor + d, 1);
  slot_z[p] = (int32_t)z;
}

constexpr int kOrderStage = 1024;     // segment length up to which k_order stages the segment in shared memory

// one warp per segment: order the segment's incidence ids ascending by counting ranks.
template <typename IdxT>
__global__ void k_order(const int32_t* __restrict__ rowptr, int64_t n_dst, const int32_t* __restrict__ slot_z,
                        const IdxT* __restrict__ other, int32_t* __restrict__ perm, int32_t* __restrict__ other_sorted) {
  pdl_prologue();
  __shared__ int32_t order_stage[8][kOrderStage];
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < n_dst; s += nwarps) {
    const int32_t lo = rowptr[s], hi = rowptr[s + 1];
    const int32_t len = hi - lo;
    if (len <= 32) {
      int32_t mine = (lane < len) ? slot_z[lo + lane] : 0x7fffffff;
      int32_t rank = 0;
#pragma unroll 4
      for (int k = 0; k < len; ++k) rank += (__shfl_sync(kFull, mine, k) < mine);
      if (lane < len) {
        perm[lo + rank] = mine;
        if (other_sorted) other_sorted[lo + rank] = (int32_t)other[mine];
      }
    } else if (len <= kOrderStage) {
      // long segment (a conjugated hyperedge of an OPV polymer, a hub atom of the transposed kNN incidence, the ghost
      // segment of a padded batch): staged once in shared memory, ranks counted from there (broadcast reads) instead of
      // len^2 global loads
      int32_t* stage = order_stage[threadIdx.x >> 5];
      __syncwarp();
      for (int32_t a = lane; a < len; a += 32) stage[a] = slot_z[lo + a];
      __syncwarp();
      for (int32_t a = lane; a < len; a += 32) {
        const int32_t mine = stage[a];
        int32_t rank = 0;
#pragma unroll 8
        for (int32_t b = 0; b < len; ++b) rank += (stage[b] < mine);
        perm[lo + rank] = mine;
        if (other_sorted) other_sorted[lo + rank] = (int32_t)other[mine];
      }
    } else {
      for (int32_t a = lo + lane; a < hi; a += 32) {
        const int32_t mine = slot_z[a];
        int32_t rank = 0;
        for (int32_t b = lo; b < hi; ++b) rank += (slot_z[b] < mine);
        perm[lo + rank] = mine;
        if (other_sorted) other_sorted[lo + rank] = (int32_t)other[mine];
      }
    }
  }
}

template <typename IdxT>
static int build(const IdxT* dst, const IdxT* other, int64_t Z, int64_t n_dst, int32_t* rowptr, int32_t* perm,
                 int32_t* dst32, int32_t* other_sorted, int32_t* status, void* ws, size_t ws_bytes,
                 cudaStream_t st) {
  EQH_CHECK_ARG(Z >= 0 && n_dst >= 0 && rowptr && (Z == 0 || (dst && perm)));
  EQH_CHECK_ARG(Z < (int64_t)1 << 31 && n_dst < (int64_t)1 << 31);
  EQH_CHECK_ARG(!(other_sorted && !other) || Z == 0);
  if (ws_bytes < eqh_incidence_workspace_bytes(Z, n_dst) || (!ws && ws_bytes)) return EQH_ERR_WORKSPACE;
  int32_t* cnt = static_cast<int32_t*>(ws);        // [n_dst + 1]  (counts, then cursors)
  int32_t* slot_z = cnt + (n_dst + 1);             // [Z]
  cudaError_t e = cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)(n_dst + 1), st);
  if (e != cudaSuccess) return (int)e;
  const int T = 256;
  if (Z > 0) launch_kernel(k_hist<IdxT>, dim3((unsigned)((Z + T - 1) / T)), dim3(T), 0, st, dst, Z, n_dst, cnt, dst32, status);
  // in-place: the scan reads cnt[i] before overwriting cursor[i] (same thread, same i)
  if (n_dst <= 16384) {
    launch_kernel(k_scan, dim3(1), dim3(1024), 0, st, cnt, n_dst, rowptr, cnt);
  } else {
    int32_t* block_sum = slot_z + Z;                 // [kScanBlocks] behind the slot list
    int64_t per = (n_dst + kScanBlocks - 1) / kScanBlocks;
    per = (per + 255) / 256 * 256;
    const int nblocks = (int)((n_dst + per - 1) / per);
    launch_kernel(k_scan_partial, dim3((unsigned)nblocks), dim3(256), 0, st, cnt, n_dst, per, block_sum);
    launch_kernel(k_scan_apply, dim3((unsigned)nblocks), dim3(256), 0, st, cnt, n_dst, per, block_sum, nblocks, rowptr, cnt);
  }
  if (Z > 0) {
    launch_kernel(k_place<IdxT>, dim3((unsigned)((Z + T - 1) / T)), dim3(T), 0, st, dst, Z, n_dst, cnt, slot_z);
    int64_t warps = n_dst;
    int64_t blocks = (warps * 32 + T - 1) / T;
    if (blocks > kSMs * 16) blocks = kSMs * 16;
    if (blocks < 1) blocks = 1;
    launch_kernel(k_order<IdxT>, dim3((unsigned)blocks), dim3(T), 0, st, rowptr, n_dst, slot_z, other, perm, other_sorted);
  }
  EQH_LAUNCH_RC();
}

// global_add_pool over a PyG batch: `batch` is sorted, so the CSR of the pooling is `ptr` itself with the identity
// permutation -- only the int64 -> int32 narrowing is left (one launch instead of the 4-kernel counting sort).
__global__ void __launch_bounds__(256) k_pool_index(const int64_t* __restrict__ ptr, int64_t nb1, const int64_t* __restrict__ batch,
                                                     int64_t n, int32_t* __restrict__ rowptr32, int32_t* __restrict__ batch32) {
  pdl_prologue();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < nb1) rowptr32[t] = (int32_t)ptr[t];
  if (t < n) batch32[t] = (int32_t)batch[t];
}

}  // namespace eqh

extern "C" {

int eqh_pool_index(const int64_t* ptr, int64_t num_graphs, const int64_t* batch, int64_t n, int32_t* rowptr32,
                   int32_t* batch32, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(ptr && batch && rowptr32 && batch32 && num_graphs >= 0 && n >= 0 && n < ((int64_t)1 << 31));
  const int64_t m = (num_graphs + 1 > n ? num_graphs + 1 : n);
  launch_kernel(k_pool_index, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), ptr,
                num_graphs + 1, batch, n, rowptr32, batch32);
  EQH_LAUNCH_RC();
}

size_t eqh_incidence_workspace_bytes(int64_t Z, int64_t n_dst) {
  if (Z < 0 || n_dst < 0) return 0;
  return sizeof(int32_t) * (size_t)(n_dst + 1 + Z + eqh::kScanBlocks);
}

int eqh_incidence_build_i64(const int64_t* dst, const int64_t* other, int64_t Z, int64_t n_dst, int32_t* rowptr,
                            int32_t* perm, int32_t* dst32, int32_t* other_sorted, int32_t* status, void* ws,
                            size_t ws_bytes, eqh_stream_t stream) {
  return eqh::build<int64_t>(dst, other, Z, n_dst, rowptr, perm, dst32, other_sorted, status, ws, ws_bytes,
                             static_cast<cudaStream_t>(stream));
}

int eqh_incidence_build_i32(const int32_t* dst, const int32_t* other, int64_t Z, int64_t n_dst, int32_t* rowptr,
                            int32_t* perm, int32_t* dst32, int32_t* other_sorted, int32_t* status, void* ws,
                            size_t ws_bytes, eqh_stream_t stream) {
  return eqh::build<int32_t>(dst, other, Z, n_dst, rowptr, perm, dst32, other_sorted, status, ws, ws_bytes,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
