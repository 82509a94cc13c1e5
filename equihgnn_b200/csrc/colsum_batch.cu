// Batched deterministic column sums: every bias gradient (nn.Linear, mlp.py:91-99) and every per-CTA partial
// buffer (LayerNorm affine gradients, phi_e weight gradients) of one layer's backward finished in ONE launch.
//
//     out_j[c] (+)= scale_j * sum_r roww_j[r] * x_j[r*ld_j + c]          j = 0 .. n_jobs-1
//
// Each job is cut into (32-column tile) x (row chunk) blocks.  A block reduces its chunk (8 interleaved row
// groups, fixed order) into ws; the block that arrives LAST at the tile's counter adds the chunk partials in
// ascending chunk order and writes the result -- the summation order never depends on which block is last, so
// reruns are bit-identical.  The last block resets the counter, so the caller zeroes `counters` only once.
#include "common.cuh"

namespace eqh {

constexpr int kMaxJobs = EQH_COLSUM_MAX_JOBS;
constexpr int kMaxChunks = 256;

struct JobPack {
  eqh_colsum_job job[kMaxJobs];
  int32_t block0[kMaxJobs + 1];   // first block of each job
  int32_t chunks[kMaxJobs];
  int32_t rpc[kMaxJobs];          // rows per chunk
  int32_t ws_off[kMaxJobs];       // floats
  int32_t cnt_off[kMaxJobs];
  int32_t n;
};

static inline int64_t chunks_for(int64_t rows) {
  int64_t c = (rows + 63) / 64;
  return c < 1 ? 1 : (c > kMaxChunks ? kMaxChunks : c);
}

__global__ void __launch_bounds__(256) k_colsum_batch(const __grid_constant__ JobPack p, float* __restrict__ ws,
                                                       int32_t* __restrict__ counters) {
  pdl_prologue();
  __shared__ float sm[8][32];
  __shared__ int s_last;
  int j = 0;
  while (j + 1 < p.n && (int)blockIdx.x >= p.block0[j + 1]) ++j;
  const eqh_colsum_job& J = p.job[j];
  const int chunks = p.chunks[j];
  const int local = blockIdx.x - p.block0[j];
  const int tile = local / chunks, chunk = local - tile * chunks;
  const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int64_t c = (int64_t)tile * 32 + cx;
  const int64_t r0 = (int64_t)chunk * p.rpc[j], r1 = min(J.rows, r0 + p.rpc[j]);
  float v = 0.f;
  if (c < J.cols) {
    const float* x = J.x + c;
    if (J.roww) {
      for (int64_t r = r0 + rg; r < r1; r += 8) v += __ldg(J.roww + r) * __ldg(x + r * J.ld);
    } else {
      for (int64_t r = r0 + rg; r < r1; r += 8) v += __ldg(x + r * J.ld);
    }
  }
  sm[rg][cx] = v;
  __syncthreads();
  float t = 0.f;
  if (rg == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][cx];
  }
  if (chunks == 1) {
    if (rg == 0 && c < J.cols) J.out[c] = J.accumulate ? J.out[c] + J.scale * t : J.scale * t;
    return;
  }
  float* part = ws + p.ws_off[j] + (int64_t)tile * chunks * 32;     // [chunks][32] per tile
  if (rg == 0) part[chunk * 32 + cx] = t;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int prev = atomicAdd(counters + p.cnt_off[j] + tile, 1);
    s_last = (prev == chunks - 1);
    if (s_last) counters[p.cnt_off[j] + tile] = 0;    // self-reset for the next launch on this stream
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float a = 0.f;
  for (int k = rg; k < chunks; k += 8) a += __ldcg(part + k * 32 + cx);
  sm[rg][cx] = a;
  __syncthreads();
  if (rg == 0 && c < J.cols) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += sm[w][cx];
    J.out[c] = J.accumulate ? J.out[c] + J.scale * s : J.scale * s;
  }
}

}  // namespace eqh

extern "C" {

size_t eqh_colsum_batch_workspace_bytes(const eqh_colsum_job* jobs, int n_jobs) {
  size_t fl = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const int64_t ch = eqh::chunks_for(jobs[j].rows);
    if (ch > 1) fl += (size_t)((jobs[j].cols + 31) / 32) * (size_t)ch * 32;
  }
  return fl * sizeof(float);
}

int64_t eqh_colsum_batch_counters(const eqh_colsum_job* jobs, int n_jobs) {
  int64_t n = 0;
  for (int j = 0; j < n_jobs; ++j) n += (jobs[j].cols + 31) / 32;
  return n;
}

int eqh_colsum_batch(const eqh_colsum_job* jobs, int n_jobs, float* ws, int32_t* counters, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(n_jobs >= 0 && n_jobs <= kMaxJobs && (n_jobs == 0 || jobs));
  JobPack p;
  int64_t blocks = 0, ws_off = 0, cnt_off = 0;
  int n = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const eqh_colsum_job& J = jobs[j];
    EQH_CHECK_ARG(J.rows >= 0 && J.cols >= 0 && J.ld >= J.cols && (J.cols == 0 || (J.x && J.out)));
    if (J.cols == 0) continue;
    const int64_t ch = chunks_for(J.rows), tiles = (J.cols + 31) / 32;
    p.job[n] = J;
    p.block0[n] = (int32_t)blocks;
    p.chunks[n] = (int32_t)ch;
    p.rpc[n] = (int32_t)((J.rows + ch - 1) / ch);
    p.ws_off[n] = (int32_t)ws_off;
    p.cnt_off[n] = (int32_t)cnt_off;
    blocks += ch * tiles;
    if (ch > 1) ws_off += tiles * ch * 32;
    cnt_off += tiles;
    EQH_CHECK_ARG(blocks < (1ll << 30) && ws_off < (1ll << 31));
    if (ch > 1) EQH_CHECK_ARG(ws && counters);
    ++n;
  }
  if (n == 0) return EQH_OK;
  p.block0[n] = (int32_t)blocks;
  p.n = n;
  launch_kernel(k_colsum_batch, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), p, ws, counters);
  EQH_LAUNCH_RC();
}

}  // extern "C"
