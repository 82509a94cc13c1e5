// Shared by the two tcgen05 3xTF32 GEMM kernels (gemm_tf32x3.cu: both operands from shared memory;
// gemm_ts.cu: A operand through registers into tensor memory): parameters, the epilogue and the split-K finish.
#pragma once
#include "tc_common.cuh"

namespace eqh {
namespace gemm {
using namespace eqh::tc;


constexpr int BM = 128, BK = 32;

#ifdef EQH_TRACE
// phase timestamps (ns, globaltimer) of CTA (0,0,0) of the most recent launch: scratch/trace_gemm.py
static __device__ unsigned long long g_trace[80];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define EQH_STAMP(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_trace[i] = gtime(); } while (0)
#else
#define EQH_STAMP(i) do { } while (0)
#endif
constexpr int kThreads = 320;           // warp 0 TMA, warp 1 MMA, warps 2-9 operand split (warps 4-7 also epilogue)
constexpr int kSplitThreads = 256;
constexpr uint32_t kTileBytes = BM * BK * 4;  // 16 KB: a 128 x 32 fp32 operand tile (any major)

// Accumulator assignment: three [128 x BN] fp32 tiles in tensor memory.  The tensor core truncates when it adds into
// its accumulator, so the error grows with the number of accumulations per tile: the big hi.hi products alternate
// between two tiles (even / odd k-steps) and the 2^-11-scaled lo.hi and hi.lo products share a third; the epilogue
// adds the three.  (Spreading the 12 products of a K block over 6 or 12 tiles was measured: the main loop is not
// limited by accumulator dependencies, and every extra tile costs 128*BN*4 bytes of TMEM reads -- 64 B/cycle -- in
// the epilogue.)
template <int BN>
struct AccMap {
  static constexpr int NA = 3;
  __device__ static __forceinline__ int hh(int ks) { return ks & 1; }
  __device__ static __forceinline__ int lh(int ks) { return 2; }
  __device__ static __forceinline__ int hl(int ks) { return 2; }
  __device__ static __forceinline__ bool first_big(int kb, int ks) { return kb == 0 && ks < 2; }
  __device__ static __forceinline__ bool first_lh(int kb, int ks) { return kb == 0 && ks == 0; }
  __device__ static __forceinline__ bool first_hl(int kb, int ks) { return false; }
};

struct Params {
  float* D;
  const float* bias;
  const float* rowscale;   // optional [M]: the bias is multiplied by rowscale[m] (empty-segment mask / segment count)
  const float* residual;   // optional [M, ldd]; may alias D (in-place accumulation)
  float out_scale, res_scale;
  float* ws;          // split-K partials [splits, M, N] (ldd == N) or nullptr
  int32_t* counters;  // split-K: one arrival counter per output tile (zero on entry, left zero)
  int splits;
  int64_t ldd;
  int M, N, K;
  int act;            // 0 none, 1 relu, 2 silu
  int kb_per_split;   // K blocks of 32 per split
  // optional epilogue extras (no split-K): `stash` [M, ldd] receives the pre-activation A.B^T + bias (what the backward of
  // SiLU needs, so act = silu costs no separate elementwise launch); `dsilu_aux` [M, ld_aux]: the result is multiplied by
  // SiLU'(dsilu_aux[m, n]) before the residual is added (the backward of SiLU fused into the dgrad GEMM).
  float* stash;
  const float* dsilu_aux;
  int64_t ld_aux;
  // second bias [N], added unscaled AFTER the activation and out_scale:  D = out_scale * act(A.B^T + bias * rowscale) + bias2
  // + res_scale * residual  (two Linears merged into one GEMM keep the outer Linear's bias separate from the inner one's,
  // which carries the row scale / mix factor)
  const float* bias2;
  // ReLU backward of the CONSUMER fused into this dgrad GEMM: after the residual is added, elements whose `relu_mask`
  // [M, ld_mask] entry (the ReLU's output) is not positive are zeroed:  D = (mask > 0) ? (... + res_scale * residual) : 0
  const float* relu_mask;
  int64_t ld_mask;
};

// A group of independent problems launched as one grid (gemm_ts.cu: k_gemm_ts_grouped)
constexpr int kMaxGroup = 8;
struct GroupParams {
  CUtensorMap tmA[kMaxGroup];
  CUtensorMap tmB[kMaxGroup];
  Params p[kMaxGroup];
  int cta_start[kMaxGroup + 1];   // first CTA of each problem in the linear grid
  int tiles[kMaxGroup];           // output tiles (m x n) of each problem; its CTAs = tiles * p.splits
  int tiles_n[kMaxGroup];
  int n;
};


// Epilogue of one 128 x BN tile, executed by warps 4-7 (TMEM lane quadrant = warp & 3):
// TMEM -> registers -> (bias, act, scales) -> padded smem (`smem`, the idle operand ring) -> coalesced stores
// (+ residual).  AM::NA accumulators of BN columns at tmem_base (see AccMap): big hi.hi tiles first, small ones after.
// ep_tid: 0..127 index of the thread among the four epilogue warps; full_parity: phase of tmem_full to wait for;
// tmem_empty (optional): arrived on by every epilogue thread once the accumulators are in registers / staged, so a
// persistent kernel can start the next tile's MMAs into this buffer while the stores are still in flight.
template <int BN, class AM, bool STORE = true>
__device__ __forceinline__ void epilogue_tile(uint8_t* smem, float* sbias, uint32_t tmem_base, uint64_t* tmem_full, int warp,
                                              int lane, int nkb, const Params& p, int m0, int n0, int ep_tid = -1,
                                              uint32_t full_parity = 0, uint64_t* tmem_empty = nullptr, int split_idx = -1) {
    if (ep_tid < 0) ep_tid = threadIdx.x - 128;
    // ---- epilogue: TMEM -> registers -> (bias, act, residual) -> padded smem -> coalesced stores
    constexpr int PITCH = BN + 4;
    float* stage_out = reinterpret_cast<float*>(smem);   // the operand ring is idle once tmem_full fires
    const int q = warp & 3;                              // TMEM lane quadrant this warp may read (warps 4-7 -> 0-3)
    const int row = q * 32 + lane;
    if (nkb > 0) {
      mbar_wait(tmem_full, full_parity);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (ep_tid == 0) EQH_STAMP(6);
    const bool final_out = (p.ws == nullptr);
    // bias tile -> shared memory once (zero outside N), so the element loop has no bounds checks
    {
      const int tt = ep_tid;
      if (tt < BN) {
        sbias[tt] = (final_out && p.bias && n0 + tt < p.N) ? __ldg(p.bias + n0 + tt) : 0.f;
        sbias[BN + tt] = (final_out && p.bias2 && n0 + tt < p.N) ? __ldg(p.bias2 + n0 + tt) : 0.f;
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const int act = final_out ? p.act : 0;
    const float rs = (final_out && p.rowscale && m0 + row < p.M) ? __ldg(p.rowscale + m0 + row) : 1.f;
    const float osc = final_out ? p.out_scale : 1.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      if (nkb > 0) {
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
        // small terms first (highest tile index), the big hi.hi tiles last; two loads in flight
        float w[32], u[32];
        tmem_ld32_nowait(ta + (AM::NA - 1) * BN, v);
        tmem_ld32_nowait(ta + (AM::NA - 2) * BN, w);
        tmem_ld_wait();
#pragma unroll
        for (int a = AM::NA - 3; a >= 0; a -= 2) {
          if (a >= 0) tmem_ld32_nowait(ta + a * BN, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
          if (a >= 1) tmem_ld32_nowait(ta + (a - 1) * BN, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += u[j];
          if (a < 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) w[j] = 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += w[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = *reinterpret_cast<const float4*>(sbias + c0 + j);   // warp-uniform address: broadcast
        v[j] += b4.x * rs; v[j + 1] += b4.y * rs; v[j + 2] += b4.z * rs; v[j + 3] += b4.w * rs;
      }
      if (final_out && p.stash != nullptr && m0 + row < p.M) {
        // pre-activation stash: each thread owns 32 consecutive columns of its row (a full 128-byte line per thread)
        float* sp = p.stash + (int64_t)(m0 + row) * p.ldd + n0 + c0;
        if ((p.ldd % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.stash) & 15) == 0) && n0 + c0 + 32 <= p.N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(sp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < p.N) sp[j] = v[j];
        }
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(sbias + BN + c0 + j);     // bias2 (zero when absent)
        *reinterpret_cast<float4*>(stage_out + row * PITCH + c0 + j) =
            make_float4(v[j] * osc + c4.x, v[j + 1] * osc + c4.y, v[j + 2] * osc + c4.z, v[j + 3] * osc + c4.w);
      }
    }
    if (tmem_empty != nullptr) {                         // all TMEM reads of this thread are complete
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(tmem_empty);
    }
    if constexpr (!STORE) {       // the caller stores the staged tile with every thread of the CTA (epilogue_store_all)
      if (ep_tid == 0) EQH_STAMP(7);
      return;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps only
    if (ep_tid == 0) EQH_STAMP(7);
    float* out = final_out ? p.D : p.ws + (size_t)(split_idx >= 0 ? split_idx : (int)blockIdx.z) * p.M * p.N;
    const int64_t ld = final_out ? p.ldd : p.N;
    const bool vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const bool has_res = final_out && p.residual != nullptr;
    const bool has_aux = final_out && p.dsilu_aux != nullptr;
    const bool has_mask = final_out && p.relu_mask != nullptr;
    // a warp instruction covers 128 / BN rows of the tile as float4s; four such groups in flight per warp
    constexpr int LPR = BN / 4;                          // lanes per row
    constexpr int RPI = 32 / LPR;                        // rows per warp instruction
    const int c = (lane % LPR) * 4, rsub = lane / LPR;
    const int n = n0 + c;
    const bool full_vec = vec_ok && (n + 4 <= p.N);
    const int rows_here = min(BM, p.M - m0);
    for (int rb = q * RPI; rb < rows_here; rb += 16 * RPI) {
      if (full_vec) {
        float4 x[4], rr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + rsub + 4 * RPI * u;
          if (r < rows_here) {
            x[u] = *reinterpret_cast<const float4*>(stage_out + r * PITCH + c);
            // plain (coherent) load: the residual may be D itself
            if (has_res) rr[u] = *reinterpret_cast<const float4*>(p.residual + (int64_t)(m0 + r) * ld + n);
            if (has_aux) {
              const float* ap = p.dsilu_aux + (int64_t)(m0 + r) * p.ld_aux + n;
              float4 a4;
              if ((p.ld_aux % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dsilu_aux) & 15) == 0)) a4 = __ldg(reinterpret_cast<const float4*>(ap));
              else a4 = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), __ldg(ap + 3));
              x[u].x *= dsilu_f(a4.x); x[u].y *= dsilu_f(a4.y); x[u].z *= dsilu_f(a4.z); x[u].w *= dsilu_f(a4.w);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + rsub + 4 * RPI * u;
          if (r < rows_here) {
            if (has_res) {
              x[u].x += p.res_scale * rr[u].x; x[u].y += p.res_scale * rr[u].y;
              x[u].z += p.res_scale * rr[u].z; x[u].w += p.res_scale * rr[u].w;
            }
            if (has_mask) {
              const float* mp = p.relu_mask + (int64_t)(m0 + r) * p.ld_mask + n;
              x[u].x = __ldg(mp) > 0.f ? x[u].x : 0.f; x[u].y = __ldg(mp + 1) > 0.f ? x[u].y : 0.f;
              x[u].z = __ldg(mp + 2) > 0.f ? x[u].z : 0.f; x[u].w = __ldg(mp + 3) > 0.f ? x[u].w : 0.f;
            }
            *reinterpret_cast<float4*>(out + (int64_t)(m0 + r) * ld + n) = x[u];
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + rsub + 4 * RPI * u;
          if (r >= rows_here) continue;
          const float* src = stage_out + r * PITCH + c;
          float* dst = out + (int64_t)(m0 + r) * ld + n;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j < p.N) {
              float x = src[j];
              if (has_aux) x *= dsilu_f(__ldg(p.dsilu_aux + (int64_t)(m0 + r) * p.ld_aux + n + j));
              if (has_res) x += p.res_scale * p.residual[(int64_t)(m0 + r) * ld + n + j];
              if (has_mask && !(__ldg(p.relu_mask + (int64_t)(m0 + r) * p.ld_mask + n + j) > 0.f)) x = 0.f;
              dst[j] = x;
            }
          }
        }
      }
    }
}

// Store phase of the epilogue executed by ALL threads of the CTA (after a CTA-wide barrier that follows the staging of
// epilogue_tile<.., STORE = false>): the 128 x BN tile staged in shared memory (pitch BN + 4) goes to global memory as
// coalesced float4 rows, + the fused SiLU-derivative factor and the residual.  With the four epilogue warps alone the
// store loop was 1.3 us of a 6.9 us CTA (scratch/trace_ts.py); the other six warps are idle by then.
template <int BN>
__device__ __forceinline__ void epilogue_store_all(const uint8_t* smem, const Params& p, int m0, int n0, int split_idx) {
  constexpr int PITCH = BN + 4;
  constexpr int C4 = BN / 4;
  const float* stage_out = reinterpret_cast<const float*>(smem);
  const bool final_out = (p.ws == nullptr);
  float* out = final_out ? p.D : p.ws + (size_t)split_idx * p.M * p.N;
  const int64_t ld = final_out ? p.ldd : p.N;
  const bool vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const bool has_res = final_out && p.residual != nullptr;
  const bool has_aux = final_out && p.dsilu_aux != nullptr;
  const bool aux_vec = has_aux && (p.ld_aux % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.dsilu_aux) & 15) == 0);
  const bool has_mask = final_out && p.relu_mask != nullptr;
  const bool mask_vec = has_mask && (p.ld_mask % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.relu_mask) & 15) == 0);
  const int rows_here = min(BM, p.M - m0);
  const int total = rows_here * C4;
  for (int e0 = threadIdx.x; e0 < total; e0 += 2 * kThreads) {
    float4 x[2], rr[2], a4[2], k4[2];
    int rr_[2], n_[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int e = e0 + u * kThreads;
      ok[u] = e < total;
      const int r = ok[u] ? e / C4 : 0, c = ok[u] ? (e - r * C4) * 4 : 0;
      rr_[u] = r; n_[u] = n0 + c;
      ok[u] = ok[u] && n_[u] < p.N;
      if (!ok[u]) continue;
      x[u] = *reinterpret_cast<const float4*>(stage_out + r * PITCH + c);
      const bool full = vec_ok && n_[u] + 4 <= p.N;
      if (has_res && full) rr[u] = *reinterpret_cast<const float4*>(p.residual + (int64_t)(m0 + r) * ld + n_[u]);   // may alias D
      if (has_aux) {
        const float* ap = p.dsilu_aux + (int64_t)(m0 + r) * p.ld_aux + n_[u];
        if (aux_vec && n_[u] + 4 <= p.N) a4[u] = __ldg(reinterpret_cast<const float4*>(ap));
        else a4[u] = make_float4(__ldg(ap), n_[u] + 1 < p.N ? __ldg(ap + 1) : 0.f, n_[u] + 2 < p.N ? __ldg(ap + 2) : 0.f,
                                 n_[u] + 3 < p.N ? __ldg(ap + 3) : 0.f);
      }
      if (has_mask) {
        const float* mp = p.relu_mask + (int64_t)(m0 + r) * p.ld_mask + n_[u];
        if (mask_vec && n_[u] + 4 <= p.N) k4[u] = __ldg(reinterpret_cast<const float4*>(mp));
        else k4[u] = make_float4(__ldg(mp), n_[u] + 1 < p.N ? __ldg(mp + 1) : 0.f, n_[u] + 2 < p.N ? __ldg(mp + 2) : 0.f,
                                 n_[u] + 3 < p.N ? __ldg(mp + 3) : 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const int r = rr_[u], n = n_[u];
      if (has_aux) {
        x[u].x *= dsilu_f(a4[u].x); x[u].y *= dsilu_f(a4[u].y); x[u].z *= dsilu_f(a4[u].z); x[u].w *= dsilu_f(a4[u].w);
      }
      float* dst = out + (int64_t)(m0 + r) * ld + n;
      if (vec_ok && n + 4 <= p.N) {
        if (has_res) {
          x[u].x += p.res_scale * rr[u].x; x[u].y += p.res_scale * rr[u].y;
          x[u].z += p.res_scale * rr[u].z; x[u].w += p.res_scale * rr[u].w;
        }
        if (has_mask) {
          x[u].x = k4[u].x > 0.f ? x[u].x : 0.f; x[u].y = k4[u].y > 0.f ? x[u].y : 0.f;
          x[u].z = k4[u].z > 0.f ? x[u].z : 0.f; x[u].w = k4[u].w > 0.f ? x[u].w : 0.f;
        }
        *reinterpret_cast<float4*>(dst) = x[u];
      } else {
        const float v[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
        const float km[4] = {has_mask ? k4[u].x : 1.f, has_mask ? k4[u].y : 1.f, has_mask ? k4[u].z : 1.f,
                             has_mask ? k4[u].w : 1.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n + j < p.N) {
            float t = v[j];
            if (has_res) t += p.res_scale * p.residual[(int64_t)(m0 + r) * ld + n + j];
            if (!(km[j] > 0.f)) t = 0.f;
            dst[j] = t;
          }
        }
      }
    }
  }
}

// Split-K finish, executed by the whole CTA after its partial tile is written (includes the block barriers).
template <int BN>
__device__ __forceinline__ void splitk_finish(const Params& p, int m0, int n0, int* s_last, int tile_idx = -1,
                                              int nsplits = -1) {
  {
    // ---- split-K finish: the CTA that arrives last at this tile's counter adds the partials in split order
    // (the order never depends on which CTA is last: deterministic) and applies the output transform
    if (threadIdx.x == 0) {
      const int tile = tile_idx >= 0 ? tile_idx : (int)(blockIdx.y * gridDim.x + blockIdx.x);
      const int prev = atomicAdd(p.counters + tile, 1);
      *s_last = (prev == (nsplits >= 0 ? nsplits : (int)gridDim.z) - 1);
      if (*s_last) p.counters[tile] = 0;         // left zero for the next launch
    }
    __syncthreads();
    if (*s_last) {
      __threadfence();
      const int64_t MN = (int64_t)p.M * p.N;
      const int rows_here = min(BM, p.M - m0);
      const int nz = nsplits >= 0 ? nsplits : (int)gridDim.z;
      const bool vec = ((p.N & 3) == 0) && ((p.ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.D) & 15) == 0) &&
                       (p.residual == nullptr || (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
      if (vec) {
        constexpr int C4 = BN / 4;
        for (int e = threadIdx.x; e < rows_here * C4; e += kThreads) {
          const int r = e / C4, c = (e - r * C4) * 4;
          if (n0 + c >= p.N) continue;
          const float* src = p.ws + (int64_t)(m0 + r) * p.N + n0 + c;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
          for (int z = 0; z < nz; ++z) {
            const float4 t = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)z * MN));
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
          }
          const int64_t o = (int64_t)(m0 + r) * p.ldd + n0 + c;
          float4 x = make_float4(p.out_scale * acc.x, p.out_scale * acc.y, p.out_scale * acc.z, p.out_scale * acc.w);
          if (p.residual) {
            const float4 rr = *reinterpret_cast<const float4*>(p.residual + o);
            x.x += p.res_scale * rr.x; x.y += p.res_scale * rr.y; x.z += p.res_scale * rr.z; x.w += p.res_scale * rr.w;
          }
          *reinterpret_cast<float4*>(p.D + o) = x;
        }
      } else {
        for (int e = threadIdx.x; e < rows_here * BN; e += kThreads) {
          const int r = e / BN, c = e - r * BN;
          if (n0 + c >= p.N) continue;
          const float* src = p.ws + (int64_t)(m0 + r) * p.N + n0 + c;
          float acc = 0.f;
#pragma unroll 8
          for (int z = 0; z < nz; ++z) acc += __ldcg(src + (int64_t)z * MN);
          const int64_t o = (int64_t)(m0 + r) * p.ldd + n0 + c;
          float x = p.out_scale * acc;
          if (p.residual) x += p.res_scale * p.residual[o];
          p.D[o] = x;
        }
      }
    }
  }
}

}  // namespace gemm
}  // namespace eqh
