// Fused per-incidence message (gather-add -> ReLU -> LayerNorm) + deterministic segmented
// reduction, forward and backward.  The [Z,C] / [Z,2C] per-incidence tensors of the reference
// (conv.py:90,96,175-176: X[vertex], E[edges], cat; mlp.py:94-96: relu_, LayerNorm) never
// touch HBM.  HBM-bound; algorithmic bytes per forward call (figure F):
//     4C(rows_A + rows_B + n_out) + 4(2Z + n_out + 1).
// One warp owns one segment; a feature row is spread over the warp (Row<V>), LayerNorm
// statistics are two warp-shuffle reductions (two-pass mean / variance, like ATen).
#include "common.cuh"

namespace eqh {

template <int V>
__device__ __forceinline__ void ln_stats(const Row<V>& t, float eps, float& mean, float& rstd) {
  constexpr float invC = 1.f / (32 * V);
  mean = warp_sum(t.sum()) * invC;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) { float d = t.v[i] - mean; q += d * d; }
  const float var = warp_sum(q) * invC;
  rstd = 1.f / sqrtf(var + eps);
}

// GA / GB: which operand is gathered through an index (the other one is the segment's own row).
// Specialising on this (instead of testing the index pointers at run time) and re-reading the LayerNorm
// affine only when a segment is finalised keeps the kernel under 64 registers -> 4 CTAs / SM, which is what
// bounds this kernel (ncu: DRAM traffic == algorithmic bytes, but only 24 % active warps at 95 registers).
template <int V, bool LN, bool GA, bool GB>
__global__ void __launch_bounds__(256, 4) k_fused_fwd(const float* __restrict__ A, const float* __restrict__ B,
                                                       const int32_t* __restrict__ ia, const int32_t* __restrict__ ib,
                                                       const int32_t* __restrict__ rowptr, int64_t n_out,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, int mode, float* __restrict__ U) {
  pdl_prologue();
  constexpr int C = 32 * V;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = warp; s < n_out; s += nwarps) {
    const int32_t lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
    Row<V> own, acc, nxt;      // own: the segment's own row of the non-gathered operand(s)
    if constexpr (!GA && !GB) {
      Row<V> o2;
      own.load(A + s * C, lane);
      o2.load(B + s * C, lane);
#pragma unroll
      for (int i = 0; i < V; ++i) own.v[i] += o2.v[i];
    } else if constexpr (!GA) {
      own.load(A + s * C, lane);
    } else if constexpr (!GB) {
      own.load(B + s * C, lane);
    } else {
      own.zero();
    }
    acc.zero();
    auto fetch = [&](int32_t k) {   // the gathered part of incidence k
      if constexpr (GA && GB) {
        Row<V> t2;
        nxt.load(A + (int64_t)__ldg(ia + k) * C, lane);
        t2.load(B + (int64_t)__ldg(ib + k) * C, lane);
#pragma unroll
        for (int i = 0; i < V; ++i) nxt.v[i] += t2.v[i];
      } else if constexpr (GA) {
        nxt.load(A + (int64_t)__ldg(ia + k) * C, lane);
      } else if constexpr (GB) {
        nxt.load(B + (int64_t)__ldg(ib + k) * C, lane);
      }
    };
    if (lo < hi) fetch(lo);
    for (int32_t k = lo; k < hi; ++k) {
      Row<V> t;
#pragma unroll
      for (int i = 0; i < V; ++i) t.v[i] = fmaxf(((GA || GB) ? nxt.v[i] : 0.f) + own.v[i], 0.f);
      if (k + 1 < hi) fetch(k + 1);   // software prefetch: next incidence requested before this one is normalised
      if constexpr (LN) {
        float mean, rstd;
        ln_stats<V>(t, eps, mean, rstd);
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] += (t.v[i] - mean) * rstd;
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] += t.v[i];
      }
    }
    const float cnt = (float)(hi - lo);
    if constexpr (LN) {  // sum_k (xhat_k * gamma + beta) = gamma * sum_k xhat_k + cnt * beta
      Row<V> g, bt;
      g.load(gamma, lane);
      bt.load(beta, lane);
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = acc.v[i] * g.v[i] + cnt * bt.v[i];
    }
    if (mode == EQH_REDUCE_MEAN) {
      const float d = fmaxf(cnt, 1.f);
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = acc.v[i] / d;
    }
    acc.store(U + s * C, lane);
  }
}

// Backward, one grouping per launch.  GA / GB / GU: A, B, dU gathered through ia / ib / iu (otherwise the
// segment's own row).  Same register discipline as the forward kernel.
template <int V, bool LN, bool DGB, bool GA, bool GB, bool GU>
__global__ void __launch_bounds__(256, DGB ? 2 : 4) k_fused_bwd(
    const float* __restrict__ A, const float* __restrict__ B, const int32_t* __restrict__ ia,
    const int32_t* __restrict__ ib, const int32_t* __restrict__ iu, const int32_t* __restrict__ rowptr, int64_t n_seg,
    const float* __restrict__ gamma, float eps, const float* __restrict__ dU, const int32_t* __restrict__ u_rowptr,
    float* __restrict__ dOp, float* __restrict__ dgb_partial) {
  pdl_prologue();
  constexpr int C = 32 * V;
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  Row<V> g, dg, db;
  if constexpr (LN) g.load(gamma, lane);
  if constexpr (DGB) { dg.zero(); db.zero(); }
  for (int64_t s = warp; s < n_seg; s += nwarps) {
    const int32_t lo = __ldg(rowptr + s), hi = __ldg(rowptr + s + 1);
    Row<V> own, uc, acc;
    own.zero();
    if constexpr (!GA) own.load(A + s * C, lane);
    if constexpr (!GB) {
      Row<V> o2;
      o2.load(B + s * C, lane);
#pragma unroll
      for (int i = 0; i < V; ++i) own.v[i] += o2.v[i];
    }
    if constexpr (!GU) {
      uc.load(dU + s * C, lane);
      const float w = u_rowptr ? 1.f / (float)max(hi - lo, 1) : 1.f;  // dU's grouping is this grouping
#pragma unroll
      for (int i = 0; i < V; ++i) uc.v[i] *= w;
    }
    acc.zero();
    // software pipeline: the gathered rows of incidence k+1 are in flight while incidence k is processed (a warp
    // otherwise has one row load outstanding at a time and the kernel is bound by the gather latency)
    Row<V> nA, nB, nU;
    float nW = 1.f;
    auto fetch = [&](int32_t k) {
      if constexpr (GA) nA.load(A + (int64_t)__ldg(ia + k) * C, lane);
      if constexpr (GB) nB.load(B + (int64_t)__ldg(ib + k) * C, lane);
      if constexpr (GU) {
        const int32_t r = __ldg(iu + k);
        nU.load(dU + (int64_t)r * C, lane);
        nW = inv_count(u_rowptr, r);
      }
    };
    // (only in the 2-CTA/SM variant: under the 64-register cap of the 4-CTA/SM variants the extra rows spill
    // and the kernel gets slower -- measured 0.61 -> 0.55 of the HBM peak)
    constexpr bool kPrefetch = DGB;
    if (kPrefetch && lo < hi) fetch(lo);
    for (int32_t k = lo; k < hi; ++k) {
      if constexpr (!kPrefetch) fetch(k);
      Row<V> pre, gk;
#pragma unroll
      for (int i = 0; i < V; ++i) pre.v[i] = own.v[i];
      if constexpr (GA) {
#pragma unroll
        for (int i = 0; i < V; ++i) pre.v[i] += nA.v[i];
      }
      if constexpr (GB) {
#pragma unroll
        for (int i = 0; i < V; ++i) pre.v[i] += nB.v[i];
      }
      if constexpr (GU) {
#pragma unroll
        for (int i = 0; i < V; ++i) gk.v[i] = nU.v[i] * nW;
      }
      if (kPrefetch && k + 1 < hi) fetch(k + 1);
      if constexpr (LN) {
        Row<V> t;
#pragma unroll
        for (int i = 0; i < V; ++i) t.v[i] = fmaxf(pre.v[i], 0.f);
        float mean, rstd;
        ln_stats<V>(t, eps, mean, rstd);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float gv = GU ? gk.v[i] : uc.v[i];
          const float xh = (t.v[i] - mean) * rstd;
          const float dxi = gv * g.v[i];
          s1 += dxi;
          s2 += dxi * xh;
          if constexpr (DGB) { dg.v[i] += gv * xh; db.v[i] += gv; }
          t.v[i] = xh;                 // keep xhat
        }
        constexpr float invC = 1.f / C;
        s1 = warp_sum(s1) * invC;
        s2 = warp_sum(s2) * invC;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float gv = GU ? gk.v[i] : uc.v[i];
          const float dt = rstd * (gv * g.v[i] - s1 - t.v[i] * s2);
          acc.v[i] += pre.v[i] > 0.f ? dt : 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] += pre.v[i] > 0.f ? (GU ? gk.v[i] : uc.v[i]) : 0.f;
      }
    }
    acc.store(dOp + s * C, lane);
  }
  if constexpr (DGB) {
    __shared__ __align__(16) float sm[8][2 * C];
    dg.store(&sm[wib][0], lane);
    db.store(&sm[wib][C], lane);
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += sm[w][c];  // fixed order
      dgb_partial[(int64_t)blockIdx.x * 2 * C + c] = v;
    }
  }
}

// Backward of the primary grouping WITH the LayerNorm affine partials (the variant every training step runs).  Segments
// are short (2-3 incidences), so the one-segment-per-warp-turn kernel above is bound by the chain of dependent loads at
// every segment start (rowptr -> index -> gathered row, plus the segment's own rows): 0.54 of the HBM peak.  Here a
// warp owns a CONTIGUOUS range of segments: its incidences are contiguous too, so the one-ahead prefetch of the gathered
// rows runs across segment boundaries, rowptr is read one segment ahead, and the next segment's own rows (A|B and dU)
// arrive by cp.async in a per-warp shared-memory buffer while the current segment is processed.
template <int V, bool GA, bool GB, bool GU>
__global__ void __launch_bounds__(256, 2) k_fused_bwd_dgb(
    const float* __restrict__ A, const float* __restrict__ B, const int32_t* __restrict__ ia,
    const int32_t* __restrict__ ib, const int32_t* __restrict__ iu, const int32_t* __restrict__ rowptr, int64_t n_seg,
    const float* __restrict__ gamma, float eps, const float* __restrict__ dU, const int32_t* __restrict__ u_rowptr,
    float* __restrict__ dOp, float* __restrict__ dgb_partial) {
  static_assert(GA != GB, "exactly one of A / B is gathered in the backward groupings");
  pdl_prologue();
  constexpr int C = 32 * V;
  extern __shared__ __align__(16) float sbuf[];      // [8 warps][2][C] staging; reused as [8][2C] for the final reduce
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t base = n_seg / nwarps, rem = n_seg - base * nwarps;          // balanced contiguous ranges
  const int64_t s0 = warp * base + min(warp, rem), s1 = s0 + base + (warp < rem ? 1 : 0);
  float* wbuf = sbuf + wib * 2 * C;
  Row<V> g, dg, db;
  g.load(gamma, lane);
  dg.zero();
  db.zero();
  const float* ownsrc = GA ? B : A;
  auto own_async = [&](int64_t s) {
    Row<V>::copy_async(wbuf, ownsrc + s * C, lane);
    if constexpr (!GU) Row<V>::copy_async(wbuf + C, dU + s * C, lane);
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  Row<V> nG, nU;                    // gathered rows of the NEXT incidence
  float nW = 1.f;
  auto fetch = [&](int32_t k) {
    if constexpr (GA) nG.load(A + (int64_t)__ldg(ia + k) * C, lane);
    else nG.load(B + (int64_t)__ldg(ib + k) * C, lane);
    if constexpr (GU) {
      const int32_t r = __ldg(iu + k);
      nU.load(dU + (int64_t)r * C, lane);
      nW = inv_count(u_rowptr, r);
    }
  };
  if (s0 < s1) {
    own_async(s0);
    int32_t lo = __ldg(rowptr + s0), hi = __ldg(rowptr + s0 + 1);
    const int32_t kend = __ldg(rowptr + s1);
    if (lo < kend) fetch(lo);
    for (int64_t s = s0; s < s1; ++s) {
      Row<V> own, uc, acc;
      asm volatile("cp.async.wait_group 0;" ::: "memory");     // each lane reads back only the chunks it copied itself
      own.load_plain(wbuf, lane);
      if constexpr (!GU) {
        uc.load_plain(wbuf + C, lane);
        const float w = u_rowptr ? 1.f / (float)max(hi - lo, 1) : 1.f;  // dU's grouping is this grouping
#pragma unroll
        for (int i = 0; i < V; ++i) uc.v[i] *= w;
      }
      if (s + 1 < s1) own_async(s + 1);
      const int32_t hi_next = s + 2 <= s1 ? __ldg(rowptr + s + 2) : hi;
      acc.zero();
      for (int32_t k = lo; k < hi; ++k) {
        Row<V> pre, gk;
#pragma unroll
        for (int i = 0; i < V; ++i) pre.v[i] = own.v[i] + nG.v[i];
        if constexpr (GU) {
#pragma unroll
          for (int i = 0; i < V; ++i) gk.v[i] = nU.v[i] * nW;
        }
        if (k + 1 < kend) fetch(k + 1);        // runs across the segment boundary
        Row<V> t;
#pragma unroll
        for (int i = 0; i < V; ++i) t.v[i] = fmaxf(pre.v[i], 0.f);
        float mean, rstd;
        ln_stats<V>(t, eps, mean, rstd);
        float s1_ = 0.f, s2_ = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float gv = GU ? gk.v[i] : uc.v[i];
          const float xh = (t.v[i] - mean) * rstd;
          const float dxi = gv * g.v[i];
          s1_ += dxi;
          s2_ += dxi * xh;
          dg.v[i] += gv * xh;
          db.v[i] += gv;
          t.v[i] = xh;                 // keep xhat
        }
        constexpr float invC = 1.f / C;
        s1_ = warp_sum(s1_) * invC;
        s2_ = warp_sum(s2_) * invC;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float gv = GU ? gk.v[i] : uc.v[i];
          const float dt = rstd * (gv * g.v[i] - s1_ - t.v[i] * s2_);
          acc.v[i] += pre.v[i] > 0.f ? dt : 0.f;
        }
      }
      acc.store(dOp + s * C, lane);
      lo = hi;
      hi = hi_next;
    }
  }
  __syncthreads();                    // every warp is done with its staging rows
  float* sm = sbuf;                   // [8][2C]
  dg.store(sm + wib * 2 * C, lane);
  db.store(sm + wib * 2 * C + C, lane);
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += sm[w * 2 * C + c];  // fixed order
    dgb_partial[(int64_t)blockIdx.x * 2 * C + c] = v;
  }
}

// Row LayerNorm of the dense MLPs (mlp.py:94-96: relu_ then LayerNorm), optionally with the ReLU folded in:
//   y = LN(relu?(x)) ; backward dx = relu'(x) * LN_backward(dy), per-CTA partials of dgamma / dbeta.
template <int V, bool RELU>
__global__ void __launch_bounds__(256) k_rowln_fwd(const float* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float eps, int64_t rows,
                                                    float* __restrict__ y, int64_t ldx, int64_t ldy) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  Row<V> g, bt;
  g.load(gamma, lane);
  bt.load(beta, lane);
  for (int64_t r = warp; r < rows; r += nwarps) {
    Row<V> t;
    t.load(x + r * ldx, lane);
    if (RELU) {
#pragma unroll
      for (int i = 0; i < V; ++i) t.v[i] = fmaxf(t.v[i], 0.f);
    }
    float mean, rstd;
    ln_stats<V>(t, eps, mean, rstd);
#pragma unroll
    for (int i = 0; i < V; ++i) t.v[i] = (t.v[i] - mean) * rstd * g.v[i] + bt.v[i];
    t.store(y + r * ldy, lane);
  }
}

template <int V, bool RELU>
__global__ void __launch_bounds__(256) k_rowln_bwd(const float* __restrict__ x, const float* __restrict__ gamma, float eps,
                                                    const float* __restrict__ dy, int64_t rows, float* __restrict__ dx,
                                                    float* __restrict__ dgb_partial, int64_t ldx, int64_t lddy, int64_t lddx,
                                                    const float* __restrict__ add) {
  pdl_prologue();
  constexpr int C = 32 * V;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  Row<V> g, dg, db;
  g.load(gamma, lane);
  dg.zero();
  db.zero();
  for (int64_t r = warp; r < rows; r += nwarps) {
    Row<V> pre, t, gy;
    pre.load(x + r * ldx, lane);
    gy.load(dy + r * lddy, lane);
#pragma unroll
    for (int i = 0; i < V; ++i) t.v[i] = RELU ? fmaxf(pre.v[i], 0.f) : pre.v[i];
    float mean, rstd;
    ln_stats<V>(t, eps, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
    Row<V> xh, dxh;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      xh.v[i] = (t.v[i] - mean) * rstd;
      dxh.v[i] = gy.v[i] * g.v[i];
      s1 += dxh.v[i];
      s2 += dxh.v[i] * xh.v[i];
      dg.v[i] += gy.v[i] * xh.v[i];
      db.v[i] += gy.v[i];
    }
    constexpr float invC = 1.f / C;
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    Row<V> o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float dt = rstd * (dxh.v[i] - s1 - xh.v[i] * s2);
      o.v[i] = (!RELU || pre.v[i] > 0.f) ? dt : 0.f;
    }
    if (add) {                       // dx = add + LayerNorm backward (a gradient that reaches x on a second path)
      Row<V> a;
      a.load_plain(add + r * lddx, lane);
#pragma unroll
      for (int i = 0; i < V; ++i) o.v[i] += a.v[i];
    }
    o.store(dx + r * lddx, lane);
  }
  __shared__ __align__(16) float sm[8][2 * C];
  dg.store(&sm[wib][0], lane);
  db.store(&sm[wib][C], lane);
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += sm[w][c];
    dgb_partial[(int64_t)blockIdx.x * 2 * C + c] = v;
  }
}

// Column sums of a tall matrix, stage 1: each CTA reduces one chunk of rows for 32 columns
// (8 interleaved row groups, fixed order); stage 2 is k_colsum over the chunk partials.
__global__ void __launch_bounds__(256) k_colsum_chunks(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                                        const float* __restrict__ roww, int64_t rows_per_chunk,
                                                        float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float sm[8][32];
  const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int64_t c = blockIdx.x * 32 + cx;
  const int64_t r0 = blockIdx.y * rows_per_chunk, r1 = min(rows, r0 + rows_per_chunk);
  float v = 0.f;
  if (c < cols)
    for (int64_t r = r0 + rg; r < r1; r += 8) v += roww ? __ldg(roww + r) * __ldg(x + r * ld + c) : __ldg(x + r * ld + c);
  sm[rg][cx] = v;
  __syncthreads();
  if (rg == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][cx];
    partial[blockIdx.y * cols + c] = t;
  }
}

// out[c] = sum_r partial[r,c], r ascending within each of 8 interleaved groups, groups combined in fixed order
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ partial, int64_t rows, int64_t cols, int64_t ld,
                                                 float* out, float scale, int accumulate) {
  pdl_prologue();
  __shared__ float sm[8][32];
  const int cx = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const int64_t c = blockIdx.x * 32 + cx;
  float v = 0.f;
  if (c < cols)
    for (int64_t r = rg; r < rows; r += 8) v += __ldg(partial + r * ld + c);
  sm[rg][cx] = v;
  __syncthreads();
  if (rg == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][cx];
    out[c] = accumulate ? out[c] + scale * t : scale * t;
  }
}

}  // namespace eqh

static bool legacy_dgb() {      // EQH_FUSED_BWD=legacy: the one-segment-per-turn kernel for the affine-partials variant too (A/B)
  static const bool v = []() {
    const char* e = getenv("EQH_FUSED_BWD");
    return e != nullptr && e[0] == 'l';
  }();
  return v;
}

extern "C" {

int eqh_partial_rows(void) { return eqh::kPartialRows; }

int eqh_fused_msg_fwd(const float* A, const float* B, const int32_t* ia, const int32_t* ib, const int32_t* rowptr,
                      int64_t n_out, int64_t C, const float* gamma, const float* beta, float eps, int mode, float* U,
                      eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(n_out >= 0 && rowptr && A && B && (n_out == 0 || U));
  EQH_CHECK_ARG((gamma == nullptr) == (beta == nullptr));
  EQH_CHECK_ARG(mode == EQH_REDUCE_SUM || mode == EQH_REDUCE_MEAN);
  EQH_CHECK_ARG(aligned16(A) && aligned16(B) && aligned16(U) && aligned16(gamma) && aligned16(beta));
  if (!supported_C(C)) return EQH_ERR_UNSUPPORTED;
  if (n_out == 0) return EQH_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (n_out + 7) / 8;
  if (blocks > (int64_t)kSMs * 16) blocks = (int64_t)kSMs * 16;
  const unsigned gb = (unsigned)blocks;
#define EQH_FWD(LNV, GAV, GBV) \
  launch_kernel(k_fused_fwd<V, LNV, GAV, GBV>, dim3(gb), dim3(256), 0, st, A, B, ia, ib, rowptr, n_out, gamma, beta, eps, mode, U)
  EQH_DISPATCH_V(C, {
    if (gamma) {
      if (ia && ib) EQH_FWD(true, true, true);
      else if (ia) EQH_FWD(true, true, false);
      else if (ib) EQH_FWD(true, false, true);
      else EQH_FWD(true, false, false);
    } else {
      if (ia && ib) EQH_FWD(false, true, true);
      else if (ia) EQH_FWD(false, true, false);
      else if (ib) EQH_FWD(false, false, true);
      else EQH_FWD(false, false, false);
    }
  });
#undef EQH_FWD
  EQH_LAUNCH_RC();
}

int eqh_fused_msg_bwd(const float* A, const float* B, const int32_t* ia, const int32_t* ib, const int32_t* iu,
                      const int32_t* rowptr, int64_t n_seg, int64_t C, const float* gamma, const float* beta,
                      float eps, const float* dU, const int32_t* u_rowptr, float* dOp, float* dgb_partial,
                      eqh_stream_t stream) {
  using namespace eqh;
  (void)beta;
  EQH_CHECK_ARG(n_seg >= 0 && rowptr && A && B && dU && (n_seg == 0 || dOp));
  EQH_CHECK_ARG(!(dgb_partial && !gamma));
  EQH_CHECK_ARG(aligned16(A) && aligned16(B) && aligned16(dU) && aligned16(dOp) && aligned16(gamma) &&
                aligned16(dgb_partial));
  if (!supported_C(C)) return EQH_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the module only ever uses two shapes: primary grouping (one operand gathered, dU own) and secondary
  // grouping (the other operand and dU gathered through the same index list)
  const bool ga = ia != nullptr, gb = ib != nullptr, gu = iu != nullptr;
  if ((ga && gb) || (!ga && !gb)) return EQH_ERR_UNSUPPORTED;
  int64_t blocks = (n_seg + 7) / 8;
  if (blocks > (int64_t)kSMs * 16) blocks = (int64_t)kSMs * 16;
  if (blocks < 1) blocks = 1;
  const unsigned grid = dgb_partial ? (unsigned)kPartialRows : (unsigned)blocks;   // partial rows: fixed, all written
  if (!dgb_partial && n_seg == 0) return EQH_OK;
#define EQH_BWD(LNV, DGBV, GAV, GBV, GUV)                                                                       \
  launch_kernel(k_fused_bwd<V, LNV, DGBV, GAV, GBV, GUV>, dim3(grid), dim3(256), 0, st, A, B, ia, ib, iu, rowptr, n_seg, gamma, eps, dU, \
                                                                 u_rowptr, dOp, dgb_partial)
#define EQH_BWD_G(LNV, DGBV)                                   \
  do {                                                         \
    if (ga && gu) EQH_BWD(LNV, DGBV, true, false, true);       \
    else if (ga) EQH_BWD(LNV, DGBV, true, false, false);       \
    else if (gu) EQH_BWD(LNV, DGBV, false, true, true);        \
    else EQH_BWD(LNV, DGBV, false, true, false);               \
  } while (0)
#define EQH_BWD_DGB(GAV, GBV, GUV)                                                                                 \
  launch_kernel(k_fused_bwd_dgb<V, GAV, GBV, GUV>, dim3(grid), dim3(256), sizeof(float) * 16 * 32 * V, st, A, B, ia, ib, iu, rowptr, \
                n_seg, gamma, eps, dU, u_rowptr, dOp, dgb_partial)
  EQH_DISPATCH_V(C, {
    if (dgb_partial && legacy_dgb()) EQH_BWD_G(true, true);
    else if (dgb_partial) {
      if (ga && gu) EQH_BWD_DGB(true, false, true);
      else if (ga) EQH_BWD_DGB(true, false, false);
      else if (gu) EQH_BWD_DGB(false, true, true);
      else EQH_BWD_DGB(false, true, false);
    }
    else if (gamma) EQH_BWD_G(true, false);
    else EQH_BWD_G(false, false);
  });
#undef EQH_BWD_DGB
#undef EQH_BWD_G
#undef EQH_BWD
  EQH_LAUNCH_RC();
}

int eqh_rowln_fwd(const float* x, const float* gamma, const float* beta, float eps, int relu, int64_t rows, int64_t C,
                  float* y, eqh_stream_t stream) {
  return eqh_rowln_fwd_ld(x, C, gamma, beta, eps, relu, rows, C, y, C, stream);
}

int eqh_rowln_fwd_ld(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int relu, int64_t rows,
                     int64_t C, float* y, int64_t ldy, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(rows >= 0 && gamma && beta && (rows == 0 || (x && y)));
  EQH_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta));
  EQH_CHECK_ARG(ldx >= C && ldy >= C && ldx % 4 == 0 && ldy % 4 == 0);
  if (!supported_C(C)) return EQH_ERR_UNSUPPORTED;
  if (rows == 0) return EQH_OK;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > (int64_t)kSMs * 16) blocks = (int64_t)kSMs * 16;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EQH_DISPATCH_V(C, {
    if (relu) launch_kernel(k_rowln_fwd<V, true>, dim3((unsigned)blocks), dim3(256), 0, st, x, gamma, beta, eps, rows, y, ldx, ldy);
    else launch_kernel(k_rowln_fwd<V, false>, dim3((unsigned)blocks), dim3(256), 0, st, x, gamma, beta, eps, rows, y, ldx, ldy);
  });
  EQH_LAUNCH_RC();
}

int eqh_rowln_bwd(const float* x, const float* gamma, float eps, int relu, const float* dy, int64_t rows, int64_t C,
                  float* dx, float* dgb_partial, eqh_stream_t stream) {
  return eqh_rowln_bwd_ld(x, C, gamma, eps, relu, dy, C, rows, C, dx, C, nullptr, dgb_partial, stream);
}

int eqh_rowln_bwd_ld(const float* x, int64_t ldx, const float* gamma, float eps, int relu, const float* dy, int64_t lddy,
                     int64_t rows, int64_t C, float* dx, int64_t lddx, const float* add, float* dgb_partial,
                     eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(rows >= 0 && gamma && dgb_partial && (rows == 0 || (x && dy && dx)));
  EQH_CHECK_ARG(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma) && aligned16(dgb_partial) && aligned16(add));
  EQH_CHECK_ARG(ldx >= C && lddy >= C && lddx >= C && ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0);
  if (!supported_C(C)) return EQH_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  EQH_DISPATCH_V(C, {
    if (relu) launch_kernel(k_rowln_bwd<V, true>, dim3(kPartialRows), dim3(256), 0, st, x, gamma, eps, dy, rows, dx, dgb_partial, ldx, lddy, lddx, add);
    else launch_kernel(k_rowln_bwd<V, false>, dim3(kPartialRows), dim3(256), 0, st, x, gamma, eps, dy, rows, dx, dgb_partial, ldx, lddy, lddx, add);
  });
  EQH_LAUNCH_RC();
}

size_t eqh_colsum_workspace_bytes(int64_t rows, int64_t cols) {
  const int64_t chunks = rows > 64 * 64 ? 64 : (rows + 63) / 64 > 0 ? (rows + 63) / 64 : 1;
  return sizeof(float) * (size_t)chunks * (size_t)cols;
}

// Column sums of a tall [rows, cols] matrix with row stride ld (bias gradients): out (+)= sum_r roww[r] * x[r,:],
// two fixed-order stages through ws.
int eqh_colsum_tall(const float* x, int64_t rows, int64_t cols, int64_t ld, const float* roww, float* out, int accumulate,
                    float* ws, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols && (cols == 0 || (x && out)));
  if (cols == 0) return EQH_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows <= 64 && !roww) {
    launch_kernel(k_colsum, dim3((unsigned)((cols + 31) / 32)), dim3(256), 0, st, x, rows, cols, ld, out, 1.f, accumulate);
    EQH_LAUNCH_RC();
  }
  EQH_CHECK_ARG(ws != nullptr);
  const int64_t chunks = rows > 64 * 64 ? 64 : (rows + 63) / 64 > 0 ? (rows + 63) / 64 : 1;
  const int64_t rpc = (rows + chunks - 1) / chunks;
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)chunks);
  launch_kernel(k_colsum_chunks, dim3(grid), dim3(256), 0, st, x, rows, cols, ld, roww, rpc, ws);
  launch_kernel(k_colsum, dim3((unsigned)((cols + 31) / 32)), dim3(256), 0, st, ws, chunks, cols, cols, out, 1.f, accumulate);
  EQH_LAUNCH_RC();
}

int eqh_colsum(const float* partial, int64_t rows, int64_t cols, int64_t ld, float* out, float scale, int accumulate,
               eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(rows >= 0 && cols >= 0 && ld >= cols && (cols == 0 || (partial && out)));
  if (cols == 0) return EQH_OK;
  launch_kernel(k_colsum, dim3((unsigned)((cols + 31) / 32)), dim3(256), 0, static_cast<cudaStream_t>(stream), partial, rows, cols, ld, out, scale,
                                                                                      accumulate);
  EQH_LAUNCH_RC();
}

}  // extern "C"
