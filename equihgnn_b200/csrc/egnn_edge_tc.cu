// EGNN edge model phi_e + neighbour pooling, forward, on the tcgen05 tensor cores.
//
//   pre  = P[i,:] + Q[j,:] + d2 * wc ;  hid = SiLU(pre)                (CUDA cores: gather, add, SiLU)
//   s2   = hid . W2^T + b2                                             (tensor cores, M=128 edges, N=16, K=Hp)
//   m_ij = SiLU(s2) ;  m_i = sum_k m_ij                                (epilogue from TMEM)
//
// Replaces egnn_layer.py:298-310,357-358.  fp32-faithful: the [128 x 32] hid tile and the [16 x 32] W2 tile of each
// K block are written to shared memory as (raw, lo = x - trunc_tf32(x)) pairs in the K-major 128B-swizzled UMMA
// layout and multiplied as lo.hi + hi.lo + hi.hi with tcgen05.mma.kind::tf32 (see gemm_tf32x3.cu), so the
// 1e-5 tolerance of the fp32 bucket holds while the 16-wide contraction -- whose W2 reads made the SIMT kernel
// shared-memory-bandwidth-bound -- costs 8 tensor-core cycles per K step.
//
// Persistent CTAs (one per SM), 128 edges (8 atoms x 16 neighbours) per tile, warp-specialised:
//   warps 0-7   producers: gather Q[j] (one 128-byte segment per edge and K block; a thread owns four neighbours of one
//               atom, so P[i] is read once per four edges), add P[i] and d2*wc, SiLU,
//               split, store swizzled; fence.proxy.async; arrive on the stage's `full` barrier (4-stage ring)
//   warp 12     MMA issuer (one thread): 4 k-steps x 3 MMAs per K block, tcgen05.commit frees the stage
//   warps 8-11  epilogue: tcgen05.ld the [128 x 16] accumulators (double-buffered in TMEM, so the epilogue of
//               tile t overlaps the production of tile t+1), + b2, SiLU, store S2 / m_ij, 16-lane shuffle sum -> m_i
#include "tc_common.cuh"

namespace eqh {
namespace edge_tc {
using namespace eqh::tc;

constexpr int kM = EQH_M_DIM;       // 16 outputs
constexpr int kNbr = EQH_KNN_K;     // 16 neighbours
constexpr int BM = 128, BK = 32, STAGES = 4;
constexpr int kProducers = 256;
constexpr int kThreads = 13 * 32;   // 8 producer warps, 4 epilogue warps, 1 MMA warp
constexpr uint32_t kABytes = BM * BK * 4;        // 16 KB
constexpr uint32_t kBBytes = 1024 * 2;           // 16 rows x 128 B = 2 KB (two 8-row swizzle groups)
constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // hid raw, hid lo, W2 raw, W2 lo = 36 KB
constexpr uint32_t kAccCols = 128;               // per TMEM buffer: 8 accumulators of 16 columns (see the MMA issuer)

// SiLU with the two SFU approximations (ex2.approx, rcp.approx: ~1 ulp each; the correctly rounded __frcp_rn costs
// ~10 instructions and made this kernel issue-bound).  |error| < 3e-7 relative, far inside the 1e-5 bar.
__device__ __forceinline__ float fast_silu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return x * r;
}

template <int PF>   // K blocks of gather prefetch held in registers: 2 on large batches (L2 / HBM latency), 1 when everything is L2-resident
__global__ void __launch_bounds__(kThreads, 1)
k_edge_fwd_tc(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ wc,
              const float* __restrict__ W2, const float* __restrict__ b2, const int32_t* __restrict__ idx,
              const float* __restrict__ d2, int64_t N, int H, int64_t ldpq, float* __restrict__ S2,
              float* __restrict__ Mij, float* __restrict__ Mi, int64_t ldmi, const int32_t* __restrict__ order) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * kStageBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_edges = N * kNbr;
  const int64_t n_tiles = (n_edges + BM - 1) / BM;
  const int nkb = H / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kProducers);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) tmem_alloc(tmem_slot, 2 * kAccCols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ============================ producers ============================
    const int t = threadIdx.x;
    const int c16 = t & 7;          // 16-byte chunk (4 floats) inside the 128-byte K block row
    const int rb = t >> 3;          // 0..31: rows 4 rb .. 4 rb + 3 of the tile (four neighbours of ONE atom: P[i] is read once)
    uint32_t it = 0;                // global K-block counter (ring position)
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const float* prow = P;
      const float* qrow[4];
      float dd[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t slot = tile * BM + 4 * rb + j;            // tile row: atom slot (slot >> 4), neighbour (slot & 15)
        ok[j] = slot < n_edges;
        // atoms are visited in `order` (the cell order of the kNN grid) when given: spatially close atoms share
        // neighbours, so the Q[j] rows gathered by one tile and by the tiles running next to it hit in L2
        const int64_t ec = ok[j] ? ((order ? (int64_t)__ldg(order + (slot >> 4)) : (slot >> 4)) << 4) + (slot & 15) : 0;
        if (j == 0) prow = P + (ec >> 4) * ldpq + c16 * 4;
        qrow[j] = Q + (int64_t)__ldg(idx + ec) * ldpq + c16 * 4;
        dd[j] = __ldg(d2 + ec);
      }
      // software pipeline, two K blocks deep: the gathers of blocks kb+1 and kb+2 are in flight while block kb is turned
      // into SiLU tiles (one block ahead left 65 % of the stall samples on the gather's scoreboard)
      struct Fetched {
        float4 q[4], p, w, wt;
      };
      auto fetch = [&](Fetched& f, int kb) {
        const int k0 = kb * BK;
#pragma unroll
        for (int j = 0; j < 4; ++j) f.q[j] = __ldg(reinterpret_cast<const float4*>(qrow[j] + k0));
        f.p = __ldg(reinterpret_cast<const float4*>(prow + k0));
        f.w = __ldg(reinterpret_cast<const float4*>(wc + k0 + c16 * 4));
        f.wt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < 128) f.wt = __ldg(reinterpret_cast<const float4*>(W2 + (int64_t)(t >> 3) * H + k0 + c16 * 4));
      };
      auto step = [&](Fetched& f, int kb) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        ++it;
        float4 cq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cq[j] = f.q[j];
        const float4 cp = f.p, cw = f.w, cwt = f.wt;
        if (kb + PF < nkb) fetch(f, kb + PF);
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * kStageBytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 4 * rb + j;
          float4 h;
          h.x = fast_silu((cp.x + cq[j].x) + dd[j] * cw.x);
          h.y = fast_silu((cp.y + cq[j].y) + dd[j] * cw.y);
          h.z = fast_silu((cp.z + cq[j].z) + dd[j] * cw.z);
          h.w = fast_silu((cp.w + cq[j].w) + dd[j] * cw.w);
          if (!ok[j]) h = make_float4(0.f, 0.f, 0.f, 0.f);
          const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c16 ^ (r & 7)) << 4);   // 128B swizzle
          *reinterpret_cast<float4*>(st + off) = h;
          *reinterpret_cast<float4*>(st + kABytes + off) = make_float4(tf32_lo(h.x), tf32_lo(h.y), tf32_lo(h.z), tf32_lo(h.w));
        }
        if (t < 128) {
          const int o = t >> 3;
          const uint32_t off = (uint32_t)o * 128u + (uint32_t)((c16 ^ (o & 7)) << 4);
          *reinterpret_cast<float4*>(st + 2 * kABytes + off) = cwt;
          *reinterpret_cast<float4*>(st + 2 * kABytes + kBBytes + off) =
              make_float4(tf32_lo(cwt.x), tf32_lo(cwt.y), tf32_lo(cwt.z), tf32_lo(cwt.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&full[s]);
      };
      Fetched fa;
      fetch(fa, 0);
      if constexpr (PF == 2) {
        Fetched fb;
        if (nkb > 1) fetch(fb, 1);
        for (int kb = 0; kb < nkb; kb += 2) {
          step(fa, kb);
          if (kb + 1 < nkb) step(fb, kb + 1);
        }
      } else {
        for (int kb = 0; kb < nkb; ++kb) step(fa, kb);
      }
    }
  } else if (warp == 12) {
    // ============================ MMA issuer (warp-uniform loop, one elected lane issues) ============================
    {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kM >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(&tmem_empty[buf], ((tcount >> 1) & 1) ^ 1);     // epilogue has drained this accumulator buffer
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * kAccCols;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t aH = smem_u32(smem + (size_t)s * kStageBytes), aL = aH + kABytes;
          const uint32_t bH = aH + 2 * kABytes, bL = bH + kBBytes;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
              const uint32_t step = (uint32_t)(kb * (BK / 8) + ks);
              const uint64_t dAh = make_desc(aH + ks * 32, 0, 1024, 2), dAl = make_desc(aL + ks * 32, 0, 1024, 2);
              const uint64_t dBh = make_desc(bH + ks * 32, 0, 1024, 2), dBl = make_desc(bL + ks * 32, 0, 1024, 2);
              // Back-to-back MMAs into the SAME accumulator serialise on the accumulate latency, and each truncating
              // accumulation adds rounding error: spread the products over 8 independent 16-column accumulators
              // (hi.hi: one per k-step of the block; lo.hi and hi.lo: two each, alternating), summed in the epilogue.
              umma_tf32(acc + (4 + (ks & 1)) * 16, dAl, dBh, idesc, step >= 2);
              umma_tf32(acc + (6 + (ks & 1)) * 16, dAh, dBl, idesc, step >= 2);
              umma_tf32(acc + ks * 16, dAh, dBh, idesc, kb != 0);
            }
            umma_commit(&empty[s]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&tmem_full[buf]);
        __syncwarp();
      }
    }
  } else {
    // ============================ epilogue (warps 8-11) ============================
    const int q = warp & 3;            // TMEM lane quadrant
    float bias[kM];
#pragma unroll
    for (int o = 0; o < kM; ++o) bias[o] = __ldg(b2 + o);
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1;
      mbar_wait(&tmem_full[buf], (tcount >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem_base + buf * kAccCols + ((uint32_t)(q * 32) << 16);
      float v[kM], w[kM];
      tmem_ld16(ta + 7 * 16, v);                     // small (lo) terms first, the hi.hi accumulators last
#pragma unroll
      for (int a = 6; a >= 0; --a) {
        tmem_ld16(ta + a * 16, w);
#pragma unroll
        for (int o = 0; o < kM; ++o) v[o] += w[o];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tmem_empty[buf]);                 // accumulators are in registers: the buffer may be reused
      const int64_t slot = tile * BM + q * 32 + lane;
      const bool ok = slot < n_edges;
      const int64_t e = ok ? ((order ? (int64_t)__ldg(order + (slot >> 4)) : (slot >> 4)) << 4) + (slot & 15) : 0;
      float m[kM];
#pragma unroll
      for (int o = 0; o < kM; ++o) {
        v[o] += bias[o];
        m[o] = ok ? silu_f(v[o]) : 0.f;
      }
      if (ok) {
        float4* s2p = reinterpret_cast<float4*>(S2 + e * kM);
#pragma unroll
        for (int o = 0; o < kM; o += 4) s2p[o >> 2] = make_float4(v[o], v[o + 1], v[o + 2], v[o + 3]);
        if (Mij) {
          float4* mp = reinterpret_cast<float4*>(Mij + e * kM);
#pragma unroll
          for (int o = 0; o < kM; o += 4) mp[o >> 2] = make_float4(m[o], m[o + 1], m[o + 2], m[o + 3]);
        }
      }
      // m_i: the 16 edges of an atom sit in 16 consecutive lanes
#pragma unroll
      for (int o = 0; o < kM; ++o) {
        float x = m[o];
        x += __shfl_xor_sync(kFull, x, 8);
        x += __shfl_xor_sync(kFull, x, 4);
        x += __shfl_xor_sync(kFull, x, 2);
        x += __shfl_xor_sync(kFull, x, 1);
        m[o] = x;
      }
      if ((lane & 15) == 0 && ok) {
        float4* mip = reinterpret_cast<float4*>(Mi + (e >> 4) * ldmi);
#pragma unroll
        for (int o = 0; o < kM; o += 4) mip[o >> 2] = make_float4(m[o], m[o + 1], m[o + 2], m[o + 3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 12) tmem_dealloc(tmem_base, 2 * kAccCols);
}

}  // namespace edge_tc
}  // namespace eqh

static int edge_fwd_tc_impl(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2, const float* b2,
                            const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij, float* Mi,
                            int64_t ldmi, const int32_t* order, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::edge_tc;
  EQH_CHECK_ARG(P && Q && wc && W2 && b2 && idx && d2 && S2 && Mi && N >= 0);
  EQH_CHECK_ARG(ldpq >= H && ldpq % 4 == 0 && ldmi >= kM && ldmi % 4 == 0);
  EQH_CHECK_ARG(aligned16(P) && aligned16(Q) && aligned16(wc) && aligned16(W2) && aligned16(S2) && aligned16(Mi) &&
                aligned16(Mij));
  if (H <= 0 || H % 64 != 0) return EQH_ERR_UNSUPPORTED;
  if (N == 0) return EQH_OK;
  constexpr size_t smem = (size_t)STAGES * kStageBytes + 128;
  const int64_t tiles = (N * kNbr + BM - 1) / BM;
  const unsigned grid = (unsigned)(tiles < kSMs ? tiles : kSMs);
  auto kern = N >= 8192 ? k_edge_fwd_tc<2> : k_edge_fwd_tc<1>;      // measured: QM9-128 45 vs 53 us, Molecule3D-1024 821 vs 668 us
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  launch_kernel(kern, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), P, Q, wc, W2, b2, idx, d2, N, (int)H, ldpq, S2,
                Mij, Mi, ldmi, order);
  EQH_LAUNCH_RC();
}

extern "C" int eqh_egnn_edge_fwd_tc(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                                    const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij,
                                    float* Mi, eqh_stream_t stream) {
  return edge_fwd_tc_impl(P, Q, H, wc, W2, b2, idx, d2, N, H, S2, Mij, Mi, EQH_M_DIM, nullptr, stream);
}

extern "C" int eqh_egnn_edge_fwd_tc_ordered(const float* P, const float* Q, const float* wc, const float* W2, const float* b2,
                                            const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2, float* Mij,
                                            float* Mi, const int32_t* order, eqh_stream_t stream) {
  return edge_fwd_tc_impl(P, Q, H, wc, W2, b2, idx, d2, N, H, S2, Mij, Mi, EQH_M_DIM, order, stream);
}

// Strided variant: P / Q are column blocks of one [N, ldpq] matrix (the merged h.[W1a|W1b]^T GEMM) and m_i is written
// into a column block of a wider row (the node MLP's input cat(LN(h), m_i), egnn_layer.py:360: row stride ldmi).
extern "C" int eqh_egnn_edge_fwd_tc_ld(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2,
                                       const float* b2, const int32_t* idx, const float* d2, int64_t N, int64_t H, float* S2,
                                       float* Mij, float* Mi, int64_t ldmi, const int32_t* order, eqh_stream_t stream) {
  return edge_fwd_tc_impl(P, Q, ldpq, wc, W2, b2, idx, d2, N, H, S2, Mij, Mi, ldmi, order, stream);
}
