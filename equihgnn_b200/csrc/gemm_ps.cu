// 3xTF32 row GEMM, persistent variant for the multi-wave problems (tens of thousands of rows: the PCQM4Mv2 /
// Molecule3D / OPV shaped batches).  Same data path as gemm_ts.cu (A: TMA -> shared memory -> registers -> tensor
// memory hi|lo, B: TMA -> shared memory hi / lo, tcgen05.mma TS form), but one CTA per SM walks the 128 x 64 output
// tiles and the accumulators are double-buffered in tensor memory, so that the epilogue of tile t (TMEM reads at
// 64 B/cycle, staging, stores: ~2 us) and the per-CTA set-up overlap the main loop of tile t+1 instead of adding to
// it -- with one tile per CTA those fixed costs were ~60 % of a CTA's life at K = 256.
//
//   warp 0       TMA producer (A and B tiles, ring continues across tiles)
//   warp 1       MMA issuer; waits for the epilogue to drain the accumulator buffer it is about to overwrite
//   warps 4-7    A rows: shared memory -> registers -> TMEM (one row per thread)
//   warps 8-11   epilogue (TMEM lane quadrant = warp % 4), own staging buffer
//   warps 12-15  B split
// Tiles are numbered n-fastest, so the CTAs that run concurrently share A row blocks in L2.
// TMEM columns: accumulator buffer b at b*192 (three 64-column tiles), A ring at 384 (2 stages x 64).
#include "gemm_common.cuh"

namespace eqh {
namespace gemm {

__device__ __forceinline__ void umma_tf32_ts2(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

__device__ __forceinline__ void tmem_st32b(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}

namespace ps {
constexpr int BN = 64;
constexpr int kThreadsPs = 512;
constexpr int SA = 2, SB = 5;
constexpr uint32_t kABytes = kTileBytes;                 // 16 KB
constexpr uint32_t kBBytes = BN * BK * 4;                // 8 KB
constexpr uint32_t kStageBytes = kABytes + 2 * kBBytes;  // 32 KB
constexpr size_t kRing = (size_t)SB * kStageBytes;
constexpr size_t kStaging = (size_t)BM * (BN + 4) * 4;   // epilogue staging, separate from the ring
constexpr size_t kSmem = kRing + kStaging + 1024 + 2 * BN * 4;
constexpr uint32_t kAccBuf = 3 * BN;                     // 192 columns per accumulator buffer
constexpr uint32_t kAcol = 2 * kAccBuf;                  // 384
static_assert(kAcol + SA * 64 <= 512, "TMEM budget");
static_assert(kSmem <= 227 * 1024, "shared memory budget");
}  // namespace ps

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(ps::kThreadsPs, 1) k_gemm_ps(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, const Params p,
                                                               int tiles_n, int n_tiles) {
  using namespace ps;
  using AM = AccMap<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* staging = smem + kRing;
  uint64_t* b_raw = reinterpret_cast<uint64_t*>(smem + kRing + kStaging);
  uint64_t* b_full = b_raw + SB;
  uint64_t* b_empty = b_full + SB;
  uint64_t* a_sfree = b_empty + SB;
  uint64_t* a_full = a_sfree + SB;
  uint64_t* a_empty = a_full + SA;
  uint64_t* tmem_full = a_empty + SA;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* sbias = reinterpret_cast<float*>(smem + kRing + kStaging + 1024);
  static_assert((4 * SB + 2 * SA + 4) * 8 + 4 <= 1024, "barrier block");

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SB; ++s) {
      mbar_init(&b_raw[s], 1);
      mbar_init(&b_full[s], 4);
      mbar_init(&b_empty[s], 1);
      mbar_init(&a_sfree[s], 4);
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(&a_full[s], 4);
      mbar_init(&a_empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  auto stageA = [&](int s) { return smem + (size_t)s * kStageBytes; };
  auto stageB = [&](int s) { return smem + (size_t)s * kStageBytes + kABytes; };
  auto stageBlo = [&](int s) { return smem + (size_t)s * kStageBytes + kABytes + kBBytes; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % SB;
        const uint32_t ph = (it / SB) & 1;
        mbar_wait(&b_empty[s], ph ^ 1);
        mbar_wait(&a_sfree[s], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&b_raw[s], kABytes + kBBytes);
          const int k0 = i * BK;
          if constexpr (A_KMAJOR) tma_load_2d(stageA(s), &tmA, &b_raw[s], k0, m0);
          else tma_load_2d(stageA(s), &tmA, &b_raw[s], m0, k0);
          if constexpr (B_KMAJOR) {
            tma_load_2d(stageB(s), &tmB, &b_raw[s], k0, n0);
          } else {
#pragma unroll
            for (int a = 0; a < BN / 32; ++a) tma_load_2d(stageB(s) + a * (BK * 128), &tmB, &b_raw[s], n0 + a * 32, k0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_KMAJOR ? 0u : 1u) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    constexpr uint32_t lboB = B_KMAJOR ? 0u : BK * 128u;
    constexpr uint32_t kstepB = B_KMAJOR ? 32u : 1024u;
    constexpr uint32_t sboB = B_KMAJOR ? 1024u : 512u;
    constexpr uint32_t ltB = B_KMAJOR ? 2u : 1u;
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1;
      mbar_wait(&tmem_empty[buf], ((tcount >> 1) & 1) ^ 1);      // the epilogue has drained this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + buf * kAccBuf;
      for (int i = 0; i < nkb; ++i, ++it) {
        const int sa = it % SA, sb = it % SB;
        mbar_wait(&b_full[sb], (it / SB) & 1);
        mbar_wait(&a_full[sa], (it / SA) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bH = smem_u32(stageB(sb)), bL = smem_u32(stageBlo(sb));
        const uint32_t aH = tmem_base + kAcol + sa * 64, aL = aH + 32;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t dBh = make_desc(bH + ks * kstepB, lboB, sboB, ltB), dBl = make_desc(bL + ks * kstepB, lboB, sboB, ltB);
            umma_tf32_ts2(acc + AM::lh(ks) * BN, aL + ks * 8, dBh, idesc, !AM::first_lh(i, ks));
            umma_tf32_ts2(acc + AM::hh(ks) * BN, aH + ks * 8, dBh, idesc, !AM::first_big(i, ks));
            umma_tf32_ts2(acc + AM::hl(ks) * BN, aH + ks * 8, dBl, idesc, !AM::first_hl(i, ks));
          }
          umma_commit(&a_empty[sa]);
          umma_commit(&b_empty[sb]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full[buf]);
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A rows: shared memory -> registers -> TMEM (hi | lo) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + kAcol;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % SB, sa = it % SA;
        mbar_wait(&b_raw[s], (it / SB) & 1);
        uint32_t v[32];
        if constexpr (A_KMAJOR) {
          const uint8_t* row = stageA(s) + r * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 t4 = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
            v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
          }
        } else {
          const uint32_t* col = reinterpret_cast<const uint32_t*>(stageA(s)) + r;
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = col[c * 128];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_sfree[s]);
        mbar_wait(&a_empty[sa], ((it / SA) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_st32b(lane_addr + sa * 64, v);
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float x = __uint_as_float(v[c]);
          v[c] = __float_as_uint(x - __uint_as_float(v[c] & 0xffffe000u));
        }
        tmem_st32b(lane_addr + sa * 64 + 32, v);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[sa]);
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ===================== epilogue =====================
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
      const uint32_t buf = tcount & 1;
      epilogue_tile<BN, AM>(staging, sbias, tmem_base + buf * kAccBuf, &tmem_full[buf], warp, lane, nkb, p, m0, n0,
                            (int)threadIdx.x - 256, (tcount >> 1) & 1, &tmem_empty[buf]);
    }
  } else if (warp >= 12) {
    // ===================== B split (hi / lo) =====================
    const int t = ((warp - 12) << 5) + lane;      // 0..127
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int i = 0; i < nkb; ++i, ++it) {
        const int s = it % SB;
        mbar_wait(&b_raw[s], (it / SB) & 1);
        const float4* bh = reinterpret_cast<const float4*>(stageB(s));
        float4* bl = reinterpret_cast<float4*>(stageBlo(s));
        constexpr int NB = kBBytes / 16 / 128;     // 4
        float4 v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) v[j] = bh[j * 128 + t];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          float4 l;
          l.x = tf32_lo(v[j].x); l.y = tf32_lo(v[j].y); l.z = tf32_lo(v[j].z); l.w = tf32_lo(v[j].w);
          bl[j * 128 + t] = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_full[s]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

template <bool AK, bool BK_>
static int launch_ps_one(const CUtensorMap& a, const CUtensorMap& b, const Params& p, int tiles_n, int n_tiles, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_gemm_ps<AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps::kSmem);
  if (e != cudaSuccess) return (int)e;
  const int grid = n_tiles < kSMs ? n_tiles : kSMs;
  launch_kernel(k_gemm_ps<AK, BK_>, dim3(grid), dim3(ps::kThreadsPs), ps::kSmem, st, a, b, p, tiles_n, n_tiles);
  EQH_LAUNCH_RC();
}

// 128 x 64 tiles, no split-K.  tmA / tmB as for the TMEM-A kernel with BN = 64.
int launch_ps(bool a_kmajor, bool b_kmajor, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p, cudaStream_t st) {
  const int tiles_n = (p.N + ps::BN - 1) / ps::BN, tiles_m = (p.M + BM - 1) / BM;
  const int n_tiles = tiles_n * tiles_m;
  if (a_kmajor && b_kmajor) return launch_ps_one<true, true>(tmA, tmB, p, tiles_n, n_tiles, st);
  if (a_kmajor && !b_kmajor) return launch_ps_one<true, false>(tmA, tmB, p, tiles_n, n_tiles, st);
  if (!a_kmajor && !b_kmajor) return launch_ps_one<false, false>(tmA, tmB, p, tiles_n, n_tiles, st);
  return launch_ps_one<false, true>(tmA, tmB, p, tiles_n, n_tiles, st);
}

}  // namespace gemm
}  // namespace eqh
