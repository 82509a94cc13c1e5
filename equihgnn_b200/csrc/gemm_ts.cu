// 3xTF32 row GEMM, second kernel: the A operand reaches the tensor core through tensor memory.
//
// gemm_tf32x3.cu is bound by shared-memory bandwidth, not by the tensor core: per 32-deep K block it moves the
// 16 KB A tile through shared memory six times (TMA write, split read, lo write, three tcgen05.mma operand reads).
// Here the A tile crosses shared memory twice (TMA write, one read): each of the 128 threads of warps 4-7 owns one
// row of the tile (= one TMEM lane), reads its 32 K values from the landed tile, splits them hi / lo in registers
// and writes both halves into tensor memory (tcgen05.st); the MMAs take A from TMEM ("TS" form) and only the B tile
// (the weights: BN x 32) is read from shared memory by the tensor core.
//
//   warp 0 (1 thread)  TMA producer: A tile (K-major: 128B-swizzled rows; MN-major: plain [32 k][128 m]) and B tile
//   warps 2,3,8,9      split the landed B tile into hi (in place) / lo, fence.proxy.async, signal
//   warps 4-7          A rows: shared memory -> registers -> TMEM hi | lo (one row per thread); then the epilogue
//   warp 1 (1 thread)  per K block 4 k-steps x 3 tcgen05.mma (A from TMEM); commits free the TMEM and smem stages
// TMEM columns: [0, 3*BN) accumulators (as in gemm_tf32x3.cu), then the A ring: 64 columns (hi | lo) per stage.
// (A first version loaded the rows straight from global memory: one 128-byte line per thread and instruction,
// which the L1 wavefront queue serialises -- slower than the shared-memory kernel.)
#include "gemm_common.cuh"

namespace eqh {
namespace gemm {

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// 32 consecutive TMEM columns of this thread's lane <- 32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
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

template <int BN>
struct TsCfg {
  static constexpr int SA = BN == 128 ? 2 : 4;                  // A stages in TMEM (64 columns each)
  static constexpr int SB = BN == 128 ? 4 : (BN == 64 ? 6 : 8); // shared-memory stages (A raw, B raw/hi, B lo)
  static constexpr uint32_t kABytes = kTileBytes;               // 16 KB
  static constexpr uint32_t kBBytes = BN * BK * 4;
  static constexpr uint32_t kStageBytes = kABytes + 2 * kBBytes;
  static constexpr uint32_t kAcol = AccMap<BN>::NA * BN; // first TMEM column of the A ring
  static constexpr size_t kRing = (size_t)SB * kStageBytes;
  static constexpr size_t kSmem = kRing + 512 + 2 * BN * 4;     // ring + barriers + bias / bias2 tiles
  static_assert(kAcol + SA * 64 <= 512, "TMEM budget");
  static_assert(kRing >= (size_t)BM * (BN + 4) * 4, "epilogue staging must fit in the B ring");
};

// One output tile (bx, by) of one problem, K range of split bz of nz; `tile` indexes the problem's split-K arrival counter.
// Called by the single-problem kernel (coordinates = blockIdx) and by the grouped kernel (coordinates decoded from a
// linear CTA index over all problems of the group).
template <int BN, bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ void gemm_ts_tile(const CUtensorMap* __restrict__ ptmA, const CUtensorMap* __restrict__ ptmB,
                                             const Params& p, int bx, int by, int bz, int nz, int tile) {
  const CUtensorMap& tmA = *ptmA;
  const CUtensorMap& tmB = *ptmB;
  using Cfg = TsCfg<BN>;
  using AM = AccMap<BN>;
  constexpr int SA = Cfg::SA, SB = Cfg::SB;
  static_assert(Cfg::kAcol == AM::NA * BN, "A ring starts after the accumulators");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* b_raw = reinterpret_cast<uint64_t*>(smem + Cfg::kRing);     // TMA landed
  uint64_t* b_full = b_raw + SB;                                        // hi/lo ready for the tensor core
  uint64_t* b_empty = b_full + SB;
  uint64_t* a_sfree = b_empty + SB;                                     // A smem tile consumed (read into registers)
  uint64_t* a_full = a_sfree + SB;
  uint64_t* a_empty = a_full + SA;
  uint64_t* tmem_full = a_empty + SA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kRing + 512);
  static_assert((4 * SB + 2 * SA + 1) * 8 + 4 <= 512, "barrier block");
  __shared__ int s_last;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {       // the descriptors live in the kernel's parameter space: start their fetch before the setup
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(ptmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(ptmB)) : "memory");
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = bx * BN, m0 = by * BM;
  if (threadIdx.x == 0) EQH_STAMP(0);
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = bz * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(kb_end - kb_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SB; ++s) {
      mbar_init(&b_raw[s], 1);
      mbar_init(&b_full[s], 4);          // one elected arrival per B-split warp
      mbar_init(&b_empty[s], 1);
      mbar_init(&a_sfree[s], 4);         // one elected arrival per A warp
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(&a_full[s], 4);          // one elected arrival per A warp
      mbar_init(&a_empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();        // everything above (barriers, tensor memory) overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) EQH_STAMP(1);

  auto stageA = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes; };
  auto stageB = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes + Cfg::kABytes; };
  auto stageBlo = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes + Cfg::kABytes + Cfg::kBBytes; };

  if (warp == 0) {
    // ===================== TMA producer (A and B) =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % SB;
        const uint32_t ph = (i / SB) & 1;
        mbar_wait(&b_empty[s], ph ^ 1);
        mbar_wait(&a_sfree[s], ph ^ 1);
        mbar_expect_tx(&b_raw[s], Cfg::kABytes + Cfg::kBBytes);
        const int k0 = (kb_begin + i) * BK;
        if constexpr (A_KMAJOR) tma_load_2d(stageA(s), &tmA, &b_raw[s], k0, m0);    // box {32 K, 128 M}, 128B swizzle
        else tma_load_2d(stageA(s), &tmA, &b_raw[s], m0, k0);                       // box {128 M, 32 K}, no swizzle
        if constexpr (B_KMAJOR) {
          tma_load_2d(stageB(s), &tmB, &b_raw[s], k0, n0);                 // box {32 K, BN}
        } else {
#pragma unroll
          for (int a = 0; a < BN / 32; ++a)                                // boxes {32 N, 32 K}
            tma_load_2d(stageB(s) + a * (BK * 128), &tmB, &b_raw[s], n0 + a * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((B_KMAJOR ? 0u : 1u) << 16) |
                                 ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t lboB = B_KMAJOR ? 0u : BK * 128u;
      constexpr uint32_t kstepB = B_KMAJOR ? 32u : 1024u;      // bytes per UMMA_K = 8
      constexpr uint32_t sboB = B_KMAJOR ? 1024u : 512u;
      constexpr uint32_t ltB = B_KMAJOR ? 2u : 1u;
      for (int i = 0; i < nkb; ++i) {
        const int sa = i % SA, sb = i % SB;
        mbar_wait(&b_full[sb], (i / SB) & 1);
        mbar_wait(&a_full[sa], (i / SA) & 1);
        if (lane == 0 && i == 0) EQH_STAMP(3);
        if (lane == 0 && i < 8) EQH_STAMP(48 + i);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bH = smem_u32(stageB(sb)), bL = smem_u32(stageBlo(sb));
        const uint32_t aH = tmem_base + Cfg::kAcol + sa * 64, aL = aH + 32;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t dBh = make_desc(bH + ks * kstepB, lboB, sboB, ltB), dBl = make_desc(bL + ks * kstepB, lboB, sboB, ltB);
            umma_tf32_ts(tmem_base + AM::lh(ks) * BN, aL + ks * 8, dBh, idesc, !AM::first_lh(i, ks));
            umma_tf32_ts(tmem_base + AM::hh(ks) * BN, aH + ks * 8, dBh, idesc, !AM::first_big(i, ks));
            umma_tf32_ts(tmem_base + AM::hl(ks) * BN, aH + ks * 8, dBl, idesc, !AM::first_hl(i, ks));
          }
          umma_commit(&a_empty[sa]);     // both stages are free once these MMAs have read them
          umma_commit(&b_empty[sb]);
        }
        __syncwarp();
        if (lane == 0 && i < 8) EQH_STAMP(56 + i);
      }
      if (lane == 0 && nkb > 0) EQH_STAMP(5);
      if (elect_one()) umma_commit(tmem_full);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A rows: shared memory -> registers -> TMEM (hi | lo) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                                   // tile row = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::kAcol;
    for (int i = 0; i < nkb; ++i) {
      const int s = i % SB, sa = i % SA;
      mbar_wait(&b_raw[s], (i / SB) & 1);
      if (r == 0 && i < 8) EQH_STAMP(32 + i);
      uint32_t v[32];
      if constexpr (A_KMAJOR) {
        const uint8_t* row = stageA(s) + r * 128;                  // 128B swizzle: chunk c sits at c ^ (r & 7)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 t4 = *reinterpret_cast<const uint4*>(row + ((c ^ (r & 7)) << 4));
          v[4 * c] = t4.x; v[4 * c + 1] = t4.y; v[4 * c + 2] = t4.z; v[4 * c + 3] = t4.w;
        }
      } else {
        const uint32_t* col = reinterpret_cast<const uint32_t*>(stageA(s)) + r;     // [32 k][128 m]: lanes are consecutive words
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = col[c * 128];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_sfree[s]);                     // the tile is in registers: TMA may refill the slot
      mbar_wait(&a_empty[sa], ((i / SA) & 1) ^ 1);
      if (r == 0 && i < 8) EQH_STAMP(64 + i);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_st32(lane_addr + sa * 64, v);                           // hi: the tensor core ignores the low 13 bits
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x = __uint_as_float(v[c]);
        v[c] = __float_as_uint(x - __uint_as_float(v[c] & 0xffffe000u));
      }
      tmem_st32(lane_addr + sa * 64 + 32, v);                      // lo
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[sa]);
      if (r == 0 && i < 8) EQH_STAMP(40 + i);
    }
    // ===================== epilogue (same warps) =====================
    epilogue_tile<BN, AM, false>(smem, sbias, tmem_base, tmem_full, warp, lane, nkb, p, m0, n0, -1, 0, nullptr, bz);
  } else {
    // ===================== B split (hi / lo): warps 2, 3, 8, 9 =====================
    const int t = ((warp < 4 ? warp - 2 : warp - 6) << 5) + lane;      // 0..127
    for (int i = 0; i < nkb; ++i) {
      const int s = i % SB;
      mbar_wait(&b_raw[s], (i / SB) & 1);
      if (t == 0 && i == 0) EQH_STAMP(2);
      if (t == 0 && i < 8) EQH_STAMP(16 + i);
      const float4* bh = reinterpret_cast<const float4*>(stageB(s));
      float4* bl = reinterpret_cast<float4*>(stageBlo(s));
      constexpr int NB = Cfg::kBBytes / 16 / 128;     // 8 (BN=128), 4, 2
      float4 v[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) v[j] = bh[j * 128 + t];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        float4 l;
        l.x = tf32_lo(v[j].x); l.y = tf32_lo(v[j].y); l.z = tf32_lo(v[j].z); l.w = tf32_lo(v[j].w);
        bl[j * 128 + t] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[s]);
      if (t == 0 && i < 8) EQH_STAMP(24 + i);
    }
  }
  if (threadIdx.x == 128) EQH_STAMP(8);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();                            // the tile is staged in shared memory, every TMEM read has completed
  if (warp == 2) tmem_dealloc(tmem_base, 512);
  epilogue_store_all<BN>(smem, p, m0, n0, bz);     // all ten warps store
  if (p.ws != nullptr) {
    __threadfence();                          // this split's partial tile is visible device-wide before the arrival
    __syncthreads();
    splitk_finish<BN>(p, m0, n0, &s_last, tile, nz);
  }
}

template <int BN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(kThreads, 1) k_gemm_ts(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB, const Params p) {
  gemm_ts_tile<BN, A_KMAJOR, B_KMAJOR>(&tmA, &tmB, p, (int)blockIdx.x, (int)blockIdx.y, (int)blockIdx.z, (int)gridDim.z,
                                       (int)(blockIdx.y * gridDim.x + blockIdx.x));
}

// Several independent GEMMs of the same operand majors in ONE launch (the weight gradients of a layer: 7 problems of 16
// tiles each would otherwise be 7 launches of a few CTA rounds each, every one with its own prologue / split-K finish
// tail).  CTA i of the linear grid belongs to problem g with cta_start[g] <= i < cta_start[g+1]; inside the problem the
// index decodes to (split, tile_m, tile_n).
template <int BN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(kThreads, 1) k_gemm_ts_grouped(const __grid_constant__ GroupParams g) {
  int i = 0;
  while (i + 1 < g.n && (int)blockIdx.x >= g.cta_start[i + 1]) ++i;
  const int local = (int)blockIdx.x - g.cta_start[i];
  const int bz = local / g.tiles[i], t = local - bz * g.tiles[i];
  const int by = t / g.tiles_n[i], bx = t - by * g.tiles_n[i];
  gemm_ts_tile<BN, A_KMAJOR, B_KMAJOR>(&g.tmA[i], &g.tmB[i], g.p[i], bx, by, bz, g.p[i].splits, t);
}

template <int BN, bool AK, bool BK_>
static int launch_one(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = TsCfg<BN>::kSmem;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(k_gemm_ts<BN, AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  launch_kernel(k_gemm_ts<BN, AK, BK_>, dim3(grid), dim3(kThreads), smem, st, a, b, p);
  EQH_LAUNCH_RC();
}

template <int BN, bool AK, bool BK_>
static int launch_group_one(const GroupParams& g, int ctas, cudaStream_t st) {
  constexpr size_t smem = TsCfg<BN>::kSmem;
  cudaError_t e = cudaFuncSetAttribute(k_gemm_ts_grouped<BN, AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  launch_kernel(k_gemm_ts_grouped<BN, AK, BK_>, dim3((unsigned)ctas), dim3(kThreads), smem, st, g);
  EQH_LAUNCH_RC();
}

int launch_ts_grouped(int BN, bool a_kmajor, bool b_kmajor, const GroupParams& g, int ctas, cudaStream_t st) {
#define EQH_TSG_MAJORS(BNV)                                                                  \
  do {                                                                                       \
    if (a_kmajor && b_kmajor) return launch_group_one<BNV, true, true>(g, ctas, st);         \
    if (a_kmajor && !b_kmajor) return launch_group_one<BNV, true, false>(g, ctas, st);       \
    if (!a_kmajor && !b_kmajor) return launch_group_one<BNV, false, false>(g, ctas, st);     \
    return launch_group_one<BNV, false, true>(g, ctas, st);                                  \
  } while (0)
  if (BN == 32) EQH_TSG_MAJORS(32);
  if (BN == 64) EQH_TSG_MAJORS(64);
  EQH_TSG_MAJORS(128);
#undef EQH_TSG_MAJORS
}

int launch_ts(int BN, bool a_kmajor, bool b_kmajor, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p,
              dim3 grid, cudaStream_t st) {
#define EQH_TS_MAJORS(BNV)                                                                \
  do {                                                                                    \
    if (a_kmajor && b_kmajor) return launch_one<BNV, true, true>(tmA, tmB, p, grid, st);   \
    if (a_kmajor && !b_kmajor) return launch_one<BNV, true, false>(tmA, tmB, p, grid, st); \
    if (!a_kmajor && !b_kmajor) return launch_one<BNV, false, false>(tmA, tmB, p, grid, st); \
    return launch_one<BNV, false, true>(tmA, tmB, p, grid, st);                        \
  } while (0)
  if (BN == 32) EQH_TS_MAJORS(32);
  if (BN == 64) EQH_TS_MAJORS(64);
  EQH_TS_MAJORS(128);
#undef EQH_TS_MAJORS
}

}  // namespace gemm
}  // namespace eqh

#ifdef EQH_TRACE
extern "C" EQH_API int eqh_debug_gemm_ts_trace(unsigned long long* host16) {
  return (int)cudaMemcpyFromSymbol(host16, eqh::gemm::g_trace, sizeof(unsigned long long) * 80);
}
#endif
