// 3xTF32 row GEMM, second kernel: the A operand never touches shared memory.
//
// gemm_tf32x3.cu is bound by shared-memory bandwidth, not by the tensor core: per 32-deep K block it moves the
// A tile through shared memory five times (TMA write, split read, lo write, and three tcgen05.mma operand reads).
// Here each of the 128 threads of warps 4-7 owns one row of the A tile (= one TMEM lane): it loads its 32 K
// values from global memory into registers, splits them hi / lo and writes both halves into tensor memory
// (tcgen05.st); the MMAs take A from TMEM ("TS" form) and only the B tile (the weights: BN x 32) goes through
// shared memory.  Shared-memory traffic per K block drops from 5*16 KB + 5*BN*128 B to 5*BN*128 B, which puts
// the main loop at the tensor-core rate (12 MMAs) for BN >= 64.
//
//   warp 0 (1 thread)  TMA producer of B tiles (128B swizzle) into a ring
//   warps 2,3,8,9      split the landed B tile into hi (in place) / lo, fence.proxy.async, signal
//   warps 4-7          A rows: global -> registers (4-deep ring, 3 K blocks in flight) -> TMEM hi | lo; then epilogue
//   warp 1 (1 thread)  per K block 4 k-steps x 3 tcgen05.mma (A from TMEM), commits free the A and B stages
// TMEM columns: [0, 3*BN) accumulators (as in gemm_tf32x3.cu), then the A ring: 64 columns (hi | lo) per stage.
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
  static constexpr int SB = BN == 128 ? 6 : 8;                  // B stages in shared memory
  static constexpr uint32_t kBBytes = BN * BK * 4;
  static constexpr uint32_t kStageBytes = 2 * kBBytes;          // B raw/hi, B lo
  static constexpr uint32_t kAcol = 3 * BN;                     // first TMEM column of the A ring
  static constexpr size_t kRing = (size_t)SB * kStageBytes;
  static constexpr size_t kSmem = kRing + 512 + BN * 4;         // ring + barriers + bias tile
  static_assert(kAcol + SA * 64 <= 512, "TMEM budget");
  static_assert(kRing >= (size_t)BM * (BN + 4) * 4, "epilogue staging must fit in the B ring");
};

constexpr int LA = 4;        // register ring of the A rows: LA-1 K blocks of loads in flight per thread

template <int BN, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(kThreads, 1) k_gemm_ts(const float* __restrict__ A, int64_t lda,
                                                          const __grid_constant__ CUtensorMap tmB, const Params p) {
  using Cfg = TsCfg<BN>;
  constexpr int SA = Cfg::SA, SB = Cfg::SB;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* b_raw = reinterpret_cast<uint64_t*>(smem + Cfg::kRing);     // TMA landed
  uint64_t* b_full = b_raw + SB;                                        // hi/lo ready for the tensor core
  uint64_t* b_empty = b_full + SB;
  uint64_t* a_full = b_empty + SB;
  uint64_t* a_empty = a_full + SA;
  uint64_t* tmem_full = a_empty + SA;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::kRing + 512);
  static_assert((3 * SB + 2 * SA + 1) * 8 + 4 <= 512, "barrier block");
  __shared__ int s_last;

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  if (threadIdx.x == 0) EQH_STAMP(0);
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(kb_end - kb_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SB; ++s) {
      mbar_init(&b_raw[s], 1);
      mbar_init(&b_full[s], 4);          // one elected arrival per B-split warp
      mbar_init(&b_empty[s], 1);
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

  auto stageB = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes; };
  auto stageBlo = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes + Cfg::kBBytes; };

  if (warp == 0) {
    // ===================== TMA producer (B) =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % SB;
        const uint32_t ph = (i / SB) & 1;
        mbar_wait(&b_empty[s], ph ^ 1);
        mbar_expect_tx(&b_raw[s], Cfg::kBBytes);
        const int k0 = (kb_begin + i) * BK;
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
    // ===================== MMA issuer =====================
    if (lane == 0) {
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
        if (i == 0) EQH_STAMP(3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t bH = smem_u32(stageB(sb)), bL = smem_u32(stageBlo(sb));
        const uint32_t aH = tmem_base + Cfg::kAcol + sa * 64, aL = aH + 32;
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t dBh = make_desc(bH + ks * kstepB, lboB, sboB, ltB), dBl = make_desc(bL + ks * kstepB, lboB, sboB, ltB);
          const uint32_t step = (uint32_t)(i * (BK / 8) + ks);
          umma_tf32_ts(tmem_base + 2 * BN, aL + ks * 8, dBh, idesc, step != 0);
          umma_tf32_ts(tmem_base + 2 * BN, aH + ks * 8, dBl, idesc, 1);
          umma_tf32_ts(tmem_base + (step & 1) * BN, aH + ks * 8, dBh, idesc, step >= 2);
        }
        umma_commit(&a_empty[sa]);     // both stages are free once these MMAs have read them
        umma_commit(&b_empty[sb]);
      }
      if (nkb > 0) EQH_STAMP(5);
      umma_commit(tmem_full);
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== A rows: global -> registers -> TMEM (hi | lo) =====================
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    const bool row_ok = m < p.M;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::kAcol;
    uint32_t a[LA][32];
    auto load = [&](int i, uint32_t* dst) {
      const int k0 = (kb_begin + i) * BK;
      if constexpr (A_KMAJOR) {
        const float* src = A + (int64_t)m * lda + k0;            // the thread's own row: 128 contiguous bytes
        if (row_ok && k0 + BK <= p.K) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c);
            dst[4 * c] = __float_as_uint(v.x); dst[4 * c + 1] = __float_as_uint(v.y);
            dst[4 * c + 2] = __float_as_uint(v.z); dst[4 * c + 3] = __float_as_uint(v.w);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) dst[c] = (row_ok && k0 + c < p.K) ? __float_as_uint(__ldg(src + c)) : 0u;
        }
      } else {
        const float* src = A + (int64_t)k0 * lda + m;            // stored [K, lda]: a warp reads 128 contiguous bytes per k
#pragma unroll
        for (int c = 0; c < 32; ++c) dst[c] = (row_ok && k0 + c < p.K) ? __float_as_uint(__ldg(src + (int64_t)c * lda)) : 0u;
      }
    };
    auto push = [&](int i, uint32_t* v) {
      const int sa = i % SA;
      mbar_wait(&a_empty[sa], ((i / SA) & 1) ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tmem_st32(lane_addr + sa * 64, v);                         // hi: the tensor core ignores the low 13 bits
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float x = __uint_as_float(v[c]);
        v[c] = __float_as_uint(x - __uint_as_float(v[c] & 0xffffe000u));
      }
      tmem_st32(lane_addr + sa * 64 + 32, v);                    // lo
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[sa]);
    };
#pragma unroll
    for (int u = 0; u < LA - 1; ++u)
      if (u < nkb) load(u, a[u]);
#pragma unroll 1
    for (int i0 = 0; i0 < nkb; i0 += LA) {
#pragma unroll
      for (int u = 0; u < LA; ++u) {
        const int i = i0 + u;
        if (i < nkb) {
          if (i + LA - 1 < nkb) load(i + LA - 1, a[(u + LA - 1) % LA]);
          push(i, a[u]);
        }
      }
    }
    // ===================== epilogue (same warps) =====================
    epilogue_tile<BN>(smem, sbias, tmem_base, tmem_full, warp, lane, nkb, p, m0, n0);
  } else {
    // ===================== B split (hi / lo): warps 2, 3, 8, 9 =====================
    const int t = ((warp < 4 ? warp - 2 : warp - 6) << 5) + lane;      // 0..127
    for (int i = 0; i < nkb; ++i) {
      const int s = i % SB;
      mbar_wait(&b_raw[s], (i / SB) & 1);
      if (t == 0 && i == 0) EQH_STAMP(2);
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
    }
  }
  if (threadIdx.x == 128) EQH_STAMP(8);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (p.ws != nullptr) __threadfence();       // this split's partial tile is visible device-wide before the arrival
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
  if (p.ws != nullptr) splitk_finish<BN>(p, m0, n0, &s_last);
}

template <int BN, bool AK, bool BK_>
static int launch_one(const float* A, int64_t lda, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = TsCfg<BN>::kSmem;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(k_gemm_ts<BN, AK, BK_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  launch_kernel(k_gemm_ts<BN, AK, BK_>, dim3(grid), dim3(kThreads), smem, st, A, lda, b, p);
  EQH_LAUNCH_RC();
}

int launch_ts(int BN, bool a_kmajor, bool b_kmajor, const float* A, int64_t lda, const CUtensorMap& tmB, const Params& p,
              dim3 grid, cudaStream_t st) {
#define EQH_TS_MAJORS(BNV)                                                                \
  do {                                                                                    \
    if (a_kmajor && b_kmajor) return launch_one<BNV, true, true>(A, lda, tmB, p, grid, st);   \
    if (a_kmajor && !b_kmajor) return launch_one<BNV, true, false>(A, lda, tmB, p, grid, st); \
    if (!a_kmajor && !b_kmajor) return launch_one<BNV, false, false>(A, lda, tmB, p, grid, st); \
    return launch_one<BNV, false, true>(A, lda, tmB, p, grid, st);                        \
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
  return (int)cudaMemcpyFromSymbol(host16, eqh::gemm::g_trace, sizeof(unsigned long long) * 16);
}
#endif
