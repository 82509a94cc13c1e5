// Row GEMMs of the dense layers next to the hot path on the 5th-generation tensor cores:
//     D[M,N] = act( A[M,K] . B[N,K]^T + bias[N] ) (+ residual[M,N])
// fp32 in, fp32 out, fp32-faithful: every operand is split a = a_hi + a_lo (a_hi = a with the low
// 13 mantissa bits cleared, a_lo = a - a_hi, exact) and the product is evaluated as
//     a_lo.b_hi + a_hi.b_lo + a_hi.b_hi          ("3xTF32", error ~2^-21 relative)
// with tcgen05.mma.kind::tf32 accumulating in TMEM.  The reference's tolerance for these layers is
// 1e-5 relative (BASELINE.json), which rules plain TF32 (2^-11) out; cuBLAS SIMT fp32 is what this
// kernel replaces (35 % of the step in profiles/r01_launches_eager_qm9b128.md).
//
// One CTA computes one 128 x BN output tile for one K-split:
//   warp 0 (1 thread)  TMA producer: cp.async.bulk.tensor tiles (128B swizzle) into a 3-stage ring
//   warps 4-7          split the landed fp32 tiles into hi (in place) / lo (second buffer) in shared
//                      memory, fence.proxy.async, signal the MMA warp; afterwards they are the epilogue
//   warp 1 (1 thread)  issues 4 k-steps x 3 tcgen05.mma per 32-wide K block, commits to the ring
//   epilogue           tcgen05.ld 32x32b -> bias / activation / residual -> padded smem -> coalesced
//                      128-bit global stores (bounds-checked; TMA zero-fills ragged M/N/K edges)
// Operands may be K-major (row-major [rows, K]) or MN-major ([K, rows], the natural layout of the
// backward GEMMs dX = dY.W and dW = dY^T.X), selected by the UMMA descriptor major bits, so no
// transposed copies are ever made.  Split-K (weight gradients: K = number of rows, tiny M x N) writes
// per-split partials; the CTA arriving last at a tile's counter adds them in split order (deterministic)
// and applies the output transform -- one launch.  The tile width (32/64/128) and the split count are
// chosen per problem so that small GEMMs still occupy every SM (plan()).
#include <algorithm>
#include <mutex>

#include "gemm_common.cuh"

namespace eqh {
namespace gemm {

// kWriteHi: store the masked hi part back (true) or let the tensor core drop the low 13 bits of the raw fp32
// tile itself (false; valid if the hardware truncates tf32 inputs -- checked by tests/test_gpu_gemm.py).
template <int BN, bool A_KMAJOR, bool B_KMAJOR, bool kWriteHi, int STAGES, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_gemm(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmB, const Params p) {
  constexpr uint32_t kABytes = kTileBytes;            // 128 x 32
  constexpr uint32_t kBBytes = BN * BK * 4;
  constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // A raw/hi, A lo, B raw/hi, B lo
  extern __shared__ __align__(1024) uint8_t smem[];     // operand ring (1024-byte aligned: 128B-swizzle atoms), then barriers
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * kStageBytes);
  uint64_t* full_split = full_raw + STAGES;
  uint64_t* empty = full_split + STAGES;
  uint64_t* tmem_full_p = empty + STAGES;
  uint32_t* tmem_slot_p = reinterpret_cast<uint32_t*>(tmem_full_p + 1);
  __shared__ int s_last;
#define tmem_full (*tmem_full_p)
#define tmem_slot (*tmem_slot_p)

  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  if (threadIdx.x == 0) EQH_STAMP(0);
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(kb_end - kb_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_split[s], kSplitThreads / 32);     // one elected arrival per split warp
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Three fp32 accumulators of BN columns each.  The tensor core adds into its accumulator with
  // truncation, so the error grows with the number of accumulations: the large hi.hi terms alternate
  // between two accumulators and the 2^-11-scaled lo terms get their own (summed in the epilogue).
  using AM = AccMap<BN>;
  constexpr uint32_t kTmemCols = AM::NA * BN <= 128 ? 128 : (AM::NA * BN <= 256 ? 256 : 512);
  static_assert(AM::NA * BN <= 512 && kTmemCols * MINB <= 512, "co-resident CTAs must fit in TMEM");
  if (warp == 2) tmem_alloc(&tmem_slot, kTmemCols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  pdl_wait();        // everything above (barriers, tensor memory) overlapped the previous kernel's tail
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) EQH_STAMP(1);

  auto stageA = [&](int s) { return smem + (size_t)s * kStageBytes; };
  auto stageAlo = [&](int s) { return smem + (size_t)s * kStageBytes + kABytes; };
  auto stageB = [&](int s) { return smem + (size_t)s * kStageBytes + 2 * kABytes; };
  auto stageBlo = [&](int s) { return smem + (size_t)s * kStageBytes + 2 * kABytes + kBBytes; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full_raw[s], kABytes + kBBytes);
        const int k0 = (kb_begin + i) * BK;
        if constexpr (A_KMAJOR) {
          tma_load_2d(stageA(s), &tmA, &full_raw[s], k0, m0);                 // box {32 K, 128 M}
        } else {
#pragma unroll
          for (int a = 0; a < BM / 32; ++a)                                   // 4 boxes {32 M, 32 K}
            tma_load_2d(stageA(s) + a * (BK * 128), &tmA, &full_raw[s], m0 + a * 32, k0);
        }
        if constexpr (B_KMAJOR) {
          tma_load_2d(stageB(s), &tmB, &full_raw[s], k0, n0);                 // box {32 K, BN}
        } else {
#pragma unroll
          for (int a = 0; a < BN / 32; ++a)
            tma_load_2d(stageB(s) + a * (BK * 128), &tmB, &full_raw[s], n0 + a * 32, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((A_KMAJOR ? 0u : 1u) << 15) |
                                 ((B_KMAJOR ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t lboA = A_KMAJOR ? 0u : BK * 128u, lboB = B_KMAJOR ? 0u : BK * 128u;
      constexpr uint32_t kstepA = A_KMAJOR ? 32u : 1024u, kstepB = B_KMAJOR ? 32u : 1024u;  // bytes per UMMA_K = 8
      constexpr uint32_t sboA = A_KMAJOR ? 1024u : 512u, sboB = B_KMAJOR ? 1024u : 512u;
      constexpr uint32_t ltA = A_KMAJOR ? 2u : 1u, ltB = B_KMAJOR ? 2u : 1u;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_split[s], ph);
        if (lane == 0 && i == 0) EQH_STAMP(3);
        if (lane == 0 && i == nkb - 1) EQH_STAMP(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aH = smem_u32(stageA(s)), aL = smem_u32(stageAlo(s));
        const uint32_t bH = smem_u32(stageB(s)), bL = smem_u32(stageBlo(s));
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / 8; ++ks) {
            const uint64_t dAh = make_desc(aH + ks * kstepA, lboA, sboA, ltA), dAl = make_desc(aL + ks * kstepA, lboA, sboA, ltA);
            const uint64_t dBh = make_desc(bH + ks * kstepB, lboB, sboB, ltB), dBl = make_desc(bL + ks * kstepB, lboB, sboB, ltB);
            umma_tf32(tmem_base + AM::lh(ks) * BN, dAl, dBh, idesc, !AM::first_lh(i, ks));
            umma_tf32(tmem_base + AM::hh(ks) * BN, dAh, dBh, idesc, !AM::first_big(i, ks));
            umma_tf32(tmem_base + AM::hl(ks) * BN, dAh, dBl, idesc, !AM::first_hl(i, ks));
          }
          umma_commit(&empty[s]);   // frees the stage when these MMAs have read it
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&tmem_full);
    }
  } else {
    // ===================== operand split (hi / lo): warps 2-9 =====================
    const int t = threadIdx.x - 64;    // 0..255
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_raw[s], ph);
      if (t == 0 && i == 0) EQH_STAMP(2);
      if (t == 0 && i == nkb - 1) EQH_STAMP(4);
      {
        float4* ah = reinterpret_cast<float4*>(stageA(s));
        float4* al = reinterpret_cast<float4*>(stageAlo(s));
        constexpr int NA = kABytes / 16 / kSplitThreads;   // 4
        float4 v[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) v[j] = ah[j * kSplitThreads + t];      // all loads first (independent)
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u); l.x = v[j].x - h.x;
          h.y = __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u); l.y = v[j].y - h.y;
          h.z = __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u); l.z = v[j].z - h.z;
          h.w = __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u); l.w = v[j].w - h.w;
          if (kWriteHi) ah[j * kSplitThreads + t] = h;
          al[j * kSplitThreads + t] = l;
        }
      }
      {
        float4* bh = reinterpret_cast<float4*>(stageB(s));
        float4* bl = reinterpret_cast<float4*>(stageBlo(s));
        constexpr int NB = kBBytes / 16 / kSplitThreads;   // 4 (BN=128), 2 (BN=64), 1 (BN=32)
        static_assert(NB >= 1, "BN >= 32");
        float4 v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) v[j] = bh[j * kSplitThreads + t];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u); l.x = v[j].x - h.x;
          h.y = __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u); l.y = v[j].y - h.y;
          h.z = __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u); l.z = v[j].z - h.z;
          h.w = __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u); l.w = v[j].w - h.w;
          if (kWriteHi) bh[j * kSplitThreads + t] = h;
          bl[j * kSplitThreads + t] = l;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_split[s]);
    }
  }
  if (warp >= 4 && warp < 8) {
    float* sbias = reinterpret_cast<float*>(smem + (size_t)STAGES * kStageBytes + 128);   // [BN] after the barriers
    epilogue_tile<BN, AM>(smem, sbias, tmem_base, &tmem_full, warp, lane, nkb, p, m0, n0);
  }
  if (threadIdx.x == 128) EQH_STAMP(8);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (p.ws != nullptr) __threadfence();       // this split's partial tile is visible device-wide before the arrival
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (p.ws != nullptr) splitk_finish<BN>(p, m0, n0, &s_last);
#undef tmem_full
#undef tmem_slot
}

// ---- host side ---------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = []() -> EncodeFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  return fn;
}

// 2-D fp32 tensor map over a row-major [rows, cols] matrix with row stride ld (floats); box = {box_cols, box_rows}
static int make_map(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_cols,
                    uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeFn fn = encode_fn();
  if (!fn) return EQH_ERR_UNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? EQH_OK : EQH_ERR_ARG;
}

template <int BN, bool AK, bool BK_, bool WH, int STAGES, int MINB>
static int launch(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (2 * kTileBytes + 2 * BN * BK * 4) + 128 + 2 * BN * 4;   // ring + barriers + bias / bias2 tiles
  static_assert((size_t)STAGES * (2 * kTileBytes + 2 * BN * BK * 4) >= (size_t)BM * (BN + 4) * 4,
                "epilogue staging must fit in the operand ring");
  static_assert(smem * MINB <= 227 * 1024, "shared memory budget");
  cudaError_t e = cudaFuncSetAttribute(k_gemm<BN, AK, BK_, WH, STAGES, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return (int)e;
  launch_kernel(k_gemm<BN, AK, BK_, WH, STAGES, MINB>, dim3(grid), dim3(kThreads), smem, st, a, b, p);
  EQH_LAUNCH_RC();
}

// gemm_ts.cu: the kernel that feeds A through registers into tensor memory (default)
int launch_ts(int BN, bool a_kmajor, bool b_kmajor, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p,
              dim3 grid, cudaStream_t st);

int launch_ts_grouped(int BN, bool a_kmajor, bool b_kmajor, const GroupParams& g, int ctas, cudaStream_t st);

// gemm_ps.cu: persistent 128 x 64 variant with double-buffered accumulators (multi-wave problems, no split-K)
int launch_ps(bool a_kmajor, bool b_kmajor, const CUtensorMap& tmA, const CUtensorMap& tmB, const Params& p, cudaStream_t st);

// Tile width and split count.  Cost model (clocks per CTA): a 32-deep K block costs ~300 (operand split of the
// 128-row A tile) + 6.4*BN (12 MMAs); the split-K finish costs ~4*BN per split in the last CTA.
static void plan(int64_t M, int64_t N, int64_t K, int splits_req, int bn_req, int* bn_out, int* splits_out,
                 bool background = false) {
  const int64_t tiles_m = (M + BM - 1) / BM, kb = (K + BK - 1) / BK;
  double best = 1e300;
  int best_bn = 128, best_s = 1;
  if (background && splits_req == 0) {
    // A GEMM issued on the side stream (weight gradients) is not on the critical path: its latency is hidden, its SM
    // time is not -- 144 CTAs for 12 us keep every SM away from the dX chain.  Minimise CTAs x duration instead:
    // ~24 K blocks per CTA (~12 us), the widest tile.
    for (int bn = 128; bn >= 32; bn >>= 1) {
      if (bn_req && bn != bn_req) continue;
      if (!bn_req && bn > 32 && bn / 2 >= N) continue;
      const int64_t t = tiles_m * ((N + bn - 1) / bn);
      int64_t s = (kb + 23) / 24;
      s = s < 1 ? 1 : (s > 32 ? 32 : s);
      const int64_t kbps = (kb + s - 1) / s, s_eff = (kb + kbps - 1) / kbps;
      const double cost = (double)(t * s_eff) * ((double)kbps * (300.0 + 6.4 * bn) + 1500.0 + (s_eff > 1 ? 4.0 * bn * s_eff : 0.0));
      if (cost < best - 1e-9) { best = cost; best_bn = bn; best_s = (int)s_eff; }
    }
    *bn_out = best_bn;
    *splits_out = best_s;
    return;
  }
  for (int bn = 128; bn >= 32; bn >>= 1) {
    if (bn_req && bn != bn_req) continue;
    if (!bn_req && bn > 32 && bn / 2 >= N) continue;               // do not pad N by more than 2x
    const int64_t t = tiles_m * ((N + bn - 1) / bn);
    const int s_max = splits_req == 0 ? (int)(kb / 2 < 1 ? 1 : (kb / 2 > 32 ? 32 : kb / 2)) : splits_req;
    for (int s = splits_req == 0 ? 1 : splits_req; s <= s_max; ++s) {
      const int64_t kbps = (kb + s - 1) / s;
      const int64_t s_eff = (kb + kbps - 1) / kbps;
      const int64_t waves = (t * s_eff + kSMs - 1) / kSMs;
      const double cost = (double)waves * ((double)kbps * (300.0 + 6.4 * bn) + 1500.0) + (s_eff > 1 ? 4.0 * bn * s_eff : 0.0);
      if (cost < best - 1e-9) { best = cost; best_bn = bn; best_s = (int)s_eff; }
    }
  }
  *bn_out = best_bn;
  *splits_out = best_s;
}

}  // namespace gemm
}  // namespace eqh

extern "C" {

#ifdef EQH_TRACE
EQH_API int eqh_debug_gemm_trace(unsigned long long* host16) {
  return (int)cudaMemcpyFromSymbol(host16, eqh::gemm::g_trace, sizeof(unsigned long long) * 16);
}
#endif

int eqh_gemm_plan(int64_t M, int64_t N, int64_t K, int splits, int flags, int* bn, int* splits_out, int64_t* tiles) {
  using namespace eqh;
  EQH_CHECK_ARG(M > 0 && N > 0 && K > 0 && splits >= 0 && bn && splits_out && tiles);
  const int code = (flags >> 9) & 3;
  gemm::plan(M, N, K, splits, (flags & 0x100) ? 128 : (code ? (16 << code) : 0), bn, splits_out, (flags & 0x8000) != 0);
  *tiles = ((M + gemm::BM - 1) / gemm::BM) * ((N + *bn - 1) / *bn);
  return EQH_OK;
}

// ---- grouped launch: n independent problems with the same operand majors ------------------------------------------
// Plan: one tile width for the group and, per problem, a split count such that every CTA owns about q K blocks; q and
// the width minimise rounds x per-CTA cost (the cost model of gemm::plan), rounds = ceil(total CTAs / SMs).
// max_ctas > 0: background plan -- the group must not occupy more than max_ctas SMs (it runs next to a latency-critical
// chain of small-grid kernels that needs the rest); the widest tile and the fewest splits that fit.
static void plan_group(const eqh_gemm_problem* pr, int n, int bn_req, int* bn_out, int* splits /*[n]*/, int max_ctas = 0) {
  using namespace eqh;
  using namespace eqh::gemm;
  double best = 1e300;
  int best_bn = 64, best_q = 1 << 30;
  int64_t kb_max = 1;
  for (int i = 0; i < n; ++i) kb_max = std::max<int64_t>(kb_max, (pr[i].K + BK - 1) / BK);
  for (int bn = 128; bn >= 32; bn >>= 1) {
    if (bn_req && bn != bn_req) continue;
    for (int s = 1; s <= 32; ++s) {
      const int64_t q = (kb_max + s - 1) / s;                 // K blocks per CTA of the deepest problem
      if (s > 1 && q < 2) break;
      // the tensor core truncates when it accumulates: bound the accumulation chain of one CTA (as the single-problem
      // planner does with its kb/2..32 splits) -- 48 K blocks keep the result within 5e-6 of fp64 at K = 30000
      if (q > 48 && s < 32 && max_ctas == 0) continue;
      int64_t ctas = 0;
      double finish = 0.0;
      for (int i = 0; i < n; ++i) {
        const int64_t kb = (pr[i].K + BK - 1) / BK;
        int64_t si = (kb + q - 1) / q;
        si = si < 1 ? 1 : (si > 32 ? 32 : si);
        const int64_t kbps = (kb + si - 1) / si, s_eff = (kb + kbps - 1) / kbps;
        ctas += ((pr[i].M + BM - 1) / BM) * ((pr[i].N + bn - 1) / bn) * s_eff;
        if (s_eff > 1) finish = std::max(finish, 4.0 * bn * (double)s_eff);
      }
      if (max_ctas > 0 && ctas > max_ctas && !(bn == 128 && s == 1)) continue;   // (the smallest plan is always allowed)
      const int64_t rounds = max_ctas > 0 ? 1 : (ctas + kSMs - 1) / kSMs;
      const double cost = (double)rounds * ((double)q * (300.0 + 6.4 * bn) + 1500.0) + finish;
      if (cost < best - 1e-9) { best = cost; best_bn = bn; best_q = (int)q; }
    }
  }
  *bn_out = best_bn;
  for (int i = 0; i < n; ++i) {
    const int64_t kb = (pr[i].K + BK - 1) / BK;
    int64_t si = (kb + best_q - 1) / best_q;
    si = si < 1 ? 1 : (si > 32 ? 32 : si);
    const int64_t kbps = (kb + si - 1) / si;
    splits[i] = (int)((kb + kbps - 1) / kbps);
  }
}

static int group_check(const eqh_gemm_problem* pr, int n) {
  if (!pr || n < 1 || n > eqh::gemm::kMaxGroup) return EQH_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    const eqh_gemm_problem& q = pr[i];
    if (!(q.A && q.B && q.D && q.M > 0 && q.N > 0 && q.K > 0 && q.M < (1 << 30) && q.N < (1 << 30) && q.K < (1 << 30)))
      return EQH_ERR_ARG;
    if (!(eqh::aligned16(q.A) && eqh::aligned16(q.B) && q.lda % 4 == 0 && q.ldb % 4 == 0 && q.ldd >= q.N)) return EQH_ERR_ARG;
  }
  return EQH_OK;
}

int eqh_gemm_grouped_plan(const eqh_gemm_problem* problems, int n, int a_kmajor, int b_kmajor, int flags, int* bn,
                          int* splits, int64_t* ws_floats, int64_t* n_counters) {
  (void)a_kmajor; (void)b_kmajor;
  using namespace eqh::gemm;
  int rc = group_check(problems, n);
  if (rc) return rc;
  EQH_CHECK_ARG(bn && splits && ws_floats && n_counters);
  const int code = (flags >> 9) & 3;
  plan_group(problems, n, code ? (16 << code) : 0, bn, splits, (flags >> 16) & 0xfff);
  int64_t ws = 0, cnt = 0;
  for (int i = 0; i < n; ++i) {
    if (splits[i] > 1) {
      ws += (int64_t)splits[i] * problems[i].M * problems[i].N;
      ws = (ws + 3) / 4 * 4;                                   // every problem's partials start 16-byte aligned
    }
    cnt += ((problems[i].M + BM - 1) / BM) * ((problems[i].N + *bn - 1) / *bn);
  }
  *ws_floats = ws;
  *n_counters = cnt;
  return EQH_OK;
}

int eqh_gemm_grouped(const eqh_gemm_problem* problems, int n, int a_kmajor, int b_kmajor, int flags, float* ws,
                     int32_t* counters, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::gemm;
  int rc = group_check(problems, n);
  if (rc) return rc;
  int BN, splits[kMaxGroup];
  const int code = (flags >> 9) & 3;
  plan_group(problems, n, code ? (16 << code) : 0, &BN, splits, (flags >> 16) & 0xfff);
  static GroupParams g;                        // ~3 KB: too large for the stack of every caller; filled and launched under a lock
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  g.n = n;
  int ctas = 0;
  int64_t ws_off = 0, cnt_off = 0;
  for (int i = 0; i < n; ++i) {
    const eqh_gemm_problem& q = problems[i];
    rc = a_kmajor ? make_map(&g.tmA[i], q.A, q.M, q.K, q.lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                  : make_map(&g.tmA[i], q.A, q.K, q.M, q.lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = b_kmajor ? make_map(&g.tmB[i], q.B, q.N, q.K, q.ldb, BK, (uint32_t)BN, CU_TENSOR_MAP_SWIZZLE_128B)
                  : make_map(&g.tmB[i], q.B, q.K, q.N, q.ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
    Params& p = g.p[i];
    const int64_t tm = (q.M + BM - 1) / BM, tn = (q.N + BN - 1) / BN;
    const int kb_total = (int)((q.K + BK - 1) / BK);
    p.D = q.D; p.ldd = q.ldd; p.residual = q.residual; p.out_scale = q.out_scale; p.res_scale = q.res_scale;
    p.M = (int)q.M; p.N = (int)q.N; p.K = (int)q.K;
    p.splits = splits[i];
    p.kb_per_split = (kb_total + splits[i] - 1) / splits[i];
    p.stash = nullptr; p.dsilu_aux = nullptr; p.ld_aux = 0; p.bias2 = nullptr; p.relu_mask = nullptr; p.ld_mask = 0;
    if (splits[i] > 1) {
      EQH_CHECK_ARG(ws && counters && !q.bias);                // the split-K finish applies scales / residual only
      p.ws = ws + ws_off;
      ws_off += (int64_t)splits[i] * q.M * q.N;
      ws_off = (ws_off + 3) / 4 * 4;
      p.bias = nullptr; p.rowscale = nullptr; p.act = 0;
    } else {
      p.ws = nullptr;
      p.bias = q.bias; p.rowscale = q.rowscale; p.act = flags & 0xff;
    }
    p.counters = counters ? counters + cnt_off : nullptr;
    cnt_off += tm * tn;
    g.cta_start[i] = ctas;
    g.tiles[i] = (int)(tm * tn);
    g.tiles_n[i] = (int)tn;
    ctas += (int)(tm * tn) * splits[i];
  }
  g.cta_start[n] = ctas;
  return launch_ts_grouped(BN, a_kmajor != 0, b_kmajor != 0, g, ctas, static_cast<cudaStream_t>(stream));
}

size_t eqh_gemm_workspace_bytes(int64_t M, int64_t N, int splits) {
  return splits > 1 ? sizeof(float) * (size_t)splits * (size_t)M * (size_t)N : 0;
}

int eqh_gemm_tf32x3(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                    int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act_flags,
                    const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                    eqh_stream_t stream) {
  return eqh_gemm_tf32x3_ex(A, a_kmajor, lda, B, b_kmajor, ldb, D, ldd, M, N, K, bias, rowscale, act_flags, residual, out_scale,
                            res_scale, splits, ws, counters, nullptr, nullptr, 0, stream);
}

int eqh_gemm_tf32x3_ex(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                       int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act_flags,
                       const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                       float* stash, const float* dsilu_aux, int64_t ld_aux, eqh_stream_t stream) {
  return eqh_gemm_tf32x3_ex2(A, a_kmajor, lda, B, b_kmajor, ldb, D, ldd, M, N, K, bias, rowscale, act_flags, residual, out_scale,
                             res_scale, splits, ws, counters, stash, dsilu_aux, ld_aux, nullptr, stream);
}

int eqh_gemm_tf32x3_ex2(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                        int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act_flags,
                        const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                        float* stash, const float* dsilu_aux, int64_t ld_aux, const float* bias2, eqh_stream_t stream) {
  return eqh_gemm_tf32x3_ex3(A, a_kmajor, lda, B, b_kmajor, ldb, D, ldd, M, N, K, bias, rowscale, act_flags, residual, out_scale,
                             res_scale, splits, ws, counters, stash, dsilu_aux, ld_aux, bias2, nullptr, 0, stream);
}

int eqh_gemm_tf32x3_ex3(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                        int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act_flags,
                        const float* residual, float out_scale, float res_scale, int splits, float* ws, int32_t* counters,
                        float* stash, const float* dsilu_aux, int64_t ld_aux, const float* bias2, const float* relu_mask,
                        int64_t ld_mask, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::gemm;
  EQH_CHECK_ARG(splits == 1 || (!stash && !dsilu_aux && !bias2 && !relu_mask));
  EQH_CHECK_ARG(!dsilu_aux || ld_aux >= N);
  EQH_CHECK_ARG(!relu_mask || ld_mask >= N);
  const int act = act_flags & 0xff;
  EQH_CHECK_ARG(A && B && D && M > 0 && N > 0 && K > 0 && splits >= 1 && act >= 0 && act <= 2);
  EQH_CHECK_ARG(M < (1 << 30) && N < (1 << 30) && K < (1 << 30));
  EQH_CHECK_ARG(aligned16(A) && aligned16(B) && lda % 4 == 0 && ldb % 4 == 0);   // TMA: 16-byte rows
  EQH_CHECK_ARG(splits == 1 || (ws && counters && !bias && act == 0));
  EQH_CHECK_ARG(ldd >= N);
  const int bn_code = (act_flags >> 9) & 3;             // 0 auto, 1..3 = 32 / 64 / 128 (tuning / tests)
  const bool write_hi = (act_flags & 0x100) != 0;       // debug/verification switch, see kWriteHi
  const bool smem_a = (act_flags & 0x1000) != 0 || write_hi;   // first kernel: both operands staged and split in shared memory
  int BN, s_plan;
  if ((act_flags & 0x8000) != 0) {      // background plan: the caller passes the split count eqh_gemm_plan returned
    plan(M, N, K, 0, write_hi ? 128 : (bn_code ? (16 << bn_code) : 0), &BN, &s_plan, true);
    EQH_CHECK_ARG(s_plan == splits);
  } else {
    plan(M, N, K, splits, write_hi ? 128 : (bn_code ? (16 << bn_code) : 0), &BN, &s_plan);
  }
  // multi-wave problems without split-K: the persistent kernel (128 x 64 tiles, epilogue overlapped with the next tile)
  const int64_t tiles64 = ((M + BM - 1) / BM) * ((N + 63) / 64);
  const bool persistent = !smem_a && splits == 1 && bn_code == 0 && (act_flags & 0x4000) == 0 && tiles64 > 2 * kSMs;
  if (persistent) BN = 64;
  const int kb_total = (int)((K + BK - 1) / BK);
  const int kbps = (kb_total + splits - 1) / splits;
  EQH_CHECK_ARG((kb_total + kbps - 1) / kbps == splits);          // every split owns at least one K block
  CUtensorMap tmA, tmB;
  int rc;
  // A: logical [M,K].  K-major: stored [M, lda] -> box {32 K, 128 M}.  MN-major: stored [K, lda] -> box {32 M, 32 K}.
  // MN-major tf32 operands must use the 128B swizzle with 32-byte atoms (UMMA layout SWIZZLE_128B_BASE32B).
  // (the TMEM-A kernel reads an MN-major A tile row by row itself: one plain box {128 M, 32 K})
  rc = a_kmajor ? make_map(&tmA, A, M, K, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                : (smem_a ? make_map(&tmA, A, K, M, lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                          : make_map(&tmA, A, K, M, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE));
  if (rc) return rc;
  rc = b_kmajor ? make_map(&tmB, B, N, K, ldb, BK, (uint32_t)BN, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map(&tmB, B, K, N, ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  Params p;
  p.D = D; p.bias = bias; p.rowscale = rowscale; p.residual = residual; p.ws = splits > 1 ? ws : nullptr; p.ldd = ldd;
  p.counters = counters; p.splits = splits;
  p.out_scale = out_scale; p.res_scale = res_scale;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.act = act;
  p.kb_per_split = kbps;
  p.stash = stash; p.dsilu_aux = dsilu_aux; p.ld_aux = ld_aux; p.bias2 = bias2; p.relu_mask = relu_mask; p.ld_mask = ld_mask;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (persistent) return launch_ps(a_kmajor != 0, b_kmajor != 0, tmA, tmB, p, st);
  if (!smem_a) return launch_ts(BN, a_kmajor != 0, b_kmajor != 0, tmA, tmB, p, grid, st);
#define EQH_GEMM_MAJORS(BNV, WHV, STG, MB)                                                            \
  do {                                                                                                \
    if (a_kmajor && b_kmajor) return launch<BNV, true, true, WHV, STG, MB>(tmA, tmB, p, grid, st);    \
    if (a_kmajor && !b_kmajor) return launch<BNV, true, false, WHV, STG, MB>(tmA, tmB, p, grid, st);  \
    if (!a_kmajor && !b_kmajor) return launch<BNV, false, false, WHV, STG, MB>(tmA, tmB, p, grid, st);\
    return launch<BNV, false, true, WHV, STG, MB>(tmA, tmB, p, grid, st);                             \
  } while (0)
  if (write_hi) {
    if (BN != 128) return EQH_ERR_UNSUPPORTED;
    EQH_GEMM_MAJORS(128, true, 3, 1);
  }
  if (BN == 32) EQH_GEMM_MAJORS(32, false, 4, 1);
  if (BN == 64) EQH_GEMM_MAJORS(64, false, 4, 1);
  EQH_GEMM_MAJORS(128, false, 3, 1);
#undef EQH_GEMM_MAJORS
}

}  // extern "C"
