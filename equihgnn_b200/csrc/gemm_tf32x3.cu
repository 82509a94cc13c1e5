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
// per-split partials that eqh_colsum finishes in a fixed order (deterministic).
#include "tc_common.cuh"

namespace eqh {
namespace gemm {
using namespace eqh::tc;

constexpr int BM = 128, BK = 32, STAGES = 3;
constexpr int kThreads = 320;           // warp 0 TMA, warp 1 MMA, warps 2-9 operand split (warps 4-7 also epilogue)
constexpr int kSplitThreads = 256;
constexpr uint32_t kTileBytes = BM * BK * 4;  // 16 KB: a 128 x 32 fp32 operand tile (any major)

struct Params {
  float* D;
  const float* bias;
  const float* rowscale;   // optional [M]: the bias is multiplied by rowscale[m] (empty-segment mask / segment count)
  const float* residual;   // optional [M, ldd]; may alias D (in-place accumulation)
  float out_scale, res_scale;
  float* ws;          // split-K partials [splits, M, N] (ldd == N) or nullptr
  int64_t ldd;
  int M, N, K;
  int act;            // 0 none, 1 relu, 2 silu
  int kb_per_split;   // K blocks of 32 per split
};

// kWriteHi: store the masked hi part back (true) or let the tensor core drop the low 13 bits of the raw fp32
// tile itself (false; valid if the hardware truncates tf32 inputs -- checked by tests/test_gpu_gemm.py).
template <int BN, bool A_KMAJOR, bool B_KMAJOR, bool kWriteHi>
__global__ void __launch_bounds__(kThreads, 1) k_gemm(const __grid_constant__ CUtensorMap tmA,
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
#define tmem_full (*tmem_full_p)
#define tmem_slot (*tmem_slot_p)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int kb_end = min(kb_total, kb_begin + p.kb_per_split);
  const int nkb = max(kb_end - kb_begin, 0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_split[s], kSplitThreads);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Three fp32 accumulators of BN columns each.  The tensor core adds into its accumulator with
  // truncation, so the error grows with the number of accumulations: the large hi.hi terms alternate
  // between two accumulators and the 2^-11-scaled lo terms get their own (summed in the epilogue).
  constexpr uint32_t kTmemCols = BN * 4;       // power of two >= 3*BN
  if (warp == 2) tmem_alloc(&tmem_slot, kTmemCols);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

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
    // ===================== MMA issuer =====================
    if (lane == 0) {
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
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t aH = smem_u32(stageA(s)), aL = smem_u32(stageAlo(s));
        const uint32_t bH = smem_u32(stageB(s)), bL = smem_u32(stageBlo(s));
#pragma unroll
        for (int ks = 0; ks < BK / 8; ++ks) {
          const uint64_t dAh = make_desc(aH + ks * kstepA, lboA, sboA, ltA), dAl = make_desc(aL + ks * kstepA, lboA, sboA, ltA);
          const uint64_t dBh = make_desc(bH + ks * kstepB, lboB, sboB, ltB), dBl = make_desc(bL + ks * kstepB, lboB, sboB, ltB);
          const uint32_t step = (uint32_t)(i * (BK / 8) + ks);
          umma_tf32(tmem_base + 2 * BN, dAl, dBh, idesc, step != 0);
          umma_tf32(tmem_base + 2 * BN, dAh, dBl, idesc, 1);
          umma_tf32(tmem_base + (step & 1) * BN, dAh, dBh, idesc, step >= 2);
        }
        umma_commit(&empty[s]);   // frees the stage when these MMAs have read it
      }
      umma_commit(&tmem_full);
    }
  } else {
    // ===================== operand split (hi / lo): warps 2-9 =====================
    const int t = threadIdx.x - 64;    // 0..255
    for (int i = 0; i < nkb; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&full_raw[s], ph);
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
        constexpr int NB = kBBytes / 16 / kSplitThreads;   // 4 (BN=128) or 2 (BN=64)
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
      mbar_arrive(&full_split[s]);
    }
  }
  if (warp >= 4 && warp < 8) {
    // ---- epilogue: TMEM -> registers -> (bias, act, residual) -> padded smem -> coalesced stores
    constexpr int PITCH = BN + 4;
    float* stage_out = reinterpret_cast<float*>(smem);   // the operand ring is idle once tmem_full fires
    const int q = warp & 3;                              // TMEM lane quadrant this warp may read (warps 4-7 -> 0-3)
    const int row = q * 32 + lane;
    if (nkb > 0) {
      mbar_wait(&tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const bool final_out = (p.ws == nullptr);
    // bias tile -> shared memory once (zero outside N), so the element loop has no bounds checks
    float* sbias = reinterpret_cast<float*>(smem + (size_t)STAGES * kStageBytes + 128);   // [BN] after the barriers
    {
      const int tt = threadIdx.x - 128;
      if (tt < BN) sbias[tt] = (final_out && p.bias && n0 + tt < p.N) ? __ldg(p.bias + n0 + tt) : 0.f;
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
        float w[32];
        tmem_ld32(ta + 2 * BN, v);                           // small terms
        if (nkb * (BK / 8) >= 2) {                           // the odd accumulator exists from the second k-step on
          tmem_ld32(ta + BN, w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
        tmem_ld32(ta, w);
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
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      } else if (act == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(stage_out + row * PITCH + c0 + j) =
            make_float4(v[j] * osc, v[j + 1] * osc, v[j + 2] * osc, v[j + 3] * osc);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps only
    float* out = final_out ? p.D : p.ws + (size_t)blockIdx.z * p.M * p.N;
    const int64_t ld = final_out ? p.ldd : p.N;
    const bool vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const bool has_res = final_out && p.residual != nullptr;
    constexpr int CPL = BN / 32;                         // columns per lane (4 for BN=128: one float4; 2 for BN=64)
    const int c = lane * CPL;
    const int n = n0 + c;
    const bool full_vec = vec_ok && (n + CPL <= p.N) && CPL == 4;
    const int rows_here = min(BM, p.M - m0);
    for (int r = q; r < rows_here; r += 4) {             // one warp per row, lanes across the BN columns
      const float* src = stage_out + r * PITCH + c;
      float* dst = out + (int64_t)(m0 + r) * ld + n;
      if (full_vec) {
        float4 x = *reinterpret_cast<const float4*>(src);
        if (has_res) {   // plain (coherent) load: the residual may be D itself
          const float4 rr = *reinterpret_cast<const float4*>(p.residual + (int64_t)(m0 + r) * ld + n);
          x.x += p.res_scale * rr.x; x.y += p.res_scale * rr.y; x.z += p.res_scale * rr.z; x.w += p.res_scale * rr.w;
        }
        *reinterpret_cast<float4*>(dst) = x;
      } else {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          if (n + j < p.N) {
            float x = src[j];
            if (has_res) x += p.res_scale * p.residual[(int64_t)(m0 + r) * ld + n + j];
            dst[j] = x;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
#undef tmem_full
#undef tmem_slot
}

// Split-K finish: D[m,n] = out_scale * sum_z ws[z,m,n] + res_scale * R[m,n], partials added in split order
// (deterministic); R may alias D.
__global__ void __launch_bounds__(256) k_splitk_reduce(const float* __restrict__ ws, int splits, int M, int N, float* D,
                                                        int64_t ldd, const float* R, float out_scale, float res_scale) {
  const int64_t total = (int64_t)M * N;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
    float acc = 0.f;
    for (int z = 0; z < splits; ++z) acc += __ldg(ws + (int64_t)z * total + i);
    const int64_t m = i / N, n = i - m * N;
    float x = out_scale * acc;
    if (R) x += res_scale * R[m * ldd + n];
    D[m * ldd + n] = x;
  }
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

template <int BN, bool AK, bool BK_>
static int launch(const CUtensorMap& a, const CUtensorMap& b, const Params& p, dim3 grid, cudaStream_t st, bool write_hi) {
  constexpr size_t smem = (size_t)STAGES * (2 * kTileBytes + 2 * BN * BK * 4) + 128 + BN * 4;   // ring + barriers + bias tile
  static_assert((size_t)STAGES * (2 * kTileBytes + 2 * BN * BK * 4) >= (size_t)BM * (BN + 4) * 4,
                "epilogue staging must fit in the operand ring");
  cudaError_t e;
  if (write_hi) {
    e = cudaFuncSetAttribute(k_gemm<BN, AK, BK_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_gemm<BN, AK, BK_, true><<<grid, kThreads, smem, st>>>(a, b, p);
  } else {
    e = cudaFuncSetAttribute(k_gemm<BN, AK, BK_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    k_gemm<BN, AK, BK_, false><<<grid, kThreads, smem, st>>>(a, b, p);
  }
  EQH_LAUNCH_RC();
}

}  // namespace gemm
}  // namespace eqh

extern "C" {

size_t eqh_gemm_workspace_bytes(int64_t M, int64_t N, int splits) {
  return splits > 1 ? sizeof(float) * (size_t)splits * (size_t)M * (size_t)N : 0;
}

int eqh_gemm_tf32x3(const float* A, int a_kmajor, int64_t lda, const float* B, int b_kmajor, int64_t ldb, float* D,
                    int64_t ldd, int64_t M, int64_t N, int64_t K, const float* bias, const float* rowscale, int act_flags,
                    const float* residual, float out_scale, float res_scale, int splits, float* ws, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::gemm;
  const int act = act_flags & 0xff;
  EQH_CHECK_ARG(A && B && D && M > 0 && N > 0 && K > 0 && splits >= 1 && act >= 0 && act <= 2);
  EQH_CHECK_ARG(M < (1 << 30) && N < (1 << 30) && K < (1 << 30));
  EQH_CHECK_ARG(aligned16(A) && aligned16(B) && lda % 4 == 0 && ldb % 4 == 0);   // TMA: 16-byte rows
  EQH_CHECK_ARG(splits == 1 || (ws && !bias && act == 0));
  EQH_CHECK_ARG(ldd >= N);
  const int BN = N <= 64 ? 64 : 128;
  CUtensorMap tmA, tmB;
  int rc;
  // A: logical [M,K].  K-major: stored [M, lda] -> box {32 K, 128 M}.  MN-major: stored [K, lda] -> box {32 M, 32 K}.
  // MN-major tf32 operands must use the 128B swizzle with 32-byte atoms (UMMA layout SWIZZLE_128B_BASE32B).
  rc = a_kmajor ? make_map(&tmA, A, M, K, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map(&tmA, A, K, M, lda, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  rc = b_kmajor ? make_map(&tmB, B, N, K, ldb, BK, (uint32_t)BN, CU_TENSOR_MAP_SWIZZLE_128B)
                : make_map(&tmB, B, K, N, ldb, 32, BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (rc) return rc;
  Params p;
  p.D = D; p.bias = bias; p.rowscale = rowscale; p.residual = residual; p.ws = splits > 1 ? ws : nullptr; p.ldd = ldd;
  p.out_scale = out_scale; p.res_scale = res_scale;
  p.M = (int)M; p.N = (int)N; p.K = (int)K; p.act = act;
  const int kb_total = (int)((K + BK - 1) / BK);
  p.kb_per_split = (kb_total + splits - 1) / splits;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool write_hi = (act_flags & 0x100) != 0;   // debug/verification switch, see kWriteHi
  auto dispatch = [&]() -> int {
#define EQH_GEMM_CASE(BNV)                                                        \
  if (a_kmajor && b_kmajor) return launch<BNV, true, true>(tmA, tmB, p, grid, st, write_hi); \
  if (a_kmajor && !b_kmajor) return launch<BNV, true, false>(tmA, tmB, p, grid, st, write_hi); \
  if (!a_kmajor && !b_kmajor) return launch<BNV, false, false>(tmA, tmB, p, grid, st, write_hi); \
  return launch<BNV, false, true>(tmA, tmB, p, grid, st, write_hi);
    if (BN == 64) { EQH_GEMM_CASE(64) }
    EQH_GEMM_CASE(128)
  };
  rc = dispatch();
  if (rc || splits == 1) return rc;
  {
    int64_t blocks = (M * N + 255) / 256;
    if (blocks > (int64_t)kSMs * 16) blocks = (int64_t)kSMs * 16;
    k_splitk_reduce<<<(unsigned)blocks, 256, 0, st>>>(ws, splits, (int)M, (int)N, D, ldd, residual, out_scale, res_scale);
  }
  EQH_LAUNCH_RC();
#undef EQH_GEMM_CASE
}

}  // extern "C"
