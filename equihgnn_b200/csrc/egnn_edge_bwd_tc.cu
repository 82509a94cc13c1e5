// EGNN edge model phi_e + neighbour pooling, BACKWARD, on the tcgen05 tensor cores.
//
// Autograd of egnn_layer.py:298-310,357-358 in the split form of egnn_edge_tc.cu
//     pre = P[i] + Q[j] + d2*wc ;  hid = SiLU(pre) ;  s2 = hid.W2^T + b2 ;  m_i = sum_k SiLU(s2)
// given dMi [N,16]:
//     dS2[e,o] = dMi[i,o] * SiLU'(S2[e,o])                                  (CUDA cores, 16 values per edge)
//     dH[e,c]  = sum_o dS2[e,o] W2[o,c]                                     (tensor cores, MMA 1)
//     dPre     = dH * SiLU'(pre) ;  dP[i] = sum_k dPre ;  dwc = sum_e d2 dPre   (CUDA cores, one pass over Q[j])
//     dW2[o,c] = sum_e dS2[e,o] hid[e,c]                                    (tensor cores, MMA 2)
// The CUDA-core kernel (egnn_edge.cu) spends 32 of its ~50 instructions per (edge, column) on the two 16-wide
// contractions; here they cost 14 small tcgen05.mma per 32 edges x 128 columns.
//
// Everything is TRANSPOSED relative to the forward kernel so that no operand ever has to be transposed in shared
// memory: a CTA owns ONE group of 128 hidden columns (= the 128 TMEM lanes) for a contiguous slice of the edges.
//   MMA 1   D1[c, e]  = W2T[c, o] . dS2[e, o]^T      A = W2^T tile (K-major, shared memory, loaded once per CTA)
//                                                    B = dS2 block [32 edges][hi 16 | lo 16] (K-major)
//   thread c (one per lane) reads D1[c, 0..31] from TMEM; Q[j, c] / P[i, c] come from row segments the warpgroup copied
//   into shared memory one block ahead (16-byte cp.async, one 512-byte row per warp instruction; per-lane
//   cp.async.bulk copies were measured first: the uniform-datapath instruction serialises over the lanes, ~90 cycles
//   per row), dP and dwc accumulate in registers in edge order (deterministic), dPre rows are coalesced
//   stores, and hid[c, e] goes straight back into TENSOR MEMORY (tcgen05.st, hi | lo) as the A operand of
//   MMA 2   D2[c, o] += hid[c, e] . dS2T[o, e]^T     A from TMEM (TS form), B = dS2 transposed [hi 16 | lo 16 rows][32 edges]
// fp32-faithful (3xTF32: lo.hi + hi.lo + hi.hi as in gemm_tf32x3.cu).  D2 is drained into fp32 registers every 8
// blocks (two TMEM buffers), so no accumulator sees more than 64 tensor-core accumulations.
//
// Warp roles (17 warps):
//   warps 0-3    dS2 producer (4 values per thread and block; its loads are issued 4 / 8 blocks ahead), writes both
//                layouts + the block's j / d2 / atom ids into an 8-deep ring; drains D2; finishes dW2 / dwc / db2 partials
//   warps 4-15   three "column" warpgroups, round-robin over the blocks: the per-(edge, column) math above
//   warp 16      MMA issuer (warp-uniform loop, one elected lane), MMA 2 lags MMA 1 by three blocks
// Grid: ceil(H/128) column groups x S edge slices (S = 148 / groups); part row s receives the slice's partial sums.
#include "tc_common.cuh"

namespace eqh {
namespace edge_bwd_tc {
using namespace eqh::tc;

constexpr int kM = EQH_M_DIM;       // 16
constexpr int kNbr = EQH_KNN_K;     // 16
constexpr int GC = 128;             // hidden columns per CTA (TMEM lanes)
constexpr int EB = 32;              // edges per block: two atoms
constexpr int NSLOT = 4;            // blocks in flight in tensor memory (D1 / hid slots)
constexpr int DSLOT = 8;            // dS2 tile + meta ring in shared memory: the producer runs this far ahead
constexpr int LAG = 3;              // MMA 2 of block b is issued after MMA 1 of block b + LAG
constexpr int kEpiWG = 3;
constexpr int kFlush = 8;           // blocks per D2 accumulation round
constexpr int kMmaWarp = 4 + 4 * kEpiWG;
constexpr int kThreads = (kMmaWarp + 1) * 32;
static_assert(kM == 16 && kNbr == 16, "layouts below assume 16 x 16");
static_assert(LAG == NSLOT - 1 && LAG >= kEpiWG, "pipeline depth");

// tensor memory columns
constexpr uint32_t kD2Col = 0;                     // 2 buffers x 32 (hi-product columns | lo-product columns)
constexpr uint32_t kD1Col = 64;                    // NSLOT x 32
constexpr uint32_t kA2Col = kD1Col + 32 * NSLOT;   // NSLOT x 64 (hid hi | hid lo)
static_assert(kA2Col + 64 * NSLOT <= 512, "TMEM budget");

// shared memory
struct Meta {
  int32_t j[EB];
  float dd[EB];
  int32_t atom[2];
  int32_t slot[EB];   // slot mode: row of the partial-dQ buffer this edge accumulates into
  uint8_t first8[4];  // ... and, bit k of byte 2 * half + (k >> 3) [k & 7]: the edge is the first of its slot (it writes)
};
constexpr uint32_t kW2TBytes = GC * 128;           // [128 c][hi 16 | lo 16] K-major, 128B swizzle
constexpr uint32_t kBBytes = EB * 128;             // 4 KB: B1 [32 e][hi 16 | lo 16], B2 [hi 16 | lo 16 rows][32 e]
constexpr uint32_t kOffB1 = kW2TBytes;
constexpr uint32_t kOffB2 = kOffB1 + DSLOT * kBBytes;
constexpr uint32_t kOffMeta = kOffB2 + DSLOT * kBBytes;
constexpr uint32_t kMetaStride = 512;
constexpr uint32_t kOffBar = kOffMeta + DSLOT * kMetaStride;
constexpr uint32_t kOffRedWc = kOffBar + 512;                       // [kEpiWG][128] floats
constexpr uint32_t kOffRedDb = kOffRedWc + kEpiWG * GC * 4;         // [128][4] floats
constexpr uint32_t kOffAcc = kOffRedDb + 128 * 4 * 4;               // [16][128] floats: dW2 of this slice (drained D2 rounds)
constexpr uint32_t kQRows = EB + 2;                                 // 32 neighbour rows of Q + the two atoms' rows of P
constexpr uint32_t kQStageBytes = kQRows * GC * 4;                  // 17 KB
constexpr uint32_t kOffQ = (kOffAcc + kM * GC * 4 + 127) / 128 * 128;
constexpr uint32_t kSmem = kOffQ + 2 * kEpiWG * kQStageBytes;      // per column warpgroup: current + next block
static_assert(kSmem <= 227 * 1024, "shared memory budget");
static_assert(sizeof(Meta) <= kMetaStride, "meta block");

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

// 16 consecutive TMEM columns of this thread's lane <- 16 registers
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

#ifdef EQH_BWDTC_TRACE
__device__ unsigned long long g_bt_trace[64];
#define TW(i, bar, par)                                  \
  do {                                                   \
    const long long t0__ = clock64();                    \
    mbar_wait(bar, par);                                 \
    tw[i] += clock64() - t0__;                           \
  } while (0)
#define TS(i)                                 \
  do {                                        \
    const long long t1__ = clock64();         \
    ts[i] += t1__ - tprev;                    \
    tprev = t1__;                             \
  } while (0)
#else
#define TW(i, bar, par) mbar_wait(bar, par)
#define TS(i)
#endif

// SiLU'(x) with the two SFU approximations (~1 ulp each; see egnn_edge_tc.cu)
__device__ __forceinline__ float fast_dsilu(float x) {
  float e, sg;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + e));
  return sg * (1.f + x * (1.f - sg));
}

// byte offset of 16-byte chunk `chunk` of row `row` in a K-major tile with 128-byte rows and the 128B swizzle
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

// ---- slot mode (dQ partial sums inside the CTA's sub-slices instead of one dPre row per edge) --------------------------
// The slice of a CTA is cut into kEpiWG contiguous SUB-SLICES, one per column warpgroup: block bl of the pipeline (the
// warpgroup bl % kEpiWG processes it) is block sub_start(bl % kEpiWG) + bl / kEpiWG of the slice.  All edges of a sub-slice
// that point at the same atom j accumulate into ONE row ("slot") of the partial buffer, in the warpgroup's own program
// order (deterministic, no atomics: one thread owns a column of a sub-slice for the whole kernel); with the atoms in grid-
// cell order a sub-slice of ~600 atoms hits each of its neighbours 5-6 times, so the buffer and the reduction that follows
// shrink by that factor.  Slots are numbered j-major (egnn_edge_slots.cu): the rows of one j are consecutive and the final
// dQ[j] is a plain segmented sum over consecutive rows.
struct SubSlice {
  int64_t b_begin;   // first block (two atoms) of the slice
  int nb;            // blocks in the slice
  int start[kEpiWG]; // first block of each sub-slice, relative to b_begin
};
__host__ __device__ __forceinline__ SubSlice sub_slices(int64_t N, int nslices, int sl) {
  SubSlice s;
  const int64_t nblk = (N + 1) >> 1;
  s.b_begin = nblk * sl / nslices;
  s.nb = (int)(nblk * (sl + 1) / nslices - s.b_begin);
  int at = 0;
  for (int w = 0; w < kEpiWG; ++w) {
    s.start[w] = at;
    at += (s.nb - w + kEpiWG - 1) / kEpiWG;      // pipeline blocks bl < nb with bl % kEpiWG == w
  }
  return s;
}

template <bool SLOTS>
__global__ void __launch_bounds__(kThreads, 1)
k_edge_bwd_tc(const float* __restrict__ P, const float* __restrict__ Q, const float* __restrict__ wc,
              const float* __restrict__ W2, const int32_t* __restrict__ idx, const float* __restrict__ d2,
              const float* __restrict__ S2, const float* __restrict__ dMi, int64_t N, int H, int ngroups, int nslices,
              float* __restrict__ dP, float* __restrict__ dPre, float* __restrict__ part, int64_t part_cols,
              const int32_t* __restrict__ order, int64_t ldpq, int64_t lddp, int64_t lddmi,
              const int32_t* __restrict__ slot_edge) {
  pdl_prologue();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* ds_full = reinterpret_cast<uint64_t*>(smem + kOffBar);   // [DSLOT] dS2 tiles + meta written
  uint64_t* slot_free = ds_full + DSLOT;                             // [DSLOT] MMA 2 done: tiles and meta reusable
  uint64_t* d1_full = slot_free + DSLOT;                             // [NSLOT] MMA 1 done
  uint64_t* a2_full = d1_full + NSLOT;                               // [NSLOT] hid written to TMEM
  uint64_t* d2_full = a2_full + NSLOT;                               // [2] accumulation round complete
  uint64_t* d2_empty = d2_full + 2;                                  // [2] drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 2);
  static_assert((2 * DSLOT + 2 * NSLOT + 4) * 8 + 4 <= 512, "barrier block");
  float* red_wc = reinterpret_cast<float*>(smem + kOffRedWc);
  float* red_db = reinterpret_cast<float*>(smem + kOffRedDb);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef EQH_BWDTC_TRACE
  long long tw[4] = {0, 0, 0, 0};
  long long ts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long t_start = clock64();
  long long tprev = t_start;
#endif
  const int g = blockIdx.x % ngroups;          // column group
  const int sl = blockIdx.x / ngroups;         // edge slice (= row of `part`)
  const int64_t nblk = (N + 1) >> 1;           // blocks of two atoms
  const int64_t b_begin = nblk * sl / nslices, b_end = nblk * (sl + 1) / nslices;
  const int nb = (int)(b_end - b_begin);
  constexpr bool slots = SLOTS;
  // slot mode: first block of the sub-slices of warpgroups 1 and 2 (sub_slices(): warpgroup w owns the pipeline blocks
  // bl < nb with bl % kEpiWG == w, as one contiguous run of the slice)
  static_assert(kEpiWG == 3, "sub-slice starts below");
  const int ss1 = (nb + 2) / 3, ss2 = ss1 + (nb + 1) / 3;
  const int64_t ld = H;

  if (threadIdx.x == 0) {
    for (int s = 0; s < DSLOT; ++s) {
      mbar_init(&ds_full[s], 4);
      mbar_init(&slot_free[s], 1);
    }
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(&d1_full[s], 1);
      mbar_init(&a2_full[s], 4);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&d2_full[b], 1);
      mbar_init(&d2_empty[b], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  // W2^T tile of this column group: row c = [W2[0..15, col] raw | their tf32 remainders]; zero rows past H
  for (int t = threadIdx.x; t < GC * kM; t += kThreads) {
    const int c = t & (GC - 1), o = t >> 7;
    const int col = g * GC + c;
    const float w = col < H ? __ldg(W2 + (int64_t)o * H + col) : 0.f;
    *reinterpret_cast<float*>(smem + sw128(c, o >> 2) + (o & 3) * 4) = w;
    *reinterpret_cast<float*>(smem + sw128(c, 4 + (o >> 2)) + (o & 3) * 4) = tf32_lo(w);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ============================ dS2 producer / D2 drain ============================
    const int t = threadIdx.x;           // 0..127 (= column c for the drain)
    const int el = t >> 2, oq = t & 3;   // edge of the block, quad of outputs
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float db[4] = {0.f, 0.f, 0.f, 0.f};
    float* acc = reinterpret_cast<float*>(smem + kOffAcc) + t;      // dW2[o, column t] of this slice at acc[o * GC]
#pragma unroll
    for (int o = 0; o < kM; ++o) acc[o * GC] = 0.f;
    auto drain = [&](uint32_t r) {
      const uint32_t buf = r & 1u;
      TW(1, &d2_full[buf], (r >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[kM];
      tmem_ld16(tmem_base + lane_off + kD2Col + buf * 32 + kM, v);     // hid_hi . dS2_lo
#pragma unroll
      for (int o = 0; o < kM; ++o) acc[o * GC] += v[o];
      tmem_ld16(tmem_base + lane_off + kD2Col + buf * 32, v);          // hid . dS2_hi
#pragma unroll
      for (int o = 0; o < kM; ++o) acc[o * GC] += v[o];
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&d2_empty[buf]);
    };
    uint32_t drained = 0;
    // The loads of a block (order -> S2 / dMi / idx / d2: two dependent round trips) are issued PD blocks (values) and
    // 2 PD blocks (atom ids) ahead of its turn and wait in registers: the producer never stalls on memory latency.
    constexpr int PD = 4;
    struct PV {
      float4 s2, gm;
      int32_t jj, atom, slot;
      float dd;
    };
    auto load_atom = [&](int bl) -> int32_t {           // atom of this thread's edge in block bl; -1: none
      if (bl >= nb) return -1;
      // slot mode: the warpgroup bl % kEpiWG walks its own contiguous sub-slice
      const int wq = bl % kEpiWG;
      const int64_t blk = slots ? (int64_t)(wq == 0 ? 0 : (wq == 1 ? ss1 : ss2)) + bl / kEpiWG : (int64_t)bl;
      const int64_t aslot = (b_begin + blk) * 2 + (el >> 4);
      if (aslot >= N) return -1;
      return order ? __ldg(order + aslot) : (int32_t)aslot;
    };
    auto load_vals = [&](int32_t atom) -> PV {
      PV v;
      v.atom = atom;
      v.s2 = v.gm = make_float4(0.f, 0.f, 0.f, 0.f);
      v.jj = 0;
      v.slot = 0;
      v.dd = 0.f;
      if (atom >= 0) {
        const int64_t e = (int64_t)atom * kNbr + (el & 15);
        v.s2 = __ldg(reinterpret_cast<const float4*>(S2 + e * kM) + oq);
        v.gm = __ldg(reinterpret_cast<const float4*>(dMi + (int64_t)atom * lddmi) + oq);
        if (oq == 0) {
          v.jj = __ldg(idx + e);
          v.dd = __ldg(d2 + e);
          if (slots) v.slot = __ldg(slot_edge + e);
        }
      }
      return v;
    };
    auto produce = [&](const PV& v, int bl) {
      const int s = bl % DSLOT;
      const uint32_t ph = (uint32_t)(bl / DSLOT) & 1u;
      float4 ds = make_float4(0.f, 0.f, 0.f, 0.f);
      if (v.atom >= 0)
        ds = make_float4(v.gm.x * fast_dsilu(v.s2.x), v.gm.y * fast_dsilu(v.s2.y), v.gm.z * fast_dsilu(v.s2.z), v.gm.w * fast_dsilu(v.s2.w));
      const float4 dl = make_float4(tf32_lo(ds.x), tf32_lo(ds.y), tf32_lo(ds.z), tf32_lo(ds.w));
      TW(0, &slot_free[s], ph ^ 1u);
      uint8_t* b1 = smem + kOffB1 + s * kBBytes;
      uint8_t* b2 = smem + kOffB2 + s * kBBytes;
      *reinterpret_cast<float4*>(b1 + sw128(el, oq)) = ds;
      *reinterpret_cast<float4*>(b1 + sw128(el, 4 + oq)) = dl;
      {
        const float hv[4] = {ds.x, ds.y, ds.z, ds.w}, lv[4] = {dl.x, dl.y, dl.z, dl.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t n = oq * 4 + u;
          *reinterpret_cast<float*>(b2 + sw128(n, el >> 2) + (el & 3) * 4) = hv[u];
          *reinterpret_cast<float*>(b2 + sw128(kM + n, el >> 2) + (el & 3) * 4) = lv[u];
        }
      }
      Meta* m = reinterpret_cast<Meta*>(smem + kOffMeta + s * kMetaStride);
      if (oq == 0) {
        m->j[el] = v.jj;
        m->dd[el] = v.dd;
        m->slot[el] = v.slot & 0x7fffffff;
        if ((el & 15) == 0) m->atom[el >> 4] = v.atom;
      }
      if (slots) {      // this warp's 8 edges (lanes 0, 4, .., 28) -> one byte of first-edge flags
        const unsigned bal = __ballot_sync(kFull, oq == 0 && v.slot < 0);
        unsigned m8 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) m8 |= ((bal >> (4 * i)) & 1u) << i;
        if (lane == 0) m->first8[warp] = (uint8_t)m8;
      }
      db[0] += ds.x; db[1] += ds.y; db[2] += ds.z; db[3] += ds.w;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[s]);
      // the round before the current one ended kFlush - 1 blocks ago (MMA 2 lags production by the LAG blocks in
      // flight): its MMAs have completed, and the issuer needs the buffer again only at the next block
      if ((bl % kFlush) == kFlush - 1 && bl >= kFlush) {
        drain((uint32_t)(bl / kFlush) - 1u);
        drained = (uint32_t)(bl / kFlush);
      }
    };
    PV pv[PD];
    int32_t an[PD];
#pragma unroll
    for (int u = 0; u < PD; ++u) an[u] = load_atom(u);
#pragma unroll
    for (int u = 0; u < PD; ++u) {
      pv[u] = load_vals(an[u]);
      an[u] = load_atom(u + PD);
    }
    for (int base = 0; base < nb; base += PD) {
#pragma unroll
      for (int u = 0; u < PD; ++u) {
        const int bl = base + u;
        if (bl < nb) {
          produce(pv[u], bl);
          pv[u] = load_vals(an[u]);
          an[u] = load_atom(bl + 2 * PD);
        }
      }
    }
    const uint32_t nrounds = (uint32_t)((nb + kFlush - 1) / kFlush);
    for (uint32_t r = drained; r < nrounds; ++r) drain(r);
    *reinterpret_cast<float4*>(red_db + t * 4) = make_float4(db[0], db[1], db[2], db[3]);
    if (g * GC + t < H) {
      float* row = part + (int64_t)sl * part_cols + g * GC + t;
#pragma unroll
      for (int o = 0; o < kM; ++o) row[(int64_t)o * H] = acc[o * GC];
    }
  } else if (warp == kMmaWarp) {
    // ============================ MMA issuer ============================
    constexpr uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(GC >> 4) << 24);
    constexpr uint32_t idesc16 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(GC >> 4) << 24);
    const uint32_t w2t = smem_u32(smem);
    for (int step = 0; step < nb + LAG; ++step) {
      if (step < nb) {
        const int sd = step % DSLOT, s = step % NSLOT;
        TW(0, &ds_full[sd], (uint32_t)(step / DSLOT) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b1 = smem_u32(smem + kOffB1 + sd * kBBytes);
        const uint32_t d1 = tmem_base + kD1Col + s * 32;
        if (elect_one()) {
          // D1[c, e] = sum_o W2T[c, o] dS2[e, o]: lo.hi, hi.lo (small terms first), then hi.hi; two k-steps of 8 each
          umma_tf32(d1, make_desc(w2t + 64, 0, 1024, 2), make_desc(b1, 0, 1024, 2), idesc32, 0);
          umma_tf32(d1, make_desc(w2t + 96, 0, 1024, 2), make_desc(b1 + 32, 0, 1024, 2), idesc32, 1);
          umma_tf32(d1, make_desc(w2t, 0, 1024, 2), make_desc(b1 + 64, 0, 1024, 2), idesc32, 1);
          umma_tf32(d1, make_desc(w2t + 32, 0, 1024, 2), make_desc(b1 + 96, 0, 1024, 2), idesc32, 1);
          umma_tf32(d1, make_desc(w2t, 0, 1024, 2), make_desc(b1, 0, 1024, 2), idesc32, 1);
          umma_tf32(d1, make_desc(w2t + 32, 0, 1024, 2), make_desc(b1 + 32, 0, 1024, 2), idesc32, 1);
          umma_commit(&d1_full[s]);
        }
        __syncwarp();
      }
      if (step >= LAG) {
        const int bl = step - LAG;
        const int s = bl % NSLOT, sd = bl % DSLOT;
        const uint32_t ph = (uint32_t)(bl / NSLOT) & 1u;
        const uint32_t r = (uint32_t)(bl / kFlush), buf = r & 1u;
        const bool first = (bl % kFlush) == 0;
        if (first) TW(2, &d2_empty[buf], ((r >> 1) & 1u) ^ 1u);     // drained two rounds ago
        TW(1, &a2_full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b2 = smem_u32(smem + kOffB2 + sd * kBBytes);
        const uint32_t aH = tmem_base + kA2Col + s * 64, aL = aH + 32;
        const uint32_t dacc = tmem_base + kD2Col + buf * 32;
        const bool last = (bl % kFlush) == kFlush - 1 || bl == nb - 1;
        if (elect_one()) {
          // D2[c, 0..15] += hid_hi.dS2_hi + hid_lo.dS2_hi ;  D2[c, 16..31] += hid_hi.dS2_lo   (summed by the drain)
#pragma unroll
          for (int ks = 0; ks < EB / 8; ++ks) {
            const uint64_t db2 = make_desc(b2 + ks * 32, 0, 1024, 2);
            umma_tf32_ts(dacc, aH + ks * 8, db2, idesc32, (first && ks == 0) ? 0u : 1u);
            umma_tf32_ts(dacc, aL + ks * 8, db2, idesc16, 1u);
          }
          umma_commit(&slot_free[sd]);
          if (last) umma_commit(&d2_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ column warpgroups ============================
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;                        // TMEM lane quadrant
    const int c = q * 32 + lane;
    const int col = g * GC + c;
    const bool cv = col < H;
    const float wcr = cv ? __ldg(wc + col) : 0.f;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const char* slot_base = reinterpret_cast<const char*>(dPre + col);      // slot mode: row s of this column at + s * ldb
    const uint32_t ldb = (uint32_t)ld * 4u;
    float dwc = 0.f;
    // Q[j] / P[i] row segments (512 B each) of the warpgroup's NEXT block are copied into shared memory with 16-byte
    // cp.async (one row per warp instruction, fully coalesced) while the current block is processed.
    const uint32_t row_bytes = (uint32_t)min(GC, H - g * GC) * 4u;
    uint8_t* qring = smem + kOffQ + wg * 2 * kQStageBytes;
    auto fetch_rows = [&](int blk, int stage) {
      const int sd = blk % DSLOT;
      TW(0, &ds_full[sd], (uint32_t)(blk / DSLOT) & 1u);
      const Meta* mm = reinterpret_cast<const Meta*>(smem + kOffMeta + sd * kMetaStride);
      if ((uint32_t)lane * 16u < row_bytes) {
        // warp q copies rows q, q+4, ..., q+28 of Q and (q < 2) the P row of atom q: all row ids first, then the copies
        const float* src[9];
#pragma unroll
        for (int i = 0; i < 8; ++i) src[i] = Q + (int64_t)mm->j[q + 4 * i] * ldpq;
        {
          const int32_t a = mm->atom[q & 1];
          src[8] = P + (int64_t)(a >= 0 ? a : 0) * ldpq;
        }
        const uint32_t d = smem_u32(qring + stage * kQStageBytes + lane * 16 + q * (GC * 4));
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + i * 4 * (GC * 4)), "l"(src[i] + g * GC + lane * 4) : "memory");
        if (q < 2)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 8 * 4 * (GC * 4)), "l"(src[8] + g * GC + lane * 4) : "memory");
      }
    };
    if (wg < nb) fetch_rows(wg, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");
    int it = 0;
    for (int bl = wg; bl < nb; bl += kEpiWG, ++it) {
      const int s = bl % NSLOT, sd = bl % DSLOT;
      const uint32_t ph = (uint32_t)(bl / NSLOT) & 1u;
      TS(0);
      if (bl + kEpiWG < nb) fetch_rows(bl + kEpiWG, (it + 1) & 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
      TS(1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");            // this block's rows have landed (the next may be in flight)
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");     // ... for all four warps of the warpgroup
      TS(2);
      const Meta* m = reinterpret_cast<const Meta*>(smem + kOffMeta + sd * kMetaStride);
      const int32_t a0 = m->atom[0], a1 = m->atom[1];
      const float* qs = reinterpret_cast<const float*>(qring + (it & 1) * kQStageBytes) + c;
      TW(2, &d1_full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d1 = tmem_base + lane_off + kD1Col + s * 32;
      const uint32_t a2 = tmem_base + lane_off + kA2Col + s * 64;
      TS(3);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int32_t at = half ? a1 : a0;
        const bool wr = cv && at >= 0;
        const float pi = wr ? qs[(EB + half) * GC] : 0.f;
        float dh[kNbr], hv[kNbr];
        // slot mode: the rows this atom's 16 edges accumulate into are fetched BEFORE the element math (one L2 round trip
        // hidden behind ~600 instructions).  The 16 neighbours of one atom are distinct, so are their slots; earlier
        // updates of a slot were made by THIS thread (its column, its sub-slice) -- program order is the only ordering
        // needed, and the volatile asm keeps the loads behind the previous half's stores.  The first edge of a slot
        // writes (no zero-fill of the buffer).
        float old[kNbr];
        if (slots && wr) {
          const uint32_t fm = reinterpret_cast<const uint16_t*>(m->first8)[half];
#pragma unroll
          for (int k = 0; k < kNbr; ++k) {
            const uint32_t sk = (uint32_t)m->slot[half * 16 + k];
            old[k] = 0.f;
            if (!((fm >> k) & 1u))
              asm volatile("ld.global.f32 %0, [%1];" : "=f"(old[k]) : "l"(slot_base + (uint64_t)sk * ldb) : "memory");
          }
        }
        tmem_ld16(d1 + half * 16, dh);
        float dpa = 0.f;
#pragma unroll
        for (int k = 0; k < kNbr; ++k) {
          const float dd = m->dd[half * 16 + k];
          const float pre = (pi + (cv ? qs[(half * 16 + k) * GC] : 0.f)) + dd * wcr;
          float ex, sg;                                   // SFU approximations (~1 ulp each), see egnn_edge_tc.cu
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(pre * -1.4426950408889634f));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(1.f + ex));
          const float hid = pre * sg;
          hv[k] = hid;
          const float dpre = dh[k] * (sg + hid * (1.f - sg));      // SiLU'(x) = s + x s (1 - s)
          dpa += dpre;
          dwc += dd * dpre;
          dh[k] = dpre;
        }
        if (wr && !slots) {                               // one branch per atom: the element loop above is straight-line code
          float* dprow = dPre + ((int64_t)at * kNbr) * ld + col;
#pragma unroll
          for (int k = 0; k < kNbr; ++k, dprow += ld) *dprow = dh[k];
          dP[(int64_t)at * lddp + col] = dpa;
        } else if (wr) {
#pragma unroll
          for (int k = 0; k < kNbr; ++k) {
            const uint32_t sk = (uint32_t)m->slot[half * 16 + k];
            asm volatile("st.global.f32 [%0], %1;" ::"l"(slot_base + (uint64_t)sk * ldb), "f"(old[k] + dh[k]) : "memory");
          }
          dP[(int64_t)at * lddp + col] = dpa;
        }
        tmem_st16(a2 + half * 16, hv);                    // hi: the tensor core ignores the low 13 bits
#pragma unroll
        for (int k = 0; k < kNbr; ++k) hv[k] = tf32_lo(hv[k]);
        tmem_st16(a2 + 32 + half * 16, hv);               // lo
      }
      TS(4);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&a2_full[s]);
      TS(5);
      asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory");     // all reads of this stage done before it is refilled
      TS(6);
    }
    red_wc[wg * GC + c] = dwc;
  }
#ifdef EQH_BWDTC_TRACE
  if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4 || warp == 8 || warp == kMmaWarp)) {
    const int r = warp == 0 ? 0 : warp == 4 ? 1 : warp == 8 ? 2 : 3;
    g_bt_trace[r * 8 + 0] = (unsigned long long)(clock64() - t_start);
    for (int i = 0; i < 4; ++i) g_bt_trace[r * 8 + 1 + i] = (unsigned long long)tw[i];
    g_bt_trace[r * 8 + 5] = (unsigned long long)nb;
    if (r == 1) for (int i = 0; i < 8; ++i) g_bt_trace[40 + i] = (unsigned long long)ts[i];
  }
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
  if (warp < 4) {
    // ---- partial sums of this (group, slice) -> row `sl` of part
    const int t = threadIdx.x;
    float* row = part + (int64_t)sl * part_cols;
    const int col = g * GC + t;
    if (col < H) {
      float w = 0.f;
#pragma unroll
      for (int k = 0; k < kEpiWG; ++k) w += red_wc[k * GC + t];
      row[(int64_t)kM * H + col] = w;
    }
    if (g == 0 && t < kM) {
      float b = 0.f;
      for (int e2 = 0; e2 < EB; ++e2) b += red_db[(e2 * 4 + (t >> 2)) * 4 + (t & 3)];
      row[(int64_t)(kM + 1) * H + t] = b;
    }
  }
}

// ---- slot numbering (integer work; deterministic: atomicMin only) ------------------------------------------------------
// position ap of the visiting order -> sub-slice u = slice * kEpiWG + warpgroup
__device__ __forceinline__ int sub_slice_of(int64_t ap, int64_t N, int nslices) {
  const int64_t nblk = (N + 1) >> 1, B = ap >> 1;
  int sl = (int)(((B + 1) * nslices + nblk - 1) / nblk) - 1;         // largest sl with nblk * sl / nslices <= B
  sl = sl < 0 ? 0 : (sl >= nslices ? nslices - 1 : sl);
  while (sl > 0 && nblk * sl / nslices > B) --sl;
  while (sl + 1 < nslices && nblk * (sl + 1) / nslices <= B) ++sl;
  const SubSlice ss = sub_slices(N, nslices, sl);
  const int rb = (int)(B - ss.b_begin);
  int w = 0;
#pragma unroll
  for (int k = 1; k < kEpiWG; ++k)
    if (rb >= ss.start[k]) w = k;
  return sl * kEpiWG + w;
}

// pass 1: mark[j * U + u] = first position-edge (ap * 16 + k) of the pair (neighbour j, sub-slice u); mark starts at INT_MAX
__global__ void __launch_bounds__(256) k_slots_mark(const int32_t* __restrict__ idx, const int32_t* __restrict__ order,
                                                    int64_t N, int nslices, int U, int32_t* __restrict__ mark) {
  pdl_prologue();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * kNbr) return;
  const int64_t ap = t / kNbr;
  const int k = (int)(t % kNbr);
  const int64_t atom = order ? __ldg(order + ap) : ap;
  const int32_t j = __ldg(idx + atom * kNbr + k);
  atomicMin(mark + (int64_t)j * U + sub_slice_of(ap, N, nslices), (int32_t)t);
}

// pass 2 (after an exclusive scan of [mark != INT_MAX] into slot_ext): slot_edge[atom * 16 + k] = slot | first-edge flag
__global__ void __launch_bounds__(256) k_slots_assign(const int32_t* __restrict__ idx, const int32_t* __restrict__ order,
                                                      int64_t N, int nslices, int U, const int32_t* __restrict__ mark,
                                                      const int32_t* __restrict__ slot_ext, int32_t* __restrict__ slot_edge) {
  pdl_prologue();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= N * kNbr) return;
  const int64_t ap = t / kNbr;
  const int k = (int)(t % kNbr);
  const int64_t atom = order ? __ldg(order + ap) : ap;
  const int32_t j = __ldg(idx + atom * kNbr + k);
  const int64_t key = (int64_t)j * U + sub_slice_of(ap, N, nslices);
  const int32_t slot = __ldg(slot_ext + key);
  slot_edge[atom * kNbr + k] = __ldg(mark + key) == (int32_t)t ? (slot | (int32_t)0x80000000) : slot;
}

static int groups_of(int64_t H) { return (int)((H + GC - 1) / GC); }
// CTAs = groups x slices, one per SM (200 KB of shared memory, the whole register file each).
// slices_of: the most slices a launch uses (= rows of `part` the callers allocate, = sub-slice keys per neighbour / 3);
// slices_for: the slices of a launch over N atoms.  Small batches leave 8 SMs to the kernels of the parallel graph branches
// (weight-gradient GEMMs, column sums, the bucketed gradient exchange), which otherwise wait for this kernel -- the longest
// of the backward's tail -- to end: QM9-128 step 1.058 -> 1.042 ms (same box; 16 / 24 free SMs: 1.042 / 1.049).  Large
// batches are bound by SM time, not by the length of the chain: all SMs (PCQM4Mv2-768 5.33 vs 5.38 ms with 8 left free).
// EQH_EDGE_BWD_FREE_SMS overrides both.
static int free_sms_env() {
  static const int v = []() {
    const char* e = getenv("EQH_EDGE_BWD_FREE_SMS");
    const int x = e ? atoi(e) : -1;
    return x > kSMs / 2 ? kSMs / 2 : x;
  }();
  return v;
}
static int slices_of(int64_t H) {
  const int s = (kSMs - (free_sms_env() > 0 ? free_sms_env() : 0)) / groups_of(H);
  return s < 1 ? 1 : s;
}
static int slices_for(int64_t H, int64_t N) {
  if (free_sms_env() >= 0) return slices_of(H);
  const int s = (kSMs - (N < 8192 ? 8 : 0)) / groups_of(H);
  return s < 1 ? 1 : s;
}

}  // namespace edge_bwd_tc
}  // namespace eqh

extern "C" {

int eqh_egnn_edge_bwd_tc_rows(int64_t H) { return H > 0 ? eqh::edge_bwd_tc::slices_of(H) : 0; }
int eqh_egnn_edge_bwd_tc_slices(int64_t H, int64_t N) { return H > 0 ? eqh::edge_bwd_tc::slices_for(H, N) : 0; }

int eqh_egnn_edge_bwd_tc(const float* P, const float* Q, const float* wc, const float* W2, const int32_t* idx,
                         const float* d2, const float* S2, const float* dMi, int64_t N, int64_t H, float* dP, float* dPre,
                         float* part, const int32_t* order, eqh_stream_t stream) {
  return eqh_egnn_edge_bwd_tc_ld(P, Q, H, wc, W2, idx, d2, S2, dMi, EQH_M_DIM, N, H, dP, H, dPre, part, order, stream);
}

// Strided variant (see eqh_egnn_edge_fwd_tc_ld): P / Q row stride ldpq, dP written with row stride lddp (a column block of
// the merged [dP | dQ] matrix), dMi read with row stride lddmi (a column block of the node MLP's input gradient).
int eqh_egnn_edge_bwd_tc_ld(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2,
                            const int32_t* idx, const float* d2, const float* S2, const float* dMi, int64_t lddmi, int64_t N,
                            int64_t H, float* dP, int64_t lddp, float* dPre, float* part, const int32_t* order,
                            eqh_stream_t stream) {
  return eqh_egnn_edge_bwd_tc_slots(P, Q, ldpq, wc, W2, idx, d2, S2, dMi, lddmi, N, H, dP, lddp, dPre, part, order, nullptr,
                                    stream);
}

// Slot numbering for the slot mode below (see k_edge_bwd_tc).  keys = N * eqh_egnn_edge_slot_keys(H) pairs (neighbour j,
// sub-slice u), j-major.  Pass 1: `mark` [keys] (filled with INT32_MAX by the caller) <- first edge of every pair; the
// caller turns [mark != INT32_MAX] into its exclusive prefix sum `slot_ext` [keys + 1]; pass 2: `slot_edge` [N*16].
// slot_ext[j * U] (U = keys per j) is then the row pointer of the final segmented sum dQ[j] = sum of the rows of j.
int eqh_egnn_edge_slot_keys(int64_t H) { return H > 0 ? eqh::edge_bwd_tc::slices_of(H) * eqh::edge_bwd_tc::kEpiWG : 0; }

int eqh_egnn_edge_slots_mark(const int32_t* idx, const int32_t* order, int64_t N, int64_t H, int32_t* mark,
                             eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::edge_bwd_tc;
  EQH_CHECK_ARG(idx && mark && N >= 0 && H > 0 && N * (int64_t)slices_of(H) * kEpiWG < ((int64_t)1 << 31));
  if (N == 0) return EQH_OK;
  launch_kernel(k_slots_mark, dim3((unsigned)((N * kNbr + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), idx,
                order, N, slices_for(H, N), slices_of(H) * kEpiWG, mark);
  EQH_LAUNCH_RC();
}

int eqh_egnn_edge_slots_assign(const int32_t* idx, const int32_t* order, int64_t N, int64_t H, const int32_t* mark,
                               const int32_t* slot_ext, int32_t* slot_edge, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::edge_bwd_tc;
  EQH_CHECK_ARG(idx && mark && slot_ext && slot_edge && N >= 0 && H > 0);
  if (N == 0) return EQH_OK;
  launch_kernel(k_slots_assign, dim3((unsigned)((N * kNbr + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), idx,
                order, N, slices_for(H, N), slices_of(H) * kEpiWG, mark, slot_ext, slot_edge);
  EQH_LAUNCH_RC();
}

// Slot mode: `slot_edge` [N*16] from eqh_egnn_edge_slots (NULL: one dPre row per edge, as above); dPre then receives one row
// per (sub-slice, neighbour) pair -- at most N*16 rows, numbered j-major.
int eqh_egnn_edge_bwd_tc_slots(const float* P, const float* Q, int64_t ldpq, const float* wc, const float* W2,
                               const int32_t* idx, const float* d2, const float* S2, const float* dMi, int64_t lddmi,
                               int64_t N, int64_t H, float* dP, int64_t lddp, float* dPre, float* part, const int32_t* order,
                               const int32_t* slot_edge, eqh_stream_t stream) {
  using namespace eqh;
  using namespace eqh::edge_bwd_tc;
  EQH_CHECK_ARG(P && Q && wc && W2 && idx && d2 && S2 && dMi && dP && dPre && part && N >= 0);
  EQH_CHECK_ARG(aligned16(S2) && aligned16(dMi) && aligned16(P) && aligned16(Q));
  EQH_CHECK_ARG(ldpq >= H && ldpq % 4 == 0 && lddp >= H && lddmi >= kM && lddmi % 4 == 0);
  if (H <= 0 || H % 64 != 0 || groups_of(H) > kSMs) return EQH_ERR_UNSUPPORTED;
  const int gr = groups_of(H), sl = slices_for(H, N);
  const int64_t part_cols = (int64_t)(kM + 1) * H + kM;
  if (sl < slices_of(H)) {       // `part` has slices_of(H) rows: the ones this launch does not write hold zeros
    cudaError_t e = cudaMemsetAsync(part + (int64_t)sl * part_cols, 0, sizeof(float) * (size_t)(slices_of(H) - sl) * (size_t)part_cols,
                                    static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return (int)e;
  }
  if (slot_edge) {
    cudaError_t e = cudaFuncSetAttribute(k_edge_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return (int)e;
    launch_kernel(k_edge_bwd_tc<true>, dim3((unsigned)(gr * sl)), dim3(kThreads), (size_t)kSmem, static_cast<cudaStream_t>(stream),
                  P, Q, wc, W2, idx, d2, S2, dMi, N, (int)H, gr, sl, dP, dPre, part, part_cols, order, ldpq, lddp, lddmi, slot_edge);
  } else {
    cudaError_t e = cudaFuncSetAttribute(k_edge_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return (int)e;
    launch_kernel(k_edge_bwd_tc<false>, dim3((unsigned)(gr * sl)), dim3(kThreads), (size_t)kSmem, static_cast<cudaStream_t>(stream),
                  P, Q, wc, W2, idx, d2, S2, dMi, N, (int)H, gr, sl, dP, dPre, part, part_cols, order, ldpq, lddp, lddmi, slot_edge);
  }
  EQH_LAUNCH_RC();
}

}  // extern "C"

#ifdef EQH_BWDTC_TRACE
extern "C" EQH_API int eqh_debug_edge_bwd_tc_trace(unsigned long long* host64) {
  return (int)cudaMemcpyFromSymbol(host64, eqh::edge_bwd_tc::g_bt_trace, sizeof(unsigned long long) * 64);
}
#endif
