// Device-side batch collation (replaces the single-process PyG `Batch.from_data_list` +
// `.to(device)` of main.py:227-229 for a dataset that is resident in HBM as a packed per-molecule store).
// Integer / byte work, bit-exact against the host collation: global atom ids are local ids + the atom offset of the
// molecule inside the batch, global hyperedge ids local ids + the hyperedge offset (HData.__inc__,
// data/utils.py:172-178), `batch` / `ptr` as PyG builds them, and -- for the CUDA-graph trainer -- the inert ghost
// molecule that pads the batch to its shape bucket (graphed.pad_batch).  One launch; every output element is written
// by exactly one thread.
#include "common.cuh"

namespace eqh {

// largest b with off[b] <= t   (off has nb+1 ascending entries, off[0] = 0)
__device__ __forceinline__ int find_mol(const int32_t* __restrict__ off, int nb, int64_t t) {
  int lo = 0, hi = nb;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)__ldg(off + mid) <= t) lo = mid; else hi = mid;
  }
  return lo;
}

struct CollateArgs {
  eqh_store st;
  const int32_t* ids;        // [B] molecule ids
  const int32_t* off;        // [3][B+1]: atom / hyperedge / incidence offsets of the selected molecules in the batch
  int32_t B;
  int64_t N, E, Z;           // real totals
  int64_t Np, Ep, Zp;        // padded totals (== real totals: no ghost molecule)
  int32_t side;              // ghost lattice side
  float ghost_off;           // ghost lattice origin
  eqh_batch out;
};

__global__ void __launch_bounds__(256) k_collate(const __grid_constant__ CollateArgs a) {
  pdl_prologue();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int B = a.B;
  const int32_t* aoff = a.off;
  const int32_t* hoff = a.off + (B + 1);
  const int32_t* zoff = a.off + 2 * (B + 1);
  const bool ghost = a.Np > a.N;
  const int64_t ga = a.Np - a.N, ge = a.Ep - a.E, gz = a.Zp - a.Z;
  // ---- atoms: x [Np,9], pos [Np,3], batch [Np]
  if (t < a.Np) {
    if (t < a.N) {
      const int b = find_mol(aoff, B, t);
      const int64_t src = __ldg(a.st.atom_off + __ldg(a.ids + b)) + (t - __ldg(aoff + b));
#pragma unroll
      for (int f = 0; f < 9; ++f) a.out.x[t * 9 + f] = __ldg(a.st.x + src * 9 + f);
#pragma unroll
      for (int d = 0; d < 3; ++d) a.out.pos[t * 3 + d] = __ldg(a.st.pos + src * 3 + d);
      a.out.batch[t] = b;
    } else {
      const int64_t g = t - a.N;
      const int64_t s = a.side;
#pragma unroll
      for (int f = 0; f < 9; ++f) a.out.x[t * 9 + f] = 0;
      a.out.pos[t * 3 + 0] = __fadd_rn(__fmul_rn((float)(g % s), 1.5f), a.ghost_off);
      a.out.pos[t * 3 + 1] = __fadd_rn(__fmul_rn((float)((g / s) % s), 1.5f), a.ghost_off);
      a.out.pos[t * 3 + 2] = __fadd_rn(__fmul_rn((float)(g / (s * s)), 1.5f), a.ghost_off);
      a.out.batch[t] = B;
    }
  }
  // ---- incidences: edge_index0 / edge_index1 [Zp]
  if (t < a.Zp) {
    if (t < a.Z) {
      const int b = find_mol(zoff, B, t);
      const int64_t src = __ldg(a.st.inc_off + __ldg(a.ids + b)) + (t - __ldg(zoff + b));
      a.out.edge_index0[t] = (int64_t)__ldg(a.st.v_loc + src) + __ldg(aoff + b);
      a.out.edge_index1[t] = (int64_t)__ldg(a.st.e_loc + src) + __ldg(hoff + b);
    } else {
      const int64_t zi = t - a.Z;
      a.out.edge_index0[t] = a.N + zi % ga;
      a.out.edge_index1[t] = a.E + zi % ge;
    }
  }
  // ---- hyperedges: edge_attr [Ep], e_order [Ep]
  if (t < a.Ep) {
    if (t < a.E) {
      const int b = find_mol(hoff, B, t);
      const int64_t src = __ldg(a.st.he_off + __ldg(a.ids + b)) + (t - __ldg(hoff + b));
      a.out.edge_attr[t] = __ldg(a.st.edge_attr + src);
      a.out.e_order[t] = __ldg(a.st.e_order + src);
    } else {
      const int64_t h = t - a.E;
      a.out.edge_attr[t] = 5;                                        // data/utils.py:145: conjugated hyperedge
      a.out.e_order[t] = gz / ge + (h < gz % ge ? 1 : 0);
    }
  }
  // ---- per molecule: n_e, y, ptr
  const int G = B + (ghost ? 1 : 0);
  if (t < G) {
    if (t < B) {
      a.out.n_e[t] = (int64_t)(__ldg(hoff + t + 1) - __ldg(hoff + t));
      a.out.y[t] = __ldg(a.st.y + __ldg(a.ids + t));
    } else {
      a.out.n_e[t] = ge;
      a.out.y[t] = 0.f;
    }
  }
  if (t <= G) a.out.ptr[t] = t <= B ? (int64_t)__ldg(aoff + t) : a.Np;
}

struct CollateCsrArgs {
  eqh_store st;
  eqh_store_csr csr;
  const int32_t* ids;
  const int32_t* off;
  int32_t B;
  int64_t N, E, Z, Np, Ep, Zp;
  eqh_incidence_out out;
};

// ghost segments: incidence zi of the gz ghost incidences belongs to segment zi % n_seg, so segment s holds
// zi = s, s + n_seg, ...: q + (s < r) of them (q = gz / n_seg, r = gz % n_seg) and starts at s * q + min(s, r)
__device__ __forceinline__ int64_t ghost_start(int64_t s, int64_t q, int64_t r) { return s * q + (s < r ? s : r); }
// sorted position p (0 <= p < gz) -> the incidence zi that lands there
__device__ __forceinline__ int64_t ghost_source(int64_t p, int64_t q, int64_t r, int64_t n_seg) {
  int64_t s, j;
  if (p < r * (q + 1)) { s = p / (q + 1); j = p - s * (q + 1); }
  else { const int64_t u = p - r * (q + 1); s = r + u / q; j = u - (u / q) * q; }     // (q >= 1 here: p >= r (q + 1) and p < gz)
  return s + j * n_seg;
}

__global__ void __launch_bounds__(256) k_collate_csr(const __grid_constant__ CollateCsrArgs a) {
  pdl_prologue();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int B = a.B;
  const int32_t* aoff = a.off;
  const int32_t* hoff = a.off + (B + 1);
  const int32_t* zoff = a.off + 2 * (B + 1);
  const int64_t ga = a.Np - a.N, ge = a.Ep - a.E, gz = a.Zp - a.Z;
  // ---- per incidence: both sorted orders and the int32 index lists
  if (t < a.Zp) {
    if (t < a.Z) {
      const int b = find_mol(zoff, B, t);
      const int32_t z0 = __ldg(zoff + b), a0 = __ldg(aoff + b), h0 = __ldg(hoff + b);
      const int64_t src = __ldg(a.st.inc_off + __ldg(a.ids + b));
      const int32_t k = (int32_t)(t - z0);
      const int32_t pe = __ldg(a.csr.perm_e_loc + src + k), pv = __ldg(a.csr.perm_v_loc + src + k);
      a.out.perm_e[t] = z0 + pe;
      a.out.vtx_by_e[t] = a0 + __ldg(a.st.v_loc + src + pe);
      a.out.perm_v[t] = z0 + pv;
      a.out.edge_by_v[t] = h0 + __ldg(a.st.e_loc + src + pv);
      a.out.vertex32[t] = a0 + __ldg(a.st.v_loc + src + k);
      a.out.edges32[t] = h0 + __ldg(a.st.e_loc + src + k);
    } else {
      const int64_t zi = t - a.Z;
      a.out.vertex32[t] = (int32_t)(a.N + zi % ga);
      a.out.edges32[t] = (int32_t)(a.E + zi % ge);
      const int64_t se = ghost_source(zi, gz / ge, gz % ge, ge);
      a.out.perm_e[t] = (int32_t)(a.Z + se);
      a.out.vtx_by_e[t] = (int32_t)(a.N + se % ga);
      const int64_t sv = ghost_source(zi, gz / ga, gz % ga, ga);
      a.out.perm_v[t] = (int32_t)(a.Z + sv);
      a.out.edge_by_v[t] = (int32_t)(a.E + sv % ge);
    }
  }
  // ---- row pointers
  if (t <= a.Ep) {
    if (t < a.E) {
      const int b = find_mol(hoff, B, t);
      const int64_t m = __ldg(a.ids + b);
      a.out.rowptr_e[t] = __ldg(zoff + b) + __ldg(a.csr.rowptr_e_loc + __ldg(a.st.he_off + m) + m + (t - __ldg(hoff + b)));
    } else {
      a.out.rowptr_e[t] = (int32_t)(a.Z + (ge > 0 ? ghost_start(t - a.E, gz / ge, gz % ge) : 0));
    }
  }
  if (t <= a.Np) {
    if (t < a.N) {
      const int b = find_mol(aoff, B, t);
      const int64_t m = __ldg(a.ids + b);
      a.out.rowptr_v[t] = __ldg(zoff + b) + __ldg(a.csr.rowptr_v_loc + __ldg(a.st.atom_off + m) + m + (t - __ldg(aoff + b)));
    } else {
      a.out.rowptr_v[t] = (int32_t)(a.Z + (ga > 0 ? ghost_start(t - a.N, gz / ga, gz % ga) : 0));
    }
  }
}

}  // namespace eqh

extern "C" int eqh_collate_incidence(const eqh_store* store, const eqh_store_csr* csr, const int32_t* ids, const int32_t* offsets,
                                     int64_t B, int64_t N, int64_t E, int64_t Z, int64_t N_pad, int64_t E_pad, int64_t Z_pad,
                                     const eqh_incidence_out* out, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(store && csr && ids && offsets && out && B > 0 && B < (1 << 24) && N >= 0 && E >= 0 && Z >= 0);
  EQH_CHECK_ARG(N_pad >= N && E_pad >= E && Z_pad >= Z && N_pad < ((int64_t)1 << 31) && Z_pad < ((int64_t)1 << 31));
  const bool ghost = N_pad > N;
  EQH_CHECK_ARG(ghost ? (E_pad > E) : (E_pad == E && Z_pad == Z));
  EQH_CHECK_ARG(csr->rowptr_e_loc && csr->perm_e_loc && csr->rowptr_v_loc && csr->perm_v_loc);
  EQH_CHECK_ARG(out->rowptr_e && out->perm_e && out->vtx_by_e && out->rowptr_v && out->perm_v && out->edge_by_v &&
                out->vertex32 && out->edges32);
  CollateCsrArgs a;
  a.st = *store; a.csr = *csr; a.ids = ids; a.off = offsets; a.B = (int32_t)B;
  a.N = N; a.E = E; a.Z = Z; a.Np = N_pad; a.Ep = E_pad; a.Zp = Z_pad; a.out = *out;
  int64_t m = N_pad + 1 > Z_pad ? N_pad + 1 : Z_pad;
  if (E_pad + 1 > m) m = E_pad + 1;
  launch_kernel(k_collate_csr, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
  EQH_LAUNCH_RC();
}

extern "C" int eqh_collate(const eqh_store* store, const int32_t* ids, const int32_t* offsets, int64_t B, int64_t N, int64_t E,
                           int64_t Z, int64_t N_pad, int64_t E_pad, int64_t Z_pad, int64_t ghost_side, float ghost_origin,
                           const eqh_batch* out, eqh_stream_t stream) {
  using namespace eqh;
  EQH_CHECK_ARG(store && ids && offsets && out && B > 0 && B < (1 << 24) && N >= 0 && E >= 0 && Z >= 0);
  EQH_CHECK_ARG(N_pad >= N && E_pad >= E && Z_pad >= Z && N_pad < ((int64_t)1 << 31) && Z_pad < ((int64_t)1 << 31));
  const bool ghost = N_pad > N;
  // a ghost molecule needs at least one atom and one hyperedge (ghost incidences attach ghost atoms to ghost hyperedges)
  EQH_CHECK_ARG(ghost ? (E_pad > E && ghost_side >= 1) : (E_pad == E && Z_pad == Z));
  EQH_CHECK_ARG(out->x && out->pos && out->edge_index0 && out->edge_index1 && out->edge_attr && out->n_e && out->e_order &&
                out->batch && out->ptr && out->y);
  CollateArgs a;
  a.st = *store; a.ids = ids; a.off = offsets; a.B = (int32_t)B;
  a.N = N; a.E = E; a.Z = Z; a.Np = N_pad; a.Ep = E_pad; a.Zp = Z_pad;
  a.side = (int32_t)ghost_side; a.ghost_off = ghost_origin; a.out = *out;
  int64_t m = N_pad > Z_pad ? N_pad : Z_pad;
  if (E_pad > m) m = E_pad;
  if (B + 2 > m) m = B + 2;
  launch_kernel(k_collate, dim3((unsigned)((m + 255) / 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), a);
  EQH_LAUNCH_RC();
}
