"""Device-side batch collation: the reference collates on the host in a single process
(`DataLoader(num_workers=0)` + PyG `Batch.from_data_list`, main.py:227-229, data/utils.py:150-178) and ships the collated
batch to the GPU every step.  Here the dataset is resident in HBM as a packed per-molecule store and a batch is
gathered on the device from the list of molecule ids (`eqh_collate`): per step the host sends the B ids and 3(B+1)
offsets it computes from its copy of the per-molecule counts (~2 KB) instead of the collated batch (~350 KB at QM9-128),
and the result lands directly in the static inputs of the captured training graph, ghost molecule included.

Bit-exact against the host collation (`synthetic.collate_host` + `graphed.pad_batch`); tests/test_gpu_collate.py.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .graphed import GHOST_OFFSET, GraphedTrainer
from .synthetic import HBatch


def local_csr(mols) -> dict:
    """Per-molecule LOCAL incidence CSR in both groupings, concatenated over the molecules (int32 numpy arrays):
    rowptr_e_loc (n_he + 1 entries per molecule), perm_e_loc (the molecule's incidences in by-hyperedge order, stable),
    rowptr_v_loc (n_atoms + 1 per molecule), perm_v_loc.  mols: tuples (n_atoms, n_he, x, pos, v_loc, e_loc, ...)."""
    rp_e, pm_e, rp_v, pm_v = [], [], [], []
    for m in mols:
        v_loc, e_loc = np.asarray(m[4], dtype=np.int64), np.asarray(m[5], dtype=np.int64)
        pm_e.append(np.argsort(e_loc, kind="stable"))
        pm_v.append(np.argsort(v_loc, kind="stable"))
        rp_e.append(np.concatenate([[0], np.cumsum(np.bincount(e_loc, minlength=m[1]))]))
        rp_v.append(np.concatenate([[0], np.cumsum(np.bincount(v_loc, minlength=m[0]))]))
    cat32 = lambda parts: np.ascontiguousarray(np.concatenate(parts).astype(np.int32))
    return {"rowptr_e_loc": cat32(rp_e), "perm_e_loc": cat32(pm_e), "rowptr_v_loc": cat32(rp_v), "perm_v_loc": cat32(pm_v)}


class MoleculeStore:
    """Packed dataset: concatenated per-molecule arrays on the device + per-molecule counts on the host."""

    def __init__(self, mols, y, device):
        self.device = torch.device(device)
        atoms = np.array([m[0] for m in mols], dtype=np.int64)
        hes = np.array([m[1] for m in mols], dtype=np.int64)
        incs = np.array([m[4].shape[0] for m in mols], dtype=np.int64)
        self.n_atoms, self.n_he, self.n_inc = atoms, hes, incs
        self.max_abs_pos = np.array([np.abs(m[3]).max() for m in mols], dtype=np.float32)
        off = lambda c: np.concatenate([[0], np.cumsum(c)]).astype(np.int64)
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.t = {
            "x": dev(np.concatenate([m[2] for m in mols])), "pos": dev(np.concatenate([m[3] for m in mols])),
            "y": dev(np.asarray(y, dtype=np.float32)),
            "v_loc": dev(np.concatenate([m[4] for m in mols]).astype(np.int32)),
            "e_loc": dev(np.concatenate([m[5] for m in mols]).astype(np.int32)),
            "edge_attr": dev(np.concatenate([m[6] for m in mols])), "e_order": dev(np.concatenate([m[7] for m in mols])),
            "atom_off": dev(off(atoms)), "he_off": dev(off(hes)), "inc_off": dev(off(incs)),
        }
        self.c_store = _lib.Store(**{k: v.data_ptr() for k, v in self.t.items()})
        self.size = len(mols)
        # per-molecule LOCAL incidence CSR in both groupings (stable orders, computed once per dataset): the collation
        # emits the batch's CSR from these by gather + offset (eqh_collate_incidence) instead of rebuilding it every step
        self.csr = {k: dev(v) for k, v in local_csr(mols).items()}
        self.c_csr = _lib.StoreCsr(**{k: v.data_ptr() for k, v in self.csr.items()})

    def shape_of(self, ids: Sequence[int]) -> Tuple[int, int, int]:
        ids = np.asarray(ids, dtype=np.int64)
        return int(self.n_atoms[ids].sum()), int(self.n_he[ids].sum()), int(self.n_inc[ids].sum())

    def step_inputs(self, ids: Sequence[int]) -> torch.Tensor:
        """Pinned int32 buffer [B + 3(B+1)]: the ids, then the atom / hyperedge / incidence offsets inside the batch."""
        ids = np.asarray(ids, dtype=np.int64)
        B = ids.shape[0]
        buf = torch.empty(B + 3 * (B + 1), dtype=torch.int32).pin_memory()
        a = buf.numpy()
        a[:B] = ids
        for k, cnt in enumerate((self.n_atoms, self.n_he, self.n_inc)):
            a[B + k * (B + 1)] = 0
            np.cumsum(cnt[ids], out=a[B + k * (B + 1) + 1:B + (k + 1) * (B + 1)])
        return buf


def empty_packed(B: int, bucket: Optional[Tuple[int, int, int]], shape: Tuple[int, int, int], device) -> HBatch:
    """Uninitialised packed batch of the given (padded) shape on the device."""
    N, E, Z = bucket if bucket is not None else shape
    G = B + (1 if bucket is not None else 0)
    i64, f32 = torch.int64, torch.float32
    like = HBatch(x=torch.empty((N, 9), dtype=i64), pos=torch.empty((N, 3), dtype=f32), edge_index0=torch.empty(Z, dtype=i64),
                  edge_index1=torch.empty(Z, dtype=i64), edge_attr=torch.empty((E, 1), dtype=i64), n_e=torch.empty(G, dtype=i64),
                  e_order=torch.empty(E, dtype=i64), batch=torch.empty(N, dtype=i64), ptr=torch.empty(G + 1, dtype=i64),
                  y=torch.empty(G, dtype=f32), num_graphs=G, num_real_graphs=B if bucket is not None else 0)
    return like.pack(device=device)


def empty_incidence(shape: Tuple[int, int, int], device) -> ops.Incidence:
    """Uninitialised incidence CSR buffers of a batch with (N, E, Z) atoms / hyperedges / incidences: views into one buffer."""
    N, E, Z = shape
    sizes = [("rowptr_e", E + 1), ("perm_e", Z), ("vtx_by_e", Z), ("rowptr_v", N + 1), ("perm_v", Z), ("edge_by_v", Z),
             ("vertex32", Z), ("edges32", Z)]
    total, offs = 0, []
    for _, n in sizes:
        offs.append(total)
        total += (n + 3) // 4 * 4               # 16-byte aligned fields
    flat = torch.empty(total, dtype=torch.int32, device=device)
    f = {name: flat[o:o + n] for (name, n), o in zip(sizes, offs)}
    inc = ops.Incidence(N, E, Z, f["rowptr_e"], f["perm_e"], f["vtx_by_e"], f["rowptr_v"], f["perm_v"], f["edge_by_v"],
                        f["vertex32"], f["edges32"])
    inc._flat = flat
    return inc


def collate_device(store: MoleculeStore, ids: Sequence[int], bucket: Optional[Tuple[int, int, int]] = None,
                   out: Optional[HBatch] = None, step_inputs: Optional[torch.Tensor] = None,
                   with_incidence: bool = True) -> HBatch:
    """Gather the molecules `ids` into one batch on the device.  bucket=(N_pad, E_pad, Z_pad) appends the ghost molecule
    (graphed.pad_batch); `out` is a packed batch of the right shape to write into (e.g. a captured graph's static
    inputs); `step_inputs` a pinned buffer from store.step_inputs(ids) (built here when omitted).
    with_incidence: also emit the batch's incidence CSR (`out.incidence`, both groupings) from the store's per-molecule
    CSRs -- the model wrappers then skip the per-step build (ops.incidence_of)."""
    B = len(ids)
    N, E, Z = store.shape_of(ids)
    Np, Ep, Zp = bucket if bucket is not None else (N, E, Z)
    if bucket is not None and (Np - N < 1 or Ep - E < 1 or Zp - Z < 0):
        raise ValueError(f"batch (N={N},E={E},Z={Z}) does not fit bucket {bucket}")
    if out is None:
        out = empty_packed(B, bucket, (N, E, Z), store.device)
    if step_inputs is None:
        step_inputs = store.step_inputs(ids)
    dev_in = step_inputs.to(store.device, non_blocking=True)
    ga = Np - N
    side = int(np.ceil(ga ** (1.0 / 3.0))) if ga > 0 else 1
    ids_np = np.asarray(ids, dtype=np.int64)
    origin = float(np.float32(GHOST_OFFSET) + store.max_abs_pos[ids_np].max()) if ga > 0 else 0.0
    c_out = _lib.Batch(**{f: getattr(out, f).data_ptr() for f in ("x", "pos", "edge_index0", "edge_index1", "edge_attr",
                                                                  "n_e", "e_order", "batch", "ptr", "y")})
    import ctypes
    ops._call("eqh_collate", 1, ctypes.byref(store.c_store), dev_in.data_ptr(), dev_in.data_ptr() + 4 * B, B, N, E, Z, Np, Ep,
              Zp, side, origin, ctypes.byref(c_out), ops._stream())
    if with_incidence:
        inc = getattr(out, "incidence", None)
        if inc is None or (inc.N, inc.E, inc.Z) != (Np, Ep, Zp):
            inc = out.incidence = empty_incidence((Np, Ep, Zp), store.device)
        c_inc = _lib.IncidenceOut(**{f: getattr(inc, f).data_ptr() for f in ("rowptr_e", "perm_e", "vtx_by_e", "rowptr_v",
                                                                             "perm_v", "edge_by_v", "vertex32", "edges32")})
        ops._call("eqh_collate_incidence", 1, ctypes.byref(store.c_store), ctypes.byref(store.c_csr), dev_in.data_ptr(),
                  dev_in.data_ptr() + 4 * B, B, N, E, Z, Np, Ep, Zp, ctypes.byref(c_inc), ops._stream())
    out._keep = dev_in          # the kernels read it asynchronously
    return out


class StoreTrainer(GraphedTrainer):
    """GraphedTrainer fed by molecule ids: collation happens on the device, straight into the graph's static inputs.
    fixed_bucket: pad every batch to this (N_pad, E_pad, Z_pad) instead of its own bucket (uniform shapes over ranks)."""

    def __init__(self, model, store: MoleculeStore, real_graphs: int, fixed_bucket: Optional[Tuple[int, int, int]] = None,
                 with_incidence: bool = True, **kw):
        super().__init__(model, real_graphs, **kw)
        self.store = store
        self.fixed_bucket = fixed_bucket
        self.with_incidence = with_incidence      # the collation emits the incidence CSR: no build kernels in the graph

    def step_from_ids(self, ids: Sequence[int], step_inputs: Optional[torch.Tensor] = None, sync: bool = True):
        shape = self.store.shape_of(ids)
        N, E, Z = shape
        bucket = self.fixed_bucket or ((N + 1 + 255) // 256 * 256, (E + 1 + 255) // 256 * 256, (Z + 511) // 512 * 512)
        key = (bucket[0], bucket[1], bucket[2], len(ids) + 1)
        bk = self.buckets.get(key)
        if bk is None:
            like = collate_device(self.store, ids, bucket, with_incidence=self.with_incidence)
            bk = self.buckets[key] = self._capture(like)
        collate_device(self.store, ids, bucket, out=bk.static, step_inputs=step_inputs, with_incidence=self.with_incidence)
        bk.graph.replay()
        self.replays += 1
        self.own_launches += bk.launches + (2 if self.with_incidence else 1)
        if self.in_graph_step:
            self.t += 1
        self._finish_outside()
        self._queue_loss(bk.loss)
        return self._read_loss(1) if sync else self._read_loss(2)
