"""Seeded synthetic molecular-hypergraph batches (SURVEY.md 8(d) profiles).

There is no network, no rdkit and no dataset in this image, so the throughput runs use
hypergraphs with the *shape* of the reference's datasets.  Per molecule the incidence list is
laid out exactly as `mol2hgraph` does (reference data/utils.py:122-145): for bond i the pairs
(begin, i), (end, i), then conjugated groups as (atom, num_bond + g) in ascending atom order
with bond-type feature 5; on collation atom ids are offset by the atoms so far and hyperedge
ids by the hyperedges so far (HData.__inc__, data/utils.py:172-178).

Profiles are stated assumptions taken from the public dataset descriptions (H-inclusive):
Positions: random tree walk, bond length ~N(1.5, 0.075) A clipped to [1.2, 1.8], centred per molecule
(molecules overlap near the origin, so kNN edges cross molecules as they do in the reference).
  qm9        a~round(N(18,3))  clip[4,29],    r~Pois(1.6), h~Bern(0.25), s~U{3..8}
  pcqm4mv2   a~round(N(29.5,7)) clip[6,51],   r~Pois(1.1), h~Pois(0.9),  s~U{4..14}
  molecule3d a~round(N(29,8))  clip[5,60],    r~Pois(1.4), h~Pois(0.8),  s~U{4..14}
  opv        a~round(N(100,30)) clip[40,200], r~Pois(8),   h~U{1..3},    s~U{15..70}
"""
from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Optional

import numpy as np
import torch

ATOM_FEATURE_DIMS = (119, 5, 12, 12, 10, 6, 6, 2, 2)  # ogb 1.3.6 get_atom_feature_dims()

PROFILES = {
    #            mean  std  lo  hi   rings  conj-kind  conj-par  s_lo s_hi
    "qm9":        (18.0, 3.0, 4, 29, 1.6, "bern", 0.25, 3, 8),
    "pcqm4mv2":   (29.5, 7.0, 6, 51, 1.1, "pois", 0.9, 4, 14),
    "molecule3d": (29.0, 8.0, 5, 60, 1.4, "pois", 0.8, 4, 14),
    "opv":        (100.0, 30.0, 40, 200, 8.0, "unif13", 0.0, 15, 70),
}


@dataclass
class HBatch:
    """The collated tensors the model wrappers read (what a PyG `Batch` of `HData` exposes)."""
    x: torch.Tensor            # [N, 9] int64 categorical atom features
    pos: torch.Tensor          # [N, 3] float32
    edge_index0: torch.Tensor  # [Z] int64 atom id per incidence
    edge_index1: torch.Tensor  # [Z] int64 hyperedge id per incidence
    edge_attr: torch.Tensor    # [E, 1] int64 bond type (0..3) / 5 for conjugated hyperedges
    n_e: torch.Tensor          # [B] int64 hyperedges per molecule
    e_order: torch.Tensor      # [E] int64 size of each hyperedge
    batch: torch.Tensor        # [N] int64 molecule id per atom (sorted)
    ptr: torch.Tensor          # [B+1] int64 atom offsets
    y: torch.Tensor            # [B] float32
    num_graphs: int = 0
    num_real_graphs: int = 0   # > 0 on ghost-padded batches (graphed.pad_batch): molecules [0, num_real_graphs) are real

    @property
    def num_atoms(self) -> int:
        return int(self.x.shape[0])

    @property
    def num_hyperedges(self) -> int:
        return int(self.edge_attr.shape[0])

    @property
    def num_incidences(self) -> int:
        return int(self.edge_index0.shape[0])

    def to(self, device, non_blocking: bool = False) -> "HBatch":
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
        return HBatch(**kw)

    def pin_memory(self) -> "HBatch":
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            kw[f.name] = v.pin_memory() if isinstance(v, torch.Tensor) else v
        return HBatch(**kw)

    def pack(self, device=None, pin: bool = False) -> "HBatch":
        """The same batch with every tensor a view into ONE contiguous buffer (16-byte aligned fields, fixed order), so
        that a whole batch moves with a single copy: pinned host -> device, or device -> the static inputs of a captured
        graph.  Batches of equal shapes have equal layouts.  The buffer is kept as `._flat`."""
        names = [f.name for f in fields(self) if isinstance(getattr(self, f.name), torch.Tensor)]
        offs, total = [], 0
        for n in names:
            t = getattr(self, n)
            total = (total + 15) // 16 * 16
            offs.append(total)
            total += t.numel() * t.element_size()
        flat = torch.empty(total, dtype=torch.uint8, device=device if device is not None else "cpu",
                           pin_memory=bool(pin and device is None))
        kw = {}
        for n, o in zip(names, offs):
            t = getattr(self, n)
            v = flat[o:o + t.numel() * t.element_size()].view(t.dtype).view(t.shape)
            v.copy_(t)
            kw[n] = v
        out = HBatch(**kw, num_graphs=self.num_graphs, num_real_graphs=self.num_real_graphs)
        out._flat = flat
        return out

    def h2d_bytes(self) -> int:
        return sum(getattr(self, f.name).numel() * getattr(self, f.name).element_size()
                   for f in fields(self) if isinstance(getattr(self, f.name), torch.Tensor))


def _molecule(rng: np.random.Generator, prof) -> tuple:
    mean, std, lo, hi, rings, ckind, cpar, s_lo, s_hi = prof
    a = int(np.clip(np.rint(rng.normal(mean, std)), lo, hi))
    # spanning tree: atom k bonds to a random earlier atom; positions follow the tree, step 1.5 A
    parent = np.array([rng.integers(0, k) for k in range(1, a)], dtype=np.int64)
    pos = np.zeros((a, 3), dtype=np.float64)
    for k in range(1, a):
        for _ in range(8):
            d = rng.normal(size=3)
            # bond length 1.5 A +- 5 % (continuous, so squared distances are tie-free in fp32)
            d *= float(np.clip(rng.normal(1.5, 0.075), 1.2, 1.8)) / np.linalg.norm(d)
            cand = pos[parent[k - 1]] + d
            if k < 2 or np.min(np.sum((pos[:k] - cand) ** 2, axis=1)) > 1.0:
                break
        pos[k] = cand
    pos -= pos.mean(axis=0, keepdims=True)
    begin = list(parent)
    end = list(range(1, a))
    have = set(zip(begin, end))
    for _ in range(int(rng.poisson(rings))):
        if a < 3:
            break
        u, v = sorted(rng.choice(a, size=2, replace=False).tolist())
        if (u, v) not in have:
            have.add((u, v))
            begin.append(u)
            end.append(v)
    nb = len(begin)
    n_idx = np.empty(2 * nb, dtype=np.int64)
    n_idx[0::2] = begin
    n_idx[1::2] = end
    e_idx = np.repeat(np.arange(nb, dtype=np.int64), 2)
    battr = rng.integers(0, 4, size=nb).astype(np.int64)
    if ckind == "bern":
        h = int(rng.random() < cpar)
    elif ckind == "pois":
        h = int(rng.poisson(cpar))
    else:
        h = int(rng.integers(1, 4))
    order = [2] * nb
    n_parts, e_parts, c_attr = [n_idx], [e_idx], []
    g = 0
    for _ in range(h):
        s = int(min(rng.integers(s_lo, s_hi + 1), a))
        if s < 2:
            continue
        members = np.sort(rng.choice(a, size=s, replace=False)).astype(np.int64)
        n_parts.append(members)
        e_parts.append(np.full(s, nb + g, dtype=np.int64))
        c_attr.append(5)
        order.append(s)
        g += 1
    feats = np.stack([rng.integers(0, d, size=a) for d in ATOM_FEATURE_DIMS], axis=1).astype(np.int64)
    return (a, nb + g, feats, pos.astype(np.float32), np.concatenate(n_parts), np.concatenate(e_parts),
            np.concatenate([battr, np.array(c_attr, dtype=np.int64)]), np.array(order, dtype=np.int64))


def make_molecules(profile: str, count: int, seed: int):
    """`count` un-collated synthetic molecules (a dataset): list of (atoms, hyperedges, x [a,9], pos [a,3],
    local atom id per incidence, local hyperedge id per incidence, edge_attr, e_order) and their targets y."""
    prof = PROFILES[profile]
    rng = np.random.default_rng(seed)
    mols = [_molecule(rng, prof) for _ in range(count)]
    return mols, rng.normal(size=count).astype(np.float32)


def collate_host(mols, y, ids) -> HBatch:
    """Host collation of the molecules `ids` of a dataset (PyG Batch.from_data_list over HData: data/utils.py:150-178)."""
    return _collate([mols[i] for i in ids], np.asarray([y[i] for i in ids], dtype=np.float32))


def make_batch(profile: str, batch_size: int, seed: int) -> HBatch:
    """One collated batch of `batch_size` synthetic molecules on the host (CPU tensors)."""
    prof = PROFILES[profile]
    rng = np.random.default_rng(seed)
    mols = [_molecule(rng, prof) for _ in range(batch_size)]
    return _collate(mols, rng.normal(size=batch_size).astype(np.float32))


def _collate(mols, y) -> HBatch:
    batch_size = len(mols)
    xs, ps, v0, v1, ea, ne, eo, bt = [], [], [], [], [], [], [], []
    ptr = [0]
    a_off = e_off = 0
    for b in range(batch_size):
        a, ne_b, feats, pos, n_idx, e_idx, attr, order = mols[b]
        xs.append(feats)
        ps.append(pos)
        v0.append(n_idx + a_off)          # HData.__inc__('edge_index0') = x.size(0)
        v1.append(e_idx + e_off)          # HData.__inc__('edge_index1') = n_e
        ea.append(attr)
        eo.append(order)
        ne.append(ne_b)
        bt.append(np.full(a, b, dtype=np.int64))
        a_off += a
        e_off += ne_b
        ptr.append(a_off)
    t = torch.from_numpy
    return HBatch(
        x=t(np.concatenate(xs)), pos=t(np.concatenate(ps)), edge_index0=t(np.concatenate(v0)),
        edge_index1=t(np.concatenate(v1)), edge_attr=t(np.concatenate(ea)).view(-1, 1),
        n_e=t(np.array(ne, dtype=np.int64)), e_order=t(np.concatenate(eo)), batch=t(np.concatenate(bt)),
        ptr=t(np.array(ptr, dtype=np.int64)), y=t(y), num_graphs=batch_size)


def bench_seed(config: int, rank: int, step: int) -> int:
    """seed = 1000*config + 10*rank + step  (SURVEY.md 8(d))."""
    return 1000 * config + 10 * rank + step


def model_args(hidden: int = 256, **over):
    """Namespace with exactly the model fields of the reference's argparse (main.py:178-202),
    defaulting to the shipped 3-D script sizes (scripts/run_qm9_3d.sh:15-30)."""
    import argparse
    d = dict(All_num_layers=3, MLP1_num_layers=2, MLP2_num_layers=2, MLP3_num_layers=2, MLP4_num_layers=2,
             MLP_hidden=hidden, output_num_layers=3, output_hidden=128, aggregate="mean", normalization="ln",
             activation="relu", dropout=0.0)
    d.update(over)
    return argparse.Namespace(**d)
