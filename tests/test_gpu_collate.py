"""GPU: device-side batch collation (eqh_collate) against the host collation -- integer / byte work, bit-exact."""
import dataclasses

import numpy as np
import pytest
import torch

from equihgnn_b200 import ops, synthetic
from equihgnn_b200.collate import MoleculeStore, StoreTrainer, collate_device
from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
from equihgnn_b200.models import EGNNEquiHNNS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _same_incidence(batch):
    """The CSR emitted by the collation (eqh_collate_incidence) == the CSR built from the collated index lists
    (eqh_incidence_build_i64): integer work, bit-exact, ghost segments included."""
    got = batch.incidence
    want = ops.build_incidence(batch.edge_index0, batch.edge_index1, batch.num_atoms, batch.num_hyperedges)
    assert (got.N, got.E, got.Z) == (want.N, want.E, want.Z)
    for f in ("rowptr_e", "perm_e", "vtx_by_e", "rowptr_v", "perm_v", "edge_by_v", "vertex32", "edges32"):
        a, b = getattr(got, f), getattr(want, f)
        assert a.dtype == b.dtype and a.shape == b.shape, f
        assert torch.equal(a, b), f


def _same(a, b):
    for f in dataclasses.fields(a):
        va, vb = getattr(a, f.name), getattr(b, f.name)
        if isinstance(va, torch.Tensor):
            assert va.dtype == vb.dtype and va.shape == vb.shape, f.name
            assert torch.equal(va.cpu(), vb.cpu()), f.name
        else:
            assert va == vb, f.name


@pytest.mark.parametrize("profile,M,B", [("qm9", 300, 128), ("opv", 40, 8), ("molecule3d", 120, 64), ("qm9", 5, 1)])
def test_collate_device_is_bit_exact(profile, M, B):
    mols, y = synthetic.make_molecules(profile, M, 77)
    store = MoleculeStore(mols, y, DEV)
    rng = np.random.default_rng(1)
    for trial in range(3):
        ids = rng.choice(M, size=B, replace=False).tolist() if trial else list(range(B))
        host = synthetic.collate_host(mols, y, ids)
        plain = collate_device(store, ids)
        _same(plain, host)                                                        # plain collation
        _same_incidence(plain)
        bucket = bucket_of(host)
        padded = collate_device(store, ids, bucket)
        _same(padded, pad_batch(host, bucket))                                    # with the ghost molecule
        _same_incidence(padded)
        tight = (host.num_atoms + 1, host.num_hyperedges + 1, host.num_incidences)   # one ghost atom / hyperedge, no ghost incidence
        t = collate_device(store, ids, tight)
        _same(t, pad_batch(host, tight))
        _same_incidence(t)
        # ghost incidences that do not divide evenly over the ghost atoms / hyperedges, and fewer of them than segments
        for extra in ((3, 2, 7), (5, 9, 4), (2, 3, 1)):
            odd = (host.num_atoms + extra[0], host.num_hyperedges + extra[1], host.num_incidences + extra[2])
            o = collate_device(store, ids, odd)
            _same(o, pad_batch(host, odd))
            _same_incidence(o)


def test_store_trainer_matches_host_fed_trainer():
    """Training from molecule ids (device collation into the graph's static inputs) == training from host batches."""
    C, B = 64, 24
    mols, y = synthetic.make_molecules("qm9", 200, 5)
    store = MoleculeStore(mols, y, DEV)
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(4)
        return EGNNEquiHNNS(1, args).to(DEV).train()

    a = StoreTrainer(fresh(), store, real_graphs=B, lr=1e-3)
    b = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3)
    rng = np.random.default_rng(3)
    for step in range(4):
        ids = rng.choice(200, size=B, replace=False).tolist()
        la = a.step_from_ids(ids)
        host = synthetic.collate_host(mols, y, ids)
        lb = b.step_from_host(pad_batch(host, bucket_of(host)).pack(pin=True), DEV)
        assert la == lb, (step, la, lb)                                          # same inputs, same graph: same bits
    for (k, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert torch.equal(p, q), k


def test_store_trainer_skips_the_incidence_build():
    """With the CSR emitted by the collation the captured step holds no incidence-build launches for the hypergraph (the
    transposed kNN list of the EGNN layer is still built in the step) and trains bit-identically to the in-graph build."""
    C, B = 64, 16
    mols, y = synthetic.make_molecules("qm9", 100, 9)
    store = MoleculeStore(mols, y, DEV)
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(6)
        return EGNNEquiHNNS(1, args).to(DEV).train()

    a = StoreTrainer(fresh(), store, real_graphs=B, lr=1e-3)
    b = StoreTrainer(fresh(), store, real_graphs=B, lr=1e-3, with_incidence=False)
    rng = np.random.default_rng(8)
    for step in range(3):
        ids = rng.choice(100, size=B, replace=False).tolist()
        assert a.step_from_ids(ids) == b.step_from_ids(ids), step
    la = next(iter(a.buckets.values())).launches
    lb = next(iter(b.buckets.values())).launches
    assert lb - la == 8, (la, lb)                 # two builds of four launches each left the graph
    for (k, p), (_, q) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert torch.equal(p, q), k
