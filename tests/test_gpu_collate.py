"""GPU: device-side batch collation (eqh_collate) against the host collation -- integer / byte work, bit-exact."""
import dataclasses

import numpy as np
import pytest
import torch

from equihgnn_b200 import synthetic
from equihgnn_b200.collate import MoleculeStore, StoreTrainer, collate_device
from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
from equihgnn_b200.models import EGNNEquiHNNS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


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
        _same(collate_device(store, ids), host)                                   # plain collation
        bucket = bucket_of(host)
        _same(collate_device(store, ids, bucket), pad_batch(host, bucket))        # with the ghost molecule
        tight = (host.num_atoms + 1, host.num_hyperedges + 1, host.num_incidences)   # one ghost atom / hyperedge, no ghost incidence
        _same(collate_device(store, ids, tight), pad_batch(host, tight))


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
