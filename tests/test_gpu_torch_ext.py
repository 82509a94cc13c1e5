"""GPU: the torch C++ extension binding (torch.ops.eqh.*, csrc/torch_ext/eqh_torch.cpp) against the ctypes binding of the
same C ABI: identical bits, same autograd."""
import os

import pytest
import torch

from conftest import load_golden
from equihgnn_b200 import ops, synthetic, torch_ext

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _ext():
    if not os.path.exists(torch_ext.LIB_PATH):
        from equihgnn_b200 import build
        build.build_torch_ext()
    torch_ext.load()


@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_scatter_matches_ctypes_binding_and_golden(reduce):
    import torch.nn.functional as F
    fx = load_golden("scatter")
    src0 = F.pad(fx["src"], (0, 16))                         # C = 48 -> 64: the kernels take multiples of 32
    src = src0.to(DEV).requires_grad_(True)
    index = fx["index"].to(DEV)
    out = torch.ops.eqh.scatter(src, index, fx["n"], reduce)
    assert torch.equal(out.cpu()[:, :48], fx[f"{reduce}_dimsize"])               # bit-exact vs the reference's scatter
    w = torch.randn_like(out)
    (out * w).sum().backward()
    src2 = src0.to(DEV).requires_grad_(True)
    out2 = ops.scatter(src2, index, dim=-2, dim_size=fx["n"], reduce=reduce)
    (out2 * w).sum().backward()
    assert torch.equal(out, out2) and torch.equal(src.grad, src2.grad)
    assert torch.equal(torch_ext.scatter(src0.to(DEV), index, reduce=reduce).cpu()[:, :48], fx[f"{reduce}_auto"])
    with pytest.raises(NotImplementedError):                 # C = 48: unsupported width is reported, not computed wrongly
        torch.ops.eqh.scatter(fx["src"].to(DEV), index, fx["n"], reduce)


def test_primitives_match_ctypes_binding():
    b = synthetic.make_batch("qm9", 32, 3)
    inc = ops.build_incidence(b.edge_index0.to(DEV), b.edge_index1.to(DEV), b.num_atoms, b.num_hyperedges)
    rowptr, perm, dst32, osort = torch.ops.eqh.incidence_build(b.edge_index1.to(DEV), b.num_hyperedges, b.edge_index0.to(DEV))
    assert torch.equal(rowptr, inc.rowptr_e) and torch.equal(perm, inc.perm_e) and torch.equal(osort, inc.vtx_by_e)
    x = torch.randn(b.num_atoms, 64, device=DEV)
    a = torch.ops.eqh.gather_reduce(x, inc.vtx_by_e, inc.rowptr_e, None, inc.E, 1)
    assert torch.equal(a, ops.gather_reduce_raw(x, inc.vtx_by_e, inc.rowptr_e, None, inc.E, ops.MEAN))
    r = torch.ops.eqh.gather_rows(x, inc.vertex32, None, inc.Z)
    assert torch.equal(r, x[b.edge_index0.to(DEV)])
    idx, d2 = torch.ops.eqh.knn(b.pos.to(DEV), 16, None, None, False)
    i2, d22 = ops.knn16(b.pos.to(DEV))
    assert torch.equal(idx, i2) and torch.equal(d2, d22)
    w, bias = torch.randn(128, 64, device=DEV) * 0.1, torch.randn(128, device=DEV)
    y = torch.ops.eqh.linear(x, w, bias, 1)
    assert torch.equal(y, ops.gemm_raw(x, True, w, True, b.num_atoms, 128, 64, bias=bias, act="relu"))
    with pytest.raises(RuntimeError):
        torch.ops.eqh.gather_reduce(x.cpu(), None, inc.rowptr_e, None, inc.E, 0)     # no CPU path
    with pytest.raises(NotImplementedError):
        torch.ops.eqh.scatter(x[:4], torch.zeros(4, dtype=torch.int64, device=DEV), 1, "max")
