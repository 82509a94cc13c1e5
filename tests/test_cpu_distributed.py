"""world_size-2 `gloo` tests (CPU) of the data-parallel plumbing (reference: Lightning DDP at
main.py:281): flat parameter / gradient buffers, ONE all-reduce of the flat gradient per step, mean over ranks
folded into the optimizer's grad_scale, zero gradient slices for the phi_x parameters that never get a gradient
(no find_unused_parameters walk), identical parameters on every rank afterwards.  The CUDA kernels are not
involved: the fused Adam launch is replaced by its torch formula for this test only."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _adam_reference(param, grad, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    g = grad * grad_scale
    if weight_decay:
        g = g + weight_decay * param
    m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    mh = m / (1 - betas[0] ** step)
    vh = v / (1 - betas[1] ** step)
    param.sub_(lr * mh / (vh.sqrt() + eps))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from equihgnn_b200 import ops, synthetic
        from equihgnn_b200.models import EGNNEquiHNNS
        from equihgnn_b200.train import Trainer
        torch.manual_seed(0)                              # same initial weights on every rank
        model = EGNNEquiHNNS(1, synthetic.model_args(hidden=32))
        ops.adam_step_ = _adam_reference                  # no GPU here: the kernel's formula in torch
        tr = Trainer(model, lr=1e-2, grad_buckets=2)
        assert tr.distributed and tr.world == world
        names = [n for n, _ in model.named_parameters()]
        # .grad fields are views of ONE buffer
        assert all((p.data_ptr() - tr.fp.flat.data_ptr()) % 128 == 0 for p in tr.fp.params)   # 128-byte aligned slices
        base = tr.fp.grad.data_ptr()
        assert all(base <= p.grad.data_ptr() < base + 4 * tr.fp.numel for p in tr.fp.params)
        p0 = tr.fp.flat.clone()
        for step in range(2):
            tr.fp.zero_grad()
            g = torch.Generator().manual_seed(100 * step + rank)        # rank-dependent "local" gradients
            for n, p in zip(names, tr.fp.params):
                if n.startswith("egnn_layer.coors_"):                  # phi_x: never receives a gradient
                    continue
                p.grad.copy_(torch.randn(p.shape, generator=g))
            local = tr.fp.grad.clone()
            tr.sync_gradients()
            gathered = [torch.empty_like(local) for _ in range(world)]
            dist.all_gather(gathered, local)
            assert torch.equal(tr.fp.grad, sum(gathered[1:], gathered[0].clone()))   # one SUM all-reduce of the flat buffer
            tr.optimizer_step()
        # every rank holds the same parameters; untouched phi_x parameters did not move (zero grad, zero moments)
        flats = [torch.empty_like(tr.fp.flat) for _ in range(world)]
        dist.all_gather(flats, tr.fp.flat)
        assert all(torch.equal(flats[0], f) for f in flats[1:])
        moved = (tr.fp.flat - p0).abs().max()
        assert float(moved) > 0
        for n, p in model.named_parameters():
            if n.startswith("egnn_layer.coors_"):
                torch.manual_seed(0)
                ref = dict(EGNNEquiHNNS(1, synthetic.model_args(hidden=32)).named_parameters())[n]
                assert torch.equal(p.detach(), ref.detach()), n
        # equals single-process Adam on the MEAN gradient
        torch.manual_seed(0)
        m2 = EGNNEquiHNNS(1, synthetic.model_args(hidden=32))
        t2 = Trainer(m2, lr=1e-2, distributed=False)
        for step in range(2):
            t2.fp.zero_grad()
            for r in range(world):
                g = torch.Generator().manual_seed(100 * step + r)
                for n, p in zip(names, t2.fp.params):
                    if n.startswith("egnn_layer.coors_"):
                        continue
                    p.grad.add_(torch.randn(p.shape, generator=g) / world)
            t2.optimizer_step()
        assert torch.allclose(t2.fp.flat, tr.fp.flat, rtol=1e-5, atol=1e-7)
        # ---- the bucketed exchange of the fused (CUDA-graph) step: the conv + readout parameters form the suffix of the
        # flat buffers that is final at the model's grad_boundary; exchanging + updating that bucket first and the rest at
        # the end must equal ONE all-reduce + ONE Adam over the whole buffer (device-side step counter: advance once)
        lo_hi = tr.bucket_ranges()
        assert len(lo_hi) == 2 and lo_hi[0][0] == 0 and lo_hi[0][1] == lo_hi[1][0] == tr.split and lo_hi[1][1] == tr.fp.numel
        first_early = next(i for i, n in enumerate(names) if n.startswith(("conv.", "mlp_out.")))
        assert tr.split == tr.fp.offsets[first_early] and tr.split % 32 == 0
        assert all(n.startswith(("conv.", "mlp_out.")) for n in names[first_early:])

        def adam_dev(param, grad, m, v, step_dev, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0,
                     advance=True):
            _adam_reference(param, grad, m, v, int(step_dev) + 1, lr, betas, eps, weight_decay, grad_scale)
            if advance:
                step_dev += 1
        ops.adam_step_dev_ = adam_dev
        tr.step_dev.fill_(tr.t)                       # continue from the two host-driven steps above
        g = torch.Generator().manual_seed(777 + rank)
        tr.fp.zero_grad()
        for n, p in zip(names, tr.fp.params):
            if not n.startswith("egnn_layer.coors_"):
                p.grad.copy_(torch.randn(p.shape, generator=g))
        want_flat, want_m, want_v = tr.fp.flat.clone(), tr.exp_avg.clone(), tr.exp_avg_sq.clone()
        gsum = tr.fp.grad.clone()
        dist.all_reduce(gsum)
        _adam_reference(want_flat, gsum, want_m, want_v, tr.t + 1, tr.lr, tr.betas, tr.eps, 0.0, 1.0 / world)
        tr._exchange_and_update(tr.split, tr.fp.numel, advance=False)      # what the boundary hook does under the backward
        assert int(tr.step_dev) == tr.t
        tr._exchange_and_update(0, tr.split, advance=True)                  # the tail of the step
        assert int(tr.step_dev) == tr.t + 1
        assert torch.equal(tr.fp.grad, gsum)
        assert torch.equal(tr.fp.flat, want_flat) and torch.equal(tr.exp_avg, want_m) and torch.equal(tr.exp_avg_sq, want_v)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_and_adam_world2(tmp_path):
    world = 2
    for attempt in range(3):          # the probed port can be taken between the probe and the rendezvous
        try:
            mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
            break
        except Exception:
            if attempt == 2:
                raise
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_bench_reference_arm_under_torchrun_prints_one_line(tmp_path):
    """`bench.py --impl reference` under a 2-process launch: rank 0 alone runs and prints one JSON line."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1", "--hidden", "32"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["warmup"] == 1 and d["steps"] == 1
    assert d["config"]["workload"].startswith("egnn_equihnns qm9-shaped batch 128/GPU, C=32")
