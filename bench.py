#!/usr/bin/env python
"""Headline benchmark: training molecules/s of egnn_equihnns on QM9-shaped synthetic molecular
hypergraphs (BASELINE.json configs[1]: batch 128 per GPU, script model sizes C=256, L=3), one
process per GPU, weak scaling.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

One JSON line on stdout (rank 0).  A "step" = forward + MSE loss + backward + gradient all-reduce + Adam on
one batch.  `value` times K steps with the batches already resident in HBM; `e2e` times the same steps through
GraphedTrainer.step_from_host (pinned host batch -> H2D -> step -> D2H loss).  Between timed steps a 256 MB
buffer is written to flush L2; every step is bracketed by CUDA events on the launching stream and the K
durations are summed, max over ranks.  The line also carries, measured in the same run with a few steps each,
the other BASELINE configurations (`workloads`), the roofline of the step's dominant kernel (the row GEMM) next to
the segmented-reduction HBM microbenchmark, the all-reduce cost and the drop-in (no Trainer) step.

The reference arm runs the UNMODIFIED reference files (offline install `baseline/_ref`) on the host
cores when that install is present and the configuration fits host memory (the verbatim EGNN backward needs a
[1,N,N,C] buffer: 5.5 GB at QM9-128, 195 GB at PCQM-768), else the oracle port (`kind` says which).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (profile, batch per GPU, BASELINE config id, method)
    "qm9_b128": ("qm9", 128, 2, "egnn_equihnns"),
    "pcqm_b768": ("pcqm4mv2", 768, 3, "egnn_equihnns"),
    "opv_b768": ("opv", 768, 4, "egnn_equihnns"),
    "mol3d_b1024": ("molecule3d", 1024, 5, "egnn_equihnns"),
    "qm9_b128_mhnnm": ("qm9", 128, 1, "mhnnm"),
}
METRIC = "training molecules/sec (egnn_equihnns, QM9-shape)"


def workload_string(name, hidden):
    """The SAME string in both arms (the driver compares `config.workload` of the two lines)."""
    profile, B, cfg, method = WORKLOADS[name]
    return (f"{method} {profile}-shaped batch {B}/GPU, C={hidden}, L=3 (BASELINE configs[{cfg - 1}]), "
            f"fwd+bwd+allreduce+Adam")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qm9_b128", choices=list(WORKLOADS))
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--n-batches", type=int, default=8, help="distinct synthetic batches cycled through")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `workloads` block, the microbenchmarks and the "
                                                            "drop-in leg (profiling runs)")
    ap.add_argument("--extra-steps", type=int, default=5, help="timed steps of each entry of the `workloads` block")
    ap.add_argument("--reference-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs (per-op dispatch)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1632.8, "bf16_tflops_sustained": 1402.4}, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU path on the host cores
# ---------------------------------------------------------------------------------------------
def _verbatim_feasible(workload, hidden):
    """The unmodified EGNN backward allocates [1,N,N,C] fp32: only the QM9-128 shapes fit host RAM."""
    profile, B, cfg, method = WORKLOADS[workload]
    if method == "mhnnm":
        return True
    n_atoms = {"qm9": 18.2, "pcqm4mv2": 30, "molecule3d": 29.5, "opv": 100}[profile] * B
    need = 4.0 * n_atoms * n_atoms * hidden * 2.5
    try:
        avail = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 16e9
    return need < 0.5 * avail


def cpu_reference_steps(workload, hidden, max_steps, budget_s, warmup=1, kind="auto"):
    """Times fwd+bwd+Adam of the reference's CPU path on the same synthetic batches, all host threads.
    kind 'reference': the unmodified reference files (oracle/ref_harness.py over baseline/_ref or /root/reference);
    kind 'port': the oracle restatement (oracle/restate.py, sparse EGNN gather).  'auto' prefers the former.
    Returns (molecules/s, cores, steps timed, seconds, kind, sample description)."""
    import torch
    import torch.nn.functional as F
    from equihgnn_b200 import synthetic
    from equihgnn_b200.models import create_model

    profile, B, cfg, method = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batches = [synthetic.make_batch(profile, B, synthetic.bench_seed(cfg, 0, s)) for s in range(2)]
    use_ref = False
    if kind in ("auto", "reference"):
        from oracle import ref_harness
        use_ref = ref_harness.available() and _verbatim_feasible(workload, hidden)
        if kind == "reference" and not use_ref:
            raise SystemExit("reference arm: the unmodified reference is not available / does not fit host memory here")
    torch.manual_seed(0)
    if use_ref:
        from oracle import ref_harness
        ref = ref_harness.load()
        model = {"egnn_equihnns": ref.EGNNEquiHNNS, "mhnnm": ref.MHNNM}[method](1, synthetic.model_args(hidden))
        model.train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-4)              # main.py:137-140

        def step(b):
            opt.zero_grad(set_to_none=True)
            loss = F.mse_loss(model(b), b.y)                              # main.py:49-51
            loss.backward()
            opt.step()
            return float(loss.detach())
        what = (f"UNMODIFIED reference files ({os.path.relpath(ref_harness.REFERENCE_ROOT, ROOT) if ref_harness.REFERENCE_ROOT.startswith(ROOT) else ref_harness.REFERENCE_ROOT}"
                f"; torch_scatter / torch_geometric / ogb stand-ins from oracle/shims), torch CPU fp32")
    else:
        from oracle import restate as R
        model = create_model(method, 1, synthetic.model_args(hidden))
        p = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()
             if "running" not in k and "num_batches" not in k}
        opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4)
        fn = {"egnn_equihnns": R.egnn_equihnns, "mhnnm": R.mhnnm}[method]

        def step(b):
            opt.zero_grad(set_to_none=True)
            loss = F.mse_loss(fn(p, b), b.y)
            loss.backward()
            opt.step()
            return float(loss.detach())
        what = "oracle port (oracle/restate.py: sparse EGNN restatement), torch CPU fp32"
    for i in range(warmup):
        step(batches[i % 2])
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < budget_s):
        step(batches[n % 2])
        n += 1
    dt = time.perf_counter() - t0
    sample = f"{n} full steps (fwd+bwd+Adam) of batch {B} in {dt:.1f} s, {what}"
    return n * B / dt, cores, n, dt, ("reference" if use_ref else "port"), sample


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, cores, n, dt, kind, sample = cpu_reference_steps(a.workload, a.hidden, a.steps, 1e9, warmup=a.warmup,
                                                          kind=a.reference_kind)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "molecules/s", "n_gpus": a.gpus, "steps": n,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(a.workload, a.hidden), "device": "host CPU"},
        "cpu_baseline": {"value": val, "unit": "molecules/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                mx.append(float(c[1]))
            except ValueError:
                continue
            for nm, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def _median_us(fn, flush_l2, reps=10, warm=3):
    import torch
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush_l2()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        fn()
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


def roofline_tensor(dev, rows, C, flush_l2):
    """The row GEMM in isolation (eqh_gemm_tf32x3, [rows, C] x [C, C], forward layout): algorithmic flops 2*M*N*K (the
    kernel executes 3x that on the TF32 pipe to stay fp32-faithful) over the median launch duration."""
    import torch
    from equihgnn_b200 import ops
    A = torch.randn(rows, C, device=dev)
    W = torch.randn(C, C, device=dev) * 0.05
    us = _median_us(lambda: ops.gemm_raw(A, True, W, True, rows, C, C), flush_l2)
    return {"rows": rows, "flops_per_launch": 2.0 * rows * C * C, "us_per_launch": us, "tflops": 2.0 * rows * C * C / us / 1e6}


def roofline_stacked(dev, C, flush_l2):
    """HBM roofline of the segmented-reduction kernels on a working set far above L2 ('the
    >=60 % HBM target is evaluated on configs 3-5 and on a stacked-batch microbenchmark (>= 1 GB per launch)'):
    the incidence of a Molecule3D-shaped batch of 1024 molecules (BASELINE configs[4]) stacked 12x, C features.
    Each kernel is launched 10 times with an L2 flush in between and timed with CUDA events around the launch
    (median); bytes are the algorithmic figures P / F of SURVEY.md 8(d)."""
    import torch
    from equihgnn_b200 import ops, synthetic
    b = synthetic.make_batch("molecule3d", 1024, synthetic.bench_seed(5, 0, 0))
    k = 12
    N0, E0 = b.num_atoms, b.num_hyperedges
    v = torch.cat([b.edge_index0 + i * N0 for i in range(k)]).to(dev)
    e = torch.cat([b.edge_index1 + i * E0 for i in range(k)]).to(dev)
    N, E = N0 * k, E0 * k
    inc = ops.build_incidence(v, e, N, E)
    Z = inc.Z
    A = torch.randn(N, C, device=dev)
    B = torch.randn(E, C, device=dev)
    src = torch.randn(Z, C, device=dev)
    gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    dU = torch.randn(N, C, device=dev)
    U, dA, dB = torch.empty(N, C, device=dev), torch.empty(N, C, device=dev), torch.empty(E, C, device=dev)
    part = torch.empty(ops.partial_rows(), 2 * C, device=dev)
    st = ops._stream
    P = lambda t: t.data_ptr()
    cases = {
        "eqh_gather_reduce[P: scatter-mean of a materialised [Z,C] tensor by hyperedge]": (
            4.0 * C * (Z + E) + 4.0 * (Z + E + 1),
            lambda: ops.gather_reduce_raw(src, inc.perm_e, inc.rowptr_e, None, E, ops.MEAN)),
        "eqh_gather_reduce[P: by vertex]": (
            4.0 * C * (Z + N) + 4.0 * (Z + N + 1),
            lambda: ops.gather_reduce_raw(src, inc.perm_v, inc.rowptr_v, None, N, ops.MEAN)),
        "eqh_gather_rows[backward of P]": (
            4.0 * C * (E + Z) + 4.0 * Z, lambda: ops.gather_rows_raw(B, inc.edges32, inc.rowptr_e, Z)),
        "eqh_gather_reduce[gather fused: V2E of MHNNSConv]": (
            4.0 * C * (N + E) + 4.0 * (Z + E + 1),
            lambda: ops.gather_reduce_raw(A, inc.vtx_by_e, inc.rowptr_e, None, E, ops.MEAN)),
        "eqh_fused_msg_fwd[F: E2V gather-add-ReLU-LN-mean]": (
            4.0 * C * (2 * N + E) + 4.0 * (Z + N + 1),
            lambda: ops._call("eqh_fused_msg_fwd", 1, P(A), P(B), None, P(inc.edge_by_v), P(inc.rowptr_v), N, C, P(gamma),
                              P(beta), 1e-5, 1, P(U), st())),
        "eqh_fused_msg_bwd[primary grouping, + LN affine partials]": (
            4.0 * C * (3 * N + E) + 4.0 * (Z + N + 1),
            lambda: ops._call("eqh_fused_msg_bwd", 1, P(A), P(B), None, P(inc.edge_by_v), None, P(inc.rowptr_v), N, C,
                              P(gamma), P(beta), 1e-5, P(dU), P(inc.rowptr_v), P(dA), P(part), st())),
        "eqh_fused_msg_bwd[secondary grouping]": (
            4.0 * C * (2 * N + 2 * E) + 4.0 * (2 * Z + E + 1),
            lambda: ops._call("eqh_fused_msg_bwd", 1, P(A), P(B), P(inc.vtx_by_e), None, P(inc.vtx_by_e), P(inc.rowptr_e),
                              E, C, P(gamma), P(beta), 1e-5, P(dU), P(inc.rowptr_v), P(dB), None, st())),
    }
    out = {}
    for name, (nbytes, fn) in cases.items():
        us = _median_us(fn, flush_l2)
        out[name] = {"bytes_per_launch": nbytes, "us_per_launch": us, "achieved_gbs": nbytes / us / 1e3}
    return {"N": N, "E": E, "Z": Z, "C": C, "kernels": out}


class Ctx:
    """Per-process state shared by the measurements."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.a = a
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != a.gpus and self.world > 1:
            raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)

    def flush_l2(self):
        self.flush_buf.fill_(1)

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_min(self, vals):
        """element-wise (max, min) over ranks of a list of floats"""
        import torch
        import torch.distributed as dist
        t = torch.tensor(vals, dtype=torch.float64, device=self.dev)
        if self.world == 1:
            return list(vals), list(vals)
        mx, mn = t.clone(), t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        return [float(v) for v in mx], [float(v) for v in mn]


def measure_workload(cx: Ctx, name: str, steps: int, warmup: int, n_batches: int, full: bool, family: bool):
    """One workload on this rank's GPU: `value` (resident batches), `e2e` (host batches through step_from_host), and -- for
    the headline workload (`full`) -- device-side collation; `family`: per-kernel figures from an eager pass."""
    import numpy as np
    import torch
    from equihgnn_b200 import ops, synthetic
    from equihgnn_b200.graphed import GraphedTrainer, common_bucket, pad_batch
    from equihgnn_b200.models import create_model
    from equihgnn_b200.train import Trainer

    a, dev, world, rank = cx.a, cx.dev, cx.world, cx.rank
    profile, B, cfg, method = WORKLOADS[name]
    torch.manual_seed(0)                                   # identical initial weights on every rank (DDP broadcast)
    model = create_model(method, 1, synthetic.model_args(a.hidden)).to(dev).train()
    raw = [synthetic.make_batch(profile, B, synthetic.bench_seed(cfg, rank, s)) for s in range(n_batches)]
    graphed = not a.eager
    res = {}
    use_store = full and graphed and method == "egnn_equihnns"
    if graphed:
        # collation pads every batch of every rank to ONE bucket with an inert ghost molecule (graphed.py): all ranks
        # replay graphs of the same shape at every step, and the gradient exchange + Adam live inside the graph
        bucket = common_bucket(raw)
        if use_store:
            # device-side collation leg (below): a dataset resident in HBM; its batches must fit the same bucket
            from equihgnn_b200.collate import MoleculeStore, StoreTrainer
            mols, ys = synthetic.make_molecules(profile, 16 * B, synthetic.bench_seed(cfg, rank, 99))
            store = MoleculeStore(mols, ys, dev)
            rng = np.random.default_rng(7 + rank)
            id_sets = [rng.choice(len(mols), size=B, replace=False) for _ in range(steps + 4)]
            shapes = [store.shape_of(ids) for ids in id_sets]
            need = [max(s_[0] for s_ in shapes) + 1, max(s_[1] for s_ in shapes) + 1, max(s_[2] for s_ in shapes)]
            mx, _ = cx.reduce_max_min([float(v) for v in need])
            bucket = tuple(max(int(-(-int(v) // g) * g), bk) for v, g, bk in zip(mx, (256, 256, 512), bucket))
            trainer = StoreTrainer(model, store, real_graphs=B, fixed_bucket=bucket, lr=1e-4, uniform_shapes=True)
        else:
            trainer = GraphedTrainer(model, real_graphs=B, lr=1e-4, uniform_shapes=True)
        host = [pad_batch(b, bucket).pack(pin=True) for b in raw]              # one pinned buffer per collated batch
        h2d_bytes = sum(h._flat.numel() for h in host) / len(host)
        res["bucket"] = list(bucket)
    else:
        trainer = Trainer(model, lr=1e-4)
        host = [b.pin_memory() for b in raw]
        h2d_bytes = sum(h.h2d_bytes() for h in host) / len(host)
    resident = [h.pack(device=dev) if graphed else h.to(dev) for h in host]

    # ---- warm-up: every distinct batch once (captures one CUDA graph per shape bucket) ----------
    for i in range(max(warmup, 3, len(resident))):
        trainer.step(resident[i % len(resident)])
    cx.barrier()

    if os.environ.get("EQH_BENCH_TRACE") and rank == 0 and full:
        # kernel timeline of three steps on rank 0 (torch.profiler / CUPTI), for scratch/trace_summary.py; not timed
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(3):
                trainer.step(resident[i % len(resident)])
            torch.cuda.synchronize()
        prof.export_chrome_trace(os.environ["EQH_BENCH_TRACE"])
    elif os.environ.get("EQH_BENCH_TRACE") and full:
        for i in range(3):
            trainer.step(resident[i % len(resident)])
    cx.barrier()

    def launches_now():
        return (trainer.own_launches + (0 if trainer.in_graph_step else trainer.t)) if graphed else ops.STATS["launches"]

    # ---- value: K steps, inputs resident in HBM -----------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    cx.barrier()
    l0 = launches_now()
    for k in range(steps):
        cx.flush_l2()
        ev[k][0].record()
        trainer.step(resident[k % len(resident)])
        ev[k][1].record()
    cx.barrier()
    launches = launches_now() - l0
    ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- e2e: host batches through the public API; the loss of step k is read while step k+1 is queued ------------
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if graphed:
        e2e_step = lambda hb: trainer.step_from_host(hb, dev, sync=False)
        e2e_flush = trainer.flush_loss
    else:
        e2e_step = lambda hb: trainer.step_from_host(hb, dev)
        e2e_flush = lambda: None
    for i in range(2):
        e2e_step(host[i % len(host)])
    e2e_flush()
    cx.barrier()
    wall0 = time.perf_counter()
    for k in range(steps):
        cx.flush_l2()
        ev2[k][0].record()
        e2e_step(host[k % len(host)])
        ev2[k][1].record()
    last_loss = e2e_flush()
    cx.barrier()
    wall_e2e = time.perf_counter() - wall0
    ms_e2e = sum(s.elapsed_time(e) for s, e in ev2)

    # ---- e2e with device-side collation: the dataset is resident in HBM as a packed
    #      per-molecule store; per step the host sends the molecule ids + offsets (pinned, ~2 KB), the batch is
    #      gathered on the device straight into the graph's static inputs, then the same step, D2H loss
    ms_store, store_bytes = 0.0, 0
    if use_store:
        st_tr = trainer                                   # the same trainer: same model, optimizer state and captured graph
        inputs = [store.step_inputs(ids) for ids in id_sets]
        for i in range(4):
            st_tr.step_from_ids(id_sets[i], inputs[i], sync=False)
        st_tr.flush_loss()
        cx.barrier()
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            cx.flush_l2()
            ev3[k][0].record()
            st_tr.step_from_ids(id_sets[4 + k], inputs[4 + k], sync=False)
            ev3[k][1].record()
        st_tr.flush_loss()
        cx.barrier()
        ms_store, store_bytes = sum(s.elapsed_time(e) for s, e in ev3), int(inputs[0].numel() * 4)

    (ms_mx, e2e_mx, store_mx), (ms_mn, _, _) = cx.reduce_max_min([ms, ms_e2e, ms_store])
    res.update({
        "value": world * B * steps / (ms_mx * 1e-3), "ms_per_step": ms_mx / steps, "steps": steps,
        "e2e": {"value": world * B * steps / (e2e_mx * 1e-3), "unit": "molecules/s", "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_mx / steps, "wall_ms_per_step": 1e3 * wall_e2e / steps,
                "last_loss": last_loss},
        "gpu_launches": launches, "rank_skew_us": 1e3 * (ms_mx - ms_mn) / steps,
        "execution": (("CUDA graph replay (%d shape bucket%s, ghost-padded batches); %s" %
                       (len(trainer.buckets), "" if len(trainer.buckets) == 1 else "s",
                        "zero-grad + fwd + bwd + bucketed all-reduce + Adam inside the graph" if trainer.in_graph_step
                        else "fwd + bwd inside the graph, all-reduce + Adam after it")) if graphed else "eager"),
        "atoms_per_batch": float(np.mean([b.num_atoms for b in raw])),
    })
    if ms_store:
        res["e2e_device_collation"] = {
            "value": world * B * steps / (store_mx * 1e-3), "unit": "molecules/s", "ms_per_step": store_mx / steps,
            "h2d_bytes_per_step": store_bytes, "d2h_bytes_per_step": 4,
            "note": "dataset (16 batches of molecules) resident in HBM as a packed per-molecule store; per step the pinned "
                    "molecule ids + offsets go H2D, eqh_collate gathers the ghost-padded batch into the graph's static "
                    "inputs, then the same step and the D2H loss"}

    # ---- per-kernel pass: the same forward+backward run eagerly (a graph replay cannot be bracketed per kernel), CUDA
    #      events around each launch of the GEMM and of the segmented-reduction family, L2 flushed between steps
    if family:
        names = ("eqh_gemm_tf32x3", "eqh_gemm_tf32x3_ex", "eqh_fused_msg_fwd", "eqh_fused_msg_bwd", "eqh_gather_reduce",
                 "eqh_gather_reduce_ld", "eqh_gather_rows", "eqh_egnn_edge_fwd_tc_ld", "eqh_egnn_edge_bwd_tc_ld")
        eager_step = trainer._fwd_bwd if graphed else trainer.forward_backward
        prof_steps = min(steps, 5)
        for i in range(2):
            eager_step(resident[i % len(resident)])
        ops.PROFILE = {n: [] for n in names}
        t_ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for k in range(prof_steps):
            cx.flush_l2()
            if k == prof_steps - 1:
                t_ev[0].record()
            eager_step(resident[k % len(resident)])
            if k == prof_steps - 1:
                t_ev[1].record()
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        fam = {}
        for n, recs in prof.items():
            tot_ms = sum(s.elapsed_time(e) for s, e, _ in recs)
            tot_w = sum(w for _, _, w in recs)
            if recs and tot_ms > 0:
                fam[n] = {"launches_per_step": len(recs) / prof_steps, "us_per_launch": 1e3 * tot_ms / len(recs),
                          "us_per_step": 1e3 * tot_ms / prof_steps, "work_per_launch": tot_w / len(recs),
                          "rate": tot_w / (tot_ms * 1e-3)}
        res["in_step"] = fam
    del trainer, resident, host
    torch.cuda.empty_cache()
    return res


def allreduce_us(cx: Ctx, numel: int):
    """One NCCL all-reduce of the flat fp32 gradient buffer (the step's only collective), median of 20 launches."""
    import torch
    import torch.distributed as dist
    if cx.world == 1:
        return 0.0
    buf = torch.zeros(numel, dtype=torch.float32, device=cx.dev)
    for _ in range(5):
        dist.all_reduce(buf)
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        dist.all_reduce(buf)
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    ts.sort()
    mx, _ = cx.reduce_max_min([ts[len(ts) // 2] * 1e3])
    return mx[0]


def dropin_eager(cx: Ctx, name: str, steps: int = 5):
    """The drop-in path a user of the reference's main.py gets WITHOUT this package's Trainer: the nn.Modules through
    per-op autograd Functions, torch.optim.Adam, a host-collated batch per step, loss.item() per step (what Lightning's
    loop does, main.py:49-63,137-140)."""
    import torch
    import torch.nn.functional as F
    from equihgnn_b200 import synthetic
    from equihgnn_b200.models import create_model
    profile, B, cfg, method = WORKLOADS[name]
    torch.manual_seed(0)
    model = create_model(method, 1, synthetic.model_args(cx.a.hidden)).to(cx.dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    host = [synthetic.make_batch(profile, B, synthetic.bench_seed(cfg, cx.rank, s)).pin_memory() for s in range(2)]

    def step(hb):
        d = hb.to(cx.dev, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(model(d), d.y)
        loss.backward()
        opt.step()
        return float(loss.item())
    for i in range(3):
        step(host[i % 2])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        step(host[k % 2])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return {"value": B * steps / dt, "unit": "molecules/s", "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "note": "one rank, wall clock: reference-style nn.Module step (per-op autograd Functions over the C ABI, "
                    "torch.optim.Adam, no CUDA graph, no Trainer), host-collated pinned batch in, loss.item() out"}


def run_b200(a):
    import torch
    import torch.distributed as dist
    from equihgnn_b200 import ops

    cx = Ctx(a)
    rank, world, dev = cx.rank, cx.world, cx.dev
    pk, peak_src = peaks()
    hbm_peak = float(pk["hbm_gbs"])
    bf16_burst = float(pk.get("bf16_tflops", 1632.8))

    sampler = ClockSampler(cx.local) if rank == 0 else None
    main = measure_workload(cx, a.workload, a.steps, a.warmup, a.n_batches, full=True, family=True)
    clocks = sampler.stop() if sampler else None

    extra = {}
    if not a.no_extra:
        for name in WORKLOADS:
            if name == a.workload:
                continue
            r = measure_workload(cx, name, a.extra_steps, 3, 2, full=False, family=(name == "mol3d_b1024"))
            extra[name] = r
    n_grad = None
    if world > 1:
        from equihgnn_b200 import synthetic
        from equihgnn_b200.models import create_model
        n_grad = sum(p.numel() for p in create_model(WORKLOADS[a.workload][3], 1, synthetic.model_args(a.hidden)).parameters())
    ar_us = allreduce_us(cx, n_grad) if n_grad else 0.0

    stacked = tens = drop = None
    if rank == 0 and not a.no_extra:
        stacked = roofline_stacked(dev, a.hidden, cx.flush_l2)
        tens = [roofline_tensor(dev, r, a.hidden, cx.flush_l2) for r in (main["bucket"][0] if "bucket" in main else 2560, 30000)]
        drop = dropin_eager(cx, a.workload)

    if rank == 0:
        profile, B, cfg, method = WORKLOADS[a.workload]
        out = {
            "metric": METRIC, "value": main["value"], "unit": "molecules/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(a.workload, a.hidden),
                       "distinct_batches": a.n_batches, "l2": "flushed between timed steps (256 MB write)",
                       "execution": main["execution"], "bucket": main.get("bucket"),
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                       "parallelism": f"dp{world}"},
            "clocks": clocks,
            "e2e": main["e2e"],
            "gpu_launches": main["gpu_launches"],
            "allreduce_us": ar_us, "rank_skew_us": main["rank_skew_us"],
        }
        if "e2e_device_collation" in main:
            out["e2e_device_collation"] = main["e2e_device_collation"]
        if drop:
            out["e2e_dropin_eager"] = drop
        if extra:
            out["workloads"] = {k: {"workload": workload_string(k, a.hidden), "value": v["value"], "unit": "molecules/s",
                                    "ms_per_step": v["ms_per_step"], "steps": v["steps"], "e2e": v["e2e"]["value"],
                                    "e2e_ms_per_step": v["e2e"]["ms_per_step"], "bucket": v.get("bucket"),
                                    "atoms_per_batch": v["atoms_per_batch"], "rank_skew_us": v["rank_skew_us"],
                                    "gpu_launches": v["gpu_launches"]} for k, v in extra.items()}
        # ---- roofline of the step's dominant kernel: the fp32-faithful row GEMM (share of the step from the eager pass)
        fam = main.get("in_step", {})
        g = [fam[n] for n in ("eqh_gemm_tf32x3", "eqh_gemm_tf32x3_ex") if n in fam]
        if g:
            g_us = sum(x["us_per_step"] for x in g)
            g_fl = sum(x["work_per_launch"] * x["launches_per_step"] for x in g)
            g_n = sum(x["launches_per_step"] for x in g)
            ach = g_fl / g_us / 1e6                       # TFLOP/s over the summed launch durations
            out["roofline"] = {
                "bound": "tensor", "kernel": "eqh_gemm_tf32x3 (fp32-faithful 3xTF32 row GEMM on tcgen05): the largest share of "
                                             "the step's kernel time",
                "achieved": ach, "peak": bf16_burst, "unit": "TFLOP/s", "frac": ach / bf16_burst,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE forward launch [2560,256]x[256,256] of this kernel
                "traffic": 2938368.0,
                "traffic_source": "CONSTANT from profiles/r02_kernels_ncu_full.md (ncu --set full, k_gemm_ts<64,1,1> at the "
                                  "headline size: 2.94 MB read, 0 written -- the compulsory A tile bytes; weights and output "
                                  "stay in L2); not re-measured by this run",
                "peak_source": peak_src + ": dense bf16 burst (kernels timed one launch at a time); a TF32 MMA runs at half "
                               "the bf16 rate and the kernel issues 3 TF32 products per algorithmic product, so this "
                               "formulation's own ceiling is peak/6",
                "frac_of_3xtf32_ceiling": ach / (bf16_burst / 6.0),
                "launches_per_step": g_n, "us_per_launch": g_us / g_n, "flops_per_launch": g_fl / g_n,
                "how": "algorithmic flops 2*M*N*K of every GEMM launch of one eager forward+backward of the headline "
                       "workload (CUDA events around each launch on the launching stream, L2 flushed between steps), "
                       "summed, over the summed launch durations",
                "in_step_family_us_per_step": {k: v["us_per_step"] for k, v in fam.items()},
            }
            if tens:
                out["roofline"]["isolated"] = {
                    f"rows={t_['rows']}": {"us_per_launch": t_["us_per_launch"], "tflops": t_["tflops"],
                                           "frac": t_["tflops"] / bf16_burst,
                                           "frac_of_3xtf32_ceiling": t_["tflops"] / (bf16_burst / 6.0)} for t_ in tens}
        if stacked:
            ks = stacked["kernels"]
            head = "eqh_gather_reduce[P: scatter-mean of a materialised [Z,C] tensor by hyperedge]"
            d = ks[head]
            out["roofline_hbm"] = {
                "bound": "hbm", "kernel": "eqh_gather_reduce (deterministic segmented mean, the torch_scatter.scatter "
                                          "replacement, SURVEY.md 8(d) figure P)",
                "achieved": d["achieved_gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": d["achieved_gbs"] / hbm_peak,
                "traffic": 832.7e6 + 353.4e6 if (stacked["Z"], stacked["C"]) == (808488, 256) else None,
                "traffic_source": "CONSTANT from profiles/r01_segmented_reduction_stacked_ncu.md (ncu --set full of this "
                                  "kernel on this workload: 832.7 MB read + 353.4 MB written per launch); not re-measured "
                                  "by this run",
                "peak_source": peak_src, "us_per_launch": d["us_per_launch"], "bytes_per_launch": d["bytes_per_launch"],
                "workload": "stacked-batch microbenchmark inside bench.py: incidence of a Molecule3D-shaped batch of 1024 "
                            "molecules stacked 12x (N=%d, E=%d, Z=%d), C=%d; 10 launches, L2 flushed between launches, "
                            "CUDA events around the launch, median" % (stacked["N"], stacked["E"], stacked["Z"], stacked["C"]),
                "family_stacked": {k: dict(v, frac=v["achieved_gbs"] / hbm_peak) for k, v in ks.items()},
            }
            m3 = extra.get("mol3d_b1024", {}).get("in_step")
            if m3:
                out["roofline_hbm"]["in_step_mol3d_b1024"] = {
                    k: {"launches_per_step": v["launches_per_step"], "us_per_launch": v["us_per_launch"],
                        "achieved_gbs": v["rate"] / 1e9, "frac": v["rate"] / 1e9 / hbm_peak}
                    for k, v in m3.items() if k.startswith(("eqh_fused_msg", "eqh_gather"))}
        if not a.no_cpu_baseline and world == 1:
            val, cores, n, dt, kind, sample = cpu_reference_steps(a.workload, a.hidden, 1000, a.cpu_seconds, kind=a.reference_kind)
            out["cpu_baseline"] = {"value": val, "unit": "molecules/s", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    # stdout carries exactly ONE JSON line: libraries that print banners on fd 1 (NCCL's version line, torchrun's
    # OMP notice) are redirected to stderr for the whole run and the result is written to the saved descriptor.
    real_stdout = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
