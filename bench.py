#!/usr/bin/env python
"""Headline benchmark: training molecules/s of egnn_equihnns on QM9-shaped synthetic molecular
hypergraphs (BASELINE.json configs[1]: batch 128 per GPU, script model sizes C=256, L=3), one
process per GPU, weak scaling.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port)

One JSON line on stdout (rank 0).  A "step" = forward + MSE loss + backward + gradient
all-reduce + Adam on one batch.  `value` times K steps with the batches already resident in HBM;
`e2e` times the same steps through Trainer.step_from_host (pinned host batch -> H2D -> step ->
D2H loss).  Between timed steps a 256 MB buffer is written to flush L2; every step is bracketed by
CUDA events on the launching stream and the K durations are summed, max over ranks.
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
NCU_TRAFFIC_BYTES = 832.7e6 + 353.4e6   # eqh_gather_reduce on the stacked workload (see roofline.traffic_source)


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
    ap.add_argument("--breakdown", action="store_true", help="print per-kernel time shares to stderr")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs (per-op dispatch)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_steps(workload, hidden, max_steps, budget_s, warmup=1):
    """Times fwd+bwd+Adam of the oracle restatement (oracle/restate.py, torch CPU fp32, all host threads)
    on the same synthetic batches.  Returns (molecules/s, cores, steps timed, seconds)."""
    import torch
    import torch.nn.functional as F
    from equihgnn_b200 import synthetic
    from equihgnn_b200.models import create_model
    from oracle import restate as R

    profile, B, cfg, method = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = create_model(method, 1, synthetic.model_args(hidden))
    p = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()
         if "running" not in k and "num_batches" not in k}
    opt = torch.optim.Adam([v for v in p.values() if v.requires_grad], lr=1e-4)
    fn = {"egnn_equihnns": R.egnn_equihnns, "mhnnm": R.mhnnm}[method]
    batches = [synthetic.make_batch(profile, B, synthetic.bench_seed(cfg, 0, s)) for s in range(2)]

    def step(b):
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(fn(p, b), b.y)
        loss.backward()
        opt.step()
        return float(loss.detach())

    for i in range(warmup):
        step(batches[i % 2])
    t0 = time.perf_counter()
    n = 0
    while n < max_steps and (n == 0 or time.perf_counter() - t0 < budget_s):
        step(batches[n % 2])
        n += 1
    dt = time.perf_counter() - t0
    return n * B / dt, cores, n, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    profile, B, cfg, method = WORKLOADS[a.workload]
    val, cores, n, dt = cpu_reference_steps(a.workload, a.hidden, a.steps, 1e9, warmup=max(1, min(a.warmup, 2)))
    sample = f"{n} full steps (fwd+bwd+Adam) of batch {B}, oracle port (sparse EGNN restatement), torch CPU fp32"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "molecules/s", "n_gpus": a.gpus, "steps": n,
        "warmup": max(1, min(a.warmup, 2)), "ms_per_step": 1e3 * dt / n, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{method} {profile}-shaped batch {B}, C={a.hidden}, L=3 (BASELINE configs[{cfg - 1}])",
                   "device": "host CPU"},
        "cpu_baseline": {"value": val, "unit": "molecules/s", "cores": cores, "kind": "port", "sample": sample},
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


def roofline_tensor(dev, rows, C, flush_l2):
    """Tensor roofline of the row GEMM that dominates the step's kernel time (eqh_gemm_tf32x3, [rows, C] x [C, C],
    forward layout): algorithmic flops 2*M*N*K (the kernel executes 3x that on the TF32 pipe to stay fp32-faithful)
    over the average launch duration (CUDA events around each launch, L2 flushed in between)."""
    import torch
    from equihgnn_b200 import ops
    A = torch.randn(rows, C, device=dev)
    W = torch.randn(C, C, device=dev) * 0.05
    for _ in range(3):
        ops.gemm_raw(A, True, W, True, rows, C, C)
    ts = []
    for _ in range(10):
        flush_l2()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        ops.gemm_raw(A, True, W, True, rows, C, C)
        s1.record()
        torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1))
    ts.sort()
    t = ts[len(ts) // 2] * 1e-3
    return {"rows": rows, "flops_per_launch": 2.0 * rows * C * C, "us_per_launch": t * 1e6, "tflops": 2.0 * rows * C * C / t / 1e12}


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
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush_l2()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            fn()
            s1.record()
            torch.cuda.synchronize()
            ts.append(s0.elapsed_time(s1))
        ts.sort()
        t = ts[len(ts) // 2] * 1e-3
        out[name] = {"bytes_per_launch": nbytes, "us_per_launch": t * 1e6, "achieved_gbs": nbytes / t / 1e9}
    return {"N": N, "E": E, "Z": Z, "C": C, "kernels": out}


def run_b200(a):
    import torch
    import torch.distributed as dist
    from equihgnn_b200 import ops, synthetic
    from equihgnn_b200.models import create_model
    from equihgnn_b200.train import Trainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    profile, B, cfg, method = WORKLOADS[a.workload]

    torch.manual_seed(0)                                   # identical initial weights on every rank (DDP broadcast)
    model = create_model(method, 1, synthetic.model_args(a.hidden)).to(dev).train()
    raw = [synthetic.make_batch(profile, B, synthetic.bench_seed(cfg, rank, s)) for s in range(a.n_batches)]
    graphed = not a.eager
    if graphed:
        # collation pads every batch to its shape bucket with an inert ghost molecule (graphed.py)
        from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch, step_input_bytes
        trainer = GraphedTrainer(model, real_graphs=B, lr=1e-4)
        host = [pad_batch(b, bucket_of(b)).pack(pin=True) for b in raw]        # one pinned buffer per collated batch
        h2d_bytes = sum(h._flat.numel() for h in host) / len(host)
    else:
        trainer = Trainer(model, lr=1e-4)
        host = [b.pin_memory() for b in raw]
        h2d_bytes = sum(h.h2d_bytes() for h in host) / len(host)
    resident = [h.pack(device=dev) if graphed else h.to(dev) for h in host]
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush_buf.fill_(1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up: every distinct batch once (captures one CUDA graph per shape bucket) ----------
    for i in range(max(a.warmup, 3, len(resident))):
        trainer.step(resident[i % len(resident)])
    barrier()

    # ---- value: K steps, inputs resident in HBM -----------------------------------------------
    def launches_now():
        return (trainer.own_launches + trainer.t) if graphed else ops.STATS["launches"]

    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    l0 = launches_now()
    for k in range(a.steps):
        flush_l2()
        ev[k][0].record()
        trainer.step(resident[k % len(resident)])
        ev[k][1].record()
    barrier()
    launches = launches_now() - l0
    ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- e2e: host batches through the public API ----------------------------------------------
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for i in range(2):
        trainer.step_from_host(host[i % len(host)], dev)
    barrier()
    wall0 = time.perf_counter()
    for k in range(a.steps):
        flush_l2()
        ev2[k][0].record()
        trainer.step_from_host(host[k % len(host)], dev)
        ev2[k][1].record()
    barrier()
    wall_e2e = time.perf_counter() - wall0
    ms_e2e = sum(s.elapsed_time(e) for s, e in ev2)

    # ---- e2e with device-side collation: the dataset is resident in HBM as a packed
    #      per-molecule store; per step the host sends the molecule ids + offsets (pinned, ~2 KB), the batch is
    #      gathered on the device straight into the graph's static inputs, then the same step, D2H loss
    store_res = None
    if graphed and method == "egnn_equihnns":
        import numpy as np
        from equihgnn_b200.collate import MoleculeStore, StoreTrainer
        mols, ys = synthetic.make_molecules(profile, 16 * B, synthetic.bench_seed(cfg, rank, 99))
        store = MoleculeStore(mols, ys, dev)
        st_tr = StoreTrainer(model, store, real_graphs=B, lr=1e-4)
        st_tr.buckets = trainer.buckets                       # same model, same captured graphs
        st_tr.exp_avg, st_tr.exp_avg_sq, st_tr.t = trainer.exp_avg, trainer.exp_avg_sq, trainer.t
        rng = np.random.default_rng(7 + rank)
        id_sets = [rng.choice(len(mols), size=B, replace=False) for _ in range(a.steps + 4)]
        inputs = [store.step_inputs(ids) for ids in id_sets]
        for i in range(4):
            st_tr.step_from_ids(id_sets[i], inputs[i])
        barrier()
        ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for k in range(a.steps):
            flush_l2()
            ev3[k][0].record()
            st_tr.step_from_ids(id_sets[4 + k], inputs[4 + k])
            ev3[k][1].record()
        barrier()
        store_res = (sum(s.elapsed_time(e) for s, e in ev3), int(inputs[0].numel() * 4))
    clocks = sampler.stop() if sampler else None

    t = torch.tensor([ms, ms_e2e, store_res[0] if store_res else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_store = float(t[0]), float(t[1]), float(t[2])

    # ---- per-kernel pass: the same step run eagerly (a graph replay cannot be bracketed per kernel),
    #      CUDA events around each launch of the segmented-reduction family, L2 flushed between steps
    roof_names = ("eqh_fused_msg_fwd", "eqh_fused_msg_bwd", "eqh_gather_reduce", "eqh_gather_rows")
    eager_step = trainer._fwd_bwd if graphed else trainer.forward_backward
    prof_steps = min(a.steps, 10)
    for i in range(2):
        eager_step(resident[i % len(resident)])
    ops.PROFILE = {n: [] for n in roof_names}
    for k in range(prof_steps):
        flush_l2()
        eager_step(resident[k % len(resident)])
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None

    # ---- roofline: (i) in-step per-kernel figures from the eager pass above, (ii) the stacked microbenchmark
    peak, peak_src = peaks()
    fam = {}
    for name, recs in prof.items():
        tot_ms, tot_bytes = 0.0, 0.0
        for s, e, nbytes in recs:
            tot_ms += s.elapsed_time(e)
            tot_bytes += nbytes
        if recs:
            fam[name] = {"launches_per_step": len(recs) / prof_steps, "us_per_launch": 1e3 * tot_ms / len(recs),
                         "achieved_gbs": tot_bytes / (tot_ms * 1e-3) / 1e9 if tot_ms > 0 else 0.0,
                         "bytes_per_launch": tot_bytes / len(recs)}
    stacked = roofline_stacked(dev, a.hidden, flush_l2) if rank == 0 else None
    tens = [roofline_tensor(dev, r, a.hidden, flush_l2) for r in (resident[0].num_atoms, 30000)] if rank == 0 else None

    if rank == 0:
        out = {
            "metric": METRIC, "value": world * B * a.steps / (ms * 1e-3), "unit": "molecules/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{method} {profile}-shaped batch {B}/GPU, C={a.hidden}, L=3 "
                                   f"(BASELINE configs[{cfg - 1}]), fwd+bwd+allreduce+Adam",
                       "distinct_batches": len(host), "l2": "flushed between timed steps (256 MB write)",
                       "execution": ("CUDA graph replay per shape bucket (%d buckets), ghost-padded batches" %
                                     len(trainer.buckets)) if graphed else "eager",
                       "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                       "parallelism": f"dp{world}"},
            "clocks": clocks,
            "e2e": {"value": world * B * a.steps / (ms_e2e * 1e-3), "unit": "molecules/s",
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / a.steps,
                    "wall_ms_per_step": 1e3 * wall_e2e / a.steps},
            "gpu_launches": launches,
        }
        if store_res:
            out["e2e_device_collation"] = {
                "value": world * B * a.steps / (ms_store * 1e-3), "unit": "molecules/s", "ms_per_step": ms_store / a.steps,
                "h2d_bytes_per_step": store_res[1], "d2h_bytes_per_step": 4,
                "note": "dataset (16 batches of molecules) resident in HBM as a packed per-molecule store; per step the "
                        "pinned molecule ids + offsets go H2D, eqh_collate gathers the ghost-padded batch into the graph's "
                        "static inputs, then the same step and the D2H loss; buckets captured on first use are timed too"}
        if stacked:
            ks = stacked["kernels"]
            head = "eqh_gather_reduce[P: scatter-mean of a materialised [Z,C] tensor by hyperedge]"
            d = ks[head]
            out["roofline"] = {
                "bound": "hbm", "kernel": "eqh_gather_reduce (deterministic segmented mean, the torch_scatter.scatter "
                                          "replacement, SURVEY.md 8(d) figure P)",
                "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": d["achieved_gbs"] / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this workload, one launch, from the
                # `ncu --set full` capture summarised in profiles/r01_segmented_reduction_stacked_ncu.md
                "traffic": NCU_TRAFFIC_BYTES if (stacked["Z"], stacked["C"]) == (808488, 256) else None,
                "traffic_source": "ncu --set full (profiles/r01_segmented_reduction_stacked_ncu.md): 832.7 MB read + 353.4 MB "
                                  "written per launch",
                "peak_source": peak_src, "us_per_launch": d["us_per_launch"],
                "bytes_per_launch": d["bytes_per_launch"],
                "workload": "stacked-batch microbenchmark inside bench.py: incidence of a Molecule3D-shaped batch of 1024 "
                            "molecules stacked 12x (N=%d, E=%d, Z=%d), C=%d; 10 launches, L2 flushed between launches, "
                            "CUDA events around the launch, median" % (stacked["N"], stacked["E"], stacked["Z"], stacked["C"]),
                "family_stacked": {k: dict(v, frac=v["achieved_gbs"] / peak) for k, v in ks.items()},
                "in_step": {"note": "same kernels inside the QM9-128 step (tensors <= 10 MB, latency-bound by construction; "
                                    "measured in an eager replay of the step with CUDA events around each launch)",
                            "family": fam}}
        if tens:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
            bf16 = float(pk.get("bf16_tflops_sustained", 1402.4))
            out["roofline_tensor"] = {
                "bound": "tensor", "kernel": "eqh_gemm_tf32x3 (fp32-faithful 3xTF32 row GEMM, [rows,C] x [C,C]): the largest "
                                             "share of the step's kernel time (profiles/r01_graph_timeline_qm9.md)",
                "achieved": tens[0]["tflops"], "peak": bf16, "unit": "TFLOP/s", "frac": tens[0]["tflops"] / bf16,
                "peak_source": "measured dense bf16, sustained (MEASURED_PEAKS.json); a TF32 MMA runs at half the bf16 rate "
                               "and the kernel issues 3 TF32 products per algorithmic product, so the ceiling of this "
                               "formulation is peak/6",
                "frac_of_3xtf32_ceiling": tens[0]["tflops"] / (bf16 / 6.0),
                "cases": {f"rows={t_['rows']}": {k: t_[k] for k in ("flops_per_launch", "us_per_launch", "tflops")} | {
                    "frac_of_3xtf32_ceiling": t_["tflops"] / (bf16 / 6.0)} for t_ in tens},
                "note": "at the QM9-128 row count the GEMM is latency bound (one 128x64 tile per CTA, 72-80 CTAs); rows=30000 "
                        "(Molecule3D-1024) runs the persistent variant; see profiles/r01_kernels_ncu_full.md for the counters"}
        if not a.no_cpu_baseline and world == 1:
            val, cores, n, dt = cpu_reference_steps(a.workload, a.hidden, 1000, a.cpu_seconds)
            out["cpu_baseline"] = {"value": val, "unit": "molecules/s", "cores": cores, "kind": "port",
                                   "sample": f"{n} full steps of batch {B} in {dt:.1f} s, oracle port (torch CPU fp32, "
                                             f"sparse EGNN restatement)"}
        print(json.dumps(out), flush=True)
    if a.breakdown and rank == 0:
        _breakdown(eager_step, resident, ops)
    if world > 1:
        dist.destroy_process_group()


def _breakdown(step_fn, resident, ops):
    import torch
    names = ["eqh_incidence_build_i64", "eqh_incidence_build_i32", "eqh_gather_reduce", "eqh_gather_rows",
             "eqh_fused_msg_fwd", "eqh_fused_msg_bwd", "eqh_colsum", "eqh_knn16", "eqh_egnn_edge_fwd", "eqh_egnn_edge_bwd",
             ]
    ops.PROFILE = {n: [] for n in names}
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record()
    reps = 5
    for i in range(reps):
        step_fn(resident[i % len(resident)])
    s1.record()
    torch.cuda.synchronize()
    total = s0.elapsed_time(s1) / reps
    prof, ops.PROFILE = ops.PROFILE, None
    print(f"[breakdown] step {total * 1e3:.0f} us (with event overhead)", file=sys.stderr)
    acc = 0.0
    for n, recs in prof.items():
        if not recs:
            continue
        t = sum(s.elapsed_time(e) for s, e, _ in recs) / reps
        acc += t
        print(f"[breakdown] {n:28s} {len(recs) / reps:5.1f} launches/step {t * 1e3:9.1f} us/step "
              f"({100 * t / total:4.1f} %)", file=sys.stderr)
    print(f"[breakdown] own kernels {acc * 1e3:.0f} us/step; remainder = torch row GEMMs / elementwise / launch gaps",
          file=sys.stderr)


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
