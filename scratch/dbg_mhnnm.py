import sys, os, traceback
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic, models
from equihgnn_b200.models import MHNNM
from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
DEV = "cuda:0"
def run(tag, **patch):
    saved = {}
    for k, v in patch.items():
        if "__" not in k: continue
        mod, name = k.split("__")
        obj = {"ops": ops, "models": models}[mod]
        saved[k] = getattr(obj, name); setattr(obj, name, v)
    try:
        torch.manual_seed(13)
        m = MHNNM(1, synthetic.model_args(hidden=64)).to(DEV).train()
        tr = GraphedTrainer(m, real_graphs=24, lr=1e-3, in_graph_step=patch.get("ingraph", True))
        b = synthetic.make_batch("qm9", 24, 700)
        l = tr.step_from_host(pad_batch(b, bucket_of(b, (64, 64, 128))).pack(pin=True), DEV)
        print(tag, "OK", l)
    except Exception as e:
        print(tag, "FAIL", str(e).splitlines()[0])
        try: torch.cuda.synchronize()
        except Exception: pass
    finally:
        for k, v in saved.items():
            if "__" not in k: continue
            mod, name = k.split("__")
            setattr({"ops": ops, "models": models}[mod], name, v)
which = sys.argv[1]
if which == "base": run("base")
if which == "nogroup": run("nogroup", ops__WGRAD_GROUPED=False)
if which == "nobn": run("nobn", models___batch_norm=lambda x, bn, n, mask: bn(x))
if which == "nobond": run("nobond", models__bond_encode=lambda emb, ea: emb(ea.reshape(-1)))
if which == "anomaly":
    torch.autograd.set_detect_anomaly(True)
    run("anomaly")
if which == "noside":
    ops.side_work.enabled = False
    run("noside")
if which == "outgraph": run("outgraph", ingraph=False)
if which == "trace":
    import threading
    orig = ops._call
    state = {"on": False}
    def traced(name, n, *a, **k):
        if state["on"]:
            print(f"  {name:28s} capturing={torch.cuda.is_current_stream_capturing()} stream={torch.cuda.current_stream().cuda_stream:#x} thread={threading.get_ident() % 10000}", flush=True)
        return orig(name, n, *a, **k)
    ops._call = traced
    import equihgnn_b200.layers.fused_conv as fc, equihgnn_b200.layers.fused_egnn as fe
    fc._call = traced; fe._call = traced
    from equihgnn_b200 import graphed
    oc = graphed.GraphedTrainer._capture
    def cap(self, like):
        orig_body = self._body
        calls = {"n": 0}
        def body(d):
            calls["n"] += 1
            state["on"] = calls["n"] == 4          # the captured call (after 3 warm-ups)
            try: return orig_body(d)
            finally: state["on"] = False
        self._body = body
        try: return oc(self, like)
        finally: self._body = orig_body
    graphed.GraphedTrainer._capture = cap
    run("trace")
if which == "trace2":
    import threading
    oj = ops._SideWork.join
    def join(self, device=None, stream=None):
        cur = torch.cuda.current_stream()
        def cap(st):
            with torch.cuda.stream(st): return torch.cuda.is_current_stream_capturing()
        pend = {k: len(v) for k, v in self.pending.items()}
        print(f"  JOIN pending={pend} target={'given' if stream is not None else 'current'} "
              f"target_stream={(stream or cur).cuda_stream:#x} target_capturing={cap(stream or cur)} "
              f"side={[ (k, hex(s.cuda_stream), cap(s)) for k, s in self.streams.items()]} thread={threading.get_ident() % 10000}", flush=True)
        try:
            return oj(self, device, stream)
        except Exception as e:
            print("  JOIN RAISED", str(e).splitlines()[0], flush=True)
            raise
    ops._SideWork.join = join
    for n, p in []: pass
    run("trace2")
if which == "hooks":
    torch.manual_seed(13)
    m = MHNNM(1, synthetic.model_args(hidden=64)).to(DEV).train()
    tr = GraphedTrainer(m, real_graphs=24, lr=1e-3)
    seen = set()
    for n, p in m.named_parameters():
        p.register_hook(lambda g, n=n: (seen.add(n), None)[1])
    b = synthetic.make_batch("qm9", 24, 700)
    try:
        tr.step_from_host(pad_batch(b, bucket_of(b, (64, 64, 128))).pack(pin=True), DEV)
    except Exception as e:
        print("FAIL", str(e).splitlines()[0])
    print("params that received a tensor gradient through autograd:", sorted(seen))
