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
            mod, name = k.split("__")
            setattr({"ops": ops, "models": models}[mod], name, v)
which = sys.argv[1]
if which == "base": run("base")
if which == "nogroup": run("nogroup", ops__WGRAD_GROUPED=False)
if which == "nobn": run("nobn", models___batch_norm=lambda x, bn, n, mask: bn(x))
if which == "nobond": run("nobond", models__bond_encode=lambda emb, ea: emb(ea.reshape(-1)))
