import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, torch.distributed as dist, torch.multiprocessing as mp
def log(rank, *a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)
def worker(rank, world, port, mode):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    log(rank, "pg up")
    from equihgnn_b200 import synthetic
    from equihgnn_b200.models import EGNNEquiHNNS
    from equihgnn_b200.graphed import GraphedTrainer, common_bucket, pad_batch
    C, B = 64, 24
    batches = [synthetic.make_batch("qm9", B, 800 + 10 * rank + i) for i in range(3)]
    bucket = common_bucket(batches, (64, 64, 128))
    log(rank, "bucket", bucket)
    host = [pad_batch(b, bucket).pack(pin=True) for b in batches]
    args = synthetic.model_args(hidden=C)
    def fresh():
        torch.manual_seed(33)
        return EGNNEquiHNNS(1, args).to(dev).train()
    a = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, uniform_shapes=True)
    b = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, in_graph_step=False) if mode == "both" else None
    for step in range(4):
        la = a.step_from_host(host[step % 3], dev)
        log(rank, "a step", step, la)
        if b is not None:
            lb = b.step_from_host(host[step % 3], dev)
            log(rank, "b step", step, lb)
    torch.cuda.synchronize()
    log(rank, "done steps")
    a.release()
    if b is not None: b.release()
    log(rank, "released")
    dist.barrier()
    log(rank, "barrier ok")
    dist.destroy_process_group()
    log(rank, "destroyed")
if __name__ == "__main__":
    mp.spawn(worker, args=(2, 29533, sys.argv[1]), nprocs=2, join=True)
    print("MAIN DONE")
