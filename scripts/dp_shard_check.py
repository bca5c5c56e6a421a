"""Vocabulary-sharded training over NCCL vs plain data-parallel training, same seeds (SURVEY 8e large catalogues).
usage: torchrun --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_shard_check.py [vocab] [hidden] [steps] [batch_per_gpu]
Both modes see the same per-rank batches; the loss trajectories and the final weights must agree to float tolerance."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from bert4rec_b200 import trainers
from bert4rec_b200.models import BERT4RecModel
from bert4rec_b200.models.components import networks

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
V = int(sys.argv[1]) if len(sys.argv) > 1 else 50003
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = dict(bench.WORKLOADS["c2"], vocab_size=V, hidden_size=H, num_attention_heads=max(2, H // 64), inner_dim=H if H == 64 else 4 * H,
         batch=int(sys.argv[4]) if len(sys.argv) > 4 else 64)
batches = [{k: v.cuda() for k, v in b.items()} for b in bench.synth_batches(w, steps, seed=100 + rank)]


def run(sharded):
    enc = networks.Bert4RecEncoder(**{k: w[k] for k in bench.ENC_KEYS}, device=f"cuda:{local}", seed=0)
    model = BERT4RecModel(enc)
    trainers.get("bert4rec", model=model).initialize_model()
    assert model.distributed
    model.vocab_sharded = sharded
    p0 = model.store.params.clone()
    losses, t = [], []
    for b in batches:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        model.reset_metrics()
        m = model.train_step(b)
        losses.append(m["loss"])
        t.append(time.perf_counter() - t0)
    return losses, model.store.params - p0, sorted(t)[len(t) // 2]


la, pa, ta = run(False)
lb, pb, tb = run(True)
diff = float((pa - pb).norm() / pa.norm())   # of the weight UPDATES (Adam's m/sqrt(v) amplifies rounding of near-zero gradients)
if rank == 0:
    print(f"world {world} V {V} H {H}: data-parallel losses {['%.5f' % x for x in la]}")
    print(f"                      vocab-sharded losses {['%.5f' % x for x in lb]}")
    print(f"relative difference of the weight updates after {steps} steps: {diff:.3e}; median step (eager, incl. metric read) dp {ta * 1e3:.2f} ms, sharded {tb * 1e3:.2f} ms")
ok = all(abs(a - b) < 2e-3 * abs(a) for a, b in zip(la, lb)) and diff < 2e-2
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
