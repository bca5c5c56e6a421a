"""Per-kernel CUDA-event breakdown of one train step (b4r_profile_*). usage: python scripts/profile_step.py [workload] [steps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bert4rec_b200 import trainers
from bert4rec_b200.models import BERT4RecModel
from bert4rec_b200.models.components import networks

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w = bench.WORKLOADS[wl]
enc = networks.Bert4RecEncoder(**{k: w[k] for k in bench.ENC_KEYS}, device="cuda:0", seed=0)
model = BERT4RecModel(enc)
trainers.get("bert4rec", model=model).initialize_model()
batches = [{k: v.cuda() for k, v in b.items()} for b in bench.synth_batches(w, 2, seed=0)]
sess = model.store.session(w["batch"], w["seq_len"], w["max_pred"])
for i in range(5):
    model.train_step(batches[i % 2])
torch.cuda.synchronize()
# un-profiled timing
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for i in range(n):
    model.train_step(batches[i % 2])
ev1.record(); torch.cuda.synchronize()
print(f"{wl}: {ev0.elapsed_time(ev1) / n * 1e3:.1f} us/step back-to-back (no flush), {w['batch'] * n / ev0.elapsed_time(ev1) * 1e3:.0f} seq/s")
rep = bench.profile_steps(model, sess, batches, n)
rows = sorted(rep.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for v in rep.values())
print(f"sum of kernel times inside the graph: {tot / n * 1e3:.1f} us/step over {sum(v[0] for v in rep.values()) / n:.0f} launches")
n_rows = int(sess.counts()[1])
for tag, (cnt, ms) in rows:
    fl, by, bound = bench.kernel_work(tag, w, n_rows)
    avg = ms / cnt
    print(f"{tag:28s} x{cnt / n:4.1f}  avg {avg * 1e3:8.1f} us  per-step {ms / n * 1e3:8.1f} us  {100 * ms / tot:5.1f}%  "
          f"{fl / (avg / 1e3) / 1e12 if fl else 0:7.1f} TF/s {by / (avg / 1e3) / 1e9 if by else 0:8.1f} GB/s")
