"""A few eager (non-graph) train steps + eval passes of one workload, for `ncu` launch lists / captures.
usage: python scripts/ncu_step.py [workload] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bert4rec_b200 import trainers
from bert4rec_b200.models import BERT4RecModel
from bert4rec_b200.models.components import networks

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
w = bench.WORKLOADS[wl]
enc = networks.Bert4RecEncoder(**{k: w[k] for k in bench.ENC_KEYS}, device="cuda:0", seed=0)
model = BERT4RecModel(enc)
trainers.get("bert4rec", model=model).initialize_model()
model.use_cuda_graph = False
batches = [{k: v.cuda() for k, v in b.items()} for b in bench.synth_batches(w, 2, seed=0)]
for i in range(n):
    model.train_step(batches[i % 2])
torch.cuda.synchronize()
print("done", wl, n)
