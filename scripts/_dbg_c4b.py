import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from bert4rec_b200 import trainers
from bert4rec_b200.models import BERT4RecModel
from bert4rec_b200.models.components import networks
w = dict(bench.WORKLOADS["c4"]); w["batch"] = int(sys.argv[1]); graph = int(sys.argv[2])
enc = networks.Bert4RecEncoder(**{k: w[k] for k in bench.ENC_KEYS}, device="cuda:0", seed=0)
model = BERT4RecModel(enc)
trainers.get("bert4rec", model=model).initialize_model()
model.use_cuda_graph = bool(graph)
batches = [{k: v.cuda() for k, v in b.items()} for b in bench.synth_batches(w, 1, seed=0)]
for i in range(4):
    model.train_step(batches[0])
    torch.cuda.synchronize(); print("step", i, time.time(), flush=True)
