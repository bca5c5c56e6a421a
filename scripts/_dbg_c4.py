import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from bert4rec_b200.engine import ParamStore
w = dict(bench.WORKLOADS["c4"]); B = int(sys.argv[1]); mode = sys.argv[2]
kw = {k: w[k] for k in bench.ENC_KEYS}
store = ParamStore(device="cuda:0", **kw); store.init_weights(0); store.ensure_training_buffers()
w["batch"] = B
b = {k: v.cuda() for k, v in bench.synth_batches(w, 1, seed=0)[0].items()}
sess = store.session(B, w["seq_len"], w["max_pred"])
def P(msg):
    torch.cuda.synchronize(); print(msg, time.time(), flush=True)
P("start " + mode)
if "f4" in mode: sess.set_flag(4, 1)
sess.select(b["masked_lm_positions"], b["masked_lm_ids"], b["masked_lm_weights"], mode=0, want_aux=True)
if "f4" in mode: sess.set_flag(4, 0)
ctr = store.step_counter if "ctr" in mode else None
seed = 0x5EEDB4A7 if "seed" in mode else 1
sess.encode(b["input_word_ids"], b["input_mask"], training=True, seed=seed, step=0, step_counter=ctr); P("encode")
