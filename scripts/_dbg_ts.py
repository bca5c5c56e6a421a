import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from bert4rec_b200.engine import ParamStore
from bert4rec_b200 import _lib
w = dict(bench.WORKLOADS[sys.argv[1]]); B = w["batch"]
kw = {k: w[k] for k in bench.ENC_KEYS}
store = ParamStore(device="cuda:0", **kw); store.init_weights(0); store.ensure_training_buffers()
b = {k: v.cuda() for k, v in bench.synth_batches(w, 1, seed=0)[0].items()}
sess = store.session(B, w["seq_len"], w["max_pred"])
for it in range(2):
    sess.select(b["masked_lm_positions"], b["masked_lm_ids"], b["masked_lm_weights"], mode=0, want_aux=True)
    sess.encode(b["input_word_ids"], b["input_mask"], training=True, seed=1, step=0)
    sess.transform(); sess.loss(); sess.backward(seed=1, step=0)
torch.cuda.synchronize()
_lib.load().b4r_fattn_dump_ts()
