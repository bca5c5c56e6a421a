"""Per-pair timeline (CTA 0, first 32 pairs) of the one-pass CE backward.  usage: B4R_CF_DEBUG=1 python scripts/ce_bwd_phases.py [workload]"""
import os, sys
os.environ.setdefault("B4R_CF_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bert4rec_b200.engine import ParamStore

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]
store = ParamStore(**{k: w[k] for k in bench.ENC_KEYS}, device="cuda:0")
store.init_weights(0)
store.ensure_training_buffers()
b = {k: v.cuda() for k, v in bench.synth_batches(w, 1, seed=0)[0].items()}
sess = store.session(w["batch"], w["seq_len"], w["max_pred"])
for it in range(3):
    sess.select(b["masked_lm_positions"], b["masked_lm_ids"], b["masked_lm_weights"], mode=0, want_aux=True)
    sess.encode(b["input_word_ids"], b["input_mask"], training=True, seed=1, step=it)
    sess.transform(); sess.loss()
    torch.cuda.synchronize()
    sess._view(store.lib.b4r_debug_buffer2(sess.h), (512,), torch.int64).zero_()
    sess.backward(seed=1, step=it)
torch.cuda.synchronize()
ts = sess._view(store.lib.b4r_debug_buffer2(sess.h), (512,), torch.int64).cpu().tolist()
t0 = min(v for v in ts if v)
us = lambda v: f"{(v - t0) / 1e3:7.2f}" if v else "      -"
print("pair | epilogue group (s_full, pre dl_empty wait, post wait, dl_full arrive) | dT issuer (dl_full seen, issued, committed) | S issuer (s_empty seen, committed)")
for p in range(32):
    g = p & 1
    e = ts[g * 128 + p * 4: g * 128 + p * 4 + 4]
    m = ts[2 * 128 + p * 4: 2 * 128 + p * 4 + 3]
    s1 = ts[3 * 128 + p * 4: 3 * 128 + p * 4 + 2]
    print(f"{p:3d} g{g} | " + " ".join(us(v) for v in e) + " | " + " ".join(us(v) for v in m) + " | " + " ".join(us(v) for v in s1))
print("item | tail start, acc_done seen, dT drained, dE drained + arrive | (next) vectors loaded (this item's pair loop starts)")
for it in range(4):
    a = ts[3 * 128 + it * 4 + 2], ts[3 * 128 + it * 4 + 3], ts[2 * 128 + it * 4 + 3], ts[3 * 128 + (it + 16) * 4 + 2]
    print(f"{it:3d} | " + " ".join(us(v) for v in a) + " | " + us(ts[3 * 128 + (it + 16) * 4 + 3]))
