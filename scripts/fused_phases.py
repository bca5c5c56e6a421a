"""Phase timeline (ns, CTA 0) of the fused encoder kernels.  usage: B4R_FUSED_DEBUG=1 python scripts/fused_phases.py [workload] [train]"""
import os, sys
os.environ.setdefault("B4R_FUSED_DEBUG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from bert4rec_b200.engine import ParamStore

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
train = (sys.argv[2] if len(sys.argv) > 2 else "1") == "1"
w = bench.WORKLOADS[wl]
store = ParamStore(**{k: w[k] for k in bench.ENC_KEYS}, device="cuda:0")
store.init_weights(0)
store.ensure_training_buffers()
b = {k: v.cuda() for k, v in bench.synth_batches(w, 1, seed=0)[0].items()}
sess = store.session(w["batch"], w["seq_len"], w["max_pred"])
which = sys.argv[3] if len(sys.argv) > 3 else "fwd"
for it in range(3):
    sess.encode(b["input_word_ids"], b["input_mask"], training=train, seed=1, step=it)
    if which == "bwd":
        torch.cuda.synchronize()
        sess._view(store.lib.b4r_debug_buffer2(sess.h), (512,), torch.int64).zero_()
        sess.select(b["masked_lm_positions"], b["masked_lm_ids"], b["masked_lm_weights"], mode=0, want_aux=True)
        sess.transform(); sess.loss(); sess.backward(seed=1, step=it)
torch.cuda.synchronize()
p = store.lib.b4r_debug_buffer2(sess.h)
ts = sess._view(p, (512,), torch.int64).cpu().tolist()
cta = [(ts[256 + 2 * i], ts[257 + 2 * i]) for i in range(128) if ts[256 + 2 * i]]
ts = ts[:256]
n = max(i for i, v in enumerate(ts) if v) + 1
ts = ts[:n]
if cta:
    t0 = min(c[0] for c in cta)
    print("CTAs", len(cta), "start spread us", (max(c[0] for c in cta) - t0) / 1e3, "last end us", (max(c[1] for c in cta) - t0) / 1e3,
          "mean dur us", sum(c[1] - c[0] for c in cta) / len(cta) / 1e3)
if which == "bwd":
    allts = sess._view(p, (512,), torch.int64).cpu().tolist()
    t0 = min(v for v in allts if v)
    print("per-warp phase stamps (us since first stamp), CTA 0: rows = warps (quad = w & 3, part = w >> 2)")
    for w in range(16):
        r = [v for v in allts[w * 32:(w + 1) * 32] if v]
        print(f"w{w:2d} " + " ".join(f"{(v - t0) / 1e3:6.2f}" for v in r[:14]))
    sys.exit(0)
print("stamps", n, "total us", (ts[-1] - ts[0]) / 1e3)
print(" ".join(f"{(ts[i + 1] - ts[i]) / 1e3:.2f}" for i in range(n - 1)))
