"""Micro-benchmark of the CE forward kernel alone. usage: python scripts/bench_ce.py [workload]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from bert4rec_b200.engine import ParamStore
wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
w = bench.WORKLOADS[wl]
store = ParamStore(**{k: w[k] for k in bench.ENC_KEYS}, device="cuda:0")
store.init_weights(0)
b = {k: v.cuda() for k, v in bench.synth_batches(w, 1, seed=0)[0].items()}
sess = store.session(w["batch"], w["seq_len"], w["max_pred"])
sess.encode(b["input_word_ids"], b["input_mask"], training=False)
sess.select(b["masked_lm_positions"], b["masked_lm_ids"], b["masked_lm_weights"], mode=0, want_aux=True)
sess.transform()
for flag in (0, 1):
    sess.set_flag(1, flag)
    for _ in range(5): sess.loss()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    e0.record()
    for _ in range(n): sess.loss()
    e1.record(); torch.cuda.synchronize()
    rows = int(sess.counts()[1])
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * rows * w["hidden_size"] * w["vocab_size"]
    print(f"{wl} umma={flag} debug={os.environ.get('B4R_CE_DEBUG','0')}: loss call (ce_fwd+finalize) {us:.1f} us, rows {rows}, {fl / us / 1e6:.1f} TF/s, "
          f"loss {float(sess.step_stats()[0] / sess.step_stats()[1]):.5f}")
if int(os.environ.get("B4R_CE_DEBUG", "0")) & 8:
    sess.set_flag(1, 1)
    sess.loss(); torch.cuda.synchronize()
    p = sess.lib.b4r_debug_buffer(sess.h)
    t = sess._view(p, (3, 16, 8), torch.int64).cpu()
    t0 = int(t[t > 0].min())
    names = {0: ["empty_ok"], 1: ["tempty_ok", "full_ok", "issued"], 2: ["top", "bar_ok", "tfull_ok", "epi_done", "arrived"]}
    for role in (0, 1, 2):
        print("role", ["producer", "mma", "epilogue"][role], names[role])
        for i in range(10):
            print("  tile", i, [int(t[role, i, k]) - t0 for k in range(len(names[role]))])
