#!/usr/bin/env python
"""bench.py -- train masked-seq/s (+ 100-negative eval seq/s) of the BERT4Rec hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

Contract (see the task statement): one JSON line on stdout from rank 0.  ``value`` = whole-job train
masked-seq/s with the step's inputs already resident in HBM; ``e2e`` = the same metric through the public API
(``BERT4RecModel.train_step``) with HOST (pinned) input buffers, the H2D copy and a D2H read of the loss inside the
timed region.  ``--impl reference`` times the CPU restatement of the reference's TF2 path (``oracle/``; TensorFlow
itself is not installable here) on the host cores for the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload table: BASELINE.json configs (SURVEY.md section 8 sizes)
WORKLOADS = {
    # C1: the reference's CPU-runnable case (ml-1m_64.json)
    "c1": dict(name="C1 ML-1m shape", vocab_size=3709, hidden_size=64, num_layers=2, num_attention_heads=2,
               max_sequence_length=200, inner_dim=256, output_dropout=0.2, attention_dropout=0.2,
               batch=256, seq_len=200, max_pred=40, mask_prob=0.2),
    # C2: Beauty shape (beauty_64.json; masked_lm_prob 0.15 per BASELINE.json) -- the headline single-GPU config
    "c2": dict(name="C2 Beauty shape", vocab_size=12004, hidden_size=64, num_layers=2, num_attention_heads=2,
               max_sequence_length=50, inner_dim=64, output_dropout=0.5, attention_dropout=0.2,
               batch=256, seq_len=50, max_pred=30, mask_prob=0.15),
    # C3: ML-20m shape, per-GPU batch 256 (data-parallel runs)
    "c3": dict(name="C3 ML-20m shape", vocab_size=26732, hidden_size=64, num_layers=2, num_attention_heads=2,
               max_sequence_length=200, inner_dim=256, output_dropout=0.1, attention_dropout=0.1,
               batch=256, seq_len=200, max_pred=40, mask_prob=0.2),
    # C4: scaled model
    "c4": dict(name="C4 scaled H256 L4", vocab_size=13047, hidden_size=256, num_layers=4, num_attention_heads=4,
               max_sequence_length=200, inner_dim=1024, output_dropout=0.1, attention_dropout=0.1,
               batch=1024, seq_len=200, max_pred=40, mask_prob=0.2),
    # C5: 1M-item synthetic catalogue, tied output embedding vocabulary-sharded over the ranks (C4's encoder)
    "c5": dict(name="C5 1M-item catalogue, vocab-sharded", vocab_size=1000003, hidden_size=256, num_layers=4, num_attention_heads=4,
               max_sequence_length=200, inner_dim=1024, output_dropout=0.1, attention_dropout=0.1,
               batch=1024, seq_len=200, max_pred=40, mask_prob=0.2),
}
ENC_KEYS = ("vocab_size", "hidden_size", "num_layers", "num_attention_heads", "max_sequence_length", "inner_dim",
            "output_dropout", "attention_dropout")


def synth_batches(w, n_batches, seed=0, eval_mode=False):
    """Seeded synthetic batches: item ids ~ truncated Zipf(1.1) over [3, V), dense (full-length) sequences, Cloze
    masking by the product's bit-exact restatement of apply_dynamic_masking_task (seed = sequence index)."""
    import torch
    from bert4rec_b200.dataloaders import dataloader_utils as du
    rng = np.random.RandomState(seed)
    V, S, P, B = w["vocab_size"], w["seq_len"], w["max_pred"], w["batch"]
    out = []
    for nb in range(n_batches):
        feats = {k: [] for k in ("labels", "input_word_ids", "input_mask", "masked_lm_ids", "masked_lm_positions", "masked_lm_weights")}
        for b in range(B):
            seq = ((rng.zipf(1.1, size=S) - 1) % (V - 3) + 3).astype(np.int64)
            labels = seq.copy()
            if eval_mode:
                ids, pos, lab = du.mask_last_token_only(seq.copy(), 1)
            else:
                ids, pos, lab = du.apply_dynamic_masking_task(seq, P, 1, [2, 0], V, selection_rate=w["mask_prob"],
                                                              mask_token_rate=1.0, random_token_rate=0.0,
                                                              seed=nb * B + b)
            k = P - len(lab)
            feats["labels"].append(labels)
            feats["input_word_ids"].append(ids)
            feats["input_mask"].append(np.ones(S, dtype=np.int64))
            feats["masked_lm_ids"].append(np.pad(lab, (0, k)))
            feats["masked_lm_positions"].append(np.pad(pos, (0, k)))
            feats["masked_lm_weights"].append(np.pad(np.ones_like(lab), (0, k)))
        out.append({k: torch.from_numpy(np.stack(v).astype(np.int64)) for k, v in feats.items()})
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def workload_config(w, world, m_valid, sample_note=None):
    """The `config` object both arms report (same workload keys; the arm-specific notes are extra keys)."""
    c = {"workload": w["name"], **{k: w[k] for k in ENC_KEYS}, "batch_per_gpu": w["batch"], "seq_len": w["seq_len"],
         "max_pred": w["max_pred"], "mask_prob": w["mask_prob"],
         "sequences": "dense (full length), Zipf(1.1) item ids", "parallelism": f"dp{world}"}
    if m_valid is not None:
        c["valid_masked_slots_per_batch"] = m_valid
    if sample_note:
        c["sample"] = sample_note
    return c


def train_flops(w, m_valid):
    """Algorithmic FLOPs of one train step: 3 x (encoder + MLM transform + tied projection)."""
    T = w["batch"] * w["seq_len"]
    H, I, S, L, V = w["hidden_size"], w["inner_dim"], w["seq_len"], w["num_layers"], w["vocab_size"]
    fwd = L * T * (8 * H * H + 4 * S * H + 4 * H * I) + 2 * m_valid * H * H + 2 * m_valid * H * V
    return 3 * fwd


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample_batch(w):
    """Bounded CPU sample of the workload: whole batches of the small configs, a slice of the big ones (seconds per step)."""
    tok = w["batch"] * w["seq_len"] * w["hidden_size"] * w["num_layers"]
    if tok <= 256 * 50 * 64 * 2:
        return w["batch"]
    return 32 if w["hidden_size"] >= 256 else 64


def cpu_reference(w, steps, warmup, sample_batch=None):
    """The reference's TF2 path restated on the CPU (oracle/model.py), fp32, all host threads."""
    import torch
    from oracle import model as om
    torch.set_num_threads(os.cpu_count())
    b = dict(w)
    if sample_batch:
        b["batch"] = sample_batch
    cfg = om.Config(**{k: w[k] for k in ENC_KEYS})
    params = om.init_params(cfg, 0)
    opt = om.AdamW({k: v for k, v in params.items() if not k.startswith("pooler")})
    batches = synth_batches(b, 2, seed=0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        om.train_step(params, cfg, batches[i % 2], opt, training=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ms = float(np.mean(times) * 1e3)
    return b["batch"] / (ms / 1e3), ms, b["batch"]


def cpu_eval_reference(w, sample_batch, reps=2):
    """The reference's evaluation path on the CPU: BERT4RecModel.rank_items (full [B,P,V] logits, gather of the 101 candidates,
    stable descending sort per slot) + the rank lookup of BERT4RecEvaluator.evaluate_batch; prebuilt candidates."""
    import torch
    from oracle import model as om, host_ops
    torch.set_num_threads(os.cpu_count())
    b = dict(w); b["batch"] = sample_batch
    cfg = om.Config(**{k: w[k] for k in ENC_KEYS})
    params = om.init_params(cfg, 0)
    batch = synth_batches(b, 1, seed=1000, eval_mode=True)[0]
    gt = batch["masked_lm_ids"][:, 0].numpy()
    cand = np.random.RandomState(5).randint(3, w["vocab_size"], size=(sample_batch, 101)).astype(np.int64)
    cand[:, 100] = gt
    items = [[cand[i].tolist()] for i in range(sample_batch)]
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        rk = om.rank_items(params, cfg, batch, items)
        _ = [host_ops.rank_of(rk[i][0].numpy(), int(gt[i])) for i in range(sample_batch)]
        ts.append(time.perf_counter() - t0)
    t = float(np.min(ts))
    return sample_batch / t, t * 1e3


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    sample = cpu_sample_batch(w)
    value, ms, bs = cpu_reference(w, args.steps, args.warmup, sample_batch=sample)
    line = {"impl": "reference", "metric": "train masked-seq/s", "value": value, "unit": "seq/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(w, args.gpus, None, sample_note=f"CPU arm: bounded sample, batch {bs} per step on rank 0"),
            "cpu_baseline": {"value": value, "unit": "seq/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} train steps of batch {bs} (CPU restatement of the TF2 path, torch fp32, {cores} threads)"},
            "e2e": {"value": value, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ B200 arm
def build_model(w, dev, dropout=None):
    from bert4rec_b200 import trainers
    from bert4rec_b200.models import BERT4RecModel
    from bert4rec_b200.models.components import networks
    kw = {k: w[k] for k in ENC_KEYS}
    if dropout is not None:
        kw.update(output_dropout=dropout, attention_dropout=dropout)
    model = BERT4RecModel(networks.Bert4RecEncoder(**kw, device=dev, seed=0))
    trainers.get("bert4rec", model=model).initialize_model()  # AdamW defaults, masked CE, [sparse_categorical_accuracy, masked_accuracy]
    return model


def working_set_bytes(model, sess):
    return int(sess.ws.numel()) + int(model.store.params.numel()) * 18


def measure(args, w, dev, world, rank, min_seconds, with_eval=True, clocks=None):
    """Timed train steps (device-resident inputs -> `value`; pinned host inputs + D2H loss read, wall clock -> `e2e`) and the
    100-negative evaluation of one workload.  Every rank runs this; numbers are the max over ranks."""
    import torch
    import torch.distributed as dist
    from bert4rec_b200.models import BERT4RecModel
    model = build_model(w, dev)
    B, S, P = w["batch"], w["seq_len"], w["max_pred"]
    n_b = 4
    host_batches = synth_batches(w, n_b, seed=rank)
    for b in host_batches:
        for k in b:
            b[k] = b[k].pin_memory()
    dev_batches = [{k: v.to(dev) for k, v in b.items()} for b in host_batches]
    m_valid = int((host_batches[0]["masked_lm_ids"] != 0).sum())
    sess = model.store.session(B, S, P)
    l2_bytes = 126 << 20
    big = working_set_bytes(model, sess) > 4 * l2_bytes       # the step's working set dwarfs L2: nothing to flush
    flush = None if big else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def rmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # launches per step: counted once on an eager (non-graph) step; a graph replay re-issues the same kernels
    model.use_cuda_graph = False
    model.train_step(dev_batches[0])
    l_a = sess.launch_count()
    model.train_step(dev_batches[0])
    launches_per_step = sess.launch_count() - l_a + 2   # + sqnorm and adamw kernels
    model.use_cuda_graph = not args.no_graph

    for i in range(args.warmup):
        model.train_step(dev_batches[i % n_b])
    barrier()
    # pilot: how many passes of --steps make the timed region at least `min_seconds`
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        model.train_step(dev_batches[i % n_b])
    e1.record(); torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) / 3e3, 1e-6)
    reps = int(rmax(max(1.0, np.ceil(min_seconds / (est * args.steps)))))
    n_timed = reps * args.steps

    # ---- device-resident inputs: CUDA events per step on the launching stream (the L2 flush, where used, is outside them)
    if clocks is not None:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_timed)]
    barrier()
    for i in range(n_timed):
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev[i][0].record()
        model.train_step(dev_batches[i % n_b])
        ev[i][1].record()
    barrier()
    clk = clocks.stop() if clocks is not None else None
    total_ms = rmax(sum(a.elapsed_time(b) for a, b in ev))

    # ---- end to end: pinned host batches through BERT4RecModel.train_step, H2D copy + D2H loss read inside; host wall clock,
    # started after the (untimed) flush has completed
    for i in range(3):
        _ = model.train_step(host_batches[i % n_b])["loss"]
    barrier()
    wall = 0.0
    for i in range(n_timed):
        if flush is not None:
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        _ = model.train_step(host_batches[i % n_b])["loss"]     # D2H read of the running loss (synchronises)
        wall += time.perf_counter() - t0
    e2e_ms = rmax(wall * 1e3)
    res = {"model": model, "sess": sess, "dev_batches": dev_batches, "host_batches": host_batches, "m_valid": m_valid,
           "ms_per_step": total_ms / n_timed, "value": world * B * n_timed / (total_ms / 1e3),
           "e2e_value": world * B * n_timed / (e2e_ms / 1e3), "e2e_ms_per_step": e2e_ms / n_timed, "timed_steps": n_timed,
           "launches": launches_per_step * n_timed, "launches_per_step": launches_per_step, "clocks": clk,
           "l2": ("working set %.1f GB per step >> 126 MB L2: no flush" % (working_set_bytes(model, sess) / 1e9)) if big
                 else "256 MiB buffer written between timed steps (untimed)",
           "h2d": BERT4RecModel._bytes_of(("input_word_ids", "input_mask", "masked_lm_positions", "masked_lm_ids", "masked_lm_weights"), host_batches[0])}
    if with_eval:
        res["eval"] = measure_eval(args, w, model, dev, world, rank, flush, barrier, rmax, min_seconds / 4)
    return res


def measure_eval(args, w, model, dev, world, rank, flush, barrier, rmax, min_seconds):
    """100-negative evaluation (leave-one-out, prebuilt [B,101] candidates, ground truth last): sequences ranked per second."""
    import torch
    B, V = w["batch"], w["vocab_size"]
    eval_batches = synth_batches(w, 2, seed=1000 + rank, eval_mode=True)
    rng = np.random.RandomState(5)
    items = []
    for b in eval_batches:
        gt = b["masked_lm_ids"][:, 0].numpy()
        cand = rng.randint(3, V, size=(B, 101)).astype(np.int64)
        cand[:, 100] = gt
        items.append((b, {k: v.to(dev) for k, v in b.items()}, torch.from_numpy(cand).pin_memory(), torch.from_numpy(cand).to(dev),
                      torch.from_numpy(gt.astype(np.int64)).pin_memory(), torch.from_numpy(gt.astype(np.int64)).to(dev)))
    for i in range(max(args.warmup, 3)):
        hb, db, hc, dc, hg, dg = items[i % 2]
        model.rank_candidates(db, dc, dg)[1].cpu()
        model.rank_candidates(hb, hc, hg)[1].cpu()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3):
        model.rank_candidates(items[i % 2][1], items[i % 2][3], items[i % 2][5])
    e1.record(); torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) / 3e3, 1e-6)
    n = int(rmax(max(args.steps, np.ceil(min_seconds / est))))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    barrier()
    for i in range(n):
        hb, db, hc, dc, hg, dg = items[i % 2]
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev[i][0].record()
        model.rank_candidates(db, dc, dg)
        ev[i][1].record()
    barrier()
    dev_ms = rmax(sum(a.elapsed_time(b) for a, b in ev))
    wall = 0.0
    for i in range(n):
        hb, db, hc, dc, hg, dg = items[i % 2]
        if flush is not None:
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.rank_candidates(hb, hc, hg)[1].cpu()          # ranks back on the host (synchronises)
        wall += time.perf_counter() - t0
    e2e_ms = rmax(wall * 1e3)
    return {"metric": "100-neg eval seq/s", "unit": "seq/s", "value": world * B * n / (dev_ms / 1e3),
            "e2e": {"value": world * B * n / (e2e_ms / 1e3), "unit": "seq/s",
                    "h2d_bytes_per_step": int(8 * (2 * B * w["seq_len"] + 2 * B * w["max_pred"] + B * 101 + B)),
                    "d2h_bytes_per_step": 4 * B},
            "ms_per_batch": dev_ms / n, "batches_timed": n, "items": items}


def eval_roofline(model, w, items, hbm, tf_burst, src, n=10):
    """Per-kernel event times of the captured ranking pass; the roofline entry is its dominant kernel."""
    import torch
    sess = model.store.session(w["batch"], w["seq_len"], w["max_pred"])
    model._graphs.clear()
    sess.profile(True)
    hb, db, hc, dc, hg, dg = items[0]
    model.rank_candidates(db, dc, dg)         # eager pass + capture (with event nodes)
    torch.cuda.synchronize()
    sess.profile_report()
    agg = {}
    for _ in range(n):
        model.rank_candidates(db, dc, dg)
        for tag, (cnt, ms) in sess.profile_report().items():
            c0, m0 = agg.get(tag, (0, 0.0))
            agg[tag] = (c0 + cnt, m0 + ms)
    sess.profile(False)
    model._graphs.clear()
    return roofline_from(agg, w, w["batch"], n, hbm, tf_burst, src, eval_mode=True)


def dp_parity_probe(w, dev, world, rank, main_model):
    """Untimed hardware check of the data-parallel step (N > 1): the flat gradient AFTER BERT4RecModel._all_reduce (NCCL, SUM over
    ranks, valid-slot count riding along) against rank 0 recomputing the GLOBAL batch on one GPU.  Dropout off (the keep masks are
    keyed by the row index inside a rank's batch), same weights everywhere."""
    import torch
    import torch.distributed as dist
    B, S, P = w["batch"], w["seq_len"], w["max_pred"]
    probe = build_model(w, dev, dropout=0.0)
    probe.load_state_dict(main_model.state_dict())
    batch = {k: v.to(dev) for k, v in synth_batches(w, 1, seed=500 + rank)[0].items()}
    sess = probe.store.session(B, S, P)
    probe.store.grads.zero_()
    probe._fwd_bwd(sess, batch, probe._stats_buf("train"))
    probe._all_reduce(sess)
    torch.cuda.synchronize()
    nt = probe.store.n_trainable
    g_dp = probe.store.grads[: nt + 1].clone()
    keys = ("input_word_ids", "input_mask", "masked_lm_positions", "masked_lm_ids", "masked_lm_weights")
    gathered = {}
    for k in keys:
        buf = [torch.empty_like(batch[k]) for _ in range(world)] if rank == 0 else None
        dist.gather(batch[k], buf, dst=0)
        if rank == 0:
            gathered[k] = torch.cat(buf, 0).contiguous()
    out = None
    if rank == 0:
        os.environ["B4R_DISABLE_P2P_ALLREDUCE"] = "1"    # this model exists on rank 0 only: no collective buffer exchange
        try:
            single = build_model(w, dev, dropout=0.0)
        finally:
            del os.environ["B4R_DISABLE_P2P_ALLREDUCE"]
        single.distributed = False
        single.load_state_dict(main_model.state_dict())
        s1 = single.store.session(B * world, S, P)
        single.store.grads.zero_()
        single._fwd_bwd(s1, gathered, single._stats_buf("train"))
        torch.cuda.synchronize()
        g1 = single.store.grads[:nt]
        n1 = float(s1.step_stats()[1])
        views_dp, views_1 = single.store.tf_views(torch.cat([g_dp[:nt], single.store.grads[nt:]])), single.store.tf_views(single.store.grads)
        worst, worst_name = 0.0, ""
        gmax = max(float(v.norm()) for v in views_1.values())
        for k, v in views_1.items():
            if k.startswith("pooler"):
                continue
            err = float((views_dp[k].double() - v.double()).norm()) / (float(v.norm()) + 1e-5 * gmax)
            if err > worst:
                worst, worst_name = err, k
        out = {"max_rel_l2_over_tensors": worst, "worst_tensor": worst_name, "global_rel_l2": float((g_dp[:nt].double() - g1.double()).norm() / g1.double().norm()),
               "valid_slots_dp": float(g_dp[nt]), "valid_slots_single": n1, "global_batch": B * world,
               "what": "flat gradient after the data-parallel all-reduce (%s) vs rank 0 recomputing the global batch on one GPU (dropout off)" % (("own NVLS multimem kernel" if probe.store.p2p.get("mc") else "own two-shot kernel over NVLink peer memory") if probe.store.p2p is not None else "NCCL")}
        del single
    del probe
    torch.cuda.empty_cache()
    return out


def allreduce_time_us(model, dev, n=30):
    """Device time of the step's gradient all-reduce alone (it is issued between the two graphs of a data-parallel step and is not
    overlapped with compute, so this is the communication time a step exposes)."""
    import torch
    sess = None
    for _ in range(5):
        model._all_reduce(sess)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        model._all_reduce(sess)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def measure_c5(args, dev, world, rank, steps=5, warmup=2):
    """BASELINE config 5: V = 1 000 003, hidden 256; train step with the tied projection SHARDED by vocabulary rows over the ranks
    (softmax-CE max / sum merged over NCCL) and full-catalogue top-10 ranking with the per-shard lists merged over NCCL."""
    import torch
    import torch.distributed as dist
    w = WORKLOADS["c5"]
    model = build_model(w, dev)
    model.vocab_sharded = world > 1
    B = w["batch"]
    batches = [{k: v.to(dev) for k, v in b.items()} for b in synth_batches(w, 2, seed=rank)]
    m_valid = int((batches[0]["masked_lm_ids"] != 0).sum())
    for i in range(warmup):
        model.train_step(batches[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        model.train_step(batches[i % 2])
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    loss = float(model.train_step(batches[0])["loss"])
    ev = {k: v.to(dev) for k, v in synth_batches(w, 1, seed=1000 + rank, eval_mode=True)[0].items()}
    ids, _ = model.top_k_items(ev, 10)
    torch.cuda.synchronize()
    e0.record()
    for i in range(3):
        ids, _ = model.top_k_items(ev, 10)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ems = float(t.item()) / 3
    fl = train_flops(w, m_valid)
    out = {"workload": w["name"], "value": world * B / (ms / 1e3), "unit": "seq/s", "ms_per_step": ms, "steps": steps, "n_gpus": world,
           "vocab_sharded": bool(model.vocab_sharded), "loss_after_warmup": loss, "ln_vocab": float(np.log(w["vocab_size"])),
           "achieved_tflops_per_gpu": fl / (ms / 1e3) / 1e12,
           "full_catalogue_top10": {"value": world * B / (ems / 1e3), "unit": "seq/s", "ms_per_batch": ems,
                                    "what": "encoder forward + per-shard top-10 over the 1M-item catalogue + NCCL all-gather + key merge"},
           "config": {k: w[k] for k in ENC_KEYS}, "batch_per_gpu": B}
    del model
    torch.cuda.empty_cache()
    return out


def run_b200(args, w, secondary):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    clocks = ClockSampler(local) if rank == 0 else None
    r = measure(args, w, dev, world, rank, args.min_seconds, with_eval=True, clocks=clocks)
    model, sess = r["model"], r["sess"]
    comm = None
    if world > 1:
        comm = {"allreduce_floats": int(model.store.n_trainable + 1), "exposed_allreduce_us_per_step": allreduce_time_us(model, dev),
                "kernel": (("b4r p2p_allreduce_nvls_kernel (one pass of multimem.ld_reduce / multimem.st through the NVSwitch multicast mapping, inside the step graph)" if model.store.p2p.get("mc") else "b4r p2p_allreduce_kernel (two-shot over NVLink peer memory, inside the step graph)") if model.store.p2p is not None else "NCCL all-reduce between the forward/backward graph and the optimizer graph"), "note": "one all-reduce of the flat fp32 gradient (+ valid-slot count) per step, timed alone (back to back)"}
        comm["p2p_error"] = model.store.p2p_error()
        t = torch.tensor([comm["exposed_allreduce_us_per_step"]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        comm["exposed_allreduce_us_per_step"] = float(t.item())
        parity = dp_parity_probe(w, dev, world, rank, model)
    line = None
    if rank == 0:
        hbm, tf_burst, tf_sus, src = peaks()
        flops = train_flops(w, r["m_valid"])
        step_ms = r["ms_per_step"]
        achieved_tf = flops / (step_ms / 1e3) / 1e12
        ev = r["eval"]
        line = {
            "metric": "train masked-seq/s", "value": r["value"], "unit": "seq/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {**workload_config(w, world, r["m_valid"]), "l2": r["l2"], "cuda_graph": not args.no_graph,
                       "timed_steps": r["timed_steps"],
                       "timing": "CUDA events per step on the launching stream, %d passes of --steps (timed region >= %.1f s); e2e by host wall clock" % (r["timed_steps"] // args.steps, args.min_seconds)},
            "e2e": {"value": r["e2e_value"], "unit": "seq/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 64,
                    "ms_per_step": r["e2e_ms_per_step"], "clock": "host wall clock around BERT4RecModel.train_step + loss read"},
            "gpu_launches": r["launches"],
            "clocks": r["clocks"],
            "step_flops": {"algorithmic_flops_per_step": flops, "achieved_tflops": achieved_tf,
                           "frac_of_bf16_sustained": achieved_tf / tf_sus, "peak_source": src},
        }
    # rank-0-only passes below: no collective may be issued from here on (the other ranks are not stepping)
    model.distributed = False
    if rank == 0:
        line["roofline"] = roofline_block(model, sess, w, r["dev_batches"], args, hbm, tf_burst, src)
        ev_items = ev.pop("items")
        ev["roofline"] = eval_roofline(model, w, ev_items, hbm, tf_burst, src)
        if world == 1 and not args.no_cpu_baseline:
            sample = cpu_sample_batch(w)
            v, ms = cpu_eval_reference(w, min(sample, 32))
            ev["cpu_baseline"] = {"value": v, "unit": "seq/s", "cores": os.cpu_count(), "kind": "port",
                                  "sample": f"rank_items + rank lookup of {min(sample, 32)} sequences x 101 candidates (CPU restatement: full [B,P,V] logits, stable sort), {ms:.0f} ms"}
        line["eval"] = ev
        if comm is not None:
            line["comm"] = comm
            line["dp_parity"] = parity
        if world == 1:
            line["host_path"] = host_path_block(w, model, synth_batches(w, 1, seed=1000, eval_mode=True)[0])
        if world == 1 and not args.no_cpu_baseline:
            sample = cpu_sample_batch(w)
            v, ms, bs = cpu_reference(w, 3, 1, sample_batch=sample)
            line["cpu_baseline"] = {"value": v, "unit": "seq/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"3 train steps of batch {bs} after 1 warm-up (CPU restatement of the TF2 path, torch fp32, {os.cpu_count()} threads), {ms:.0f} ms/step"}
    else:
        r["eval"].pop("items", None)
    del model, sess, r
    torch.cuda.empty_cache()
    # ---- the other single-GPU BASELINE configurations, same method, shorter timed regions (N = 1 only)
    if world == 1 and secondary:
        line["configs"] = []
        for name in secondary:
            w2 = WORKLOADS[name]
            r2 = measure(args, w2, dev, 1, 0, min(args.min_seconds, 0.25), with_eval=True)
            hbm, tf_burst, tf_sus, src = peaks()
            fl = train_flops(w2, r2["m_valid"])
            r2["model"].distributed = False
            e2 = r2["eval"]; e2.pop("items")
            line["configs"].append({"workload": w2["name"], "value": r2["value"], "unit": "seq/s", "ms_per_step": r2["ms_per_step"],
                                    "e2e": r2["e2e_value"], "timed_steps": r2["timed_steps"], "gpu_launches_per_step": r2["launches_per_step"],
                                    "achieved_tflops": fl / (r2["ms_per_step"] / 1e3) / 1e12, "l2": r2["l2"],
                                    "eval": {"value": e2["value"], "e2e": e2["e2e"]["value"], "unit": "seq/s"},
                                    "roofline": roofline_block(r2["model"], r2["sess"], w2, r2["dev_batches"], args, hbm, tf_burst, src, brief=True)})
            del r2
            torch.cuda.empty_cache()
    if world > 1 and (world == 8 or os.environ.get("B4R_BENCH_C5")) and not args.workload:
        c5 = measure_c5(args, dev, world, rank)     # every rank takes part (collectives inside)
        if rank == 0:
            line["c5"] = c5
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def host_path_block(w, model, eval_batch):
    """Host data path of the loop: Cloze masking and negative sampling on the C++ threads of libb4r.so, timed
    on this box's cores next to the per-sequence Python (the product's restatement of the reference functions; the
    reference's own list-scan versions are slower still), and the 100-negative evaluation INCLUDING the host sampling."""
    import torch
    from bert4rec_b200.dataloaders import host_native as hn, dataloader_utils as du, samplers
    from bert4rec_b200.evaluation import BERT4RecEvaluator
    V, S, P, B = w["vocab_size"], w["seq_len"], w["max_pred"], w["batch"]
    rng = np.random.RandomState(11)
    n = 64 * B
    vals = rng.randint(3, V, size=n * S).astype(np.int64)
    off = np.arange(n + 1, dtype=np.int64) * S
    seeds = np.arange(n, dtype=np.uint64)
    args = (S, P, 1, [2, 0], V, w["mask_prob"], 1.0, 0.0)
    hn.cloze_mask_batch((vals[:B * S], off[:B + 1]), *args, seeds=seeds[:B])
    t0 = time.perf_counter(); hn.cloze_mask_batch((vals, off), *args, seeds=seeds); t_native = time.perf_counter() - t0
    m = 2 * B
    t0 = time.perf_counter()
    for i in range(m):
        du.apply_dynamic_masking_task(vals[i * S:(i + 1) * S], P, 1, [2, 0], V, w["mask_prob"], 1.0, 0.0, seed=i)
    t_py = time.perf_counter() - t0
    sm = samplers.get("random", vocab=list(range(3, V)), sample_size=100, seed=3)
    ev = BERT4RecEvaluator(sampler=sm)
    ev.build_candidates(eval_batch, as_array=True)
    t0 = time.perf_counter()
    for _ in range(4):
        ev.build_candidates(eval_batch, as_array=True)
    t_samp = (time.perf_counter() - t0) / 4
    hist = eval_batch["labels"][0].tolist()
    t0 = time.perf_counter()
    for i in range(32):
        sm.sample(without=hist + [5 + i])
    t_samp_py = (time.perf_counter() - t0) / 32
    ev.evaluate_batch(model, eval_batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        ev.evaluate_batch(model, eval_batch)
    torch.cuda.synchronize()
    t_eval = (time.perf_counter() - t0) / 4
    # seed=None in the reference = fresh entropy for every call: every request needs its own full permutation of the pool
    ev_fresh = BERT4RecEvaluator(sampler=samplers.get("random", vocab=list(range(3, V)), sample_size=100, seed=None))
    ev_fresh.build_candidates(eval_batch, as_array=True)
    t0 = time.perf_counter()
    for _ in range(2):
        ev_fresh.build_candidates(eval_batch, as_array=True)
    t_samp_fresh = (time.perf_counter() - t0) / 2
    ev_fresh.evaluate_batch(model, eval_batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        ev_fresh.evaluate_batch(model, eval_batch)
    torch.cuda.synchronize()
    t_eval_fresh = (time.perf_counter() - t0) / 2
    return {"threads": os.cpu_count(),
            "cloze_masking_seq_per_s": {"native_batch": n / t_native, "python_per_sequence": m / t_py},
            "negative_sampling_seq_per_s": {"native_batch_fixed_seed": B / t_samp, "native_batch_fresh_seeds": B / t_samp_fresh,
                                            "python_per_sequence": 1.0 / t_samp_py,
                                            "sampler": "RandomSampler(100 of V, without = history + [gt]), exact numpy legacy stream; "
                                                       "a fixed seed shares one shuffle per pool size across the batch"},
            "eval_with_host_sampling_seq_per_s": {"fixed_seed": B / t_eval, "fresh_seeds": B / t_eval_fresh},
            "note": "bit-exact with the reference's python `random` / np.random streams (tests/test_host_native.py)"}


def kernel_work(tag, w, n_rows, eval_mode=False):
    """Algorithmic (flops, bytes, bound) of ONE launch of the kernel behind a profile tag.
    Bytes are the COMPULSORY traffic only: inputs read once + outputs written once; per-CTA partial buffers a kernel creates for
    its own reductions are not algorithmic work and are not counted."""
    B, S, H, I, V, N = w["batch"], w["seq_len"], w["hidden_size"], w["inner_dim"], w["vocab_size"], w["num_attention_heads"]
    T, M, Vp, e = B * S, n_rows, (V + 127) // 128 * 128, 2
    L = w["num_layers"]
    saved = T * e * (H + L * (3 * H + 5 * H + 2 * I))   # activations the fused forward writes for backward
    enc = L * T * (8 * H * H + 4 * S * H + 4 * H * I)
    t = {
        # fused whole-encoder kernels: FLOPs of SURVEY 8d's encoder formula; bytes = ids + saved activations once (+ d(out) in the backward)
        "enc_fwd_fused": (enc, T * 8 + (saved if not eval_mode else T * H * e * L) + L * 4 * T * 4, "hbm"),
        "enc_bwd_fused": (2 * enc, saved + 2 * T * H * 4 + L * 4 * T * 4, "hbm"),
        "ce_fwd_umma": (2 * M * H * V, M * H * e + V * H * e + V * 4 + 3 * M * 4, "tensor"),
        # recompute passes: only the useful GEMM (dT = dl E, dE = dl^T t) is counted, not the recomputed logits
        "ce_bwd_umma:dT": (2 * M * H * V, M * H * e + V * H * e + V * 4 + M * H * 4, "tensor"),
        "ce_bwd_umma:dE": (2 * M * H * V, M * H * e + V * H * e + V * H * 4 + V * 4, "tensor"),
        "ce_bwd_fused": (4 * M * H * V, M * H * e + V * H * e + V * 4 + M * H * 4 + V * H * 4, "tensor"),
        "sqnorm+adamw": (0, 32 * (V * H + 64 * H + L * (4 * H * H + 2 * H * I) + H * H + V), "hbm"),
        "embed_ln_fwd": (0, T * (8 + 2 * H * e), "hbm"),
        # LN backward of the embedding rows: fp32 residual-stream gradient + bf16 branch gradient in, fp32 dx rows out (in place)
        "embed_bwd": (0, T * (8 + H * e + H * 4 + H * e + H * 4), "hbm"),
        # fixed-order per-item sums of the dx rows: every row gathered once + (at most) every table row read and written once
        "table_grad": (0, T * (H * 4 + 8) + 2 * min(V, T) * H * 4, "hbm"),
        "gemm:qkv": (2 * T * H * 3 * H, T * H * e + 3 * H * H * e + T * 3 * H * e, "tensor"),
        "attn_fwd": (4 * T * S * H, T * 3 * H * e + T * H * e + T * N * 4, "tensor"),
        "rowln:attn_out": (2 * T * H * H, 3 * T * H * e + H * H * e + T * H * e, "tensor"),
        "gemm:ffn1_gelu": (2 * T * H * I, T * H * e + H * I * e + 2 * T * I * e, "tensor"),
        "rowln:ffn2": (2 * T * I * H, T * I * e + I * H * e + 3 * T * H * e, "tensor"),
        "rowln:mlm_transform": (2 * M * H * H, 4 * M * H * e + H * H * e, "tensor"),
        "ce_fwd": (2 * M * H * V, M * H * e + V * H * e + V * 4 + 3 * M * 4, "tensor"),
        "ce_dlogits": (2 * M * H * V, M * H * e + V * H * e + V * 4 + M * Vp * e, "hbm"),
        "gemm:ce_dT": (2 * M * V * H, M * Vp * e + V * H * e + M * H * 4, "hbm"),
        "wgrad:ce_dE": (2 * M * V * H, M * Vp * e + M * H * e + V * H * 4, "hbm"),
        "colsum:vbias": (0, M * Vp * e + V * 4, "hbm"),
        "ln_bwd": (0, T * H * (4 + e + e + 4 + e), "hbm"),   # fp32 stream + bf16 branch + bf16 pre-LN in; fp32 stream (in place) + bf16 branch out
        "attn_bwd": (10 * T * S * H, T * 3 * H * e * 2 + 2 * T * H * e, "tensor"),
        "wgrad:w2": (2 * T * I * H, T * I * e + T * H * e + I * H * 4, "tensor"),
        "wgrad:w1": (2 * T * I * H, T * I * e + T * H * e + I * H * 4, "tensor"),
        "wgrad:wo": (2 * T * H * H, 2 * T * H * e + H * H * 4, "tensor"),
        "wgrad:wqkv": (2 * T * H * 3 * H, T * H * e + T * 3 * H * e + 3 * H * H * 4, "tensor"),
        "gemm:ffn2_dgrad_gelu": (2 * T * H * I, T * H * e + 2 * T * I * e + I * H * e, "tensor"),
        "gemm:ffn1_dgrad": (2 * T * I * H, T * I * e + I * H * e + T * H * e, "tensor"),
        "gemm:attn_out_dgrad": (2 * T * H * H, 2 * T * H * e, "tensor"),
        "gemm:qkv_dgrad": (2 * T * 3 * H * H, T * 3 * H * e + 3 * H * H * e + T * H * e, "tensor"),
        "colsum:bqkv": (0, T * 3 * H * e, "hbm"),
        # evaluation: fused gather-dot over 101 candidates + stable rank (101 H e + 101 * 4 + H e per sequence + ids)
        "rank_candidates": (2 * M * 101 * H, M * (101 * H * e + 101 * 4 + H * e + 101 * 8 + 4), "hbm"),
    }
    return t.get(tag, (0, 0, "hbm"))


def profile_steps(model, sess, batches, n, flush=None):
    """Per-kernel device times of n train steps INSIDE the CUDA graph: the step is re-captured with profiling on, so
    every launch is bracketed by external event-record nodes and each replay re-times it (b4r_profile_*).
    Returns {tag: (launches, total_ms)} summed over the n replays."""
    import torch
    model.use_cuda_graph = True
    model._graphs.clear()
    sess.profile(True)
    model.train_step(batches[0])          # eager pass + capture (with event nodes)
    torch.cuda.synchronize()
    sess.profile_report()                 # discard the eager records
    agg = {}
    for i in range(n):
        if flush is not None:
            flush.fill_(i & 0xFF)
        model.train_step(batches[0])      # the profiled graph is the one captured for this batch's buffers
        for tag, (cnt, ms) in sess.profile_report().items():
            c0, m0 = agg.get(tag, (0, 0.0))
            agg[tag] = (c0 + cnt, m0 + ms)
    sess.profile(False)
    model._graphs.clear()                 # the profiled graph references destroyed events: never replay it again
    return agg


def roofline_from(rep, w, n_rows, n, hbm, tf_burst, src, eval_mode=False, brief=False):
    """Roofline entry = the kernel with the largest share of the pass among those with a stated algorithmic workload; every
    kernel of the breakdown carries BOTH fractions (tensor and HBM) of the measured peaks."""
    # "token_sort" is an integer side branch that runs beside the whole backward; the events around it measure how long it waited
    # for SM slots between the main path's kernels, not work on the step's critical path: it is listed in no share
    rows = sorted(((tag, cnt, tot) for tag, (cnt, tot) in rep.items() if tag != "token_sort"), key=lambda r: -r[2])
    total = sum(r[2] for r in rows)
    breakdown = []
    for tag, cnt, tot in rows[:(6 if brief else 16)]:
        fl, by, bound = kernel_work(tag, w, n_rows, eval_mode)
        ms = tot / cnt
        tf, gb = (fl / (ms / 1e3) / 1e12 if fl else 0.0), (by / (ms / 1e3) / 1e9 if by else 0.0)
        breakdown.append({"kernel": tag, "launches_per_step": cnt / n, "avg_ms": ms, "share": tot / total, "bound": bound,
                          "tflops": tf, "gbs": gb, "frac_tensor": tf / tf_burst, "frac_hbm": gb / hbm})
    known = [r for r in rows if any(kernel_work(r[0], w, n_rows, eval_mode)[:2])]
    tag, cnt, tot = (known or rows)[0]
    fl, by, bound = kernel_work(tag, w, n_rows, eval_mode)
    ms = tot / cnt
    tf, gb = fl / (ms / 1e3) / 1e12, by / (ms / 1e3) / 1e9
    ach, peak, unit = (tf, tf_burst, "TFLOP/s") if bound == "tensor" else (gb, hbm, "GB/s")
    traffic = NCU_TRAFFIC.get((w["name"][:2], tag))
    return {"kernel": tag, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
            "frac_tensor": tf / tf_burst, "frac_hbm": gb / hbm, "traffic": traffic, "peak_source": src, "launch_ms": ms,
            "share_of_step": tot / total, "algorithmic_flops": fl, "algorithmic_bytes": by, "kernel_ms_per_step": total / n,
            "timing": "CUDA events recorded as graph nodes around every launch, averaged over %d replays" % n,
            "breakdown": breakdown}


def roofline_block(model, sess, w, dev_batches, args, hbm, tf_burst, src, brief=False):
    """Per-kernel CUDA-event timing inside graph replays of the train step (profile_steps)."""
    n = 10
    model.distributed = False      # rank-0-only profiling pass: no collective (the other ranks are not stepping)
    rep = profile_steps(model, sess, dev_batches, n)
    return roofline_from(rep, w, int(sess.counts()[1]), n, hbm, tf_burst, src, brief=brief)


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/), keyed by
# (config, kernel tag); cold-cache captures
NCU_TRAFFIC = {
    ("C2", "enc_bwd_fused"): 39_311_872 + 396_288,       # profiles/r01_ncu_full_c2.md
    ("C2", "enc_fwd_fused"): 1_029_376 + 227_328,
    ("C2", "ce_fwd_umma"): 1_879_040,
    ("C2", "ce_bwd_fused"): 1_999_616 + 809_472,
    ("C4", "attn_bwd"): 555_666_432 + 280_639_232,       # profiles/r02_ncu_c4_backward.md (final kernel)
    ("C4", "attn_fwd"): 336_830_208 + 109_373_952,       # profiles/r02_ncu_fattn_c4.md
    ("C4", "ln_bwd"): 421_081_344 + 273_888_000,         # profiles/r02_ncu_c4_backward.md
    ("C4", "ce_fwd_umma"): 27_901_696 + 6_400,           # profiles/r02_ncu_ce_fwd_c4.md
}


_REAL_STDOUT = None


def _quiet_stdout():
    """Libraries (NCCL's version banner, torch warnings) may write to fd 1; the contract is ONE JSON line on stdout.
    fd 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c4 (the largest single-GPU BASELINE configuration) on 1 GPU, c3 (the data-parallel one) on N > 1")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="lower bound of the timed region (passes of --steps are repeated)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C1 / C2 entries of the `configs` array (N = 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    name = args.workload or ("c4" if max(world, args.gpus) == 1 else "c3")
    w = WORKLOADS[name]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        secondary = [] if (args.no_secondary or args.workload) else [c for c in ("c1", "c2") if c != name]
        run_b200(args, w, secondary)


if __name__ == "__main__":
    main()
