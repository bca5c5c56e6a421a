"""CPU suite: the C-ABI library loads and exports every symbol include/*.h declares (b4r.h = product ABI, b4r_debug.h = test helpers) (no compute calls), the flat
parameter layout is sane, and the reference-facing factories keep their contract."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = "".join(open(os.path.join(ROOT, "include", f)).read() for f in sorted(os.listdir(os.path.join(ROOT, "include"))) if f.endswith(".h"))
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b4r_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bert4rec_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/*.h but not exported by libb4r.so"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.b4r_version() == 100


def test_param_layout_and_error_reporting():
    from bert4rec_b200 import _lib
    lib = _lib.load()
    cfg = _lib.Config(3709, 64, 2, 2, 200, 256, 0.2, 0.2)
    n = lib.b4r_param_entries(C.byref(cfg), None, 0)
    ents = (_lib.ParamEntry * n)()
    assert lib.b4r_param_entries(C.byref(cfg), ents, n) == n
    names = [e.name.decode() for e in ents]
    assert names[0] == "word_embeddings" and "layer_1/wqkv" in names and "head/output_bias" in names
    nd, nt, na = C.c_int64(), C.c_int64(), C.c_int64()
    assert lib.b4r_param_counts(C.byref(cfg), C.byref(nd), C.byref(nt), C.byref(na)) == 0
    assert 0 < nd.value < nt.value < na.value and nd.value % 8 == 0 and nt.value % 8 == 0
    real = sum(e.numel for e in ents if e.group < 2)
    assert real == 3709 * 64 + 200 * 64 + 2 * (64 * 192 + 192 + 64 * 64 + 64 + 4 * 64 + 64 * 256 + 256 + 256 * 64 + 64) \
        + 64 * 64 + 64 + 2 * 64 + 3709 + 2 * 64                    # = the reference's trainable variable count
    for e in ents:
        assert e.offset % 8 == 0 and (e.group == 0) == (e.offset < nd.value)
    assert lib.b4r_session_workspace_bytes(C.byref(cfg), 256, 200, 40) > 0
    bad = _lib.Config(100, 96, 2, 2, 50, 64, 0.1, 0.1)
    assert lib.b4r_param_entries(C.byref(bad), None, 0) < 0 and b"hidden_size" in lib.b4r_last_error()
    assert lib.b4r_session_workspace_bytes(C.byref(cfg), 4, 300, 4) == 0 and b"seq_len" in lib.b4r_last_error()


def test_factories_and_schedule():
    from bert4rec_b200 import trainers, evaluation, tokenizers
    from bert4rec_b200.trainers import optimizers
    from bert4rec_b200.dataloaders import samplers
    from oracle import model as om
    for fn in (trainers.get, evaluation.get, tokenizers.get, optimizers.get, samplers.get):
        with pytest.raises(ValueError):
            fn("unknown-id")
    opt = optimizers.get("adamw")
    assert isinstance(opt, optimizers.AdamWeightDecay) and isinstance(opt.learning_rate, optimizers.WarmUp)
    hp = opt.hparams_struct()
    assert (hp.num_train_steps, hp.num_warmup_steps) == (400000, 100) and abs(hp.epsilon - 1e-6) < 1e-12
    assert abs(hp.clip_norm - 5.0) < 1e-9 and abs(hp.weight_decay_rate - 0.01) < 1e-9
    for step in (0, 1, 50, 99, 100, 101, 200000, 400000, 500000):
        assert opt.learning_rate(step) == om.lr_schedule(step)
    assert opt.learning_rate(0) == 0.0                                   # first step changes only Adam's moments
    assert opt._do_use_weight_decay("transformer/layer_0/intermediate/kernel")
    assert not opt._do_use_weight_decay("transformer/layer_0/output_layer_norm/gamma")
    assert not opt._do_use_weight_decay("cls/predictions/output_bias/bias")
    ev = evaluation.get("bert4rec", sampler=samplers.RandomSampler(vocab=[3, 4, 5], sample_size=2))
    assert [m.name for m in ev.get_metrics()] == ["Valid Ranks", "NDCG@1", "NDCG@5", "NDCG@10", "HR@1", "HR@5", "HR@10", "MAP"]
    with pytest.raises(ValueError):
        evaluation.get("bert4rec", sampler=samplers.RandomSampler()).evaluate(None, [])


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: constructing the encoder on a non-CUDA device raises."""
    import torch
    from bert4rec_b200 import _lib
    from bert4rec_b200.models.components import networks
    with pytest.raises((_lib.B4RError, RuntimeError, AssertionError)):
        networks.Bert4RecEncoder(vocab_size=100, hidden_size=64, num_layers=1, num_attention_heads=2,
                                 max_sequence_length=16, inner_dim=64, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            networks.Bert4RecEncoder(vocab_size=100, hidden_size=64, num_layers=1, num_attention_heads=2,
                                     max_sequence_length=16, inner_dim=64, device="cuda:0")


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "bert4rec_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the B200 arm): exactly one JSON line on stdout with the
    contract's keys, timed on the oracle port (the reference itself cannot be installed here: TensorFlow is absent)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "train masked-seq/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_p2p_allreduce_entry_validates_its_arguments_without_a_gpu():
    """b4r_p2p_allreduce_f32 (the data-parallel gradient all-reduce over NVLink peer memory) rejects worlds it does not support and
    null buffers before touching the device; its exchange needs >= 2 GPUs and is checked by scripts/p2p_allreduce_check.py and the
    `dp_parity` entry of the multi-GPU bench line."""
    import ctypes as C
    from bert4rec_b200 import _lib
    lib = _lib.load()
    assert lib.b4r_p2p_allreduce_max_world() >= 8
    one = C.c_void_p(16)
    assert lib.b4r_p2p_allreduce_f32(None, one, None, 0, 4, 0, 2, one, None) != 0
    assert lib.b4r_p2p_allreduce_f32(one, one, None, 0, 4, 0, 1, one, None) != 0
    assert lib.b4r_p2p_allreduce_f32(one, one, None, 0, 4, 5, 4, one, None) != 0
    assert lib.b4r_p2p_allreduce_f32(one, one, None, 2, 4, 0, 2, one, None) != 0
    assert b"16-byte" in lib.b4r_last_error()
