"""CPU suite, world_size 2 over gloo: the host-side math of the multi-GPU paths.
(1) batch data-parallel: SUM-loss gradients and valid-slot counts all-reduced, then divided by the GLOBAL count,
    equal the single-process gradient of the mean loss (what BERT4RecModel.train_step does over NCCL);
(2) vocab-sharded softmax-CE partials (max, sumexp, label logit, arg-max) merged across shards equal the unsharded
    result; sharded "items beating the ground truth" counts add up to the unsharded rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om
from tests.helpers import make_batch


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = om.Config(vocab_size=157, hidden_size=64, num_layers=1, num_attention_heads=2, max_sequence_length=16, inner_dim=64)
    params = om.init_params(cfg, 0)
    batch = make_batch(8, 16, 4, 157, seed=2)
    half = {k: v[rank * 4:(rank + 1) * 4] for k, v in batch.items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    logits = om.model_forward(leaves, cfg, half)["mlm_logits"]
    y = half["masked_lm_ids"]
    mask = (y != 0)
    per = torch.logsumexp(logits, -1) - torch.gather(logits, -1, y.unsqueeze(-1)).squeeze(-1)
    loss_sum = (per * mask).sum()
    names = [k for k in leaves if not k.startswith("pooler")]
    grads = torch.autograd.grad(loss_sum, [leaves[k] for k in names])
    flat = torch.cat([g.reshape(-1) for g in grads])
    count = mask.sum().float().reshape(1)
    dist.all_reduce(flat)
    dist.all_reduce(count)
    flat = flat / count
    # ---- vocab-sharded CE partial merge on the full batch
    full = om.model_forward(params, cfg, batch)["mlm_logits"].reshape(-1, 157)
    lo, hi = (0, 80) if rank == 0 else (80, 157)
    sh = full[:, lo:hi]
    mx = sh.max(-1).values
    se = torch.exp(sh - mx[:, None]).sum(-1)
    yy = batch["masked_lm_ids"].reshape(-1)
    in_shard = (yy >= lo) & (yy < hi)
    lab = torch.where(in_shard, sh.gather(-1, (yy - lo).clamp(0, hi - lo - 1)[:, None]).squeeze(-1), torch.full_like(mx, -float("inf")))
    gmx = mx.clone(); dist.all_reduce(gmx, op=dist.ReduceOp.MAX)
    gse = se * torch.exp(mx - gmx); dist.all_reduce(gse)
    glab = lab.clone(); dist.all_reduce(glab, op=dist.ReduceOp.MAX)
    lse = gmx + torch.log(gse)
    beat = ((sh > glab[:, None]) | ((sh == glab[:, None]) & (torch.arange(lo, hi)[None, :] < yy[:, None]))).sum(-1)
    dist.all_reduce(beat)
    if rank == 0:
        ret["flat"] = flat
        ret["lse"], ret["lab"], ret["beat"] = lse, glab, beat
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_gradients_and_vocab_sharded_ce_over_gloo():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    cfg = om.Config(vocab_size=157, hidden_size=64, num_layers=1, num_attention_heads=2, max_sequence_length=16, inner_dim=64)
    params = om.init_params(cfg, 0)
    batch = make_batch(8, 16, 4, 157, seed=2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    logits = om.model_forward(leaves, cfg, batch)["mlm_logits"]
    loss = om.masked_sparse_ce(batch["masked_lm_ids"], logits)
    names = [k for k in leaves if not k.startswith("pooler")]
    ref = torch.cat([g.reshape(-1) for g in torch.autograd.grad(loss, [leaves[k] for k in names])])
    assert torch.allclose(ret["flat"], ref, rtol=1e-4, atol=1e-7)
    full = logits.detach().reshape(-1, 157)
    yy = batch["masked_lm_ids"].reshape(-1)
    assert torch.allclose(ret["lse"], torch.logsumexp(full, -1), atol=1e-5)
    assert torch.equal(ret["lab"], full.gather(-1, yy[:, None]).squeeze(-1))
    sg = full.gather(-1, yy[:, None])
    ref_beat = ((full > sg) | ((full == sg) & (torch.arange(157)[None, :] < yy[:, None]))).sum(-1)
    assert torch.equal(ret["beat"], ref_beat)


def _eval_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bert4rec_b200.evaluation.bert4rec_evaluator import default_bert4rec_metrics
    from bert4rec_b200.evaluation.base_evaluator import BaseEvaluator

    class _E(BaseEvaluator):
        def evaluate(self, model, test_data):
            return self._metrics

    import numpy as np
    from bert4rec_b200.dataloaders import samplers
    ev = _E(default_bert4rec_metrics(), samplers.get("random", sample_size=5, vocab=list(range(3, 50))))
    ranks = np.random.RandomState(7).randint(1, 102, size=200)
    mine = ranks[rank::world]                      # this rank's share of the sequences
    for m in ev.get_metrics():
        m.update_many(mine)
    out = ev.all_reduce_metrics()
    if rank == 0:
        ret["merged"] = dict(out)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_evaluation_metrics_merge_over_gloo():
    """Every rank ranks its own sequences; one all-reduce of the partial sums gives the single-process metrics."""
    import numpy as np
    from bert4rec_b200.evaluation.bert4rec_evaluator import default_bert4rec_metrics
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_eval_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    ranks = np.random.RandomState(7).randint(1, 102, size=200)
    ref = default_bert4rec_metrics()
    for m in ref:
        m.reset()
        m.update_many(ranks)
    for m in ref:
        assert abs(ret["merged"][m.name] - m.result()) < 1e-12, (m.name, ret["merged"][m.name], m.result())


def test_shard_range_partitions_the_catalogue():
    """Host logic of the vocabulary-sharded projection: contiguous, disjoint, covering slices (engine.shard_range)."""
    from bert4rec_b200.engine import shard_range
    for V in (1, 7, 1203, 12004, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(V, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == V
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(lo <= hi for lo, hi in spans)
            assert max(hi - lo for lo, hi in spans) == -(-V // world)
