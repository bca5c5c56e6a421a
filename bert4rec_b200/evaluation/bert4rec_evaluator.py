"""Leave-one-out evaluator with sampled negatives (reference: bert4rec/evaluation/bert4rec_evaluator.py:24-120).

Per slot with ``masked_lm_weights == 1``: candidates = ``sampler.sample(without=labels + [gt])`` + ``[gt]`` (the ground
truth is always the LAST candidate, :98-104); the model ranks the candidates; rank = 1-based position of the ground
truth; every metric is updated with that rank.  The candidate lists are built on the host with the reference's exact
sampler semantics (bit-exact with a fixed seed); scoring + ranking run in ONE fused kernel over the whole batch
(``b4r_rank_candidates``) instead of per-position tf.gather/argsort dispatches, and only the integer ranks come
back to the host."""
from typing import Union

import numpy as np
import torch

from bert4rec_b200.dataloaders import samplers
from .base_evaluator import BaseEvaluator
from .evaluation_metrics import EvaluationMetric, Counter, NDCG, HR, MAP


def default_bert4rec_metrics():
    return [Counter(name="Valid Ranks"), NDCG(1), NDCG(5), NDCG(10), HR(1), HR(5), HR(10), MAP()]


bert4rec_evaluation_metrics = default_bert4rec_metrics()


class BERT4RecEvaluator(BaseEvaluator):
    def __init__(self, metrics: list = None, sampler: Union[str, samplers.BaseSampler] = "pop_random", dataloader=None):
        if metrics is None:
            metrics = bert4rec_evaluation_metrics
        if isinstance(sampler, str):
            cfg = {"sample_size": 100}
            if dataloader is not None:
                vocab = dataloader.tokenizer.get_vocab()
                cfg.update({"source": dataloader.create_item_list_tokenized(),
                            "vocab": dataloader.tokenizer.tokenize(vocab)})
            sampler = samplers.get(sampler, **cfg)
        super().__init__(metrics, sampler, dataloader)
        self.last_ranks = None

    def evaluate(self, model, test_data):
        if self.dataloader is None and not self.sampler.is_fully_prepared():
            raise ValueError("The evaluator has to be either initialized with a dataloader or a fully prepared "
                             "sampler has to be given.")
        for batch in test_data:
            self.evaluate_batch(model, batch)
        return self._metrics

    def build_candidates(self, test_batch, as_array=False):
        """Host side of evaluate_batch: returns (candidate lists in slot order, ground truths).  All slots of the batch go
        through ONE ``sampler.sample_batch`` call (C++ threads of libb4r.so for the shipped samplers, bit-exact with the
        per-slot ``sample(without=history + [gt])`` calls of the reference).  ``as_array``: candidates as an int64
        [n_slots, sample_size + 1] array when every list has the full length (else the lists)."""
        w = np.asarray(torch.as_tensor(test_batch["masked_lm_weights"]).cpu()) != 0
        ids = np.asarray(torch.as_tensor(test_batch["masked_lm_ids"]).cpu())
        labels = np.asarray(torch.as_tensor(test_batch["labels"]).cpu()).astype(np.int64)
        bs, ps = np.nonzero(w)
        if bs.size == 0:
            return [], []
        gts = ids[bs, ps].astype(np.int64)
        n, S = bs.size, labels.shape[1]
        without = np.concatenate([labels[bs], gts[:, None]], axis=1)          # history + [gt] per slot
        arr, lens = self.sampler.sample_batch((without.reshape(-1), np.arange(n + 1, dtype=np.int64) * (S + 1)), as_array=True)
        if as_array and bool((lens == arr.shape[1]).all()):
            return np.concatenate([arr, gts[:, None]], axis=1), gts
        cands = [arr[i, :lens[i]].tolist() + [int(gts[i])] for i in range(n)]
        return cands, gts.tolist()

    def evaluate_batch(self, model, test_batch: dict):
        cands, gts = self.build_candidates(test_batch, as_array=True)
        if len(cands) == 0:
            return
        lengths = {cands.shape[1]} if isinstance(cands, np.ndarray) else {len(c) for c in cands}
        if len(lengths) == 1:
            cand = torch.as_tensor(np.asarray(cands, dtype=np.int64))
            _, rank = model.rank_candidates(test_batch, cand, torch.as_tensor(np.asarray(gts, dtype=np.int64)))
            ranks = rank.cpu().numpy().astype(np.int64)
        else:  # ragged candidate lists (e.g. a popular sampler that ran out of items): generic API path
            w = np.asarray(torch.as_tensor(test_batch["masked_lm_weights"]).cpu()) != 0
            per_seq, k = [], 0
            for b in range(w.shape[0]):
                n = int(w[b].sum())
                per_seq.append(cands[k:k + n])
                k += n
            rankings = model.rank_items(test_batch, per_seq)
            flat = [r for row in rankings for r in row]
            ranks = np.array([int(np.where(r.numpy() == g)[0][0]) + 1 for r, g in zip(flat, gts)], dtype=np.int64)
        self.last_ranks = ranks
        for metric in self._metrics:
            metric.update_many(ranks)
