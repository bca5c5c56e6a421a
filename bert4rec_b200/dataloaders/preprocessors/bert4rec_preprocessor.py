"""Per-sequence model-input layout (reference: bert4rec_preprocessor.py:47-168): tokenize, truncate (most recent
``max_seq_len`` items when finetuning / evaluating, random window otherwise), Cloze or last-token masking, right-pad
with PAD=0.  Output keys and dtypes (all int64): labels, input_word_ids, input_mask [max_seq_len];
masked_lm_ids, masked_lm_positions, masked_lm_weights [max_predictions_per_seq]."""
import random

import numpy as np
import torch

from .. import dataloader_utils


class BERT4RecPreprocessor:
    tokenizer = None
    max_seq_len: int = None
    max_predictions_per_seq: int = None
    mask_token_id: int = None
    unk_token_id: int = None
    pad_token_id: int = None
    masked_lm_rate: float = None
    mask_token_rate: float = None
    random_token_rate: float = None

    @classmethod
    def set_properties(cls, **kwargs):
        """Only the given (non-None) properties are updated."""
        for k, v in kwargs.items():
            if not hasattr(cls, k):
                raise TypeError(f"unknown property {k}")
            if v is not None:
                setattr(cls, k, v)

    @classmethod
    def process_element(cls, sequence, apply_mlm: bool, finetuning: bool, seed: int = None) -> dict:
        tokens = cls.tokenizer.tokenize(sequence)
        S, P = cls.max_seq_len, cls.max_predictions_per_seq
        if finetuning or len(tokens) <= S:
            window = tokens[-S:]
        else:
            start = random.randint(0, len(tokens) - S)
            window = tokens[start:start + S]
        ids = np.array(window, dtype=np.int64)
        mask = np.ones_like(ids)
        labels = ids.copy()
        out = {}
        if apply_mlm:
            if finetuning:
                ids, pos, lab = dataloader_utils.mask_last_token_only(ids, cls.mask_token_id)
            else:
                ids, pos, lab = dataloader_utils.apply_dynamic_masking_task(
                    ids, P, cls.mask_token_id, [cls.unk_token_id, cls.pad_token_id], cls.tokenizer.get_vocab_size(),
                    selection_rate=cls.masked_lm_rate, mask_token_rate=cls.mask_token_rate,
                    random_token_rate=cls.random_token_rate, seed=seed)
            w = np.ones_like(lab)
            short = P - lab.shape[0]
            if short > 0:
                lab, pos, w = (np.pad(a, (0, short), constant_values=cls.pad_token_id) for a in (lab, pos, w))
            out["masked_lm_ids"], out["masked_lm_positions"], out["masked_lm_weights"] = lab, pos, w
        short = S - ids.shape[0]
        if short > 0:
            ids, mask, labels = (np.pad(a, (0, short), constant_values=cls.pad_token_id) for a in (ids, mask, labels))
        out["labels"], out["input_word_ids"], out["input_mask"] = labels, ids, mask
        return out

    @classmethod
    def process_batch(cls, sequences, apply_mlm: bool = True, finetuning: bool = False, seeds=None, n_threads: int = 0) -> dict:
        """``process_element`` for a list of raw sequences at once -> dict of int64 arrays ``[n, max_seq_len]`` /
        ``[n, max_predictions_per_seq]``.  The Cloze-masking case (the per-epoch work of training) runs on the C++ threads
        of libb4r.so (``b4r_host_cloze_mask_batch``): with ``seeds[i]`` it is bit-identical to
        ``process_element(sequences[i], True, False, seed=seeds[i])``; ``seeds=None`` draws fresh entropy per sequence, which
        is what the reference does (it never passes a seed, bert4rec_preprocessor.py:82-90)."""
        if not (apply_mlm and not finetuning):
            els = [cls.process_element(seq, apply_mlm, finetuning) for seq in sequences]
            keys = els[0].keys() if els else ()
            return {k: np.stack([np.asarray(e[k], dtype=np.int64) for e in els]) for k in keys}
        from .. import host_native
        S = cls.max_seq_len
        windows = []
        for seq in sequences:
            tokens = cls.tokenizer.tokenize(seq)
            if len(tokens) <= S:
                windows.append(tokens)
            else:
                start = random.randint(0, len(tokens) - S)
                windows.append(tokens[start:start + S])
        return host_native.cloze_mask_batch(
            windows, S, cls.max_predictions_per_seq, cls.mask_token_id, [cls.unk_token_id, cls.pad_token_id],
            cls.tokenizer.get_vocab_size(), cls.masked_lm_rate, cls.mask_token_rate, cls.random_token_rate, seeds=seeds,
            pad_token_id=cls.pad_token_id, n_threads=n_threads)

    @classmethod
    def process_dataset(cls, ds, apply_mlm: bool, finetuning: bool):
        return [cls.process_element(seq, apply_mlm, finetuning) for seq in ds]

    @classmethod
    def prepare_inference(cls, data) -> dict:
        """History (list of raw items) -> batch-of-one model input with a MASK appended as the last token."""
        if type(data) is not list:
            raise ValueError("To prepare data for inference, please simply put in an unprocessed sequence of data "
                             "(i.e. a list of strings).")
        seq = data[-cls.max_seq_len + 1:]
        seq.append("[UNK]")
        el = cls.process_element(seq, True, True)
        return {k: torch.from_numpy(np.asarray(v, dtype=np.int64)).unsqueeze(0) for k, v in el.items()}
