"""Ranker app (reference: bert4rec/apps/ranker.py): rank of one item, for a raw history, relative to the whole vocabulary or
to a list of items.  Mirrors the reference line by line, INCLUDING its sign convention: when the model returns ``mlm_logits``
(it does for the inference input, which carries ``masked_lm_positions``) the reference ranks by ``-mlm_logits`` in descending
order (ranker.py:30, :58), i.e. the least probable item gets rank 1.  A serving path that wants the conventional order uses
``BERT4RecModel.rank_items`` / ``full_catalogue_ranks`` (device-side, no logits materialised)."""
import torch


class Ranker:
    def __init__(self, ranker_model, dataloader=None):
        from .inference import InferenceDataloader
        self.ranker_model = ranker_model
        self.dataloader = dataloader if dataloader is not None else InferenceDataloader(max_seq_len=512)

    def __call__(self, sequence: list, rank_item: str, rank_items: list = None):
        model_input = self.dataloader.prepare_inference(sequence)
        predictions = self.ranker_model(model_input, training=False)
        vocab_logits = -predictions["mlm_logits"][:, -1]                      # ranker.py:30
        tokenized_rank_items = None
        if rank_items is not None:
            tokenized_rank_items = torch.as_tensor(self.dataloader.tokenizer.tokenize(list(rank_items)), dtype=torch.int64,
                                                   device=vocab_logits.device)
            vocab_logits = vocab_logits[:, tokenized_rank_items]                # ranker.py:33
        rank_item_token = self.dataloader.tokenizer.tokenize(rank_item)
        sorted_indexes = torch.argsort(vocab_logits[0], descending=True, stable=True)
        vocab_ranking = sorted_indexes if tokenized_rank_items is None else tokenized_rank_items[sorted_indexes]
        hits = (vocab_ranking == rank_item_token).nonzero()
        if hits.numel() == 0:
            raise IndexError(f"\"{rank_item}\" is not among the ranked items")   # np.where(...)[0][0] of the reference
        rank = int(hits[0, 0]) + 1
        assert_string = f"Rank of \"{rank_item}\" is {rank} in the given sequence:\n{sequence}\n"
        if rank_items is not None:
            assert_string += f"relative to {len(rank_items)} other elements:\n{rank_items}"
        else:
            assert_string += "relative to the whole vocabulary"
        return rank, assert_string
