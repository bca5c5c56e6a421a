"""Recommender app (reference: bert4rec/apps/recommender.py): the most probable next item for a raw history, never one the
user has already seen (prediction mask of -inf on the history's item ids)."""
import torch


class Recommender:
    def __init__(self, recommender_model, dataloader=None):
        from .inference import InferenceDataloader
        self.recommender_model = recommender_model
        self.dataloader = dataloader if dataloader is not None else InferenceDataloader(max_seq_len=512)

    def __call__(self, sequence: list):
        model = self.recommender_model
        model_input = self.dataloader.prepare_inference(sequence)
        seen = self.dataloader.get_tokenizer().tokenize(list(sequence))
        if getattr(model, "prediction_mask", None) is None and hasattr(model, "top_k_items") and len(set(seen)) < 120:
            # fused full-catalogue top-k on the device (no [V] logits row): best item that is not in the history
            ids, _ = model.top_k_items(model_input, 1, exclude=[seen])
            return self.dataloader.tokenizer.detokenize(int(ids[0, 0]))
        out = model(model_input, training=False)
        logits = out["mlm_logits"][0, 0].clone()        # [MASK] is the only (last) prediction slot of the inference input
        logits[torch.as_tensor(sorted(set(seen)), dtype=torch.int64, device=logits.device)] = -float("inf")
        if getattr(model, "prediction_mask", None) is not None:
            logits += model.prediction_mask
        return self.dataloader.tokenizer.detokenize(int(torch.argmax(logits)))
