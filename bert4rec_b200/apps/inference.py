"""The part of the reference's dataloader the apps use (bert4rec/dataloaders/base_dataloader.py: ``tokenizer``,
``get_tokenizer()``, ``prepare_inference()``), without any dataset behind it."""
from bert4rec_b200 import tokenizers
from bert4rec_b200.dataloaders.preprocessors import BERT4RecPreprocessor


class InferenceDataloader:
    def __init__(self, tokenizer="simple", max_seq_len: int = 512, max_predictions_per_seq: int = 1,
                 mask_token="[MASK]", unk_token="[UNK]", pad_token="[PAD]"):
        self.tokenizer = tokenizers.get(tokenizer)
        self.max_seq_len, self.max_predictions_per_seq = max_seq_len, max(1, max_predictions_per_seq)
        self._special = (pad_token, mask_token, unk_token)   # ids 0, 1, 2 (bert4rec_dataloader.py:35-43)
        self.tokenizer.tokenize(list(self._special))

    def get_tokenizer(self):
        return self.tokenizer

    def prepare_inference(self, sequence):
        P = BERT4RecPreprocessor
        P.set_properties(tokenizer=self.tokenizer, max_seq_len=self.max_seq_len, max_predictions_per_seq=self.max_predictions_per_seq,
                         mask_token_id=self.tokenizer.tokenize(self._special[1]), unk_token_id=self.tokenizer.tokenize(self._special[2]),
                         pad_token_id=self.tokenizer.tokenize(self._special[0]), masked_lm_rate=0.2, mask_token_rate=1.0,
                         random_token_rate=0.0)
        return P.prepare_inference(list(sequence))
