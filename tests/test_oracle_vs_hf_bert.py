"""CPU suite: the floating-point oracle against an INDEPENDENT implementation of the same architecture.

The reference's arithmetic lives in tf-models-official / Keras layers that cannot run here (parity with TF itself
stays unpinned).  Those layers (`TransformerEncoderBlock`, `MaskedLM`) are Google BERT's post-LayerNorm encoder block and MLM
head; Hugging Face `BertForMaskedLM` (installed, PyTorch) is the other widely used port of the same network.  Loading the
oracle's weights (TF variable names / shapes) into it and comparing outputs pins the oracle's restatement of
the embedding stage, attention with the key-padding mask, both LayerNorms (eps 1e-12), erf-GELU FFN, pooler, the tied MLM head
and the masked CE to an independent code base -- on CPU, fp32, eval mode (dropout is checked elsewhere with replayed masks)."""
import pytest
import torch

from oracle import model as om
from tests.helpers import make_batch

transformers = pytest.importorskip("transformers")


def _hf_from_oracle(cfg, p):
    from transformers import BertConfig, BertForMaskedLM
    H, N = cfg.hidden_size, cfg.num_attention_heads
    hc = BertConfig(vocab_size=cfg.vocab_size, hidden_size=H, num_hidden_layers=cfg.num_layers, num_attention_heads=N,
                    intermediate_size=cfg.inner_dim, max_position_embeddings=cfg.max_sequence_length, type_vocab_size=1,
                    hidden_act="gelu", layer_norm_eps=1e-12, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                    pad_token_id=0, tie_word_embeddings=True)
    m = BertForMaskedLM(hc).eval()
    sd = {}
    sd["bert.embeddings.word_embeddings.weight"] = p["word_embeddings/embeddings"]
    sd["bert.embeddings.position_embeddings.weight"] = p["position_embedding/embeddings"]
    sd["bert.embeddings.token_type_embeddings.weight"] = torch.zeros(1, H)        # the reference has no segment embedding
    sd["bert.embeddings.LayerNorm.weight"] = p["embeddings/layer_norm/gamma"]
    sd["bert.embeddings.LayerNorm.bias"] = p["embeddings/layer_norm/beta"]
    for i in range(cfg.num_layers):
        t, h = f"transformer/layer_{i}/", f"bert.encoder.layer.{i}."
        for n in ("query", "key", "value"):
            sd[h + f"attention.self.{n}.weight"] = p[t + f"self_attention/{n}/kernel"].reshape(H, H).t()   # [H,N,D] -> [out,in]
            sd[h + f"attention.self.{n}.bias"] = p[t + f"self_attention/{n}/bias"].reshape(H)
        sd[h + "attention.output.dense.weight"] = p[t + "self_attention/attention_output/kernel"].reshape(H, H).t()   # [N,D,H]
        sd[h + "attention.output.dense.bias"] = p[t + "self_attention/attention_output/bias"]
        sd[h + "attention.output.LayerNorm.weight"] = p[t + "self_attention_layer_norm/gamma"]
        sd[h + "attention.output.LayerNorm.bias"] = p[t + "self_attention_layer_norm/beta"]
        sd[h + "intermediate.dense.weight"] = p[t + "intermediate/kernel"].t()
        sd[h + "intermediate.dense.bias"] = p[t + "intermediate/bias"]
        sd[h + "output.dense.weight"] = p[t + "output/kernel"].t()
        sd[h + "output.dense.bias"] = p[t + "output/bias"]
        sd[h + "output.LayerNorm.weight"] = p[t + "output_layer_norm/gamma"]
        sd[h + "output.LayerNorm.bias"] = p[t + "output_layer_norm/beta"]
    sd["cls.predictions.transform.dense.weight"] = p["cls/predictions/transform/dense/kernel"].t()
    sd["cls.predictions.transform.dense.bias"] = p["cls/predictions/transform/dense/bias"]
    sd["cls.predictions.transform.LayerNorm.weight"] = p["cls/predictions/transform/LayerNorm/gamma"]
    sd["cls.predictions.transform.LayerNorm.bias"] = p["cls/predictions/transform/LayerNorm/beta"]
    sd["cls.predictions.bias"] = p["cls/predictions/output_bias/bias"]
    sd["cls.predictions.decoder.bias"] = p["cls/predictions/output_bias/bias"]
    sd["cls.predictions.decoder.weight"] = p["word_embeddings/embeddings"]          # tied output projection
    missing, unexpected = m.load_state_dict({k: v.clone().contiguous() for k, v in sd.items()}, strict=False)
    assert not unexpected and all("position_ids" in k or "token_type_ids" in k for k in missing), (missing, unexpected)
    return m


@pytest.mark.parametrize("shape", [(211, 64, 2, 2, 24, 128), (97, 128, 1, 4, 17, 256)])
def test_oracle_matches_huggingface_bert(shape):
    V, H, L, N, S, I = shape
    cfg = om.Config(vocab_size=V, hidden_size=H, num_layers=L, num_attention_heads=N, max_sequence_length=S, inner_dim=I)
    p = om.init_params(cfg, 0)
    g = torch.Generator().manual_seed(1)
    for k, v in p.items():      # non-trivial biases / LayerNorm parameters, larger weights: every term matters
        if k.endswith(("bias", "beta")):
            v.copy_(torch.randn(v.shape, generator=g) * 0.2)
        elif k.endswith("gamma"):
            v.copy_(1.0 + torch.randn(v.shape, generator=g) * 0.2)
        else:
            v.mul_(4.0)
    B, P = 5, 6
    batch = make_batch(B, S, P, V, seed=9)
    out = om.model_forward(p, cfg, batch, training=False)
    hf = _hf_from_oracle(cfg, p)
    with torch.no_grad():
        base = hf.bert(input_ids=batch["input_word_ids"], attention_mask=batch["input_mask"],
                       token_type_ids=torch.zeros_like(batch["input_word_ids"]))
        seq = base.last_hidden_state
        logits_all = hf.cls(seq)                                                     # [B, S, V]
    valid = batch["input_mask"].bool()
    # padded QUERY rows are computed by both (only keys are masked); compare every row
    assert float((seq - out["sequence_output"]).abs().max()) < 2e-5
    pos = batch["masked_lm_positions"]
    logits = torch.gather(logits_all, 1, pos.unsqueeze(-1).expand(B, P, V))
    assert float((logits - out["mlm_logits"]).abs().max()) < 1e-4
    # pooler: tanh(dense(first token)) (bert4rec_encoder.py:224-226)
    with torch.no_grad():
        from transformers.models.bert.modeling_bert import BertPooler
        pooler = BertPooler(hf.config)
        pooler.dense.weight.copy_(p["pooler_transform/kernel"].t()); pooler.dense.bias.copy_(p["pooler_transform/bias"])
        assert float((pooler(seq) - out["pooled_output"]).abs().max()) < 2e-5
    # masked sparse CE (trainer_utils.py:12-23) == token-level cross entropy ignoring the padded slots
    y = batch["masked_lm_ids"]
    ref_loss = torch.nn.functional.cross_entropy(logits.reshape(-1, V), torch.where(y != 0, y, torch.full_like(y, -100)).reshape(-1),
                                                 ignore_index=-100)
    assert abs(float(om.masked_sparse_ce(y, out["mlm_logits"])) - float(ref_loss)) < 1e-5
    assert valid.any()


def test_oracle_gradients_match_huggingface_bert():
    """Same comparison for the gradients of the masked CE w.r.t. every weight (autograd through both implementations)."""
    V, H, L, N, S, I = 151, 64, 2, 2, 20, 128
    cfg = om.Config(vocab_size=V, hidden_size=H, num_layers=L, num_attention_heads=N, max_sequence_length=S, inner_dim=I)
    p = om.init_params(cfg, 3)
    for k, v in p.items():
        if not k.endswith(("bias", "beta", "gamma")):
            v.mul_(4.0)
    batch = make_batch(4, S, 5, V, seed=2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    y = batch["masked_lm_ids"]
    loss = om.masked_sparse_ce(y, om.model_forward(leaves, cfg, batch, training=False)["mlm_logits"])
    loss.backward()
    hf = _hf_from_oracle(cfg, p).train()          # dropout probabilities are 0: train mode only enables autograd paths
    seq = hf.bert(input_ids=batch["input_word_ids"], attention_mask=batch["input_mask"],
                  token_type_ids=torch.zeros_like(batch["input_word_ids"])).last_hidden_state
    logits = torch.gather(hf.cls(seq), 1, batch["masked_lm_positions"].unsqueeze(-1).expand(4, 5, V))
    hloss = torch.nn.functional.cross_entropy(logits.reshape(-1, V), torch.where(y != 0, y, torch.full_like(y, -100)).reshape(-1),
                                              ignore_index=-100)
    hloss.backward()
    hp = dict(hf.named_parameters())
    pairs = {
        "word_embeddings/embeddings": hp["bert.embeddings.word_embeddings.weight"].grad,       # tied: gather + projection parts
        "position_embedding/embeddings": hp["bert.embeddings.position_embeddings.weight"].grad,
        "transformer/layer_0/intermediate/kernel": hp["bert.encoder.layer.0.intermediate.dense.weight"].grad.t(),
        "transformer/layer_1/output/kernel": hp["bert.encoder.layer.1.output.dense.weight"].grad.t(),
        "transformer/layer_0/self_attention/value/kernel": hp["bert.encoder.layer.0.attention.self.value.weight"].grad.t().reshape(H, N, H // N),
        "transformer/layer_1/self_attention/attention_output/kernel": hp["bert.encoder.layer.1.attention.output.dense.weight"].grad.t().reshape(N, H // N, H),
        "transformer/layer_0/self_attention_layer_norm/gamma": hp["bert.encoder.layer.0.attention.output.LayerNorm.weight"].grad,
        "cls/predictions/transform/dense/kernel": hp["cls.predictions.transform.dense.weight"].grad.t(),
        "cls/predictions/output_bias/bias": hp["cls.predictions.bias"].grad,
    }
    for k, g in pairs.items():
        ref = leaves[k].grad
        assert float((ref - g).norm()) <= 1e-4 * float(g.norm()) + 1e-7, k
