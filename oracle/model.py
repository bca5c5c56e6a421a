"""Oracle restatement of the reference's floating-point path (torch, CPU, fp32 or fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **parity unpinned**: the
arithmetic restated here lives in tensorflow==2.10.0 / keras==2.10.0 /
tf-models-official==2.10.1 (Pipfile.lock pins), none of which is vendored under
/root/reference or installable here, and no reference test checks a single
floating-point value of this path (bert4rec_encoder_tests.py:152-153).  The
restatement follows the published semantics of those layers (SURVEY.md
Appendix A) at the reference's call sites, op for op, *including* the tensors
the B200 path never materialises (the [B,S,S] attention mask, the [B,N,S,S]
scores, the [B,P,V] logits over ALL P slots) -- which is also what makes it the
honest CPU baseline for bench.py.

Weights are a flat dict keyed by the TF variable names with TF shapes:
    word_embeddings/embeddings [V,H]; position_embedding/embeddings [Smax,H];
    embeddings/layer_norm/{gamma,beta} [H];
    transformer/layer_{i}/self_attention/{query,key,value}/{kernel [H,N,D], bias [N,D]};
    transformer/layer_{i}/self_attention/attention_output/{kernel [N,D,H], bias [H]};
    transformer/layer_{i}/self_attention_layer_norm/{gamma,beta};
    transformer/layer_{i}/intermediate/{kernel [H,I], bias [I]};
    transformer/layer_{i}/output/{kernel [I,H], bias [H]};
    transformer/layer_{i}/output_layer_norm/{gamma,beta};
    pooler_transform/{kernel [H,H], bias [H]};
    cls/predictions/transform/dense/{kernel [H,H], bias [H]};
    cls/predictions/transform/LayerNorm/{gamma,beta}; cls/predictions/output_bias/bias [V].
"""
import math
import re
from dataclasses import dataclass

import torch

LN_EPS = 1e-12  # bert4rec_encoder.py:116-117; TransformerEncoderBlock norm_epsilon default; MaskedLM LayerNorm


@dataclass
class Config:
    vocab_size: int
    hidden_size: int = 64
    num_layers: int = 2
    num_attention_heads: int = 2
    max_sequence_length: int = 200
    inner_dim: int = 256
    output_dropout: float = 0.1
    attention_dropout: float = 0.1


def param_shapes(cfg: Config):
    """Variable name -> shape, in creation order (bert4rec_encoder.py:102-153, bert4rec_model.py:76-81)."""
    V, H, L, N, I = cfg.vocab_size, cfg.hidden_size, cfg.num_layers, cfg.num_attention_heads, cfg.inner_dim
    D = H // N
    s = {}
    s["word_embeddings/embeddings"] = (V, H)
    s["position_embedding/embeddings"] = (cfg.max_sequence_length, H)
    s["embeddings/layer_norm/gamma"] = (H,)
    s["embeddings/layer_norm/beta"] = (H,)
    for i in range(L):
        p = f"transformer/layer_{i}/"
        for n in ("query", "key", "value"):
            s[p + f"self_attention/{n}/kernel"] = (H, N, D)
            s[p + f"self_attention/{n}/bias"] = (N, D)
        s[p + "self_attention/attention_output/kernel"] = (N, D, H)
        s[p + "self_attention/attention_output/bias"] = (H,)
        s[p + "self_attention_layer_norm/gamma"] = (H,)
        s[p + "self_attention_layer_norm/beta"] = (H,)
        s[p + "intermediate/kernel"] = (H, I)
        s[p + "intermediate/bias"] = (I,)
        s[p + "output/kernel"] = (I, H)
        s[p + "output/bias"] = (H,)
        s[p + "output_layer_norm/gamma"] = (H,)
        s[p + "output_layer_norm/beta"] = (H,)
    s["pooler_transform/kernel"] = (H, H)
    s["pooler_transform/bias"] = (H,)
    s["cls/predictions/transform/dense/kernel"] = (H, H)
    s["cls/predictions/transform/dense/bias"] = (H,)
    s["cls/predictions/transform/LayerNorm/gamma"] = (H,)
    s["cls/predictions/transform/LayerNorm/beta"] = (H,)
    s["cls/predictions/output_bias/bias"] = (V,)
    return s


def init_params(cfg: Config, seed=0, dtype=torch.float32):
    """TruncatedNormal(0.02) for tables/kernels/pooler (bert4rec_encoder.py:73-74,106,112,145,152),
    glorot-uniform MLM dense (bert4rec_model.py:43,79), zeros biases, ones/zeros LN."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("gamma"):
            t = torch.ones(shape, dtype=torch.float64)
        elif name.endswith("beta") or name.endswith("bias"):
            t = torch.zeros(shape, dtype=torch.float64)
        elif name == "cls/predictions/transform/dense/kernel":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        else:
            t = torch.empty(shape, dtype=torch.float64)
            torch.nn.init.trunc_normal_(t, mean=0.0, std=0.02, a=-0.04, b=0.04, generator=g)
        out[name] = t.to(dtype)
    return out


def layer_norm(x, gamma, beta, eps=LN_EPS):
    """Keras LayerNormalization non-fused path (eps < 1.001e-5): biased variance, rsqrt(var+eps)."""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    inv = torch.rsqrt(var + eps) * gamma
    return x * inv + (beta - mean * inv)


def gelu_erf(x):
    """keras.activations.gelu(approximate=False)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _drop(x, rate, keep_masks, key, training):
    """Inverted dropout.  With ``keep_masks`` the caller supplies the 0/1 keep mask for each
    site (so the oracle can replay the CUDA path's Philox masks); without, torch's RNG."""
    if not training or rate <= 0.0:
        return x
    if keep_masks is not None:
        m = keep_masks[key].to(x.dtype).reshape(x.shape)
    else:
        m = (torch.rand_like(x) >= rate).to(x.dtype)
    return x * m / (1.0 - rate)


def encoder_forward(params, cfg: Config, input_word_ids, input_mask, training=False, keep_masks=None,
                    return_intermediates=False):
    """Bert4RecEncoder.call (bert4rec_encoder.py:186-231)."""
    dt = params["word_embeddings/embeddings"].dtype
    B, S = input_word_ids.shape
    H, N = cfg.hidden_size, cfg.num_attention_heads
    D = H // N
    inter = {}
    # OnDeviceEmbedding (:199) + PositionEmbedding (:207) + add/LN/dropout (:209-211)
    x = params["word_embeddings/embeddings"][input_word_ids.long()]
    x = x + params["position_embedding/embeddings"][:S].unsqueeze(0)
    x = layer_norm(x, params["embeddings/layer_norm/gamma"], params["embeddings/layer_norm/beta"])
    x = _drop(x, cfg.output_dropout, keep_masks, "emb", training)
    inter["emb"] = x
    # SelfAttentionMask (:216): mask[b,i,j] = input_mask[b,j], materialised [B,S,S]
    attn_mask = torch.ones(B, S, 1, dtype=dt) * input_mask.to(dt).reshape(B, 1, S)
    outs = []
    for i in range(cfg.num_layers):
        p = f"transformer/layer_{i}/"
        q = torch.einsum("abc,cde->abde", x, params[p + "self_attention/query/kernel"]) + params[p + "self_attention/query/bias"]
        k = torch.einsum("abc,cde->abde", x, params[p + "self_attention/key/kernel"]) + params[p + "self_attention/key/bias"]
        v = torch.einsum("abc,cde->abde", x, params[p + "self_attention/value/kernel"]) + params[p + "self_attention/value/bias"]
        q = q * (1.0 / math.sqrt(float(D)))
        scores = torch.einsum("aecd,abcd->acbe", k, q)  # [B,N,Sq,Sk]
        scores = scores + (1.0 - attn_mask[:, None, :, :]) * -1e9
        probs = torch.softmax(scores, dim=-1)
        probs = _drop(probs, cfg.attention_dropout, keep_masks, f"l{i}.attn", training)
        ctx = torch.einsum("acbe,aecd->abcd", probs, v)
        attn = torch.einsum("abcd,cde->abe", ctx, params[p + "self_attention/attention_output/kernel"]) \
            + params[p + "self_attention/attention_output/bias"]
        attn = _drop(attn, cfg.output_dropout, keep_masks, f"l{i}.attn_out", training)
        y = layer_norm(x + attn, params[p + "self_attention_layer_norm/gamma"], params[p + "self_attention_layer_norm/beta"])
        h = torch.einsum("abc,cd->abd", y, params[p + "intermediate/kernel"]) + params[p + "intermediate/bias"]
        h = gelu_erf(h)
        o = torch.einsum("abc,cd->abd", h, params[p + "output/kernel"]) + params[p + "output/bias"]
        o = _drop(o, cfg.output_dropout, keep_masks, f"l{i}.ffn_out", training)
        x = layer_norm(o + y, params[p + "output_layer_norm/gamma"], params[p + "output_layer_norm/beta"])
        outs.append(x)
        if return_intermediates:
            inter[f"l{i}.q"], inter[f"l{i}.k"], inter[f"l{i}.v"] = q, k, v
            inter[f"l{i}.ctx"], inter[f"l{i}.y"] = ctx.reshape(B, S, H), y
    pooled = torch.tanh(outs[-1][:, 0, :] @ params["pooler_transform/kernel"] + params["pooler_transform/bias"])
    res = dict(sequence_output=outs[-1], pooled_output=pooled, encoder_outputs=outs)
    if return_intermediates:
        res["intermediates"] = inter
    return res


def masked_lm(params, sequence_output, masked_lm_positions):
    """tfm.nlp.layers.MaskedLM (bert4rec_model.py:76-81,143): ALL P slots, padded slots gather position 0."""
    B, S, H = sequence_output.shape
    P = masked_lm_positions.shape[1]
    flat = masked_lm_positions.long() + (torch.arange(B).unsqueeze(1) * S)
    g = sequence_output.reshape(B * S, H)[flat.reshape(-1)]
    t = gelu_erf(g @ params["cls/predictions/transform/dense/kernel"] + params["cls/predictions/transform/dense/bias"])
    t = layer_norm(t, params["cls/predictions/transform/LayerNorm/gamma"], params["cls/predictions/transform/LayerNorm/beta"])
    logits = t @ params["word_embeddings/embeddings"].t() + params["cls/predictions/output_bias/bias"]
    return logits.reshape(B, P, -1)


def model_forward(params, cfg, batch, training=False, keep_masks=None):
    """BERT4RecModel.call (bert4rec_model.py:110-149)."""
    out = encoder_forward(params, cfg, batch["input_word_ids"], batch["input_mask"], training, keep_masks)
    if "masked_lm_positions" in batch:
        out["mlm_logits"] = masked_lm(params, out["sequence_output"], batch["masked_lm_positions"])
    return out


def masked_sparse_ce(y_true, logits, pad_token=0):
    """MaskedSparseCategoricalCrossentropy.call (trainer_utils.py:12-23)."""
    mask = (y_true != pad_token)
    lse = torch.logsumexp(logits, dim=-1)
    picked = torch.gather(logits, -1, y_true.long().unsqueeze(-1)).squeeze(-1)
    per = lse - picked
    m = mask.to(per.dtype)
    return (per * m).sum() / m.sum()


def masked_accuracy(y_true, logits):
    """trainer_utils.py:49-60 (tf.argmax -> first index of the maximum)."""
    pred = first_argmax(logits)
    mask = y_true != 0
    match = (pred == y_true.long()) & mask
    return match.float().sum() / mask.float().sum()


def sparse_categorical_accuracy(y_true, logits):
    """keras.metrics.SparseCategoricalAccuracy over ALL slots (bert4rec_trainer.py:30)."""
    return (first_argmax(logits) == y_true.long()).float().mean()


def first_argmax(x):
    mx = x.max(dim=-1, keepdim=True).values
    V = x.shape[-1]
    idx = torch.arange(V).expand_as(x)
    return torch.where(x == mx, idx, torch.full_like(idx, V)).min(dim=-1).values


# ----------------------------------------------------------------------------- optimizer
def lr_schedule(step, init_lr=1e-4, num_train_steps=400000, num_warmup_steps=100, end_lr=0.0):
    """WarmUp.__call__ (adam_w_optimizer.py:22-36) over PolynomialDecay(power=1) (optimizers/__init__.py:38-46).
    fp32 scalar arithmetic as TF does it."""
    f32 = torch.float32
    s = torch.tensor(float(step), dtype=f32)
    if num_warmup_steps and float(s) < float(num_warmup_steps):
        return float(torch.tensor(init_lr, dtype=f32) * (s / torch.tensor(float(num_warmup_steps), dtype=f32)))
    st = torch.minimum(s, torch.tensor(float(num_train_steps), dtype=f32))
    p = st / torch.tensor(float(num_train_steps), dtype=f32)
    return float((torch.tensor(init_lr, dtype=f32) - end_lr) * (1 - p) + end_lr)


def uses_weight_decay(name, exclude=("LayerNorm", "layer_norm", "bias")):
    """AdamWeightDecay._do_use_weight_decay (adam_w_optimizer.py:154-168)."""
    return not any(re.search(r, name) for r in exclude)


def clip_by_global_norm(grads, clip=5.0):
    """tf.clip_by_global_norm: g * clip * min(1/norm, 1/clip)."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).to(next(iter(grads.values())).dtype)
    scale = clip * torch.minimum(1.0 / gn, torch.tensor(1.0 / clip, dtype=gn.dtype))
    return {k: g * scale for k, g in grads.items()}, gn


class AdamW:
    """AdamWeightDecay.apply_gradients (adam_w_optimizer.py:100-136) over Keras optimizer_v2 Adam
    (ResourceApplyAdam, epsilon-hat form)."""

    def __init__(self, params, init_lr=1e-4, num_train_steps=400000, num_warmup_steps=100, end_lr=0.0,
                 weight_decay_rate=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-6, clip_norm=5.0):
        self.hp = dict(init_lr=init_lr, num_train_steps=num_train_steps, num_warmup_steps=num_warmup_steps, end_lr=end_lr)
        self.wd, self.b1, self.b2, self.eps, self.clip = weight_decay_rate, beta_1, beta_2, epsilon, clip_norm
        self.iterations = 0
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}

    def apply(self, params, grads):
        """grads: name -> tensor; names absent (pooler) are skipped like Keras skips None grads."""
        if self.clip > 0:
            grads, gn = clip_by_global_norm(grads, self.clip)
        lr = lr_schedule(self.iterations, **self.hp)
        t = self.iterations + 1
        alpha = lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
        for k, g in grads.items():
            w = params[k]
            if self.wd and uses_weight_decay(k):
                w -= lr * w * self.wd
            self.m[k] += (g - self.m[k]) * (1 - self.b1)
            self.v[k] += (g * g - self.v[k]) * (1 - self.b2)
            w -= alpha * self.m[k] / (torch.sqrt(self.v[k]) + self.eps)
        self.iterations += 1
        return lr


def train_step(params, cfg, batch, opt: AdamW, keep_masks=None, training=True):
    """BERT4RecModel.train_step (bert4rec_model.py:151-173): fwd(training) -> loss -> grads -> apply."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    out = model_forward(leaves, cfg, batch, training=training, keep_masks=keep_masks)
    y = batch["masked_lm_ids"]
    loss = masked_sparse_ce(y, out["mlm_logits"])
    names = [k for k in leaves]
    gs = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    grads = {k: g for k, g in zip(names, gs) if g is not None}
    with torch.no_grad():
        lr = opt.apply(params, grads)
        metrics = dict(loss=float(loss),
                       sparse_categorical_accuracy=float(sparse_categorical_accuracy(y, out["mlm_logits"])),
                       masked_accuracy=float(masked_accuracy(y, out["mlm_logits"])))
    return metrics, grads, lr


# ----------------------------------------------------------------------------- ranking
@torch.no_grad()
def rank_items(params, cfg, batch, items=None):
    """BERT4RecModel.rank_items (bert4rec_model.py:203-240): full-vocab logits for all P slots, keep
    slots with weight 1, per slot gather candidates + stable descending sort."""
    logits = model_forward(params, cfg, batch, training=False)["mlm_logits"]
    B, P, V = logits.shape
    w = batch["masked_lm_weights"].bool() if "masked_lm_weights" in batch else torch.ones(B, P, dtype=torch.bool)
    rankings = []
    for b in range(B):
        row = []
        slot = 0
        for p in range(P):
            if not w[b, p]:
                continue
            tok = logits[b, p]
            if items is not None:
                cand = torch.as_tensor(items[b][slot], dtype=torch.long)
                s = tok[cand]
                order = torch.argsort(s, descending=True, stable=True)
                row.append(cand[order])
            else:
                row.append(torch.argsort(tok, descending=True, stable=True))
            slot += 1
        rankings.append(row)
    return rankings
