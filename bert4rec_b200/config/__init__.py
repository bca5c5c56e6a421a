"""Encoder hyper-parameter sets of the reference's 13 training configs
(``bert4rec/config/bert4rec_train_configs/{dataset}_{64,128,256}.json``), kept as one table.
``encoder_config("ml-1m_64")`` returns the kwargs for ``networks.Bert4RecEncoder(vocab_size, **cfg)``."""

# name: (hidden, inner, max_seq_len, heads, layers, attention_dropout, output_dropout)
_TABLE = {
    "beauty_64": (64, 64, 50, 2, 2, 0.2, 0.5),
    "beauty_128": (128, 512, 50, 4, 2, 0.2, 0.5),
    "beauty_256": (256, 1024, 50, 8, 2, 0.2, 0.5),
    "ml-1m_64": (64, 256, 200, 2, 2, 0.2, 0.2),
    "ml-1m_128": (128, 512, 200, 4, 2, 0.2, 0.5),
    "ml-1m_256": (256, 512, 200, 8, 2, 0.2, 0.5),
    "ml-20m_64": (64, 256, 200, 2, 2, 0.1, 0.1),
    "ml-20m_128": (128, 512, 200, 4, 2, 0.1, 0.1),
    "ml-20m_256": (256, 1024, 200, 8, 2, 0.1, 0.1),
    "reddit_128": (128, 512, 200, 4, 2, 0.1, 0.1),
    "steam_64": (64, 256, 50, 2, 2, 0.1, 0.1),
    "steam_128": (128, 512, 50, 4, 2, 0.1, 0.1),
    "steam_256": (256, 1024, 50, 8, 2, 0.2, 0.2),
}


def encoder_config(name: str) -> dict:
    if name.endswith(".json"):
        name = name[:-5]
    if name not in _TABLE:
        raise FileNotFoundError(f"No config named {name}; known: {sorted(_TABLE)}")
    h, i, s, n, l, ad, od = _TABLE[name]
    return dict(attention_dropout=ad, output_dropout=od, hidden_size=h, inner_dim=i, max_sequence_length=s,
                num_attention_heads=n, num_layers=l)


def available_configs():
    return sorted(_TABLE)
