"""Host-side driver of libb4r: flat parameter store + shape-specialised sessions.

torch is used for device memory, streams and (in the trainer) torch.distributed only; every kernel on the path is
launched through the C ABI (include/b4r.h).  Nothing here falls back to torch ops for compute.
"""
import ctypes as C
import math
import os

import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dl_view(t):
    """Zero-copy DLPack hand-off: validates the tensor through the C ABI's DLPack adapter (device, contiguity) and
    returns the raw device pointer the compute entry points take."""
    cap = torch.utils.dlpack.to_dlpack(t)
    C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
    C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
    p = C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor")
    v = _lib.DLView()
    check(_lib.load().b4r_dl_view_of(C.c_void_p(p), C.byref(v)))
    # the capsule still owns the DLManagedTensor; run its deleter by re-importing it (keeps the storage alive in `t`)
    torch.utils.dlpack.from_dlpack(cap)
    return v


def _dl_ptr(t):
    """Device pointer of an input tensor handed over as a DLPack capsule (checked by the C ABI's adapter: CUDA, row-major)."""
    return C.c_void_p(dl_view(t).data) if t is not None else C.c_void_p(0)


class ParamStore:
    """Flat fp32 master parameters + bf16 shadow + (optionally) gradient / Adam moment buffers on one device."""

    def __init__(self, vocab_size, hidden_size, num_layers, num_attention_heads, max_sequence_length, inner_dim,
                 output_dropout=0.1, attention_dropout=0.1, device="cuda:0"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.B4RError("bert4rec_b200 runs on CUDA devices only (no CPU fallback)")
        self.cfg = _lib.Config(vocab_size, hidden_size, num_layers, num_attention_heads, max_sequence_length,
                               inner_dim, float(output_dropout), float(attention_dropout))
        n = self.lib.b4r_param_entries(C.byref(self.cfg), None, 0)
        if n < 0:
            raise ValueError(self.lib.b4r_last_error().decode())
        ents = (_lib.ParamEntry * n)()
        self.lib.b4r_param_entries(C.byref(self.cfg), ents, n)
        self.entries = {e.name.decode(): (e.offset, e.numel, e.rows, e.cols, e.group) for e in ents}
        nd, nt, na = C.c_int64(), C.c_int64(), C.c_int64()
        check(self.lib.b4r_param_counts(C.byref(self.cfg), C.byref(nd), C.byref(nt), C.byref(na)))
        self.n_decay, self.n_trainable, self.n_total = nd.value, nt.value, na.value
        with torch.cuda.device(self.device):
            check(self.lib.b4r_device_check(self.device.index or 0))
            self.params = torch.zeros(self.n_total, dtype=torch.float32, device=self.device)
            self.shadow = torch.zeros(self.n_total, dtype=torch.bfloat16, device=self.device)
        self.grads = None
        self.p2p = None      # symmetric-memory handles of the peer-memory gradient all-reduce (data-parallel NCCL worlds)
        self.m = self.v = None
        self.step_counter = None
        self.sessions = {}
        self.generation = 0     # bumped whenever sessions (and their workspaces) are re-created: captured graphs are stale

    # ---- hyper-parameters
    @property
    def V(self): return self.cfg.vocab_size
    @property
    def H(self): return self.cfg.hidden_size
    @property
    def L(self): return self.cfg.num_layers
    @property
    def N(self): return self.cfg.num_heads
    @property
    def I(self): return self.cfg.inner_dim

    def _alloc_grads(self):
        """The flat gradient buffer.  In a NCCL world of 2..16 ranks it is allocated in symmetric memory and exchanged with the other
        ranks (a collective: every rank creates its training buffers at the same point), so that the data-parallel all-reduce can be
        the library's own peer-memory kernel (csrc/k_p2p.cu); otherwise (single process, gloo, B4R_DISABLE_P2P_ALLREDUCE=1, or a
        platform without peer access) a plain tensor, reduced by torch.distributed.all_reduce."""
        self.p2p = None
        import torch.distributed as dist
        if (dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
                and 2 <= dist.get_world_size() <= self.lib.b4r_p2p_allreduce_max_world()
                and not os.environ.get("B4R_DISABLE_P2P_ALLREDUCE")):
            try:
                import torch.distributed._symmetric_memory as symm
                with torch.cuda.device(self.device):
                    grads = symm.empty(self.n_total, dtype=torch.float32, device=self.device)
                    flags = symm.empty(3 * self.lib.b4r_p2p_allreduce_max_world(), dtype=torch.int32, device=self.device)
                    grads.zero_(); flags.zero_()
                    torch.cuda.synchronize(self.device)
                    hg = symm.rendezvous(grads, dist.group.WORLD)
                    hf = symm.rendezvous(flags, dist.group.WORLD)
                    state = torch.zeros(8, dtype=torch.int32, device=self.device)
                    torch.cuda.synchronize(self.device)
                dist.barrier()     # every rank's flags are zero before anyone signals
                # NVSwitch multicast mapping of the gradient buffer (0 when the platform has none): the one-pass NVLS variant
                mc = 0 if os.environ.get("B4R_DISABLE_NVLS") else int(getattr(hg, "multicast_ptr", 0) or 0)
                self.p2p = dict(hg=hg, hf=hf, flags=flags, state=state, rank=dist.get_rank(), world=dist.get_world_size(), mc=mc)
                return grads
            except Exception as e:   # noqa: BLE001
                import warnings
                warnings.warn(f"peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using torch.distributed.all_reduce")
                self.p2p = None
        return torch.zeros(self.n_total, dtype=torch.float32, device=self.device)

    def all_reduce_grads(self, n_floats, stream=None):
        """SUM over ranks of grads[:n_floats], in place, on the current stream (capturable when the peer-memory kernel is in use)."""
        if self.p2p is None:
            torch.distributed.all_reduce(self.grads[:n_floats])
            return
        p = self.p2p
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        check(self.lib.b4r_p2p_allreduce_f32(C.c_void_p(p["hg"].buffer_ptrs_dev), C.c_void_p(p["hf"].buffer_ptrs_dev), C.c_void_p(p["mc"]), 0,
                                             int(n_floats), p["rank"], p["world"], _ptr(p["state"]), C.c_void_p(st)))

    def p2p_error(self):
        """0, or the code a peer-memory all-reduce left behind when a rank did not arrive within its bounded wait (synchronises)."""
        return 0 if self.p2p is None else int(self.p2p["state"][7].item())

    def ensure_training_buffers(self):
        if self.grads is None:
            self.grads = self._alloc_grads()
            self.m = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
            self.v = torch.zeros(self.n_trainable, dtype=torch.float32, device=self.device)
            self.step_counter = torch.zeros(1, dtype=torch.int64, device=self.device)
            self.adam_scratch = torch.zeros(self.lib.b4r_adamw_scratch_floats(), dtype=torch.float32, device=self.device)
            self.lr_out = torch.zeros(2, dtype=torch.float32, device=self.device)
            if self.sessions:
                for s in self.sessions.values():
                    s.close()
                self.sessions = {}
                self.generation += 1

    # ---- views
    def seg(self, name, buf=None):
        off, numel, rows, cols, _ = self.entries[name]
        t = (self.params if buf is None else buf)[off:off + numel]
        return t.view(rows, cols) if cols else t

    def tf_views(self, buf=None):
        """TF variable name -> tensor *view* with the TF shape (SURVEY.md Appendix A name map)."""
        H, N = self.H, self.N
        D = H // N
        out = {}
        s = lambda n: self.seg(n, buf)
        out["word_embeddings/embeddings"] = s("word_embeddings")
        out["position_embedding/embeddings"] = s("position_embedding")
        out["embeddings/layer_norm/gamma"] = s("emb_ln/gamma")
        out["embeddings/layer_norm/beta"] = s("emb_ln/beta")
        for i in range(self.L):
            p, q = f"transformer/layer_{i}/", f"layer_{i}/"
            for k, n in enumerate(("query", "key", "value")):
                out[p + f"self_attention/{n}/kernel"] = s(q + "wqkv")[:, k * H:(k + 1) * H].unflatten(1, (N, D))
                out[p + f"self_attention/{n}/bias"] = s(q + "bqkv")[k * H:(k + 1) * H].view(N, D)
            out[p + "self_attention/attention_output/kernel"] = s(q + "wo").view(N, D, H)
            out[p + "self_attention/attention_output/bias"] = s(q + "bo")
            out[p + "self_attention_layer_norm/gamma"] = s(q + "ln1/gamma")
            out[p + "self_attention_layer_norm/beta"] = s(q + "ln1/beta")
            out[p + "intermediate/kernel"] = s(q + "w1")
            out[p + "intermediate/bias"] = s(q + "b1")
            out[p + "output/kernel"] = s(q + "w2")
            out[p + "output/bias"] = s(q + "b2")
            out[p + "output_layer_norm/gamma"] = s(q + "ln2/gamma")
            out[p + "output_layer_norm/beta"] = s(q + "ln2/beta")
        out["pooler_transform/kernel"] = s("pooler/w")
        out["pooler_transform/bias"] = s("pooler/b")
        out["cls/predictions/transform/dense/kernel"] = s("head/wt")
        out["cls/predictions/transform/dense/bias"] = s("head/bt")
        out["cls/predictions/transform/LayerNorm/gamma"] = s("head/ln/gamma")
        out["cls/predictions/transform/LayerNorm/beta"] = s("head/ln/beta")
        out["cls/predictions/output_bias/bias"] = s("head/output_bias")
        return out

    def state_dict(self):
        return {k: v.detach().clone().cpu() for k, v in self.tf_views().items()}

    def grad_dict(self):
        return {k: v.detach().clone().cpu() for k, v in self.tf_views(self.grads).items()}

    def load_state_dict(self, sd):
        views = self.tf_views()
        missing = [k for k in views if k not in sd]
        if missing:
            raise KeyError(f"missing weights: {missing[:4]}...")
        with torch.no_grad():
            for k, v in views.items():
                v.copy_(torch.as_tensor(sd[k]).to(device=self.device, dtype=torch.float32).reshape(v.shape))
        self.sync_shadow()

    def sync_shadow(self):
        with torch.no_grad():
            self.shadow.copy_(self.params)  # device-side cast; weights loading is off the hot path

    def init_weights(self, seed=0, mlm_initializer="glorot_uniform"):
        """TruncatedNormal(0.02) tables / kernels / pooler, glorot-uniform MLM dense, zero biases, unit LN
        (bert4rec_encoder.py:73-74,106,112,145,152; bert4rec_model.py:43,79)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        with torch.no_grad():
            for name, v in self.tf_views().items():
                if name.endswith("gamma"):
                    v.fill_(1.0)
                elif name.endswith("beta") or name.endswith("bias"):
                    v.zero_()
                elif name == "cls/predictions/transform/dense/kernel" and mlm_initializer == "glorot_uniform":
                    lim = math.sqrt(6.0 / (v.shape[0] + v.shape[1]))
                    v.copy_(((torch.rand(v.shape, generator=g) * 2 - 1) * lim).to(self.device))
                else:
                    t = torch.empty(tuple(v.shape))
                    torch.nn.init.trunc_normal_(t, 0.0, 0.02, -0.04, 0.04, generator=g)
                    v.copy_(t.to(self.device))
        self.sync_shadow()

    # ---- sessions
    def session(self, batch, seq_len, max_pred):
        key = (batch, seq_len, max_pred)
        s = self.sessions.get(key)
        if s is None:
            s = Session(self, batch, seq_len, max_pred)
            self.sessions[key] = s
        return s

    # ---- optimizer
    def adamw_step(self, hp, count=None, grad_scale=1.0):
        self.ensure_training_buffers()
        check(self.lib.b4r_adamw_step(_ptr(self.params), _ptr(self.shadow), _ptr(self.grads), _ptr(self.m), _ptr(self.v),
                                      self.n_decay, self.n_trainable, C.byref(hp), _ptr(count), float(grad_scale),
                                      _ptr(self.step_counter), _ptr(self.adam_scratch), _ptr(self.lr_out), _stream()))


class Session:
    """Static-shape (batch, seq_len, max_predictions_per_seq) execution context over a ParamStore."""

    def __init__(self, store, batch, seq_len, max_pred):
        self.store, self.lib = store, store.lib
        self.B, self.S, self.P = batch, seq_len, max_pred
        self.Mcap = batch * max_pred + batch
        nbytes = self.lib.b4r_session_workspace_bytes(C.byref(store.cfg), batch, seq_len, max_pred)
        if nbytes == 0:
            raise ValueError(self.lib.b4r_last_error().decode())
        with torch.cuda.device(store.device):
            self.ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=store.device)
            off = (-self.ws.data_ptr()) % 256
            self._ws_ptr = self.ws.data_ptr() + off
            h = C.c_void_p()
            check(self.lib.b4r_session_create(C.byref(store.cfg), batch, seq_len, max_pred, _ptr(store.params),
                                              _ptr(store.shadow), _ptr(store.grads), C.c_void_p(self._ws_ptr), nbytes,
                                              C.byref(h)))
        self.h = h
        self._keep = []

    def close(self):
        if self.h:
            self.lib.b4r_session_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- steps (all enqueue on torch's current stream)
    def encode(self, ids, mask, training=False, seed=0, step=0, step_counter=None):
        assert ids.dtype == torch.int64 and mask.dtype == torch.int64 and ids.is_contiguous() and mask.is_contiguous()
        assert tuple(ids.shape) == (self.B, self.S), (ids.shape, self.B, self.S)
        self._keep = [ids, mask]
        check(self.lib.b4r_encode(self.h, _dl_ptr(ids), _dl_ptr(mask), int(bool(training)), int(seed), int(step),
                                  _ptr(step_counter), _stream()))

    def select(self, positions, ids=None, weights=None, mode=0, want_aux=False):
        self._keep += [positions, ids, weights]
        check(self.lib.b4r_mlm_select(self.h, _dl_ptr(positions), _dl_ptr(ids), _dl_ptr(weights), mode, int(want_aux), _stream()))

    def transform(self):
        check(self.lib.b4r_mlm_transform(self.h, _stream()))

    def loss(self, stats=None):
        check(self.lib.b4r_mlm_loss(self.h, _ptr(stats), _stream()))

    def logits(self, n_rows):
        out = torch.empty(n_rows, self.store.V, dtype=torch.float32, device=self.store.device)
        check(self.lib.b4r_mlm_logits(self.h, _ptr(out), _stream()))
        return out

    def backward(self, seed=0, step=0, step_counter=None):
        check(self.lib.b4r_backward(self.h, int(seed), int(step), _ptr(step_counter), _stream()))

    def backward_from_dt(self, dt, seed=0, step=0, step_counter=None):
        """Backward from an externally supplied gradient of the transformed rows (vocabulary-sharded projection)."""
        assert dt.dtype == torch.float32 and dt.is_contiguous() and tuple(dt.shape) == (self.Mcap, self.store.H)
        check(self.lib.b4r_backward_from_dt(self.h, _ptr(dt), int(seed), int(step), _ptr(step_counter), _stream()))

    def pooled_output(self):
        out = torch.empty(self.B, self.store.H, dtype=torch.float32, device=self.store.device)
        check(self.lib.b4r_pooled_output(self.h, _ptr(out), _stream()))
        return out

    def rank_candidates(self, cand, gt=None, want_ranking=True, want_scores=False, hist=None):
        n, Cn = cand.shape
        dev = self.store.device
        ranking = torch.empty(n, Cn, dtype=torch.int64, device=dev) if want_ranking else None
        scores = torch.empty(n, Cn, dtype=torch.float32, device=dev) if want_scores else None
        rank = torch.zeros(n, dtype=torch.int32, device=dev)
        check(self.lib.b4r_rank_candidates(self.h, _dl_ptr(cand), _dl_ptr(gt), n, Cn, _ptr(ranking), _ptr(scores), _ptr(rank),
                                           _ptr(hist), _stream()))
        return ranking, scores, rank

    def topk_full(self, k, v_begin=0, v_end=None, t_rows=None, n_rows=None, want_keys=False):
        """The k best catalogue items of every selected row (or of the external bf16 rows ``t_rows``) over the vocabulary shard
        [v_begin, v_end): (ids int64 [n, k], scores fp32 [n, k][, keys uint64-as-int64 [n, k]]), best first, lower id first among
        equal scores; no logits are materialised (b4r_topk_full)."""
        dev = self.store.device
        v_end = self.store.V if v_end is None else v_end
        n = int(n_rows if n_rows is not None else (t_rows.shape[0] if t_rows is not None else self.Mcap))
        if t_rows is not None:
            assert t_rows.dtype == torch.bfloat16 and t_rows.is_contiguous() and t_rows.shape[1] == self.store.H
        nbytes = self.lib.b4r_topk_scratch_bytes(n, int(v_begin), int(v_end), int(k))
        scratch = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
        ids = torch.empty(n, k, dtype=torch.int64, device=dev)
        scores = torch.empty(n, k, dtype=torch.float32, device=dev)
        keys = torch.empty(n, k, dtype=torch.int64, device=dev) if want_keys else None
        check(self.lib.b4r_topk_full(self.h, _ptr(t_rows), n, int(v_begin), int(v_end), int(k), _ptr(scratch), _ptr(keys), _ptr(ids),
                                     _ptr(scores), _stream()))
        self._keep_topk = scratch
        return (ids, scores, keys) if want_keys else (ids, scores)

    def rank_full(self, n_rows, v_begin=0, v_end=None):
        beat = torch.zeros(max(n_rows, 1), dtype=torch.int32, device=self.store.device)
        check(self.lib.b4r_rank_full(self.h, v_begin, self.store.V if v_end is None else v_end, _ptr(beat), _stream()))
        return beat[:n_rows]

    def rank_full_ext(self, t_rows, labels, gt_scores, counts2, v_begin, v_end, beat):
        """Items of the shard [v_begin, v_end) ranking ahead of each external row's ground truth, added into ``beat``."""
        assert t_rows.dtype == torch.bfloat16 and labels.dtype == torch.int32 and gt_scores.dtype == torch.float32
        assert counts2.dtype == torch.int32 and beat.dtype == torch.int32 and t_rows.is_contiguous()
        check(self.lib.b4r_rank_full_ext(self.h, _ptr(t_rows), _ptr(labels), _ptr(gt_scores), _ptr(counts2), t_rows.shape[0],
                                         int(v_begin), int(v_end), _ptr(beat), _stream()))

    # ---- introspection (zero-copy views over the workspace)
    def _view(self, ptr, shape, dtype):
        n = 1
        for d in shape:
            n *= d
        esz = torch.empty(0, dtype=dtype).element_size()
        off = ptr - self.ws.data_ptr()
        return self.ws[off:off + n * esz].view(dtype).view(*shape)

    def sequence_output(self, layer=-1):
        p = self.lib.b4r_sequence_output(self.h, layer)
        return self._view(p, (self.B, self.S, self.store.H), torch.bfloat16)

    def mlm_hidden(self):
        return self._view(self.lib.b4r_mlm_hidden(self.h), (self.Mcap, self.store.H), torch.bfloat16)

    def counts(self):
        return self._view(self.lib.b4r_mlm_counts(self.h), (2,), torch.int32)

    def rows(self):
        return self._view(self.lib.b4r_mlm_rows(self.h), (self.Mcap,), torch.int32)

    def step_stats(self):
        return self._view(self.lib.b4r_step_stats(self.h), (8,), torch.float32)

    def labels(self):
        return self._view(self.lib.b4r_mlm_labels(self.h), (self.Mcap,), torch.int32)

    def row_weights(self):
        return self._view(self.lib.b4r_mlm_row_weights(self.h), (self.Mcap,), torch.float32)

    def row_mult(self):
        return self._view(self.lib.b4r_mlm_row_mult(self.h), (self.Mcap,), torch.int32)

    def attn_keep_mask(self, layer):
        """[B, N, S, S] uint8 keep mask of the attention-prob dropout of the last training forward."""
        w = C.c_int()
        p = self.lib.b4r_attn_keep_bits(self.h, layer, C.byref(w))
        N, S, W = self.store.N, self.S, w.value
        words = self._view(p, (self.B, N, S, W), torch.int64)
        bits = torch.arange(64, device=words.device, dtype=torch.int64)
        m = ((words.unsqueeze(-1) >> bits) & 1).reshape(self.B, N, S, W * 64)[..., :S]
        return m.to(torch.uint8)

    def layer_tensor(self, layer, name):
        """Saved activation of one encoder layer (see b4r_layer_tensor); zero-copy view over the workspace."""
        cols, f32 = C.c_int(), C.c_int()
        p = self.lib.b4r_layer_tensor(self.h, layer, name.encode(), C.byref(cols), C.byref(f32))
        if not p:
            raise KeyError(name)
        rows = self.B * self.S * (self.store.N if name == "lse" else 1)
        return self._view(p, (rows, cols.value), torch.float32 if f32.value else torch.bfloat16)

    def launch_count(self):
        return self.lib.b4r_launch_count(self.h)

    def set_flag(self, flag, value):
        check(self.lib.b4r_session_set_flag(self.h, int(flag), int(value)))

    def profile(self, on=True):
        check(self.lib.b4r_profile_enable(self.h, int(on)))

    def profile_report(self):
        """{tag: (count, total_ms)} of every kernel launched through the session since profiling was enabled."""
        buf = C.create_string_buffer(1 << 16)
        check(self.lib.b4r_profile_report(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            tag, cnt, ms = line.rsplit(" ", 2)
            out[tag] = (int(cnt), float(ms))
        return out


def topk_merge(keys):
    """keys: int64 view of the uint64 keys [nlists, n_rows, k] (per shard / per rank) -> (ids, scores) of the k best per row."""
    assert keys.dtype == torch.int64 and keys.is_contiguous() and keys.dim() == 3
    nl, n, k = keys.shape
    ids = torch.empty(n, k, dtype=torch.int64, device=keys.device)
    scores = torch.empty(n, k, dtype=torch.float32, device=keys.device)
    check(_lib.load().b4r_topk_merge(_ptr(keys), nl, n, k, None, _ptr(ids), _ptr(scores), _stream()))
    return ids, scores


def shard_range(vocab, world, rank):
    """Catalogue rows [lo, hi) of the tied projection owned by ``rank`` (contiguous, equal ceil(V/world) slices)."""
    per = (vocab + world - 1) // world
    return min(vocab, rank * per), min(vocab, (rank + 1) * per)


class VocabShard:
    """One rank's slice [v_begin, v_end) of the tied output projection over the masked-slot rows of all ``n_ranks`` ranks
    (SURVEY 8e large catalogues; b4r_shard_* in include/b4r.h).  The collectives stay with the caller."""

    def __init__(self, store, n_ranks, rows_per_rank, v_begin, v_end):
        self.store, self.lib = store, store.lib
        self.n, self.Mcap, self.cap = n_ranks, rows_per_rank, n_ranks * rows_per_rank
        self.v_begin, self.v_end = int(v_begin), int(v_end)
        store.ensure_training_buffers()
        H, V = store.H, store.V
        nbytes = self.lib.b4r_shard_workspace_bytes(H, V, n_ranks, rows_per_rank, self.v_begin, self.v_end)
        if nbytes == 0:
            raise ValueError(self.lib.b4r_last_error().decode())
        with torch.cuda.device(store.device):
            self.ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=store.device)
            off = (-self.ws.data_ptr()) % 256
            h = C.c_void_p()
            check(self.lib.b4r_shard_create(H, V, n_ranks, rows_per_rank, self.v_begin, self.v_end,
                                            _ptr(store.seg("word_embeddings", store.shadow)), _ptr(store.seg("head/output_bias")),
                                            _ptr(store.seg("word_embeddings", store.grads)),
                                            _ptr(store.seg("head/output_bias", store.grads)),
                                            C.c_void_p(self.ws.data_ptr() + off), nbytes, C.byref(h)))
        self.h = h
        self._keep = None

    def close(self):
        if self.h:
            self.lib.b4r_shard_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pack(self, rows, labels, weights, mult, counts):
        """All-gathered session buffers: rows bf16 [n, Mcap, H]; labels/mult int32, weights fp32 [n, Mcap]; counts int32 [n, 2]."""
        H = self.store.H
        assert rows.dtype == torch.bfloat16 and rows.numel() == self.cap * H and rows.is_contiguous()
        assert labels.dtype == torch.int32 and mult.dtype == torch.int32 and weights.dtype == torch.float32
        assert labels.numel() == self.cap and mult.numel() == self.cap and weights.numel() == self.cap
        assert counts.dtype == torch.int32 and counts.numel() == 2 * self.n and counts.is_contiguous()
        self._keep = (rows, labels, weights, mult, counts)
        check(self.lib.b4r_shard_pack(self.h, _ptr(rows), _ptr(labels), _ptr(weights), _ptr(mult), _ptr(counts), _stream()))

    def partial(self, out=None):
        if out is None:
            out = torch.empty(self.cap, 6, dtype=torch.float32, device=self.store.device)
        check(self.lib.b4r_shard_ce_partial(self.h, _ptr(out), _stream()))
        return out

    def merge(self, parts, global_batch, stats=None):
        assert parts.dtype == torch.float32 and parts.is_contiguous() and parts.numel() % (self.cap * 6) == 0
        check(self.lib.b4r_shard_ce_merge(self.h, _ptr(parts), parts.numel() // (self.cap * 6), int(global_batch), _ptr(stats),
                                          _stream()))

    def backward(self, dt_out=None, zero_all=True):
        if dt_out is None:
            dt_out = torch.empty(self.n, self.Mcap, self.store.H, dtype=torch.float32, device=self.store.device)
        assert dt_out.dtype == torch.float32 and dt_out.is_contiguous() and dt_out.numel() == self.cap * self.store.H
        check(self.lib.b4r_shard_ce_backward(self.h, _ptr(dt_out), int(bool(zero_all)), _stream()))
        return dt_out

    def _view(self, ptr, n, dtype):
        off = ptr - self.ws.data_ptr()
        return self.ws[off:off + n * torch.empty(0, dtype=dtype).element_size()].view(dtype)

    def step_stats(self):
        return self._view(self.lib.b4r_shard_step_stats(self.h), 8, torch.float32)

    def lse(self):
        return self._view(self.lib.b4r_shard_lse(self.h), self.cap, torch.float32)

    def counts(self):
        return self._view(self.lib.b4r_shard_counts(self.h), 2, torch.int32)


def dropout_keep_mask(rows, cols, rate, seed, site, layer, step, device):
    out = torch.empty(rows, cols, dtype=torch.uint8, device=device)
    check(_lib.load().b4r_dropout_keep_mask(_ptr(out), rows, cols, float(rate), int(seed), site, layer, int(step), _stream()))
    return out
