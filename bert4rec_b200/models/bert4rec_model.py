"""BERT4RecModel on the B200 path: encoder + tied-embedding masked-item head, custom train/test step, ranking.

Mirrors the reference's ``BERT4RecModel`` (bert4rec/models/bert4rec_model.py:27-240): same constructor, same batch
dict keys (:15-22), same output dict keys (:117-149), ``compile`` / ``train_step`` / ``test_step`` / ``fit`` /
``rank_items``.  Differences that follow from the B200 design: the [B,P,V] logits are only materialised
when ``__call__`` is asked for them; ``train_step`` never builds them (fused projection + softmax-CE), keeps the
running metrics on the device and returns a lazy mapping (reading a value synchronises); ``rank_items`` scores only
the candidate ids (fused gather-dot) instead of the whole vocabulary.
"""
import collections.abc
import os
import copy
import ctypes as C
import pathlib
import time

import numpy as np
import torch

from bert4rec_b200 import _lib
from bert4rec_b200.models.components import networks

_ENCODER_CONFIG_FILE_NAME = "encoder_config.json"
_META_CONFIG_FILE_NAME = "meta_config.json"
_TOKENIZER_VOCAB_FILE_NAME = "vocab.txt"
_MODEL_WEIGHTS_FILES_PREFIX = "model_weights"

# PAD, MASK, UNK -- BERT4RecDataloader(0, 0)._SPECIAL_TOKEN_IDS in the reference (bert4rec_dataloader.py:35-43)
SPECIAL_TOKEN_IDS = [0, 1, 2]
BATCH_KEYS = ("labels", "input_word_ids", "input_mask", "masked_lm_ids", "masked_lm_positions", "masked_lm_weights")
_STAGED_KEYS = ("input_word_ids", "input_mask", "masked_lm_positions", "masked_lm_ids", "masked_lm_weights")


class History:
    """keras.callbacks.History stand-in: ``history`` maps metric name -> list of per-epoch values."""

    def __init__(self):
        self.history = {}
        self.epoch = []
        self.params = {}

    def _append(self, epoch, logs):
        self.epoch.append(epoch)
        for k, v in logs.items():
            self.history.setdefault(k, []).append(v)


class _PinnedRing:
    """A few pinned host buffers used round-robin as the source of asynchronous H2D copies.  A buffer is only rewritten after
    the copy that last read it has completed (CUDA event recorded behind that copy): the host may run many steps ahead of
    the stream (lazy metrics, graph replays), and a single staging buffer would be overwritten while its copy is pending."""

    def __init__(self, make, depth=3):
        self._make, self._depth, self._slots, self._next = make, depth, [], 0

    def acquire(self):
        """-> (buffer(s), release): fill the buffers, enqueue the copy, then call release() on the copy's stream."""
        if len(self._slots) < self._depth:
            slot = [self._make(), None]
            self._slots.append(slot)
        else:
            slot = self._slots[self._next]
            self._next = (self._next + 1) % self._depth
            if slot[1] is not None:
                slot[1].synchronize()
        def release(slot=slot):
            if slot[1] is None:
                slot[1] = torch.cuda.Event()
            slot[1].record()
        return slot[0], release


class StepMetrics(collections.abc.Mapping):
    """Lazy view of the device-side running metrics (Keras ``{m.name: m.result()}``, bert4rec_model.py:173).
    Values are read from the device (one small D2H copy, synchronising) on first access."""

    _pinned = {}   # one pinned host landing buffer per device: the D2H read of the 16 floats without a pageable staging copy

    def __init__(self, stats, names):
        self._stats, self._names, self._cache = stats, names, None

    def _load(self):
        if self._cache is None:
            dev = self._stats.device
            host = StepMetrics._pinned.get(dev)
            if host is None:
                host = StepMetrics._pinned[dev] = torch.empty(16, dtype=torch.float32).pin_memory()
            host.copy_(self._stats.detach(), non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            s = host.tolist()
            vals = {"loss": s[5] / s[6] if s[6] else 0.0,
                    "sparse_categorical_accuracy": s[3] / s[4] if s[4] else 0.0,
                    "masked_accuracy": s[7] / s[8] if s[8] else 0.0}
            self._cache = {k: vals[k] for k in self._names}
        return self._cache

    def __getitem__(self, k):
        return self._load()[k]

    def __iter__(self):
        return iter(self._names)

    def __len__(self):
        return len(self._names)

    def __repr__(self):
        return repr(self._load())


class BERT4RecModel:
    def __init__(self,
                 encoder: networks.Bert4RecEncoder,
                 customized_masked_lm=None,
                 mlm_activation="gelu",
                 mlm_initializer="glorot_uniform",
                 name: str = "bert4rec",
                 special_token_ids: list = SPECIAL_TOKEN_IDS,
                 **kwargs):
        if customized_masked_lm is not None:
            raise NotImplementedError("customized_masked_lm: only the tied-embedding MaskedLM head is implemented")
        if mlm_activation != "gelu":
            raise NotImplementedError("mlm_activation: only exact-erf 'gelu'")
        self.name = name
        self._config = {"encoder": encoder, "customized_masked_lm": customized_masked_lm,
                        "mlm_activation": mlm_activation, "mlm_initializer": mlm_initializer, "name": name}
        self.encoder = encoder
        self.store = encoder.store
        self.vocab_size = encoder.get_config()["vocab_size"]
        self.prediction_mask = None  # the reference builds a special-token mask and then disables it (:101-102)
        self.inputs = dict(encoder.inputs, masked_lm_positions=None)
        self.optimizer = None
        self.loss = None
        self.compiled_metrics = None
        self._metric_names = ("loss",)
        self._want_sca = False
        self._stats = {}
        self._staging = {}
        self._pinned_plans = {}
        self._packed = {}
        self._seed = 0x5EEDB4A7
        self._host_step = 0
        self.stop_training = False
        self.distributed = False
        self.vocab_sharded = False
        self._shards = {}
        self.use_cuda_graph = True
        self._graphs = {}
        self._store_gen = self.store.generation
        # data-parallel step as ONE graph with the NCCL all-reduce captured inside: opt-in only (B4R_DP_SINGLE_GRAPH=1).  Measured
        # 308 vs 316 us at N=2, but a live graph holding NCCL work hung the process at teardown on this stack.
        self._dp_single_graph = None if os.environ.get("B4R_DP_SINGLE_GRAPH") else False

    @property
    def identifier(self):
        return "bert4rec"

    @property
    def device(self):
        return self.store.device

    # ------------------------------------------------------------------ input staging
    def _stage(self, inputs, keys, persistent=False):
        """Moves the int64 input tensors named by ``keys`` to the device.  Host tensors are packed into ONE pinned
        staging buffer and copied with a single async H2D copy.  Device tensors are used in place unless
        ``persistent`` (CUDA-graph replays need fixed addresses): then they are gathered into the same persistent
        device buffer with one D2D copy kernel."""
        vals = [v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v)) for v in (inputs[k] for k in keys)]
        all_cuda = all(v.is_cuda for v in vals)
        if all_cuda and not persistent:
            return {k: (v if v.dtype == torch.int64 and v.is_contiguous() else v.to(torch.int64).contiguous())
                    for k, v in zip(keys, vals)}
        shapes = [tuple(v.shape) for v in vals]
        sizes = [v.numel() for v in vals]
        total = sum(sizes)
        key = (keys, tuple(shapes))
        st = self._staging.get(key)
        if st is None:
            host = _PinnedRing(lambda: torch.empty(total, dtype=torch.int64).pin_memory())
            dev = torch.empty(total, dtype=torch.int64, device=self.device)
            views, off = {}, 0
            for k, shp, n in zip(keys, shapes, sizes):   # the per-key device views are fixed: built once
                views[k] = dev[off:off + n].view(shp)
                off += n
            st = self._staging[key] = (host, dev, views)
        host_ring, dev, views = st
        if keys == _STAGED_KEYS and not any(v.is_cuda for v in vals) and os.environ.get("B4R_NO_PACKED_H2D") is None:
            # host batch of a train / test step: compact transfer (ids / positions / labels as int32, the 0/1 arrays as bytes: 2.9x
            # fewer PCIe bytes than the int64 tensors), ONE H2D copy, widened on the device into the persistent int64 views
            n_tok, n_pred = sizes[0], sizes[2]
            pk = self._packed.get(key)
            if pk is None:
                nbytes = self.store.lib.b4r_packed_inputs_bytes(n_tok, n_pred)
                n32 = n_tok + 2 * n_pred

                def make():
                    hp = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                    h32 = hp[: 4 * n32].view(torch.int32)
                    return (hp, (h32[:n_tok], h32[n_tok:n_tok + n_pred], h32[n_tok + n_pred:]),
                            (hp[4 * n32: 4 * n32 + n_tok], hp[4 * n32 + n_tok:]))
                pk = self._packed[key] = (_PinnedRing(make), torch.empty(nbytes, dtype=torch.uint8, device=self.device))
            ring, dp = pk
            (hp, (h_ids, h_pos, h_mlm), (h_mask, h_w)), release = ring.acquire()   # waits for the copy that last read this buffer
            h_ids.copy_(vals[0].reshape(-1)); h_mask.copy_(vals[1].reshape(-1))
            h_pos.copy_(vals[2].reshape(-1)); h_mlm.copy_(vals[3].reshape(-1)); h_w.copy_(vals[4].reshape(-1))
            dp.copy_(hp, non_blocking=True)
            release()
            from bert4rec_b200 import _lib
            import ctypes as C
            _lib.check(self.store.lib.b4r_unpack_inputs(
                C.c_void_p(dp.data_ptr()), n_tok, n_pred, *[C.c_void_p(views[k].data_ptr()) for k in keys],
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
            return views
        if not all_cuda and all(v.dtype == torch.int64 and v.is_contiguous() and not v.is_cuda for v in vals):
            # pinned host tensors: straight DMA of every tensor into its device view (no packing pass over the batch)
            ptrs = tuple(v.data_ptr() for v in vals)
            plan = self._pinned_plans.get((key, ptrs))
            if plan is None and all(v.is_pinned() for v in vals):
                import ctypes as C
                n = len(vals)
                plan = ((C.c_void_p * n)(*[views[k].data_ptr() for k in keys]), (C.c_void_p * n)(*ptrs),
                        (C.c_size_t * n)(*[8 * s_ for s_ in sizes]), n, vals)   # (vals kept alive: the plan is keyed by their addresses)
                if len(self._pinned_plans) > 64:
                    self._pinned_plans.clear()
                self._pinned_plans[(key, ptrs)] = plan
            if plan is not None:
                from bert4rec_b200 import _lib
                _lib.check(self.store.lib.b4r_h2d_copy_many(plan[0], plan[1], plan[2], plan[3],
                                                            torch.cuda.current_stream(self.device).cuda_stream))
                return views
        if all_cuda:
            torch.cat([v.reshape(-1).to(torch.int64) for v in vals], out=dev)
        else:
            host, release = host_ring.acquire()
            if all(v.dtype == torch.int64 for v in vals):
                torch.cat([v.reshape(-1) for v in vals], out=host)   # one packed host copy into the pinned buffer
            else:
                off = 0
                for v, n in zip(vals, sizes):
                    host[off:off + n].copy_(v.reshape(-1))           # dtype-converting host copy
                    off += n
            dev.copy_(host, non_blocking=True)   # ONE H2D copy (five separate small copies measured slower)
            release()
        return views

    @staticmethod
    def _bytes_of(keys, inputs):
        """Bytes the staging path copies host -> device for a host batch (packed form for the five step inputs)."""
        if tuple(keys) == _STAGED_KEYS and os.environ.get("B4R_NO_PACKED_H2D") is None:
            n_tok = int(np.prod(tuple(inputs["input_word_ids"].shape)))
            n_pred = int(np.prod(tuple(inputs["masked_lm_positions"].shape)))
            return 4 * (n_tok + 2 * n_pred) + n_tok + n_pred
        return sum(int(np.prod(tuple(inputs[k].shape))) * 8 for k in keys)

    # ------------------------------------------------------------------ forward (API parity)
    def __call__(self, inputs, training=None, mask=None):
        return self.call(inputs, training=training, mask=mask)

    def call(self, inputs, training=None, mask=None):
        if isinstance(inputs, (list, tuple)):
            names = ["input_word_ids", "input_mask", "masked_lm_positions"]
            inputs = dict(zip(names, inputs))
        if not isinstance(inputs, dict):
            raise ValueError("inputs should be a dict with keys input_word_ids, input_mask[, masked_lm_positions]")
        keys = ("input_word_ids", "input_mask") + (("masked_lm_positions",) if "masked_lm_positions" in inputs else ())
        d = self._stage(inputs, keys)
        B, S = d["input_word_ids"].shape
        P = d["masked_lm_positions"].shape[1] if "masked_lm_positions" in d else 0
        sess = self.store.session(B, S, P)
        sess.encode(d["input_word_ids"], d["input_mask"], training=bool(training),
                    seed=self._seed, step=self._host_step)
        L = self.store.L
        outs = [sess.sequence_output(l).float() for l in range(L)]
        outputs = dict(sequence_output=outs[-1], pooled_output=sess.pooled_output(), encoder_outputs=outs)
        if P:
            sess.select(d["masked_lm_positions"], None, None, mode=2)
            sess.transform()
            outputs["mlm_logits"] = sess.logits(B * P).view(B, P, self.vocab_size)
        return outputs

    # ------------------------------------------------------------------ compile / steps
    def compile(self, optimizer=None, loss=None, metrics=None, **kwargs):
        from bert4rec_b200.trainers import optimizers as _opt, trainer_utils
        if optimizer is None or isinstance(optimizer, str):
            optimizer = _opt.get(optimizer or "adamw")
        if not isinstance(optimizer, _opt.AdamWeightDecay):
            raise NotImplementedError("only the AdamWeightDecay optimizer is implemented on the B200 path")
        if loss is None:
            loss = trainer_utils.MaskedSparseCategoricalCrossentropy()
        if not isinstance(loss, trainer_utils.MaskedSparseCategoricalCrossentropy) or loss.pad_token != 0:
            raise NotImplementedError("only MaskedSparseCategoricalCrossentropy(pad_token=0) is fused into the step")
        names = ["loss"]
        for m in (metrics or []):
            n = getattr(m, "name", None) or getattr(m, "__name__", None) or str(m)
            if n in ("sparse_categorical_accuracy", "masked_accuracy"):
                names.append(n)
            else:
                raise NotImplementedError(f"metric {n!r} is not computed by the fused step")
        groups = {k: (2 if k.startswith("pooler_transform") else (0 if k.endswith(("kernel", "embeddings")) else 1))
                  for k in self.store.tf_views()}
        optimizer.check_layout(groups)
        self.optimizer, self.loss, self.compiled_metrics = optimizer, loss, list(metrics or [])
        self._metric_names = tuple(names)
        self._want_sca = "sparse_categorical_accuracy" in names
        self.store.ensure_training_buffers()
        self._hp = optimizer.hparams_struct()
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            self.distributed = True
            # the global valid-slot count travels in the SAME all-reduce as the gradients: element n_trainable of the
            # flat gradient buffer (first element of the frozen pooler segment, which has no gradient)
            nt = self.store.n_trainable
            self._count = self.store.grads[nt:nt + 1]
            # large catalogues: the tied output projection is sharded by vocabulary rows over the ranks
            self.vocab_sharded = bool(kwargs.get("vocab_sharded", os.environ.get("B4R_VOCAB_SHARDED")))

    @property
    def metrics_names(self):
        return list(self._metric_names)

    def _stats_buf(self, phase):
        b = self._stats.get(phase)
        if b is None:
            b = self._stats[phase] = torch.zeros(16, dtype=torch.float32, device=self.device)
        return b

    def reset_metrics(self, phase=None):
        for k, b in self._stats.items():
            if phase is None or k == phase:
                b.zero_()

    def _check_graphs(self):
        """Captured graphs reference session workspaces: drop them when the store re-created its sessions (compile() after an
        evaluation allocates the gradient buffers and rebuilds every session)."""
        if self._store_gen != self.store.generation:
            self._graphs.clear()
            self._graph_inputs = {}
            self._store_gen = self.store.generation

    def train_step(self, inputs):
        """fwd(training) -> fused CE -> backward -> [NCCL allreduce] -> clip + AdamW (bert4rec_model.py:151-173)."""
        if self.optimizer is None:
            raise RuntimeError("compile() the model (or trainer.initialize_model()) before train_step")
        self._check_graphs()
        # Device-resident int64 inputs are consumed IN PLACE: the captured graph is keyed by their addresses (a training
        # loop cycling over a cached, device-resident dataset replays with no copy at all).  Host inputs, and device
        # inputs once more than 8 address sets have been seen, go through the persistent staging buffer.
        vals = [inputs[k] for k in _STAGED_KEYS]
        in_place = self.use_cuda_graph and all(torch.is_tensor(v) and v.is_cuda and v.dtype == torch.int64 and v.is_contiguous()
                                               for v in vals)
        gkey_ptrs = tuple(v.data_ptr() for v in vals) if in_place else ()
        if in_place and (vals[0].shape[0], vals[0].shape[1], vals[2].shape[1]) + gkey_ptrs not in self._graphs \
                and sum(1 for k in self._graphs if len(k) > 3 and k[0] != "rank") >= 8:
            in_place, gkey_ptrs = False, ()
        if in_place:
            d = dict(zip(_STAGED_KEYS, vals))
            self._graph_inputs = getattr(self, "_graph_inputs", {})
        else:
            d = self._stage(inputs, _STAGED_KEYS, persistent=self.use_cuda_graph)
        B, S = d["input_word_ids"].shape
        P = d["masked_lm_positions"].shape[1]
        sess = self.store.session(B, S, P)
        stats = self._stats_buf("train")
        if self.distributed and self.vocab_sharded:
            self._fwd_bwd_sharded(sess, d, stats)   # eager: five collectives inside the step
            self._all_reduce(sess)
            self._count.copy_(self._shard_of(sess).step_stats()[1:2])   # global valid slots (same on every rank)
            self._update(self._count)
        elif not self.use_cuda_graph:
            self._fwd_bwd(sess, d, stats)
            self._reduce_and_update(sess)
        else:
            gkey = (B, S, P) + gkey_ptrs
            if in_place:
                self._graph_inputs[gkey] = vals   # keep the captured tensors alive as long as the graph
            g = self._graphs.get(gkey)
            if g is None:
                # first step of this shape runs eagerly (lazy one-time initialisation), then the same launch
                # sequence is captured once and replayed for every later step.  Single GPU: one graph for the
                # whole step.  Data-parallel: forward+backward graph, eager NCCL all-reduce, optimizer graph.
                self._fwd_bwd(sess, d, stats)
                self._reduce_and_update(sess)
                torch.cuda.synchronize(self.device)
                g1, g2 = None, None
                if self.distributed and (self._dp_single_graph is not False or self.store.p2p is not None):
                    # data-parallel: try ONE graph for the whole step with the NCCL all-reduce captured inside it
                    # (NCCL collectives are capturable); on any failure fall back to graph / eager all-reduce / graph
                    try:
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            self._fwd_bwd(sess, d, stats)
                            self._reduce_and_update(sess)
                        g1 = g
                        self._dp_single_graph = True
                    except Exception as e:   # noqa: BLE001
                        self._dp_single_graph = False
                        torch.cuda.synchronize(self.device)
                        import warnings
                        warnings.warn(f"NCCL all-reduce could not be captured into the step graph ({e}); using two graphs")
                if g1 is None:
                    g1 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g1):
                        self._fwd_bwd(sess, d, stats)
                        if not self.distributed:
                            self._reduce_and_update(sess)
                    if self.distributed:
                        g2 = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g2):
                            self._update(self._count)
                self._graphs[gkey] = (g1, g2)
            else:
                g1, g2 = g
                g1.replay()
                if g2 is not None:
                    self._all_reduce(sess)
                    g2.replay()
        return StepMetrics(stats, self._metric_names)

    def _all_reduce(self, sess):
        # batch data-parallel: SUM gradients of the SUM loss and the valid-slot counts over ranks (NCCL), so that
        # the normaliser is the GLOBAL number of valid masked slots (trainer_utils.py:22)
        # the library's own two-shot kernel over NVLink peer mappings when the gradient buffer lives in symmetric memory
        # (engine.ParamStore._alloc_grads), else one NCCL all-reduce
        self.store.all_reduce_grads(self.store.n_trainable + 1)

    def _update(self, count):
        self.store.adamw_step(self._hp, count=count, grad_scale=1.0)

    def _reduce_and_update(self, sess):
        if self.distributed:
            self._all_reduce(sess)
            self._update(self._count)
        else:
            self._update(sess.step_stats()[1:2])

    def _fwd_bwd(self, sess, d, stats):
        """The launch sequence of forward + backward (all on torch's current stream; CUDA-graph capturable:
        the dropout counter and the learning-rate schedule read the device-side iteration counter)."""
        ctr = self.store.step_counter
        # the slot compaction depends on the inputs only: it is enqueued first, as a branch beside the encoder forward
        sess.set_flag(4, 1)
        sess.select(d["masked_lm_positions"], d["masked_lm_ids"], d["masked_lm_weights"], mode=0, want_aux=self._want_sca)
        sess.set_flag(4, 0)
        sess.encode(d["input_word_ids"], d["input_mask"], training=True, seed=self._seed, step=0, step_counter=ctr)
        sess.transform()
        sess.loss(stats)
        sess.backward(seed=self._seed, step=0, step_counter=ctr)
        if self.distributed:
            self._count.copy_(sess.step_stats()[1:2])   # inside the captured graph: rides along with the gradients

    def _shard_of(self, sess):
        import torch.distributed as dist
        from bert4rec_b200.engine import VocabShard, shard_range
        sh = self._shards.get(sess.Mcap)
        if sh is None:
            world, me = dist.get_world_size(), dist.get_rank()
            lo, hi = shard_range(self.store.V, world, me)
            if lo >= hi:
                raise ValueError(f"vocabulary of {self.store.V} rows cannot be sharded over {world} ranks")
            sh = self._shards[sess.Mcap] = VocabShard(self.store, world, sess.Mcap, lo, hi)
            dev, H, n = self.device, self.store.H, world
            sh.buf = {"rows": torch.empty(n, sess.Mcap, H, dtype=torch.bfloat16, device=dev),
                      "meta_in": torch.empty(3 * sess.Mcap + 2, dtype=torch.int32, device=dev),
                      "meta": torch.empty(n, 3 * sess.Mcap + 2, dtype=torch.int32, device=dev),
                      "part": torch.empty(n * sess.Mcap, 6, dtype=torch.float32, device=dev),
                      "parts": torch.empty(n, n * sess.Mcap, 6, dtype=torch.float32, device=dev),
                      "dt_all": torch.empty(n, sess.Mcap, H, dtype=torch.float32, device=dev),
                      "dt": torch.empty(sess.Mcap, H, dtype=torch.float32, device=dev)}
        return sh

    def _fwd_bwd_sharded(self, sess, d, stats):
        """Forward + backward with the tied output projection SHARDED by vocabulary rows: all-gather the
        transformed rows (+ labels / weights / counts) of every rank, score them against the local catalogue slice, exchange
        and merge the per-row softmax partials, run the projection backward on the slice (its table-gradient slice is complete)
        and reduce-scatter the partial row gradients back to the ranks that own the rows."""
        import torch.distributed as dist
        ctr = self.store.step_counter
        sess.select(d["masked_lm_positions"], d["masked_lm_ids"], d["masked_lm_weights"], mode=0, want_aux=self._want_sca)
        sess.encode(d["input_word_ids"], d["input_mask"], training=True, seed=self._seed, step=0, step_counter=ctr)
        sess.transform()
        sh = self._shard_of(sess)
        b, M = sh.buf, sess.Mcap
        torch.cat([sess.labels(), sess.row_weights().view(torch.int32), sess.row_mult(), sess.counts()], out=b["meta_in"])
        dist.all_gather_into_tensor(b["rows"], sess.mlm_hidden())
        dist.all_gather_into_tensor(b["meta"], b["meta_in"])
        meta = b["meta"]
        # the pack kernel wants each field contiguous over the ranks
        labels = meta[:, :M].contiguous()
        weights = meta[:, M:2 * M].contiguous().view(torch.float32)
        mult = meta[:, 2 * M:3 * M].contiguous()
        counts = meta[:, 3 * M:3 * M + 2].contiguous()
        sh.pack(b["rows"], labels, weights, mult, counts)
        sh.partial(b["part"])
        dist.all_gather_into_tensor(b["parts"], b["part"])
        sh.merge(b["parts"], sess.B * sh.n, stats)
        sh.backward(b["dt_all"], zero_all=True)
        dist.reduce_scatter_tensor(b["dt"], b["dt_all"])
        sess.backward_from_dt(b["dt"], seed=self._seed, step=0, step_counter=ctr)

    def test_step(self, inputs):
        """fwd(inference) -> fused CE + accuracies, no update (bert4rec_model.py:175-192)."""
        d = self._stage(inputs, _STAGED_KEYS)
        B, S = d["input_word_ids"].shape
        P = d["masked_lm_positions"].shape[1]
        sess = self.store.session(B, S, P)
        stats = self._stats_buf("test")
        sess.encode(d["input_word_ids"], d["input_mask"], training=False)
        sess.select(d["masked_lm_positions"], d["masked_lm_ids"], d["masked_lm_weights"], mode=0, want_aux=self._want_sca)
        sess.transform()
        sess.loss(stats)
        return StepMetrics(stats, self._metric_names)

    def evaluate(self, x, steps=None, return_dict=True):
        self.reset_metrics("test")
        res = None
        for i, batch in enumerate(x):
            if steps is not None and i >= steps:
                break
            res = self.test_step(batch)
        out = dict(res) if res is not None else {}
        return out if return_dict else [out[k] for k in self._metric_names]

    def fit(self, x=None, validation_data=None, epochs=1, callbacks=None, steps_per_epoch=None,
            validation_steps=None, verbose=0, initial_epoch=0, **kwargs):
        """Minimal keras ``Model.fit``: per epoch reset metrics, run train_step over ``x``, evaluate
        ``validation_data`` (metrics prefixed ``val_``), call the callbacks, return a History."""
        history = History()
        callbacks = list(callbacks or [])
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        self.stop_training = False
        for cb in callbacks:
            getattr(cb, "on_train_begin", lambda logs=None: None)()
        for epoch in range(initial_epoch, epochs):
            for cb in callbacks:
                getattr(cb, "on_epoch_begin", lambda e, logs=None: None)(epoch)
            self.reset_metrics("train")
            t0 = time.time()
            last, n = None, 0
            for i, batch in enumerate(x):
                if steps_per_epoch is not None and i >= steps_per_epoch:
                    break
                last = self.train_step(batch)
                n += 1
            logs = dict(last) if last is not None else {}     # (reads the metrics: the device has finished the epoch)
            err = self.store.p2p_error()
            if err:
                raise RuntimeError(f"data-parallel all-reduce: a rank did not arrive within the bounded wait (code {err}); "
                                   "the gradients of this epoch are not trustworthy")
            if validation_data is not None:
                val = self.evaluate(validation_data, steps=validation_steps)
                logs.update({f"val_{k}": v for k, v in val.items()})
            history._append(epoch, logs)
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs} - {n} steps - {time.time() - t0:.2f}s - "
                      + " - ".join(f"{k}: {v:.4f}" for k, v in logs.items()), flush=True)
            for cb in callbacks:
                getattr(cb, "on_epoch_end", lambda e, logs=None: None)(epoch, logs)
            if self.stop_training:
                break
        for cb in callbacks:
            getattr(cb, "on_train_end", lambda logs=None: None)()
        self.history = history
        return history

    # ------------------------------------------------------------------ ranking
    def _encode_for_ranking(self, encoder_input):
        keys = ("input_word_ids", "input_mask", "masked_lm_positions") + \
               (("masked_lm_weights",) if "masked_lm_weights" in encoder_input else ())
        d = self._stage(encoder_input, keys)
        B, S = d["input_word_ids"].shape
        P = d["masked_lm_positions"].shape[1]
        sess = self.store.session(B, S, P)
        sess.encode(d["input_word_ids"], d["input_mask"], training=False)
        if "masked_lm_weights" in d:
            sess.select(d["masked_lm_positions"], None, d["masked_lm_weights"], mode=1)
        else:
            sess.select(d["masked_lm_positions"], None, None, mode=2)
        sess.transform()
        return sess, d

    def _stage_aux(self, name, t):
        """int64 auxiliary input (candidate lists, ground truth): device tensors in place, host tensors through a
        persistent pinned + device buffer pair (stable addresses for graph replays)."""
        t = torch.as_tensor(t)
        if t.is_cuda:
            return t if (t.dtype == torch.int64 and t.is_contiguous()) else t.to(torch.int64).contiguous()
        key = (name, tuple(t.shape))
        st = self._staging.get(key)
        if st is None:
            shape = tuple(t.shape)
            st = self._staging[key] = (_PinnedRing(lambda: torch.empty(shape, dtype=torch.int64).pin_memory()),
                                       torch.empty(shape, dtype=torch.int64, device=self.device))
        ring, dev = st
        host, release = ring.acquire()
        host.copy_(t)
        dev.copy_(host, non_blocking=True)
        release()
        return dev

    def rank_candidates(self, encoder_input, candidates, ground_truth=None, want_ranking=False, hist=None):
        """Fast path of ``rank_items`` for rectangular candidate lists: ``candidates`` int64 [n_slots, C] (row order =
        row-major order of the slots with weight 1), ``ground_truth`` int64 [n_slots].  Returns (ranking or None,
        ranks int32 [n_slots]) as device tensors.  The launch sequence (encoder forward, slot selection, MLM
        transform, fused gather-dot + rank) is captured per set of input buffers and replayed; the returned tensors
        are the graph's output buffers, valid until the next call with the same buffers."""
        self._check_graphs()
        keys = ("input_word_ids", "input_mask", "masked_lm_positions") + \
               (("masked_lm_weights",) if "masked_lm_weights" in encoder_input else ())
        d = self._stage(encoder_input, keys)
        cand = self._stage_aux("cand", candidates)
        gt = self._stage_aux("gt", ground_truth) if ground_truth is not None else None
        B, S = d["input_word_ids"].shape
        P = d["masked_lm_positions"].shape[1]
        sess = self.store.session(B, S, P)

        def run():
            sess.set_flag(4, 1)   # the slot selection runs as a branch beside the encoder forward
            if "masked_lm_weights" in d:
                sess.select(d["masked_lm_positions"], None, d["masked_lm_weights"], mode=1)
            else:
                sess.select(d["masked_lm_positions"], None, None, mode=2)
            sess.set_flag(4, 0)
            sess.encode(d["input_word_ids"], d["input_mask"], training=False)
            sess.transform()
            ranking, _, rank = sess.rank_candidates(cand, gt, want_ranking=want_ranking, hist=hist)
            return ranking, rank

        if not self.use_cuda_graph:
            return run()
        gkey = ("rank", B, S, P, tuple(cand.shape), bool(want_ranking), hist.data_ptr() if hist is not None else 0,
                cand.data_ptr(), gt.data_ptr() if gt is not None else 0) + tuple(d[k].data_ptr() for k in keys)
        g = self._graphs.get(gkey)
        if g is None:
            if sum(1 for k in self._graphs if k and k[0] == "rank") >= 8:
                return run()          # too many distinct buffer sets: stay eager
            eager_out = run()         # this call's result (and its histogram update) comes from the eager pass
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = run()
            self._graphs[gkey] = (graph, out, [d, cand, gt, hist])
            return eager_out
        return self._replay_rank(gkey)

    def _replay_rank(self, gkey):
        graph, out, _ = self._graphs[gkey]
        graph.replay()
        return out

    def full_catalogue_ranks(self, encoder_input, ground_truth, group=None):
        """1-based rank of every slot's ground truth over the WHOLE catalogue (``rank_items(items=None)`` + the evaluator's
        rank lookup, bert4rec_model.py:235-240, bert4rec_evaluator.py:112-117), without materialising any logits.
        With an initialised process group the vocabulary is SHARDED over the ranks (large catalogues): the hidden
        rows, labels and ground-truth scores of all ranks are all-gathered, every rank counts the items of its shard that
        rank ahead (``b4r_rank_full_ext``), one all-reduce(SUM) of the counts gives the exact ranks."""
        import torch.distributed as dist
        sess, _ = self._encode_for_ranking(encoder_input)
        gt = self._stage_aux("gt_full", ground_truth)
        n = int(gt.shape[0])
        cap = sess.Mcap
        # ground-truth scores with the gather-dot kernel (one candidate per slot)
        _, score, _ = sess.rank_candidates(gt.view(-1, 1).contiguous(), None, want_ranking=False, want_scores=True)
        dev = self.device
        t = torch.zeros(cap, self.store.H, dtype=torch.bfloat16, device=dev)
        t[:n] = sess.mlm_hidden()[:n]
        lab = torch.zeros(cap, dtype=torch.int32, device=dev); lab[:n] = gt.to(torch.int32)
        sc = torch.zeros(cap, dtype=torch.float32, device=dev); sc[:n] = score.view(-1)
        cnt = torch.tensor([n, n], dtype=torch.int32, device=dev)
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            beat = torch.zeros(cap, dtype=torch.int32, device=dev)
            sess.rank_full_ext(t, lab, sc, cnt, 0, self.store.V, beat)
            return beat[:n] + 1
        me = dist.get_rank(group)
        ts = [torch.empty_like(t) for _ in range(world)]
        labs = [torch.empty_like(lab) for _ in range(world)]
        scs = [torch.empty_like(sc) for _ in range(world)]
        cnts = [torch.empty_like(cnt) for _ in range(world)]
        dist.all_gather(ts, t, group=group); dist.all_gather(labs, lab, group=group)
        dist.all_gather(scs, sc, group=group); dist.all_gather(cnts, cnt, group=group)
        V = self.store.V
        per = (V + world - 1) // world
        v_lo, v_hi = me * per, min(V, (me + 1) * per)
        beat = torch.zeros(world, cap, dtype=torch.int32, device=dev)
        if v_lo < v_hi:
            for q in range(world):
                sess.rank_full_ext(ts[q], labs[q], scs[q], cnts[q], v_lo, v_hi, beat[q])
        dist.all_reduce(beat, group=group)
        return beat[me, :n] + 1

    def top_k_items(self, encoder_input, k, group=None, exclude=None):
        """The ``k`` most probable catalogue items of every prediction slot: ``rank_items(items=None)[...][:k]`` of the reference
        (bert4rec_model.py:235-236: argsort over the whole vocabulary, descending, lower id first among equal logits) without
        materialising logits.  Returns (ids int64 [n_slots, k], scores fp32 [n_slots, k]) on the device, best first.
        ``exclude``: optional list (per slot) of item ids never to return (apps.Recommender masks the user's history): the kernel is
        asked for k + max(len) items and the excluded ones are dropped on the host side of the ranking.
        With an initialised process group the catalogue is SHARDED over the ranks: the hidden rows of all ranks are
        all-gathered, every rank selects the k best of its vocabulary slice for every row (``b4r_topk_full``), the per-rank lists
        are all-gathered and merged (``b4r_topk_merge``: a key comparison, exact, lowest id first)."""
        import torch.distributed as dist
        from bert4rec_b200.engine import shard_range, topk_merge
        sess, _ = self._encode_for_ranking(encoder_input)
        n = int(sess.counts()[0])
        extra = max((len(set(e)) for e in exclude), default=0) if exclude is not None else 0
        kk = min(k + extra, self.store.V)
        if kk > 128:
            raise ValueError(f"top_k_items: k + excluded items = {kk} exceeds the kernel's limit of 128")
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            ids, scores = sess.topk_full(kk, n_rows=n)
        else:
            me = dist.get_rank(group)
            H = self.store.H
            nmax = torch.tensor([n], dtype=torch.int64, device=self.device)
            dist.all_reduce(nmax, op=dist.ReduceOp.MAX, group=group)
            cap = int(nmax.item())                       # live rows of the fullest rank: only those travel and are scored
            mine = torch.zeros(cap, H, dtype=torch.bfloat16, device=self.device)
            mine[:n] = sess.mlm_hidden()[:n]
            rows = torch.empty(world, cap, H, dtype=torch.bfloat16, device=self.device)
            dist.all_gather_into_tensor(rows, mine, group=group)
            lo, hi = shard_range(self.store.V, world, me)
            if lo < hi:
                _, _, keys = sess.topk_full(kk, lo, hi, t_rows=rows.view(world * cap, H), want_keys=True)
            else:
                keys = torch.zeros(world * cap, kk, dtype=torch.int64, device=self.device)
            allk = torch.empty(world, world * cap, kk, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allk, keys, group=group)
            ids, scores = topk_merge(allk)
            ids, scores = ids[me * cap: me * cap + n], scores[me * cap: me * cap + n]
        if exclude is None:
            return ids[:, :k], scores[:, :k]
        ids_h, sc_h = ids.cpu(), scores.cpu()
        out_i = torch.full((n, k), -1, dtype=torch.int64)
        out_s = torch.full((n, k), -float("inf"))
        for r in range(n):
            ban = set(int(x) for x in exclude[r])
            keep = [j for j in range(kk) if int(ids_h[r, j]) not in ban and int(ids_h[r, j]) >= 0][:k]
            out_i[r, :len(keep)] = ids_h[r, keep]
            out_s[r, :len(keep)] = sc_h[r, keep]
        return out_i.to(self.device), out_s.to(self.device)

    def rank_items(self, encoder_input: dict, items: list = None):
        """Reference semantics (bert4rec_model.py:203-240): one ranking per slot whose ``masked_lm_weights`` is 1;
        with ``items`` (list per sequence of list per slot of candidate ids) the candidates sorted by descending
        logit (stable: lower list index first on ties); without, the whole vocabulary sorted by logit."""
        sess, d = self._encode_for_ranking(encoder_input)
        B = d["input_word_ids"].shape[0]
        counts = sess.counts().cpu()
        n = int(counts[0])
        if "masked_lm_weights" in d:
            per_seq = (d["masked_lm_weights"] != 0).sum(1).cpu().tolist()
        else:
            per_seq = [d["masked_lm_positions"].shape[1]] * B
        flat_rankings = [None] * n
        if items is not None and len(items) and type(items[0]) is list:
            flat = [items[b][j] for b in range(B) for j in range(per_seq[b])]
            by_len = {}
            for i, lst in enumerate(flat):
                by_len.setdefault(len(lst), []).append(i)
            hidden = sess.mlm_hidden()
            for Cn, idxs in by_len.items():
                cand = torch.tensor([flat[i] for i in idxs], dtype=torch.int64)
                if len(idxs) == n:
                    ranking, _, _ = sess.rank_candidates(cand.to(self.device), None, want_ranking=True)
                else:  # ragged candidate lists: score each length group on a compacted copy of the hidden rows
                    ranking = self._rank_rows(sess, hidden, idxs, cand)
                ranking = ranking.cpu()
                for r, i in enumerate(idxs):
                    flat_rankings[i] = ranking[r]
        else:
            order = self._full_ranking(sess, n)
            for i in range(n):
                flat_rankings[i] = order[i]
        out, k = [], 0
        for b in range(B):
            out.append(flat_rankings[k:k + per_seq[b]])
            k += per_seq[b]
        return out

    def _rank_rows(self, sess, hidden, idxs, cand):
        sel = torch.tensor(idxs, device=self.device)
        saved = hidden[: len(idxs)].clone()
        hidden[: len(idxs)] = hidden[sel]
        ranking, _, _ = sess.rank_candidates(cand.to(self.device), None, want_ranking=True)
        hidden[: len(idxs)] = saved
        return ranking

    def _full_ranking(self, sess, n):
        """items=None: argsort of the whole vocabulary per slot.  Off the measured path: materialises the logits with
        the GEMM kernel and orders them with a stable descending sort on the device."""
        logits = sess.logits(n)
        return torch.sort(logits, dim=-1, descending=True, stable=True).indices.cpu()

    # ------------------------------------------------------------------ config / weights
    def get_config(self):
        return dict(self._config)

    @classmethod
    def from_config(cls, config, custom_object=None):
        return cls(**config)

    def state_dict(self):
        return self.store.state_dict()

    def load_state_dict(self, sd):
        self.store.load_state_dict(sd)

    def save_weights(self, path):
        path = pathlib.Path(str(path))
        path.parent.mkdir(parents=True, exist_ok=True)
        target = path if path.suffix == ".npz" else path.with_name(path.name + ".npz")
        np.savez(target, **{k: v.numpy() for k, v in self.state_dict().items()})
        return target

    def load_weights(self, path):
        path = pathlib.Path(str(path))
        target = path if path.suffix == ".npz" else path.with_name(path.name + ".npz")
        with np.load(target) as z:
            self.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
        return self

    @property
    def trainable_variables(self):
        return [v for k, v in self.store.tf_views().items() if not k.startswith("pooler_transform")]
