"""CPU oracle for the BERT4Rec train + ranking hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` arms may import it, and only as the checker or the timed
CPU baseline.  The product package (``bert4rec_b200``) never imports it.

Pinning status (see DESIGN.md "Oracle"):

* integer / host parts (masking, samplers, tokenizer ids, metrics, rank
  semantics): PINNED against the reference's own functions, executed from
  ``/root/reference`` under a stub ``tensorflow`` module by
  ``oracle/gen_golden.py`` -> ``tests/golden/*.json``; and against the
  reference's known-answer tests (``tests/evaluators_tests/
  evaluation_metrics_tests.py:12-103``).
* floating-point network / loss / optimizer arithmetic: **parity unpinned** --
  the arithmetic lives in tensorflow==2.10.0, keras==2.10.0 and
  tf-models-official==2.10.1, none of which is present under /root/reference
  or installable here; the restatement follows the published layer semantics
   and the reference's call sites.  It IS cross-checked
  against an independent implementation of the same architecture (Hugging Face
  ``BertForMaskedLM`` with the oracle's weights: outputs, loss and gradients,
  ``tests/test_oracle_vs_hf_bert.py``); parity with TensorFlow itself remains
  unpinned.
"""
