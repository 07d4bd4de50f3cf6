"""CPU oracle for the dpr-scale bi-encoder training path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import this package; the product (``dpr_scale_b200``) never does and has no CPU fallback.

Parity status: the reference pins no values for this path (its tests assert shapes only,
dpr_scale/models/tests/test_models.py:52-54; nothing tests dpr_task.py).  The oracle is therefore pinned
against outputs of the reference itself, generated in the authoring container by
``tests/golden/make_golden.py`` (imports /root/reference unmodified) and committed under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks restatement == golden.
"""
