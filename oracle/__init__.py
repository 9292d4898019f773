"""CPU oracle for the steganographic coder step.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it, and there only as the checker or as the
CPU baseline being timed -- never as the path being shipped.

Parity status: PINNED by live differential runs against the unmodified
reference (``oracle/ref_harness.py`` imports ``/root/reference/code_base`` and
``/root/reference/src`` in the build container) and by the committed golden
vectors under ``tests/golden/`` that those runs generated
(``oracle/make_golden.py``).  The reference's own test-suite holds no
token-level golden vectors for this path; its only
known-answer test, ``_select_cutoff_k([.4,.35,.25], .1, 50) == 3`` /
``topk=2 -> 2`` (tests/codec/test_arithmetic_threshold.py:43-58), is
reproduced in ``tests/test_oracle.py``.
"""
