"""CPU oracle for the MI critic / estimator hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker (or as the timed CPU baseline), never on the product path.

Parity status
-------------
* ``dv_bound_loss`` / ``infonce_bound_loss`` / ``create_mi_pairs``: PINNED —
  ``tests/golden/*.npz`` were produced by executing the reference's own
  functions (``oracle/make_golden.py`` imports them verbatim from
  ``/root/reference``), and ``oracle.matrix_oracle`` is checked against those
  vectors and, when ``/root/reference`` is present, against the live reference.
* row / symmetric InfoNCE and the dot / bilinear critics do not exist in the
  reference (its critic is a concat-MLP, its "infonce" is DV + log N): for those
  the oracle is PARITY UNPINNED by the reference; it is anchored to it only
  through the DV path (same score matrix, same mask, same pair ordering).
* ``oracle.chunked_oracle`` (the same matrix form with the scores formed 4096 rows at a time, so that it runs at
  B = 65536 on whatever device the inputs live on) is pinned to ``oracle.matrix_oracle`` by
  ``tests/test_oracle.py::test_chunked_oracle_equals_matrix_oracle``.
"""
