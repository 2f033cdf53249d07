"""ORACLE - TEST INFRASTRUCTURE ONLY (see ``oracle/pandapower_nr.py`` header).

CPU restatement (NumPy/SciPy fp64) of the reference hot path: pandapower 2.7.0's default
``runpp`` plus MAPDN's ``VoltageControl`` env logic. PARITY UNPINNED (no pandapower, no
reference tests/golden vectors available) - pinned instead by closed-form / literature /
independent-solver checks in ``tests/test_oracle_*.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package; the product (``mapdn_b200``) never does.
"""
