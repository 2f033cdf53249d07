"""ORACLE - TEST INFRASTRUCTURE ONLY (see ``oracle/pandapower_nr.py`` header).

CPU restatement (NumPy/SciPy fp64) of the reference hot path: pandapower 2.7.0's default
``runpp`` plus MAPDN's ``VoltageControl`` env logic.

* Env logic (``voltage_control_ref.py``): PINNED by trajectories the reference's own, unmodified env code produced in
  this container (``ref_harness.py`` / ``ref_scenarios.py`` -> ``tests/golden/ref_env_*.npz``,
  ``tests/test_reference_golden.py``).
* Power flow (``pandapower_nr.py``, ``c/nr_dense.c``): PARITY UNPINNED against pandapower itself (not installable
  here, the reference ships no tests / golden vectors) - pinned instead by published load-flow results (IEEE 33- and
  69-bus feeders, Stagg & El-Abiad 5-bus, Saadat Ex. 6.7), closed forms and independent solvers
  (``tests/test_literature_kats*.py``, ``tests/test_oracle_*.py``); ``scripts/pin_with_pandapower.py`` is the
  one-command upgrade wherever pandapower is importable.

Only ``tests/``, ``scripts/make_*golden*.py``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / parity legs
may import this package; the product (``mapdn_b200``) never does.
"""
