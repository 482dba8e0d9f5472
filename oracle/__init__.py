"""ORACLE -- test infrastructure, never product code.

CPU FP64 restatement of the reference (sail-sg/jrystal) energy+gradient path, used as the
checker by ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of
``bench.py``.  Nothing under ``jrystal_b200/`` may import this package.
See ``oracle/reference_port.py`` for the pin status ("parity unpinned" at the jax_xc
boundary) and ``DESIGN.md``.
"""
