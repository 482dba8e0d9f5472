"""Parity, host-logic and C-ABI tests of jrystal_b200 (see tests/conftest.py for the gpu marker)."""
