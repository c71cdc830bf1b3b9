"""CPU parity oracle of hydpy_b200 — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
package.  The product (`hydpy_b200/`) must never do so.
"""
