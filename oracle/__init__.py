"""CPU oracle for the BitDance image-generation hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this package. Nothing under ``bitdance_b200/`` imports it; the product path has no CPU fallback.
"""
