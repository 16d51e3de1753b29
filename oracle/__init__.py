"""CPU oracle for the quantization hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``quantizers_b200/`` may import this package.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
"""
