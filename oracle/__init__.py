"""Parity oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``dexterous_rl_manipulation_b200``) never does; it fails loudly without its CUDA extension.
"""
