"""CPU oracle for the differentiable ray-march (TEST INFRASTRUCTURE, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
See oracle/cpu_ref.c for what is restated, oracle/ti_shim.py for how it is pinned against the reference's own source, and
oracle/taichi_probe.py for what stays unpinned (the real Taichi compiler's rounding).
"""
