"""CPU oracle for the differentiable ray-march (TEST INFRASTRUCTURE, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package.
See oracle/cpu_ref.c and oracle/torch_ref.py for what is restated and why parity is "unpinned".
"""
