"""CPU oracle for the dip-benchmark hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  See oracle/dip_oracle.c for what it restates.
"""
