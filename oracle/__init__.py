"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's (Refefer/Dampr) algorithm for the hot path plus the
deterministic synthetic-input generators.  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package; the product
(dampr_b200/) never does.
"""
