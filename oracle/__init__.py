"""ORACLE package — test infrastructure only.

CPU restatements of the reference algorithms on the association hot path, used as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (mmmot_b200) must never import anything from here.
"""
