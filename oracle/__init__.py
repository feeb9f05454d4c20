"""ORACLE — test infrastructure only.

CPU restatements of the reference's hot path, used by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
the CHECKER.  The product package (unopose_b200) never imports this package.
"""
