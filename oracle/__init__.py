"""ORACLE — test infrastructure only (CPU restatement of the reference path). PARITY UNPINNED.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
