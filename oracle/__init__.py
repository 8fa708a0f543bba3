"""oracle/ -- CPU oracle for the point-cloud front end.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (detmatch_b200/) never does.
"""
