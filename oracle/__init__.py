"""CPU oracle for the mdpy nonbonded hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
package; the product (mdpy_b200/) never does.  See oracle/mdpy_oracle.c for what is a restatement of
the reference (pinned by tests/golden/) and what is float64 truth for physics the reference lacks
(PME, erfc direct space, CHARMM switch: parity unpinned by the reference).
"""
