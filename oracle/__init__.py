"""TEST INFRASTRUCTURE ONLY -- CPU oracle of rnascan's scoring path (see oracle/oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package; rnascan_b200 never does.
"""
