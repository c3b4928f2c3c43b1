"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SAC-RCBF safety hot path.

Nothing in the product package ``sac_rcbf_b200`` imports from here.  Allowed importers:
tests/, __graft_entry__.smoke(), bench.py (cpu_baseline leg and --impl reference).
"""
