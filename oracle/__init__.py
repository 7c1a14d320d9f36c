"""CPU oracle of the EMDR2 retrieve-and-read hot path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never from emdr2_b200/.
"""
