"""CPU oracle for the Minigrid hot path — TEST INFRASTRUCTURE ONLY (see mg_oracle.h).

Nothing in ``minigrid_b200`` imports this package. Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs may.
"""
