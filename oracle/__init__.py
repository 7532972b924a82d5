"""TEST INFRASTRUCTURE ONLY - the parity checker.  Import from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs only.  Nothing in mysteryann_b200/ imports this."""
