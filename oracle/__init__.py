"""oracle/ -- CPU restatement of the gym-fx env.step() hot path.  TEST INFRASTRUCTURE ONLY:
may be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package (gym_fx_b200)."""
