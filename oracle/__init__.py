"""CPU oracle of the ORB front end -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this package.  See oracle/orb_oracle.h for what it restates and how
it is pinned."""
