"""CPU oracle of the rVPM P2P path -- TEST INFRASTRUCTURE (see vpm_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  The product never does."""
