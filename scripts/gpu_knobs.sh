#!/bin/bash
# environment-only knobs of the Fock build on two workloads
run() { for w in "def2-tzvp 4,4,4" "cc-pvtz 4,4,4"; do echo "[$1] $w: $(env $1 python scripts/fock_profile.py $w 1 2>/dev/null | tail -1)"; done; }
run "LB200_FOCK_STREAMS=4"
run "LB200_FOCK_STREAMS=8"
run "LB200_FOCK_BUCKETS=1,4"
run "LB200_FOCK_BUCKETS=1,9"
run "LB200_FOCK_BUCKETS=1,3,12"
run "LB200_FOCK_BUCKETS=1,6,24"
run "LB200_FOCK_BUCKETS=1"
