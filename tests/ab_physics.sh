#!/bin/bash
# A/B of the launch-policy knobs on the cfg2 workload (tuning; results never depend on them)
run() { echo "== $*"; env "$@" python tests/tune_physics.py 2>&1 | tail -1; }
run SS_X=0
run SS_SYNC=1
run SS_SETS=1
run SS_SETS=3
run SS_NOSORT=1
