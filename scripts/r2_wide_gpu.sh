#!/bin/bash
# Runs ON THE GPU BOX: coalescing re-rank on narrower rows.
run() { echo "== $*"; env "$@" timeout 300 python scripts/prof_run.py --n $N --d $D --k $K --gen $G --reps 3 2>&1 | python scripts/short_stats.py | tail -2; }
N=1000000 D=128 K=17 G=sift_like
run PQH_TC_WIDE_D=256
run PQH_TC_WIDE_D=64
N=100000
run PQH_TC_WIDE_D=256
run PQH_TC_WIDE_D=64
N=2000000 D=96 K=33 G=deep_like
run PQH_TC_WIDE_D=256
run PQH_TC_WIDE_D=64
