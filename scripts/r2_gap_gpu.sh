#!/bin/bash
# Runs ON THE GPU BOX: what the filter costs on top of the MMA pipeline (250k of 1M rows as
# queries) -- PQH_TC_DEBUG=1: the filter warps release every tile untouched; 2: they copy the
# tile and take the hot path only (no appends, no shrinks); 0: the product.  (With 1 or 2 the
# results come from the exact kernel afterwards.)  Then list capacity 512 against 256.
run() { echo "== $*"; env "$@" timeout 300 python scripts/prof_run.py --n 1000000 --rows ${R:-250000} --reps 2 2>&1 | python scripts/short_stats.py | tail -1; }
run PQH_TC_DEBUG=0
run PQH_TC_DEBUG=2
run PQH_TC_DEBUG=1
R=1000000 run PQH_TC_GCAP=256
R=1000000 run PQH_TC_GCAP=512 PQH_TC_SLICE_MB=4096
