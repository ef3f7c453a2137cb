#!/bin/bash
# Runs ON THE GPU BOX: staging threads / chunk size of the pageable host copies.
for T in 8 12; do for C in 512 1024 2048; do
  echo "== threads $T chunk $C KB"; PQH_STAGE_THREADS=$T PQH_STAGE_CHUNK_KB=$C timeout 300 python scripts/e2e_host.py 2>&1 | tail -2
done; done
