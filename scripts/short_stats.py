"""stdin: lines printed by scripts/prof_run.py -> the few fields that matter, one line each."""
import ast
import sys

for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = ast.literal_eval(line)
        print("  prep %.2f gemm %.2f rerank %.2f exact %.2f wall %.2f | cand/row %.1f pass2 %d exact_rows %d "
              "i8 %d launches %d" % (d["ms_prep"], d["ms_gemm"], d["ms_rerank"], d["ms_exact"],
                                     d["wall_ms"], d["candidates"] / max(1, d["rows"]),
                                     d["rows_second_pass"], d["rows_exact"], d["inexact_operands"],
                                     d["launches"]))
    elif line:
        print("  " + line)
