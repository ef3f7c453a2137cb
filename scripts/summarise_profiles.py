"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/ (run in the build container)."""
import collections
import csv
import io
import json
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
# ---- launch list ----
txt = open("gpurun_out/launches_%s.csv" % tag).read()
rd = list(csv.reader(io.StringIO("\n".join(l for l in txt.splitlines() if l.startswith('"')))))
hdr = rd[0]
ik, iv, im, iu = (hdr.index(x) for x in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
agg = collections.OrderedDict()
for r in rd[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
    a = agg.setdefault(r[ik].split("(")[0].replace("void ", ""), [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
with open("profiles/%s_launches_bench_sift1m.csv" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 "
            "--warmup 1 --no-cpu (cold-cache, serialised: compare shares)\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('"%s",%d,%.3f,%.4f\n' % (k, c, ms, ms / tot))
# ---- full capture of the top kernel ----
raw = subprocess.check_output(["ncu", "-i", "gpurun_out/%s_knn_tc_sift1m.ncu-rep" % tag, "--page",
                               "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.avg.per_second", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
with open("profiles/%s_knn_tc_kernel_sift1m_ncu.csv" % tag, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 1 -c 1 "
            "python bench.py --steps 2 --warmup 1 --no-cpu\nmetric,unit,value\n")
    for w in want:
        if w in d:
            f.write('"%s","%s","%s"\n' % (w, d[w][0], d[w][1]))
            print(w, d[w])


def tobytes(k):
    u, v = d[k]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9,
                                        "Tbyte": 1e12}[u]


traffic = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
json.dump({"sift1m": traffic,
           "source": "profiles/%s_knn_tc_kernel_sift1m_ncu.csv (dram__bytes_read.sum + "
                     "dram__bytes_write.sum of one knn_tc_kernel launch)" % tag},
          open("profiles/roofline_traffic.json", "w"), indent=1)
print("traffic GB", traffic / 1e9)
