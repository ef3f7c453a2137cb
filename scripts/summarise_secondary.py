"""Tracked summaries (profiles/) of the ncu --set full captures of the secondary kernels."""
import csv
import io
import subprocess

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
for rep, out, cmd in [
        ("r1_knn_tcs_gist1m", "r1_knn_tcs_kernel_gist1m_ncu.csv",
         "-k regex:knn_tcs_kernel -s 2 -c 1 python bench.py --workload gist1m --steps 1 --warmup 1 --no-cpu"),
        ("r1_knn_rerank_sift1m", "r1_knn_rerank_kernel_sift1m_ncu.csv",
         "-k regex:knn_rerank_kernel -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu"),
        ("r1_knn_rerank_gist1m", "r1_knn_rerank_kernel_gist1m_ncu.csv",
         "-k regex:knn_rerank_kernel -s 2 -c 1 python bench.py --workload gist1m --steps 1 --warmup 1 --no-cpu")]:
    raw = subprocess.check_output(["ncu", "-i", "gpurun_out/%s.ncu-rep" % rep, "--page", "raw", "--csv"],
                                  text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    print("==", rep, d.get("Kernel Name"))
    with open("profiles/" + out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on %s\nmetric,unit,value\n" % cmd)
        for w in WANT:
            if w in d:
                f.write('"%s","%s","%s"\n' % (w, d[w][0], d[w][1]))
                print(w, d[w])
