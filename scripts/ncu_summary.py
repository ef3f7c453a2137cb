"""ncu report -> tracked summary under profiles/ (run in the build container, no GPU needed).

  python scripts/ncu_summary.py <gpurun_out/x.ncu-rep> <profiles/out.csv> "<command that was profiled>"

Writes the metrics the roofline discussion uses (duration, tensor-pipe activity, DRAM bytes, L2,
issue slots, stall reasons, TMEM loads, registers) and prints the DRAM traffic in bytes.
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_tmem_ldt.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out, what = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, one launch of: %s\n" % what)
        f.write("# kernel: %s\n" % d.get("Kernel Name", ("", "?"))[1])
        f.write("metric,unit,value\n")
        for w in WANT:
            if w in d:
                f.write('"%s","%s","%s"\n' % (w, d[w][0], d[w][1]))
                print(w, d[w])

    def tobytes(k):
        u, v = d[k]
        return float(v.replace(",", "")) * UNIT[u]

    print("dram traffic bytes", tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum"))


if __name__ == "__main__":
    main()
