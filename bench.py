#!/usr/bin/env python
"""bench.py -- exact kNN queries/s on B200 for the workloads BASELINE.json names.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" is one whole all-pairs kNN job (every query row against the whole base).
N > 1 is launched by torch.distributed.run, one rank per GPU: rank 0 holds the
vectors, the base is broadcast once over NCCL (NVLink), query rows are split
evenly over the ranks and the per-rank neighbour lists are gathered to rank 0
(SURVEY.md 8e).  Rank 0 prints ONE JSON line:

  value      queries/s with the vectors already resident in HBM (device-timed,
             CUDA events, max over ranks)
  e2e        queries/s through the host-pointer C ABI (pqh_knn_search_rows, what
             knn_full_thread calls) with PAGEABLE host buffers, as the reference's caller
             hands them over (compute_nn_fast.c:471-481): staged H2D + search + D2H inside
             the timed region.  For N > 1 the same single call runs on rank 0 with the
             library's own multi-GPU path (PQH_DEVICES = all N GPUs: every GPU uploads 1/N
             of the base over its own link, one ncclAllGather over NVLink, query rows split
             over the GPUs, results written straight into the caller's arrays)
  roofline   the tcgen05 distance-GEMM kernel against the measured dense bf16/fp16 peak
  cpu_baseline  the Yael-shaped CPU implementation (oracle/knn_blas.c) on a bounded
             sample of the same workload, on this box's host cores (N == 1 only)

--impl reference times that CPU implementation alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (generator, n, d, num_nn)   -- k passed to the search is num_nn + 1 (self + num_nn)
    "sift100k": ("sift_like", 100_000, 128, 16),
    "sift1m": ("sift_like", 1_000_000, 128, 16),
    "gist1m": ("gist_like", 1_000_000, 960, 16),
    "deep10m": ("deep_like", 10_000_000, 96, 32),
}
DESCRIPTIONS = {
    "sift100k": "synthetic SIFT-shaped 100k x 128 fvecs, k=16, single block",
    "sift1m": "SIFT1M-shaped 1M x 128 fvecs all-pairs kNN k=16",
    "gist1m": "GIST1M-shaped 1M x 960 fvecs kNN k=16",
    "deep10m": "Deep10M-shaped 10M x 96 fvecs kNN k=32",
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region (B200_PROFILING.md).

    Sampled in-process through NVML every few milliseconds (a timed region of a short workload
    is over before a freshly started `nvidia-smi -lms` prints its first line), one sample taken
    synchronously when the region starts; `nvidia-smi` is the fallback when NVML cannot be
    loaded.  The device is found by its PCI bus id, so CUDA_VISIBLE_DEVICES does not matter."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []      # nvidia-smi lines
        self.samples = []   # NVML: (sm_mhz, power_w, reasons bitmask)
        self.proc = None
        self.gpu = gpu_index
        self.nvml = None
        self.handle = None
        self.smax = None
        self.stop_flag = threading.Event()
        self.thread = None

    def _nvml_open(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            pr = torch.cuda.get_device_properties(self.gpu)
            bus = f"{pr.pci_domain_id:08X}:{pr.pci_bus_id:02X}:{pr.pci_device_id:02X}.0"
            self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            return True
        except Exception:
            self.nvml = None
            return False

    def _nvml_sample(self):
        n = self.nvml
        try:
            sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
            rs = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            self.samples.append((sm, pw, rs))
        except Exception:
            pass

    def _nvml_loop(self):
        while not self.stop_flag.wait(0.004):
            self._nvml_sample()

    def start(self):
        if self._nvml_open():
            self._nvml_sample()
            self.thread = threading.Thread(target=self._nvml_loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    @staticmethod
    def _summary(sm, smax, power, reasons, source):
        # under-load samples: the upper half of the SM clock readings
        sm = sorted(sm)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons), "source": source}

    def stop(self):
        if self.nvml is not None:
            self._nvml_sample()      # one more while the last step is still warm
            self.stop_flag.set()
            self.thread.join(timeout=1)
            n = self.nvml
            bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown,
                    "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown,
                    "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
            reasons = {name for name, b in bits.items() if any(r & b for _, _, r in self.samples)}
            return self._summary([s[0] for s in self.samples], self.smax,
                                 [s[1] for s in self.samples], reasons, "nvml")
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return self._summary(sm, smax, power, reasons, "nvidia-smi")


def config_of(workload, gpus):
    """The `config` object of the JSON line: the same keys and values for both arms."""
    gen, n, d, num_nn = WORKLOADS[workload]
    per = (n + gpus - 1) // gpus
    return {"workload": DESCRIPTIONS[workload], "n": n, "d": d, "k": num_nn,
            "rows_per_gpu": per, "parallelism": f"query-rows/{gpus}",
            "l2": "inputs (%.0f MB) larger than L2, no flush needed" % (n * d * 4 / 1e6)}


def make_data(workload):
    from pq_huffman_b200 import synth
    gen, n, d, num_nn = WORKLOADS[workload]
    return getattr(synth, gen)(n, d), num_nn + 1


def cpu_sample(x, k, seconds_target, threads):
    """Time the Yael-shaped CPU kNN on a bounded sample: a slice of query rows x whole base."""
    from tests import oracle_ffi
    blas = oracle_ffi.BlasBaseline()
    n = x.shape[0]
    probe = min(n, 1024)
    t = time.perf_counter()
    blas.knn_rows(x, None, k, 0, probe, n_thread=threads)
    dt = max(time.perf_counter() - t, 1e-6)
    rows = int(min(n, max(probe, probe * seconds_target / dt)))
    rows = max(threads, rows // threads * threads)
    t = time.perf_counter()
    blas.knn_rows(x, None, k, 0, rows, n_thread=threads)
    dt = time.perf_counter() - t
    return rows / dt, rows, dt


def run_reference(args):
    """--impl reference: the reference's CPU path (Yael-shaped stand-in), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    x, k = make_data(args.workload)
    threads = os.cpu_count() or 1
    n, d = x.shape
    per_step = max(2.0, min(30.0, 90.0 / max(1, args.steps + args.warmup)))
    vals, rows_used = [], 0
    for i in range(args.warmup + args.steps):
        qps, rows, dt = cpu_sample(x, k, per_step, threads)
        if i >= args.warmup:
            vals.append((rows, dt))
        rows_used = rows
    tot_rows = sum(r for r, _ in vals)
    tot_dt = sum(t for _, t in vals)
    qps = tot_rows / tot_dt
    sample = (f"{rows_used} query rows x full {n}-row base per step (of {n} rows), "
              f"linear in query rows")
    line = {
        "impl": "reference", "metric": "exact kNN queries/sec", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n / qps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(args.workload, args.gpus),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "Yael v438 is not vendored/installed; oracle/knn_blas.c restates "
                                 "its sgemm + per-query heap algorithm on scipy's OpenBLAS"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from pq_huffman_b200 import dist_knn, knn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # control-plane barriers that keep the GPUs free (an NCCL barrier spins on the device)
        ctl = dist.new_group(backend="gloo")
    knn.init([local])

    gen, n, d, num_nn = WORKLOADS[args.workload]
    k = num_nn + 1
    q_lo, q_hi, per = dist_knn.row_slice(n, world, rank)
    rows = q_hi - q_lo

    # ---- data: rank 0 generates, pinned; the other ranks only receive the broadcast ----
    host_x = x_np = None
    if rank == 0:
        x_np, _ = make_data(args.workload)           # plain numpy memory: pageable
        host_x = torch.from_numpy(x_np).pin_memory()
    d_x = torch.empty((n, d), dtype=torch.float32, device=dev)
    if rank == 0:
        d_x.copy_(host_x, non_blocking=True)
    dist_knn.broadcast_base(d_x, src=0)
    torch.cuda.synchronize()

    d_assign = torch.empty((max(rows, 1), k), dtype=torch.int32, device=dev)
    d_dis = torch.empty((max(rows, 1), k), dtype=torch.float32, device=dev)
    wbytes = knn.workspace_bytes(max(rows, 1), n, d, k)
    d_work = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    # a real (non-null) stream: the library runs its kernels on the handle it is given, and
    # torch.cuda.Event only sees work on the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)

    def device_step():
        knn.search_device(local, n, n, d, k, d_x.data_ptr(), d_x.data_ptr(), True, q_lo, q_hi,
                          d_assign.data_ptr(), d_dis.data_ptr(), d_work.data_ptr(), wbytes,
                          stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed arm ----
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gemm_ms, launches, st_acc = 0.0, 0, None
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
        st = knn.last_stats()
        gemm_ms += st["ms_gemm"]
        launches += st["launches"]
        st_acc = st
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms, gemm_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, gemm_ms_max = float(t[0]), float(t[1])
    value = n * args.steps / (ms_max / 1e3)

    # ---- MMA-only ceiling of the GEMM kernel (N == 1): same launch, filter switched off ----
    mma_only_ms = None
    if world == 1 and not args.no_ceiling:
        os.environ["PQH_TC_DEBUG"] = "1"
        knn.init([local])
        vals = []
        for _ in range(2):      # (every row then goes to the exact kernel: ~15 s per run at 1M)
            device_step()
            torch.cuda.synchronize()
            vals.append(knn.last_stats()["ms_gemm"])
        mma_only_ms = vals[1]
        del os.environ["PQH_TC_DEBUG"]
        knn.init([local])

    # ---- end-to-end arm: pageable host pointers through the C ABI, on rank 0 ----
    e2e_steps = max(1, min(args.steps, 3))
    e2e_s, e2e_stats = 0.0, None
    if ctl is not None:
        torch.cuda.synchronize()
        dist.barrier(group=ctl)
    if rank == 0:
        del host_x
        knn.init(list(range(world)))      # the library's own multi-GPU context (NCCL inside)
        out_a = np.empty((n, k), dtype=np.int32)
        out_d = np.empty((n, k), dtype=np.float32)
        lib = knn._lib.load()

        def e2e_step():
            # the call a user of the reference makes: host buffers in, host results out
            rc = lib.pqh_knn_search_rows(n, n, d, k, x_np.ctypes.data, x_np.ctypes.data, 1, 0, n,
                                         out_a.ctypes.data, out_d.ctypes.data)
            if rc:
                raise RuntimeError(lib.pqh_last_error().decode())

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        e2e_s = time.perf_counter() - t0
        e2e_stats = knn.last_stats()
        assert out_a[n // 2, 0] == n // 2 and e2e_stats["rows"] == n
    if ctl is not None:
        dist.barrier(group=ctl)
    e2e_value = n * e2e_steps / e2e_s if rank == 0 else 0.0

    if rank == 0:
        peaks, peak_src = load_peaks()
        flops_per_launch = 2.0 * rows * n * d          # algorithmic: 2*Nb*d per query
        launches_gemm = args.steps
        achieved = (flops_per_launch / (gemm_ms_max / launches_gemm / 1e3) / 1e12
                    if gemm_ms_max > 0 else 0.0)
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        # which tensor-core flavour ran: uint8-valued data takes the exact INT8 path
        int8_path = st_acc["inexact_operands"] == 2
        operands = ("u8 x s8 -> s32 (tcgen05.mma kind::i8, exact scores)" if int8_path else
                    "f16 x f16 -> f32 (tcgen05.mma kind::f16)")
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get(args.workload + ("_i8" if int8_path else ""))
        line = {
            "metric": "exact kNN queries/sec", "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": ("f32 (results: exact FP32 re-rank; candidate filter: int8 tensor cores, "
                      "s32 accumulate)" if int8_path else
                      "f32 (results: exact FP32 re-rank; candidate filter: f16 tensor cores, "
                      "f32 accumulate)"),
            "data": "synthetic",
            "config": config_of(args.workload, world),
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": n * d * 4,
                    "d2h_bytes_per_step": n * k * 8, "steps": e2e_steps,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "buffers": "pageable",
                    "api": "pqh_knn_search_rows (host pointers)" if world == 1 else
                           f"pqh_knn_search_rows on rank 0, library context over {world} GPUs: "
                           "split H2D + ncclAllGather + per-GPU search + D2H into the caller's "
                           "arrays"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor",
                         # d > 144 takes the streaming kernel (A and B k-blocks through one ring)
                         "kernel": "knn_tcs_kernel" if d > 144 else "knn_tc_kernel",
                         "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "peak_source": f"{peak_src} bf16_tflops_sustained (kernel runs inside a "
                                        "long step)",
                         "ms_per_launch": gemm_ms_max / launches_gemm, "traffic": traffic,
                         # the same launch with the filter epilogue switched off: what the MMA
                         # pipeline alone takes here (operand format and clocks included)
                         "ms_mma_only": mma_only_ms,
                         "frac_of_mma_only": (mma_only_ms / (gemm_ms_max / launches_gemm)
                                              if mma_only_ms else None),
                         "operands": operands,
                         # the contract's denominator stays the measured bf16 peak; the INT8
                         # pipe's own nominal dense peak on B200 is twice the 16-bit one
                         "peak_operand_format_nominal": 4500.0 if int8_path else 2250.0,
                         "frac_operand_format_nominal":
                             achieved / (4500.0 if int8_path else 2250.0)},
            "phases_ms": {key: st_acc[key] for key in ("ms_prep", "ms_gemm", "ms_rerank",
                                                       "ms_exact")},
            "rows_exact_fallback": st_acc["rows_exact"],
            "candidates_per_row": st_acc["candidates"] / max(1, st_acc["rows"]),
            "clocks": clocks,
        }
        if e2e_stats:
            line["e2e"]["phases_ms"] = {key: e2e_stats[key] for key in
                                        ("ms_h2d", "ms_kernels", "ms_d2h", "ms_total")}
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            qps, srows, dt = cpu_sample(x_np, k, args.cpu_seconds, threads)
            line["cpu_baseline"] = {
                "value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                "sample": f"{srows} query rows x full {n}-row base in {dt:.1f} s "
                          "(oracle/knn_blas.c, Yael-shaped sgemm + heaps)"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """Libraries (NCCL's version banner, OpenBLAS warnings) print to fd 1; the contract is ONE
    JSON line on stdout.  Everything else goes to stderr; emit() writes the line to the real fd."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="sift1m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ceiling", action="store_true", help="skip the MMA-only ceiling runs")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
