#!/usr/bin/env python
"""bench.py — pivots/s and tableau-update HBM GB/s of the simplex pivot loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[3], the configuration the metric is quoted on): synthetic
dense LP, 20,000 constraints x 40,000 variables, max with <= rows (SURVEY.md §8d generator,
seed 0), FP64 tableau of 6.4 GB resident in HBM.  One *step* = `--pivots-per-step` pivots of
the running solve (getEntering / getLeaving / pivot on the device, LPSolver.java:101-112); the
default loop is the blocked one (16 pivots share one pass over the tableau, lps_blocked.cuh).

The JSON line carries, beside the contract keys:
  value      pivots/s, whole job, inputs already in HBM, CUDA-event time on the library's
             stream (max over ranks)
  e2e        the same metric through the reference-facing call path with HOST buffers:
             lps_load (H2D of the whole tableau from pinned memory) + run + read-back of
             b, c, v and the position map, wall clock around synchronous calls
  roofline   tableau-update kernel: algorithmic bytes 16(m+1)(n+1) per launch / its mean
             CUDA-event duration over the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle's C binary64 twin of LPState.pivotConcurrently (kind "port";
             the Java reference cannot run: no JVM) on the box's host cores, bounded sample

`--impl reference` times that CPU port alone on the same config (all host threads).
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

METRIC = "pivots_per_sec"
UNIT = "pivots/s"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel on the 20000x40000
# workload, from the committed ncu --set full captures (profiles/): key = (config, pivots per launch)
NCU_TRAFFIC = {("n1", 1): 12.756e9, ("n1", 16): 13.040e9}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--m", type=int, default=20000)
    ap.add_argument("--n", type=int, default=40000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--pivots-per-step", type=int, default=256)
    ap.add_argument("--e2e-pivots", type=int, default=1024)
    ap.add_argument("--cpu-pivots", type=int, default=0, help="0 = sized for ~10-30 s")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--loop-mode", type=int, default=0,
                    help="0 auto (blocked loop at this size), 1 three kernels per pivot, 2 persistent loop, 5 blocked loop")
    ap.add_argument("--block", type=int, default=0, help="pivots per tableau pass of the blocked loop (0 = library default, 1 = off)")
    return ap.parse_args()


def workload_name(m, n):
    return "synthetic dense LP %dx%d (max, <= rows, seed-generated, FP64 tableau %.2f GB)" % (
        m, n, 8.0 * (m + 1) * (n + 1) / 1e9)


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def roofline_block(bytes_pp, pivots, upd_ms, upd_n, kernel_names, peak, peak_src, traffic=None):
    """The dominant kernel's roofline entry.

    One launch of the pass kernel applies `pivots/upd_n` pivots (1 for the pivot-per-pass kernels,
    block_pivots for the blocked loop).  `achieved` is the contract's figure: ALGORITHMIC bytes per
    launch = 16(m+1)(n+1) per pivot x pivots per launch, over the mean launch time — it exceeds the
    HBM peak when several pivots share one pass, which is the point of the blocked loop.  The kernel's
    real DRAM rate (one read + one write of the tableau per launch) is `dram_achieved` / `dram_frac`."""
    upd_n = max(int(upd_n), 1)
    avg_ms = upd_ms / upd_n
    per_launch = pivots / upd_n
    achieved = bytes_pp * per_launch / (avg_ms * 1e-3) / 1e9 if avg_ms else 0.0
    dram = bytes_pp / (avg_ms * 1e-3) / 1e9 if avg_ms else 0.0
    blocked = per_launch > 1.5
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": kernel_names[1 if blocked else 0], "launches": upd_n, "avg_ms": avg_ms,
            "peak_source": peak_src, "pivots_per_launch": per_launch,
            "bytes_per_launch": int(bytes_pp * per_launch), "dram_bytes_per_launch": int(bytes_pp),
            "dram_achieved": dram, "dram_frac": dram / peak, "dram_frac_of_8tbs": dram / 8000.0,
            "frac_of_8tbs": achieved / 8000.0,
            "note": ("blocked loop: %.1f pivots are replayed per pass, so the pass moves 16(m+1)(n+1) bytes once for "
                     "all of them; `achieved` counts the per-pivot algorithmic bytes (SURVEY 8d), `dram_achieved` the "
                     "bytes the kernel really moves; at this block size the pass is FP64-issue-bound, not HBM-bound"
                     % per_launch) if blocked else "one pivot per pass"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples, self.proc, self.index, self.first = [], None, index, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def wait_first(self, timeout=10.0):
        """block until nvidia-smi has answered once (it can take seconds on an 8-GPU box)"""
        t0 = time.perf_counter()
        while self.proc and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        """the timed region starts here: earlier samples (warm-up) are dropped unless nothing else arrives"""
        self.first = len(self.samples)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = self.samples[self.first:]
        note = "sampled during the timed region"
        if len(timed) < 2:                       # a very short timed region: fall back to the warm-up steps too
            timed, note = self.samples, "timed region shorter than the sampling period: warm-up samples included"
        for s in timed:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "note": note}


# ------------------------------------------------------------------------------------------------
def cpu_port_rate(A, b, c, pivots, threads):
    """pivots/s of the oracle's C twin of pivotConcurrently on host cores (bench-only use of oracle/)."""
    from oracle import tier_f
    st = tier_f.TierFState(A, b, c, nthreads=threads)
    t0 = time.perf_counter()
    status, k = st.run(pivots)
    dt = time.perf_counter() - t0
    return (k / dt if dt > 0 else 0.0), k, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path, all host threads, same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import tier_f
    threads = tier_f.lib().tf_max_threads()
    m, n = args.m, args.n
    A, b, c = tier_f.gen_dense_feasible(m, n, args.seed, nthreads=threads)
    st = tier_f.TierFState(A, b, c, nthreads=threads)
    # one pivot moves 16*m*n bytes through host DRAM: bound a step to ~1-2 s
    t0 = time.perf_counter()
    st.run(1)
    t1 = time.perf_counter() - t0
    per_step = max(1, min(args.pivots_per_step, int(1.0 / max(t1, 1e-4))))
    for _ in range(args.warmup):
        st.run(per_step)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, k = st.run(per_step)
        done += k
    dt = time.perf_counter() - t0
    value = done / dt
    sample = "%d steps x %d pivots of the %dx%d solve (one pivot = %.1f GB of host DRAM traffic)" % (
        args.steps, per_step, m, n, 16.0 * m * n / 1e9)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(m, n), "pivots_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C binary64 port of LPState.pivotConcurrently (oracle/tier_f.c); the Java reference "
                "cannot run (no JVM in the image); an upper bound on its BigDecimal speed",
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    import linear_programming_solver_b200 as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # keep stdout to the one JSON line: NCCL's banner ("NCCL version ...") goes there at VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1:
        from linear_programming_solver_b200 import sharded
        return sharded.bench_sharded(args, dist, rank, world, local_rank, METRIC, UNIT, workload_name,
                                     measured_peak, ClockSampler, roofline_block)

    m, n, P = args.m, args.n, args.pivots_per_step
    kw = dict(device=local_rank, time_kernels=True, loop_mode=args.loop_mode, block_pivots=args.block)
    if args.variant >= 0:
        kw["update_variant"] = args.variant
    st = L.LPState.synthetic_dense(m, n, args.seed, 1000, **kw)
    bytes_pp = st.algorithmic_bytes_per_pivot()

    def barrier():
        torch.cuda.synchronize()

    # nvidia-smi needs about a second to start answering: launch it before the warm-up and mark where
    # the timed region begins, so that short timed regions still get their samples
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for _ in range(args.warmup):
        st.run(P)
    barrier()
    sampler.mark()
    dev_ms, upd_ms, upd_n, launches, pivots = 0.0, 0.0, 0, 0, 0
    t0 = time.perf_counter()
    restarts = 0
    for _ in range(args.steps):
        need = P
        while need > 0:
            r = st.run(need)
            dev_ms += r.device_ms
            upd_ms += r.update_ms
            upd_n += r.update_launches
            launches += r.kernel_launches
            pivots += r.npivots
            need -= r.npivots
            if r.verdict != 3:          # the solve reached a verdict inside the timed region: solve it again
                if r.npivots == 0 and restarts > 0:
                    raise SystemExit("the synthetic LP does not pivot (verdict %d)" % r.verdict)
                st.regenerate()
                restarts += 1
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    if pivots != args.steps * P:
        raise SystemExit("pivot count mismatch: %d != %d" % (pivots, args.steps * P))
    value = pivots / (dev_ms / 1e3)
    peak, peak_src = measured_peak()
    rl = roofline_block(bytes_pp, pivots, upd_ms, upd_n, ("lps::k_update", "lps::kb_flush"), peak, peak_src,
                        traffic=NCU_TRAFFIC.get(("n1", round(pivots / max(upd_n, 1)))))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(m, n), "pivots_per_step": P, "seed": args.seed,
                   "loop": "blocked: %.1f pivots per tableau pass" % rl["pivots_per_launch"]
                           if rl["pivots_per_launch"] > 1.5 else "one tableau pass per pivot",
                   "l2": "tableau (6.4 GB) is far larger than the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the library's stream around each step",
                   "restarts": restarts},
        "gpu_launches": int(launches),
        "loop_gbs": bytes_pp * pivots / (dev_ms * 1e-3) / 1e9,
        "frac_of_8tbs": bytes_pp * pivots / (dev_ms * 1e-3) / 1e9 / 8000.0,
        "loop_dram_gbs": bytes_pp * upd_n / (dev_ms * 1e-3) / 1e9,
        "wall_s": wall,
        "roofline": rl,
        "clocks": clocks,
    }
    st.close()

    # ---- host copy of the same input (untimed): generated on the device, read back ----
    need_host = not (args.no_e2e and args.no_cpu_baseline)
    if need_host:
        from linear_programming_solver_b200.lp_state import _dp
        A_pin = torch.empty((m, n), dtype=torch.float64, pin_memory=True)
        A_host = A_pin.numpy()
        g = L.LPState.synthetic_dense(m, n, args.seed, 1000, device=local_rank)
        g._ck(g._lib.lps_read_A(g._h, _dp(A_host), n), "read_A")
        b_host, c_host = g.b, g.c
        g.close()
    # ---- e2e: host buffers in, results out, through the reference-facing call sequence ----
    if not args.no_e2e:
        Pe = args.e2e_pivots
        best = None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s = L.LPState(A_host, b_host, c_host, m, n, device=local_rank, loop_mode=args.loop_mode,
                          block_pivots=args.block)                                # H2D of the tableau
            r = s.run(Pe)
            out_b, out_c, out_v, out_pos = s.b, s.c, s.v, s.positions         # D2H of the result
            dt = time.perf_counter() - t0
            s.close()
            best = dt if best is None else min(best, dt)
        line["e2e"] = {"value": Pe / best, "unit": UNIT,
                       "h2d_bytes_per_step": int(8 * (m * n + m + n)),
                       "d2h_bytes_per_step": int(8 * (m + n + 1) + 4 * (m + n)),
                       "pivots_per_call": Pe, "seconds_per_call": best,
                       "what": "LPState(A,b,c) from pinned host memory + run(%d) + read b,c,v,positions" % Pe}
    # ---- CPU baseline on the same input (bounded sample) ----
    if not args.no_cpu_baseline:
        from oracle import tier_f
        threads = tier_f.lib().tf_max_threads()
        A_cpu = np.array(A_host, copy=True)
        rate1, k1, dt1 = cpu_port_rate(A_cpu, b_host.copy(), c_host.copy(), 2, threads)
        want = args.cpu_pivots or max(3, min(60, int(15.0 * rate1)))
        A_cpu[...] = A_host
        rate, k, dt = cpu_port_rate(A_cpu, b_host.copy(), c_host.copy(), want, threads)
        line["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d pivots of the same %dx%d LP in %.1f s, C binary64 twin of "
                      "LPState.pivotConcurrently with %d threads (reference THREAD_AMOUNT is 4)" % (k, m, n, dt, threads)}
        # SURVEY 8d also asks for the reference's own thread count (THREAD_AMOUNT = 4, LPState.java:22);
        # a short extra sample, never allowed to cost the line
        try:
            if threads > 4:
                A_cpu[...] = A_host
                rate4, k4, dt4 = cpu_port_rate(A_cpu, b_host.copy(), c_host.copy(), max(2, min(12, int(4.0 * rate))), 4)
                line["cpu_baseline"]["value_4_threads"] = rate4
                line["cpu_baseline"]["sample_4_threads"] = "first %d pivots in %.1f s with 4 threads" % (k4, dt4)
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"]["value_4_threads"] = None
            line["cpu_baseline"]["sample_4_threads"] = "failed: %s" % ex
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
