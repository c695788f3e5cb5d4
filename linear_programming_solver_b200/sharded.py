"""Row-sharded LPState: one rank per GPU of one NVSwitch box (include/lps_b200.h, lps_shard_*).

`partition(m, world, rank)` is the reference's own block split (LPState.java:222-223).  The
per-pivot exchange happens inside the CUDA kernels over peer memory; `torch.distributed` is used
only for plumbing: exchanging the 64-byte CUDA IPC handles at set-up, gathering results and
the max-over-ranks timing.
"""
from __future__ import annotations

import ctypes
import time
from ctypes import byref, c_int, c_void_p
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from .exceptions import LpsError
from .lp_state import LPState, _dp


def partition(m: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [from, to) of rank `rank`: from = k*m/T, to = (k+1)*m/T (LPState.java:222-223)."""
    return (rank * m) // world, ((rank + 1) * m) // world


def owner_of(row: int, m: int, world: int) -> int:
    for k in range(world):
        lo, hi = partition(m, world, k)
        if lo <= row < hi:
            return k
    raise ValueError("row out of range")


def reduce_candidates(cands):
    """Lexicographic (ratio, global row) minimum over the ranks' candidates, the rule the kernels
    apply (lowest row wins ties — LPState.java:292-303).  cands: iterable of (ratio, row) with
    row = -1 for "no candidate".  Returns the winning (ratio, row) or (inf, -1)."""
    best = (float("inf"), -1)
    for ratio, row in cands:
        if row < 0:
            continue
        if best[1] < 0 or ratio < best[0] or (ratio == best[0] and row < best[1]):
            best = (ratio, row)
    return best


class ShardedLPState(LPState):
    """The local shard of a row-sharded LPState.  Construct on every rank, then `attach`."""

    def __init__(self, m_total: int, n: int, rank: int, world: int, A_local=None, b_local=None, c=None,
                 v: float = 0.0, synthetic_seed: Optional[int] = None, pos_permille: int = 1000,
                 epsilon: float = LPState.DEF_EPSILON, inf: float = LPState.DEF_INF, device: int = -1,
                 time_kernels: bool = False, loop_mode: int = 0, block_pivots: int = 0,
                 synthetic_kind: int = N.LPS_GEN_DENSE, update_variant: int = -1, panel_ctas: int = 0,
                 pass_chunk_rows: int = 0):
        self._lib = N.load()
        self._h = c_void_p()
        self._names0 = None
        self.rank, self.world, self.m_total = rank, world, m_total
        self.row0, self.row1 = partition(m_total, world, rank)
        opts = N.default_options()
        opts.epsilon, opts.inf, opts.device, opts.time_kernels = epsilon, inf, device, int(time_kernels)
        opts.loop_mode = int(loop_mode)
        opts.block_pivots = int(block_pivots)
        opts.update_variant = int(update_variant)
        opts.panel_ctas, opts.pass_chunk_rows = int(panel_ctas), int(pass_chunk_rows)
        rc = self._lib.lps_create(byref(self._h), byref(opts))
        if rc != N.LPS_OK:
            raise LpsError(rc, "lps_create: " + self._lib.lps_status_string(rc).decode())
        if synthetic_seed is not None:
            self._ck(self._lib.lps_shard_generate_lp(self._h, int(synthetic_kind), m_total, n, rank, world,
                                                     synthetic_seed, pos_permille), "lps_shard_generate_lp")
        else:
            mloc = self.row1 - self.row0
            A_local = np.ascontiguousarray(np.asarray(A_local, dtype=np.float64).reshape(mloc, n))
            b_local = np.ascontiguousarray(np.asarray(b_local, dtype=np.float64).reshape(mloc))
            c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(n))
            self._ck(self._lib.lps_shard_load(self._h, m_total, n, rank, world, _dp(A_local), max(n, 1),
                                              _dp(b_local), _dp(c), float(v)), "lps_shard_load")

    # -- exchange set-up --------------------------------------------------------------------
    def export_ipc(self) -> bytes:
        buf = ctypes.create_string_buffer(64)
        self._ck(self._lib.lps_shard_export(self._h, buf), "lps_shard_export")
        return buf.raw

    def attach_ipc(self, handles) -> None:
        blob = b"".join(handles)
        assert len(blob) == 64 * self.world
        self._ck(self._lib.lps_shard_attach_ipc(self._h, ctypes.c_char_p(blob)), "lps_shard_attach_ipc")

    def comm_ptr(self) -> int:
        p = c_void_p()
        self._ck(self._lib.lps_shard_comm_ptr(self._h, byref(p)), "lps_shard_comm_ptr")
        return p.value

    def attach_ptrs(self, ptrs) -> None:
        arr = (c_void_p * self.world)(*ptrs)
        self._ck(self._lib.lps_shard_attach_ptrs(self._h, arr), "lps_shard_attach_ptrs")

    def attach_via(self, dist) -> None:
        """Exchange IPC handles through torch.distributed (plumbing) and attach."""
        if self.world == 1:
            return
        handles = [None] * self.world
        dist.all_gather_object(handles, self.export_ipc())
        self.attach_ipc(handles)
        dist.barrier()

    # -- global views -------------------------------------------------------------------------
    @property
    def positions(self) -> np.ndarray:
        out = np.empty(self.m_total + self.n, dtype=np.int32)
        self._ck(self._lib.lps_read_positions(self._h, out.ctypes.data_as(ctypes.POINTER(c_int))), "read_positions")
        return out

    def gather_b(self, dist=None) -> np.ndarray:
        local = self.b
        if self.world == 1 or dist is None:
            return local
        parts = [None] * self.world
        dist.all_gather_object(parts, local)
        return np.concatenate(parts)

    def primal(self, nvars: int, dist=None) -> np.ndarray:
        b = self.gather_b(dist)
        pos = self.positions
        x = np.zeros(nvars)
        n = self.n
        for p in range(n, n + self.m_total):
            if pos[p] < nvars:
                x[pos[p]] = b[p - n]
        return x


# ---------------------------------------------------------------------------------------------
def bench_sharded(args, dist, rank, world, local_rank, B):
    """bench.py's N > 1 arm: the same LP row-sharded over `world` GPUs (strong scaling).  `B` is the bench
    module (helpers shared with the single-GPU arm: roofline record, clock sampler, digests)."""
    import json

    import torch

    m, n, P = args.m, args.n, args.pivots_per_step
    kw = dict(loop_mode=args.loop_mode, block_pivots=args.block, panel_ctas=args.panel_ctas)
    if args.variant >= 0:
        kw["update_variant"] = args.variant
    st = ShardedLPState(m, n, rank, world, synthetic_seed=args.seed, device=local_rank, time_kernels=True, **kw)
    st.attach_via(dist)
    bytes_pp_local = st.algorithmic_bytes_per_pivot()        # this rank's rows (+ objective replica)
    bytes_pp_global = 16 * (m + 1) * (n + 1)
    mloc = st.row1 - st.row0

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()

    sampler = B.ClockSampler(local_rank)     # started before the warm-up: nvidia-smi takes a second to answer
    sampler.start()
    if rank == 0:
        sampler.wait_first()
    total = 0
    for _ in range(args.warmup):
        total += st.run(P).npivots
    barrier()
    sampler.mark()
    dev_ms, upd_ms, upd_n, launches, pivots = 0.0, 0.0, 0, 0, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = st.run(P)
        dev_ms += r.device_ms
        upd_ms += r.update_ms
        upd_n += r.update_launches
        launches += r.kernel_launches
        pivots += r.npivots
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    total += pivots
    fp64_peak = st.measure_fp64_issue_rate(100.0)
    t = torch.tensor([dev_ms, upd_ms / max(upd_n, 1)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, upd_avg_ms = float(t[0]), float(t[1])
    tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
    dist.all_reduce(tl)
    tf = torch.tensor([fp64_peak], dtype=torch.float64, device="cuda")
    dist.all_reduce(tf, op=dist.ReduceOp.MIN)

    # parity: every rank holds the whole pivot log (it must be the same one) and its rows of b
    logs = [None] * world
    dist.all_gather_object(logs, B.digest_log(st.pivot_log))
    b_global = st.gather_b(dist)
    parity = None
    if rank == 0:
        parity = B.parity_block(m, n, args.seed, total, st.pivot_log, b_global)
        parity["ranks_agree_on_the_log"] = bool(len(set(logs)) == 1)
        if not parity["ranks_agree_on_the_log"]:
            parity["ok"] = False

    # e2e: every rank loads ITS rows from pinned host memory, runs, reads its part of the result
    e2e = None
    if not args.no_e2e:
        A_pin = torch.empty((mloc, n), dtype=torch.float64, pin_memory=True)
        A_host = A_pin.numpy()
        g = ShardedLPState(m, n, rank, world, synthetic_seed=args.seed, device=local_rank)
        g._ck(g._lib.lps_read_A(g._h, _dp(A_host), n), "read_A")
        b_host, c_host = g.b, g.c
        g.close()
        Pe = args.e2e_pivots
        best = None
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            s = ShardedLPState(m, n, rank, world, A_host, b_host, c_host, device=local_rank, **kw)
            s.attach_via(dist)
            r2 = s.run(Pe)
            out = (s.b, s.c, s.v, s.positions)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.barrier()
            s.close()
            best = float(tt[0]) if best is None else min(best, float(tt[0]))
        e2e = {"value": Pe / best, "unit": B.UNIT, "h2d_bytes_per_step": int(8 * (m * n + m + world * n)),
               "d2h_bytes_per_step": int(8 * (m + world * (n + 1)) + 4 * world * (m + n)),
               "pivots_per_call": Pe, "seconds_per_call": best,
               "what": "per rank: shard load from pinned host memory + IPC attach + run(%d) + read b,c,v,positions" % Pe}
    loop_desc = st.loop_description()       # after the timed region: the split the runs settled on
    bad = False
    if rank == 0:
        peak, peak_src = B.measured_peak()
        value = pivots / (dev_ms_max / 1e3)
        per_launch = round(pivots / max(upd_n, 1))
        kernels = ("lps::ks_update (per rank)", loop_desc.split(": ", 1)[-1] + " (per rank)")
        rl = B.roofline_block(bytes_pp_local, pivots, upd_avg_ms * max(upd_n, 1), upd_n, kernels, peak, peak_src,
                              traffic=B.ncu_traffic("kb_step_n%d" % world if per_launch > 1 else "ks_update_n%d" % world),
                              fp64_peak=float(tf[0]), cells=(mloc + 1) * (n + 1))
        line = {
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": B.config_block(m, n, args.seed),
            "details": {"pivots_per_step": P, "loop": loop_desc,
                        "sharding": "rows [k*m/G,(k+1)*m/G) per rank, objective row replicated",
                        "exchange": "ratio candidates + scaled pivot row pushed into peer memory over NVLink "
                                    "inside the kernels (no NCCL in the loop)",
                        "l2": "per-rank shard (%.2f GB) is larger than the 126 MB L2" % (bytes_pp_local / 2e9),
                        "timing": "CUDA events on each rank's stream, max over ranks"},
            "gpu_launches": int(tl[0]),
            "loop_gbs": bytes_pp_global * pivots / (dev_ms_max * 1e-3) / 1e9,
            "frac_of_8tbs_per_gpu": bytes_pp_global * pivots / (dev_ms_max * 1e-3) / 1e9 / 8000.0 / world,
            "loop_dram_gbs_per_gpu": bytes_pp_local * upd_n / (dev_ms_max * 1e-3) / 1e9,
            "wall_s": wall,
            "roofline": rl,
            "parity": parity,
            "clocks": clocks,
        }
        if e2e:
            line["e2e"] = e2e
        print(json.dumps(line), flush=True)
        bad = parity["ok"] is False
    st.close()
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        raise SystemExit("PARITY FAILURE: the %d-rank run's pivot log / b column differ from the committed digests" % world)
