"""bench.py's pure-Python pieces (no GPU): the roofline entry of a one-pivot-per-pass loop and of the
blocked loop, the one-line JSON contract keys of the reference arm's fields, the clock sampler's
parsing."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_roofline_one_pivot_per_pass():
    bytes_pp = 16 * 20001 * 40001
    rl = bench.roofline_block(bytes_pp, pivots=1000, upd_ms=1911.0, upd_n=1000, kernel_names=("k_update", "kb_step"),
                              peak=6544.3, peak_src="measured")
    assert rl["kernel"] == "k_update" and rl["pivots_per_launch"] == 1
    assert abs(rl["achieved"] - bytes_pp / 1.911e-3 / 1e9) < 1e-6
    assert rl["achieved"] == rl["pivot_equiv_achieved"]
    assert abs(rl["frac"] - rl["achieved"] / 6544.3) < 1e-12
    assert rl["bound"] == "hbm" and rl["unit"] == "GB/s"


def test_roofline_blocked_loop_frac_is_physical_and_the_pivot_credit_is_separate():
    """`frac` must be a physical fraction of the HBM roof; the per-pivot credit of SURVEY 8d
    (which exceeds the peak when pivots share a pass) lives in pivot_equiv_*; the FP64 roof is beside it"""
    bytes_pp = 16 * 20001 * 40001
    cells = 20001 * 40001
    rl = bench.roofline_block(bytes_pp, pivots=5120, upd_ms=320 * 2.7, upd_n=320, kernel_names=("k_update", "kb_step"),
                              peak=6544.3, peak_src="measured", traffic=13.04e9, fp64_peak=15.7e12, cells=cells)
    assert rl["kernel"] == "kb_step" and rl["pivots_per_launch"] == 16
    assert rl["bytes_per_launch"] == bytes_pp
    assert abs(rl["achieved"] - bytes_pp / 2.7e-3 / 1e9) < 1e-6
    assert 0.0 < rl["frac"] < 1.0 and abs(rl["frac"] - rl["achieved"] / 6544.3) < 1e-12
    assert abs(rl["pivot_equiv_achieved"] / rl["achieved"] - 16.0) < 1e-9 and rl["pivot_equiv_frac"] > 1.0
    assert rl["traffic"] == 13.04e9
    f = rl["fp64"]
    assert abs(f["achieved_tinst_s"] - 2 * 16 * cells / 2.7e-3 / 1e12) < 1e-9 and 0.0 < f["frac"] < 1.0
    assert rl["binding_roof"] == "hbm" and 0.0 < rl["frac_of_binding_roof"] < 1.0


def test_parity_block_compares_with_the_committed_table(tmp_path, monkeypatch):
    import json
    log = [(1, 2), (3, 4)]
    b = [1.0, 2.5]
    table = {"m": 2, "n": 3, "seed": 0, "made_by": "test",
             "digests": {"2": {"log_sha256": bench.digest_log(log), "b_sha256": bench.digest_b(b)}}}
    f = tmp_path / "d.json"
    f.write_text(json.dumps(table))
    monkeypatch.setattr(bench, "DIGESTS", str(f))
    assert bench.parity_block(2, 3, 0, 2, log, b)["ok"] is True
    assert bench.parity_block(2, 3, 0, 2, [(1, 2), (3, 5)], b)["ok"] is False
    assert bench.parity_block(2, 3, 0, 2, log, [1.0, 2.25])["ok"] is False
    assert bench.parity_block(2, 3, 0, 4, log + log, b)["ok"] is None          # no committed digest for this count
    assert bench.parity_block(2, 3, 1, 2, log, b)["ok"] is None                # another seed
    assert bench.parity_block(2, 3, 0, 2, log, b, usable=False, why="x")["ok"] is None


def test_both_arms_use_the_same_config_block():
    assert bench.config_block(20000, 40000, 0) == {"workload": bench.workload_name(20000, 40000), "seed": 0}
    assert bench.host_threads() >= 1


def test_clock_sampler_parses_nvidia_smi_lines():
    s = bench.ClockSampler(0)
    s.proc = object()          # pretend a process was started; stop() only terminates real ones
    s.samples = ["1965, 1965, 400.1, Not Active, Not Active, Not Active, Not Active",
                 "1740, 1965, 995.0, Not Active, Not Active, Not Active, Active",
                 "1755, 1965, 990.0, Not Active, Not Active, Not Active, Active",
                 "garbage"]
    s.first = 1

    class _P:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    s.proc = _P()
    out = s.stop()
    assert out["sm_max_mhz"] == 1965.0 and out["samples"] == 2
    assert out["sm_mhz"] in (1740.0, 1755.0) and out["reasons"] == ["sw_power_cap"]


def test_workload_name_states_the_size():
    assert "20000x40000" in bench.workload_name(20000, 40000) and "6.40 GB" in bench.workload_name(20000, 40000)


def test_committed_digest_table_is_anchored_on_the_cpu_twin():
    """tests/golden/bench_c4_seed0_digests.json comes from the pivot-per-pass GPU kernels; its first
    checkpoints must equal the binary64 CPU twin's (tools/make_bench_digests.py cpu), committed beside it"""
    import json
    g = os.path.join(ROOT, "tests", "golden")
    tab = json.load(open(os.path.join(g, "bench_c4_seed0_digests.json")))
    cpu = json.load(open(os.path.join(g, "bench_c4_seed0_digests_cpu_twin.json")))
    assert (tab["m"], tab["n"], tab["seed"]) == (cpu["m"], cpu["n"], cpu["seed"]) == (20000, 40000, 0)
    assert len(cpu["digests"]) >= 8 and max(int(k) for k in cpu["digests"]) == tab["cpu_twin_checked_up_to"]
    for k, v in cpu["digests"].items():
        assert tab["digests"][k] == v
    # the driver's default runs land on committed checkpoints: (warmup + steps) * 256 pivots
    for warm, steps in [(3, 20), (5, 20), (3, 2), (1, 2)]:
        assert str((warm + steps) * 256) in tab["digests"]


def test_committed_ncu_traffic_covers_every_bench_line():
    """roofline.traffic comes from profiles/r02_traffic.json (ncu --set full captures, one per shard size):
    every key bench.py / sharded.py ask for is there and within 10 % of the algorithmic bytes of that shard"""
    for world in (1, 2, 4, 8):
        t = bench.ncu_traffic("kb_step_n%d" % world)
        algorithmic = 16 * (20000 // world + 1) * 40001
        assert t is not None and 0.95 * algorithmic < t < 1.10 * algorithmic, (world, t, algorithmic)
    assert bench.ncu_traffic("k_update_n1") is not None and bench.ncu_traffic("kb_step_c3") is not None
    assert bench.ncu_traffic("no_such_kernel") is None
