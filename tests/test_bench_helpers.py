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
    rl = bench.roofline_block(bytes_pp, pivots=1000, upd_ms=1911.0, upd_n=1000, kernel_names=("k_update", "kb_flush"),
                              peak=6544.3, peak_src="measured")
    assert rl["kernel"] == "k_update" and rl["pivots_per_launch"] == 1
    assert abs(rl["achieved"] - bytes_pp / 1.911e-3 / 1e9) < 1e-6
    assert rl["achieved"] == rl["dram_achieved"]
    assert abs(rl["frac"] - rl["achieved"] / 6544.3) < 1e-12
    assert rl["bound"] == "hbm" and rl["unit"] == "GB/s"


def test_roofline_blocked_loop_credits_every_pivot_of_the_pass():
    bytes_pp = 16 * 20001 * 40001
    rl = bench.roofline_block(bytes_pp, pivots=5120, upd_ms=320 * 2.7, upd_n=320, kernel_names=("k_update", "kb_flush"),
                              peak=6544.3, peak_src="measured", traffic=13.04e9)
    assert rl["kernel"] == "kb_flush" and rl["pivots_per_launch"] == 16
    assert rl["bytes_per_launch"] == 16 * bytes_pp and rl["dram_bytes_per_launch"] == bytes_pp
    assert abs(rl["achieved"] / rl["dram_achieved"] - 16.0) < 1e-9       # algorithmic bytes vs bytes really moved
    assert rl["dram_frac"] < 1.0 < rl["frac"]
    assert rl["traffic"] == 13.04e9


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
