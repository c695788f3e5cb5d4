"""Host logic of the look-ahead step's SM split (include/lps_b200.h: lps_plan_split_model /
lps_plan_split_tuned): pure arithmetic, no device.  The split never changes results (GPU tests sweep it);
these tests pin its behaviour as a controller."""
import pytest

from linear_programming_solver_b200 import _native as N

G = 148          # SMs of a B200


@pytest.fixture(scope="module")
def lib():
    return N.load()


def _rows(world):
    return 20000 // world


def test_model_gives_the_panel_more_sms_on_smaller_shards(lib):
    picks = [lib.lps_plan_split_model(G, 16, w, _rows(w), 40016) for w in (1, 2, 4, 8)]
    assert picks == sorted(picks) and picks[0] < picks[-1]
    assert 2 <= picks[0] <= 12 and 24 <= picks[-1] <= G // 2
    # never the whole grid, never nothing
    for w in (1, 2, 4, 8):
        for rows, ld in ((10, 16), (2500, 40016), (60000, 120016)):
            p = lib.lps_plan_split_model(G, 16, w, rows, ld)
            assert 1 <= p <= G - 1
    assert lib.lps_plan_split_model(1, 16, 1, 100, 112) < 0          # LPS_ERR_INVALID
    assert lib.lps_plan_split_model(G, 0, 1, 100, 112) < 0


def test_tuner_moves_towards_the_slower_role_and_stops_when_balanced(lib):
    # panel-bound: 16 pivots x 40 us = 640 us against a 400 us pass -> more panel CTAs
    up = lib.lps_plan_split_tuned(G, 16, 8, 30, 40.0, 400.0)
    assert 30 < up <= 40                                              # at most a third of the way per run
    # pass-bound: 16 x 20 us = 320 us against 900 us -> fewer
    down = lib.lps_plan_split_tuned(G, 16, 8, 30, 20.0, 900.0)
    assert 20 <= down < 30
    # balanced within the 2 % hysteresis: stays
    assert lib.lps_plan_split_tuned(G, 16, 8, 40, 35.0, 560.0) == 40
    # nonsense measurements leave the split alone
    assert lib.lps_plan_split_tuned(G, 16, 8, 40, 0.0, 560.0) == 40
    assert lib.lps_plan_split_tuned(G, 16, 8, 40, 35.0, float("nan")) == 40
    assert lib.lps_plan_split_tuned(G, 16, 8, G, 35.0, 560.0) == G       # not a split: returned as given
    assert lib.lps_plan_split_tuned(1, 16, 8, 40, 35.0, 560.0) < 0


@pytest.mark.parametrize("world,sync_us,per_kcell_us,pass_full_us", [(1, 10.0, 11.5, 2400.0), (4, 22.0, 12.0, 780.0),
                                                                     (8, 21.0, 15.0, 400.0)])
def test_tuner_converges_on_a_synthetic_machine(lib, world, sync_us, per_kcell_us, pass_full_us):
    """Closed loop against a machine whose roles follow panel(p) = sync + k * cells / p and
    pass(p) = full * G / (G - p): from a bad start the controller settles, in a few runs, on a split
    whose step time is within 3 % of the best one, and then stops moving."""
    cells_k = 1e-3 * (_rows(world) + 1 + 40016)

    def panel(p):
        return sync_us + per_kcell_us * cells_k / p

    def pas(p):
        return pass_full_us * G / (G - p)

    def step(p):
        return max(16 * panel(p), pas(p))

    best = min(step(p) for p in range(2, G // 2 + 1))
    for start in (2, 70):
        p = start
        seen = []
        for _ in range(12):
            p = lib.lps_plan_split_tuned(G, 16, world, p, panel(p), pas(p))
            seen.append(p)
        assert step(p) <= 1.03 * best, (start, seen, step(p), best)
        assert seen[-1] == seen[-2] == seen[-3], seen               # settled
