"""Host logic of srb_conv_wgrad_batched (csrc/wgrad_umma.cu plan_launches), checked without a GPU through
srb_wgrad_plan: how the 64x64 weight-gradient blocks of a training step are grouped into launches and split
over CTAs.  Invariants: every block appears once, a launch holds <= 74 blocks and <= #SMs CTAs, no block gets
more CTAs than tiles, and the CTAs of one launch carry about the same work."""
import ctypes as C
import math

import pytest


def _plan(blocks, num_sms=148):
    from srb200 import lib
    h = lib.load()
    n = len(blocks)
    arr = lambda vals: (C.c_int32 * n)(*vals)
    tiles, k1 = arr([b[0] for b in blocks]), arr([b[1] for b in blocks])
    t_out, l_out, s_out = arr([0] * n), arr([0] * n), arr([0] * n)
    lib.check(h.srb_wgrad_plan(num_sms, n, tiles, k1, t_out, l_out, s_out), "srb_wgrad_plan")
    return [(abs(t_out[i]), int(t_out[i] < 0), l_out[i], s_out[i]) for i in range(n)]


def _cost(tiles, k1):
    return tiles * (0.3 if k1 else 1.0)


def _launches(plan):
    out = {}
    for tiles, k1, li, s in plan:
        out.setdefault(li, []).append((tiles, k1, s))
    return [out[k] for k in sorted(out)]


# (tiles, is_1x1) per block of one training step at batch 16 x 48x48 LR, x4
RCAN = [(288, 0)] * 417 + [(1152, 0)] * 4 + [(4608, 0)]          # 411 body + head + body-end + 4 up1 | 4 up2 | tail
RDN = [(288, 0)] * (16 * 36 + 6) + [(288, 1)] * (16 * 9 + 16) + [(1152, 0)] * 4 + [(4608, 0)]
EDSR = [(288, 0)] * 38 + [(1152, 0)] * 4 + [(4608, 0)]


@pytest.mark.parametrize("name,blocks", [("rcan", RCAN), ("rdn", RDN), ("edsr", EDSR)])
def test_training_step_plans_are_balanced(name, blocks):
    plan = _plan(blocks)
    assert sorted((t, k) for t, k, _, _ in plan) == sorted(blocks)                 # every block exactly once
    for launch in _launches(plan):
        ctas = sum(s for _, _, s in launch)
        assert len(launch) <= 74 and ctas <= 148
        assert all(1 <= s <= min(t, 96) for t, _, s in launch)
        per_cta = [_cost(math.ceil(t / s), k) for t, k, s in launch]
        ideal = sum(_cost(t, k) for t, k, _ in launch) / 148
        # the slowest CTA is within 35 % of a perfectly even split (a launch that cannot fill the SMs is
        # bounded by its largest single tile count instead)
        assert max(per_cta) <= max(1.35 * ideal, 1.0) + 1e-9, (name, max(per_cta), ideal)
    if name == "rcan":
        first = _launches(plan)[0]
        assert (4608, 0) in [(t, k) for t, k, _ in first] and sum(s for _, _, s in first) == 148
        # body layers: two CTAs each, 74 per launch
        full = [l for l in _launches(plan) if len(l) == 74]
        assert len(full) == 5 and all(s == 2 for l in full for _, _, s in l)


def test_small_batches_use_the_whole_device():
    plan = _plan([(288, 0)])
    assert plan == [(288, 0, 0, 96)]                       # one block: as many CTAs as the cap allows
    plan = _plan([(2, 0)] * 3)
    assert [s for _, _, _, s in plan] == [2, 2, 2]          # never more CTAs than tiles
    plan = _plan([(288, 0)] * 10, num_sms=8)
    assert all(sum(s for _, _, s in l) <= 8 or len(l) == 1 for l in _launches(plan))


def test_plan_rejects_bad_input():
    from srb200 import lib
    h = lib.load()
    one = (C.c_int32 * 1)(0)
    out = (C.c_int32 * 1)(0)
    assert h.srb_wgrad_plan(148, 1, one, one, out, out, out) != 0
    assert b"tiles" in h.srb_last_error()
