"""CPU checks of the gather plan (job classification + record layout) the persistent CUDA gather kernel works from.

The kernel trusts the host for everything it does not re-check on the device: that every window of a staged tile lies
inside its TMA box, that a seam tile's box wraps around the left/right border and its records are re-based onto the
unwrapped box, that the warps flagged in shareMask really keep their source column down their four rows, that a
thread's four records sit in one output column, and that every output pixel has exactly one record.  All of that is
host code (csrc/gather_plan.cpp) and is verified here without a GPU, against the plain row-major sampling records
(T360B200_hostPlanSamples), which tests/test_host_plan.py pins against the reference.
"""
import numpy as np
import pytest

import transform360_b200 as t360
from tests.golden.cases import FULL, SMALL, plane_dims

KIND_SHIFT, PLANE_SHIFT, ROW_MASK = 24, 28, (1 << 24) - 1
BOX_W = {0: 192, 1: 240}


def box_h(k, cls):
    return (112 if cls == 0 else 144) if k == 8 else (64 if cls == 0 else 96)


def _plan(case, plane):
    ctx = t360.make_context(**case["ov"])
    iw, ih, ow, oh, _ = plane_dims(case, plane)
    hp = t360.HostPlan(ctx, iw, ih, ow, oh)
    return ctx, hp, iw, ih


def _unpack(records):
    x, y = records[..., 0].astype(np.int64), records[..., 1].astype(np.int64)
    col_in_seg = (x >> 27) & 31
    col0 = ((x & ((1 << 27) - 1)) ^ (1 << 26)) - (1 << 26)  # sign-extend 27 bits
    return col_in_seg, col0, y >> 10, y & 1023


CASES = [("small", n, p) for n in ("cube_cubic", "cube_linear", "cube_lanczos", "cube_cubic_odd", "lp_tiles", "lr_stereo", "rotated",
                                   "eac_tb_lanczos", "cube_to_equirect", "cfg1", "barrel") for p in (0, 1)] + \
        [("full", "cfg2", 0), ("full", "cfg2", 1), ("full", "cfg4", 1)]


@pytest.mark.parametrize("group,name,plane", CASES)
def test_gather_plan_invariants(group, name, plane):
    case = (SMALL if group == "small" else FULL)[name]
    ctx, hp, iw, ih = _plan(case, plane)
    k = hp.kernel_size
    g = hp.gather_plan()
    s = hp.samples.astype(np.int64)  # [mapH][mapW][2] = {col0, row0 << 10 | phase}, row-major
    mh, mw = s.shape[:2]
    tpr, th = g["tiles_per_row"], g["tile_h"]
    assert th == (64 if k == 8 else 32) and tpr == -(-mw // 32) and g["tile_rows"] == -(-mh // th)
    colseg, col0, row0, phase = _unpack(g["records"])

    # -- which tiles are re-based (seam), from the job list
    jobs = g["jobs"]
    staged_plan = k >= 2 and ctx.output_layout not in (t360.LAYOUT_BARREL, t360.LAYOUT_BARREL_SPLIT)
    assert (jobs is not None) == staged_plan
    seam_box = {}
    if jobs is not None:
        cnt = g["counts"]
        assert len(jobs) == tpr * g["tile_rows"] == cnt["class0"] + cnt["class1"] + cnt["seam"] + cnt["general"]
        kinds = (jobs[:, 1] >> KIND_SHIFT) & 15
        assert list(kinds) == sorted(kinds, key=lambda v: {2: 0, 3: 1, 1: 2, 0: 3}[int(v)]), "launch order: general, seam, class 1, class 0"
        assert (np.bincount(kinds, minlength=4)[[0, 1, 3, 2]] == [cnt["class0"], cnt["class1"], cnt["seam"], cnt["general"]]).all()
        assert ((jobs[:, 1] >> PLANE_SHIFT) == 0).all()
        seen = set()
        for ox, oy, boxxy, share in jobs:
            tile = ((oy & ROW_MASK) // th) * tpr + ox // 32
            assert tile not in seen and ox % 32 == 0 and (oy & ROW_MASK) % th == 0
            seen.add(tile)
            if (oy >> KIND_SHIFT) & 15 == 3:
                seam_box[tile] = boxxy & 0xFFFF
        assert len(seen) == len(jobs)

    # -- every output pixel has exactly one record, with its own phase / row and (re-based) first column
    for ty in range(g["tile_rows"]):
        for tx in range(tpr):
            tile = ty * tpr + tx
            y0, x0 = ty * th, tx * 32
            hh, ww = min(th, mh - y0), min(32, mw - x0)
            cs, c0, r0, ph = colseg[tile, :hh], col0[tile, :hh], row0[tile, :hh], phase[tile, :hh]
            want = s[y0:y0 + hh, x0:x0 + ww]
            if ww == 32:
                assert (np.sort(cs, axis=1) == np.arange(32)).all(), "a row segment must hold every column once"
                # a thread (lane) keeps one output column through the four rows of its block
                blk = cs[:hh // 4 * 4].reshape(-1, 4, 32)
                assert (blk == blk[:, :1]).all(), "lane order must be constant inside a 32 x 4 block"
            else:
                assert (cs[:, :ww] == np.arange(ww)).all() and (g["records"][tile, :hh, ww:] == 0).all()
            got_rowphase = np.take_along_axis((r0 << 10) | ph, np.argsort(cs[:, :ww] if ww < 32 else cs, axis=1), axis=1)[:, :ww]
            got_col0 = np.take_along_axis(c0, np.argsort(cs[:, :ww] if ww < 32 else cs, axis=1), axis=1)[:, :ww]
            assert (got_rowphase == want[..., 1]).all()
            if tile in seam_box:
                bx = seam_box[tile]
                assert (got_col0 % iw == want[..., 0] % iw).all() and (got_col0 >= bx).all() and (got_col0 + k <= bx + BOX_W[0]).all()
            else:
                assert (got_col0 == want[..., 0]).all()
            if hh < th:
                assert (g["records"][tile, hh:] == 0).all(), "rows below the plane are padded with zero records"

    if jobs is None:
        return
    # -- boxes and share masks
    for ox, oy, boxxy, share in jobs:
        kind, yy = (oy >> KIND_SHIFT) & 15, oy & ROW_MASK
        blk = s[yy:yy + th, ox:ox + 32]
        c, r = blk[..., 0], blk[..., 1] >> 10
        bx, by = boxxy & 0xFFFF, boxxy >> 16
        if kind in (0, 1):
            assert bx % 16 == 0 and c.min() >= bx and c.max() + k <= bx + BOX_W[kind] and bx + BOX_W[kind] <= iw + BOX_W[kind]
            assert c.min() >= 0 and c.max() + k <= iw, "a staged tile never needs BORDER_WRAP"
            assert r.min() >= by >= 0 and r.max() + k <= by + box_h(k, kind) and r.max() + k <= ih
        elif kind == 3:
            assert iw % 16 == 0 and bx % 16 == 0 and bx < iw < bx + BOX_W[0], "the box of a seam tile wraps around the border"
            rel = (c % iw - bx) % iw
            assert rel.max() + k <= BOX_W[0] and r.min() >= by >= 0 and r.max() + k <= by + box_h(k, 0) and r.max() + k <= ih
        if kind == 2:
            assert boxxy == 0 and share == 0  # general tiles: taps through L1, nothing for the host to promise
            continue
        # share mask: bit w <=> all four rows exist and every column keeps its first column, 1-2 source rows apart
        for w in range(th // 4):
            ya = yy + 4 * w
            if k < 4 or ya + 4 > mh:
                ok = False
            else:
                q = s[ya:ya + 4, ox:ox + 32]
                d = np.diff(q[..., 1] >> 10, axis=0)
                ok = bool(((d == 1) | (d == 2)).all() and (q[..., 0] == q[:1, :, 0]).all())
            assert bool((share >> w) & 1) == ok, f"tile ({ox},{yy}) warp {w}"


def test_tile_counts_of_the_headline_plan():
    """cfg2 (8K equirect -> 3840x2560 cubemap, bicubic): the numbers DESIGN.md quotes."""
    for plane, want in ((0, dict(class0=8808, class1=432, seam=112, general=248)), (1, dict(class0=2168, class1=104, seam=56, general=72))):
        _, hp, _, _ = _plan(FULL["cfg2"], plane)
        assert hp.gather_plan()["counts"] == want


def test_gather_plan_invariants_on_random_contexts():
    """The same invariants over random contexts (all layouts, stereo, rotation, off-centre incl. NaN map entries, scale
    factors, odd sizes, input widths with and without whole 16-byte columns)."""
    from tests.test_host_plan import _random_context
    rng = np.random.default_rng(500)
    checked = 0
    for _ in range(30):
        ov = _random_context(rng)
        iw, ih = int(rng.integers(200, 700)) * 2, int(rng.integers(100, 300)) * 2
        ow, oh = int(rng.integers(40, 200)) * 2 + int(rng.random() < 0.3), int(rng.integers(30, 150)) * 2 + int(rng.random() < 0.3)
        if rng.random() < 0.5:
            iw = (iw + 15) // 16 * 16
        SMALL["__random"] = dict(ov=ov, inp=(iw, ih), out=(ow, oh))
        try:
            test_gather_plan_invariants("small", "__random", int(rng.integers(0, 2)))
            checked += 1
        except ValueError:  # the planner refuses what the reference refuses
            pass
        finally:
            SMALL.pop("__random", None)
    assert checked >= 25
