"""CPU checks of the gather plan (job classification + record layout) the persistent CUDA gather kernel works from.

The kernel trusts the host for everything it does not re-check on the device: that every window of a staged job lies
inside its TMA box, that a seam tile's box wraps around the left/right border and its records are re-based onto the
unwrapped box, that a share job's columns really keep their source column with row steps of 0-2, that the slot field
of a record addresses the weights of the pixel's phase in the shared-memory image, and that every output pixel is
produced exactly once.  All of that is host code (csrc/gather_plan.cpp) and is verified here without a GPU: the
compact records are decoded exactly the way gather_frame.cu decodes them and compared with the plain row-major
sampling records (T360B200_hostPlanSamples), which tests/test_host_plan.py pins against the reference.
"""
import numpy as np
import pytest

import transform360_b200 as t360
from tests.golden.cases import FULL, SMALL, plane_dims

KIND_SHIFT, PLANE_SHIFT, ROW_MASK = 24, 28, (1 << 24) - 1
CLASS0, CLASS1, GENERAL, SHARE_STAY, SHARE, SEAM = 0, 1, 2, 3, 4, 7
SKIP = 0x8000
ORDER = {GENERAL: 0, SEAM: 1, CLASS1: 2, SHARE_STAY: 3, SHARE: 4, CLASS0: 5}
SLOT_MASK = 0x7FF0
INTERP = {2: t360.LINEAR, 4: t360.CUBIC, 8: t360.LANCZOS4}


def box_w(kind):
    return 192 if kind in (SHARE, SHARE_STAY) else (240 if kind == CLASS1 else 208)


def box_h(k, kind):
    if k == 8:
        return 80 if kind in (SHARE, SHARE_STAY) else (128 if kind == CLASS1 else 72)
    return 72 if kind in (SHARE, SHARE_STAY) else (96 if kind == CLASS1 else 64)


def box_variant_rows(k, kind, v):
    """kernels.cuh boxVariantRows: the box a job names is one of three heights of its class (class 1: one)."""
    h = box_h(k, kind)
    if kind == CLASS1 or v == 0:
        return h
    return h - 8 * v if kind in (SHARE, SHARE_STAY) else h - 8 - 8 * v


def share_rows(k):
    return 8


def copies_of(k):
    return 2 if k == 4 else 1


def _plan(case, plane):
    ctx = t360.make_context(**case["ov"])
    iw, ih, ow, oh, _ = plane_dims(case, plane)
    hp = t360.HostPlan(ctx, iw, ih, ow, oh)
    return ctx, hp, iw, ih


class WeightImage:
    """Reads a pixel's weights the way the kernel does: vector v of the slot at byte v * stride + field."""

    def __init__(self, k):
        self.k = k
        self.table = t360.remap_table(INTERP[k]).reshape(1024, k * k)
        self.image = t360.weight_image(INTERP[k])
        self.stride = (8192 if k == 2 else 16384) * copies_of(k)
        assert self.image.size == (8192 if k == 2 else (k * k // 8) * 16384 * copies_of(k))

    def weights_at(self, field):
        """field: int array of record slot fields -> int16 [n][k*k]"""
        field = np.asarray(field, np.int64)
        img16 = self.image.view(np.int16)
        if self.k == 2:
            idx = (field >> 1)[:, None] // 2 + np.arange(4)[None, :]
            return img16[idx]
        nv = self.k * self.k // 8
        if self.k == 8:  # XOR-diagonal image (kernels.cuh): vector m of a slot sits at field ^ (m * 0x4010)
            idx = (field[:, None, None] ^ (np.arange(nv)[None, :, None] * 0x4010)) // 2 + np.arange(8)[None, None, :]
            return img16[idx].reshape(len(field), -1)
        idx = (field[:, None, None] + np.arange(nv)[None, :, None] * self.stride) // 2 + np.arange(8)[None, None, :]
        return img16[idx].reshape(len(field), -1)

    def check(self, field, phase):
        assert ((np.asarray(field) & ~SLOT_MASK) == 0).all()
        assert (self.weights_at(field) == self.table[np.asarray(phase)]).all(), "slot field does not address the pixel's weights"


def _check_full_records(g, s, k):
    """Full records (general kernels / general jobs): every pixel once, lane order constant inside a 32 x 4 block."""
    mh, mw = s.shape[:2]
    tpr, th = g["tiles_per_row"], g["tile_h"]
    assert th == (64 if k == 8 else 32) and tpr == -(-mw // 32) and g["tile_rows"] == -(-mh // th)
    x, y = g["records"][..., 0].astype(np.int64), g["records"][..., 1].astype(np.int64)
    colseg = (x >> 27) & 31
    col0 = ((x & ((1 << 27) - 1)) ^ (1 << 26)) - (1 << 26)  # sign-extend 27 bits
    for ty in range(g["tile_rows"]):
        for tx in range(tpr):
            tile = ty * tpr + tx
            y0, x0 = ty * th, tx * 32
            hh, ww = min(th, mh - y0), min(32, mw - x0)
            cs, c0, rp = colseg[tile, :hh], col0[tile, :hh], y[tile, :hh]
            want = s[y0:y0 + hh, x0:x0 + ww]
            if ww == 32:
                assert (np.sort(cs, axis=1) == np.arange(32)).all(), "a row segment must hold every column once"
                blk = cs[:hh // 4 * 4].reshape(-1, 4, 32)
                assert (blk == blk[:, :1]).all(), "lane order must be constant inside a 32 x 4 block"
                order = np.argsort(cs, axis=1)
            else:
                assert (cs[:, :ww] == np.arange(ww)).all() and (g["records"][tile, :hh, ww:] == 0).all()
                order = np.argsort(cs[:, :ww], axis=1)
            assert (np.take_along_axis(rp, order, axis=1)[:, :ww] == want[..., 1]).all()
            assert (np.take_along_axis(c0, order, axis=1)[:, :ww] == want[..., 0]).all()
            if hh < th:
                assert (g["records"][tile, hh:] == 0).all(), "rows below the plane are padded with zero records"


CASES = [("small", n, p) for n in ("cube_cubic", "cube_linear", "cube_lanczos", "cube_cubic_odd", "lp_tiles", "lr_stereo", "rotated",
                                   "eac_tb_lanczos", "eac_mono_cubic", "cube_to_equirect", "equirect_to_equirect_rot", "cfg1", "barrel")
         for p in (0, 1)] + [("full", "cfg2", 0), ("full", "cfg2", 1), ("full", "cfg4", 1)]


@pytest.mark.parametrize("group,name,plane", CASES)
def test_gather_plan_invariants(group, name, plane):
    case = (SMALL if group == "small" else FULL)[name]
    ctx, hp, iw, ih = _plan(case, plane)
    k = hp.kernel_size
    g = hp.gather_plan()
    s = hp.samples.astype(np.int64)  # [mapH][mapW][2] = {col0, row0 << 10 | phase}, row-major
    mh, mw = s.shape[:2]
    _check_full_records(g, s, k)

    jobs = g["jobs"]
    staged_plan = k >= 2 and ctx.output_layout not in (t360.LAYOUT_BARREL, t360.LAYOUT_BARREL_SPLIT)
    assert (jobs is not None) == staged_plan
    if jobs is None:
        return
    cnt = g["counts"]
    kinds = (jobs[:, 1] >> KIND_SHIFT) & 15
    assert list(kinds) == sorted(kinds, key=lambda v: ORDER[int(v)]), "launch order: general, seam, class 1, share (with stays), share, class 0"
    by_kind = np.bincount(kinds, minlength=8)
    assert (by_kind[:3] == [cnt["class0"], cnt["class1"], cnt["general"]]).all() and by_kind[3] + by_kind[4] == cnt["share"] and by_kind[7] == cnt["seam"]
    assert ((jobs[:, 1] >> PLANE_SHIFT) == 0).all()
    if k < 4:
        assert cnt["share"] == 0
    wimg = WeightImage(k)
    compact = g["compact"]
    produced = np.zeros((mh, mw), np.int32)
    next_offset = 0
    for ox, oy, boxxy, rec_off in jobs:
        kind, y0 = (oy >> KIND_SHIFT) & 15, oy & ROW_MASK
        bx, by, variant = boxxy & 0xFFF0, boxxy >> 16, boxxy & 15  # variant: which of the class's box heights is loaded
        quad, ox = (ox & 7) - 1, ox & ~7  # 16 x 16 quadrant of the tile this job covers (-1: all of it)
        assert ox % 32 == 0 and y0 % 32 == 0 and quad < 4 and (quad < 0 or kind == CLASS0)
        if kind == GENERAL:
            assert boxxy == 0 and rec_off == 0  # taps through L1 from the full records: nothing for the host to promise
            produced[y0:y0 + 32, ox:ox + 32] += 1
            continue
        pitch, bh = box_w(kind), box_h(k, kind)
        assert bx % 16 == 0 and rec_off == next_offset, "records are laid out in launch order, 16-byte units"
        assert variant < 3 and (kind != CLASS1 or variant == 0)
        bh = box_variant_rows(k, kind, variant)  # what the kernel loads: every window must lie inside it (checked below) ...
        lower = box_variant_rows(k, kind, variant + 1) if variant < 2 and kind != CLASS1 else 0  # ... but not inside the next lower box
        rows_used = 0
        if kind in (SHARE, SHARE_STAY):
            R = share_rows(k)
            sh, nwords = 4 * R, R // 8 * 128 + 32  # job height; 32-bit words per warp
            assert ox % 64 == 0 and y0 % sh == 0 and ox + 64 <= mw and y0 + sh <= mh
            words = compact[rec_off * 4:rec_off * 4 + 8 * nwords].reshape(8, nwords).astype(np.int64)
            next_offset += 8 * nwords // 4
            for w in range(8):
                wx, wy = w & 1, w >> 1
                px = words[w, :R // 8 * 128].reshape(R // 8, 32, 4)  # [block of 8 rows][lane][word]
                hdr = words[w, R // 8 * 128:]                       # [lane]
                col, off = hdr >> 27, hdr & 0x7FFF
                assert ((hdr & ((1 << 27) - 1)) == off).all() and (np.sort(col) == np.arange(32)).all()
                rec = np.stack([px[j >> 3, :, (j >> 1) & 3] >> (16 * (j & 1)) & 0xFFFF for j in range(R)], axis=1)  # [lane][row]
                field = rec & SLOT_MASK
                if kind == SHARE:  # bit 0: the window moves two source rows instead of one
                    d = (rec & 1) + 1
                    d[:, 0] -= 1
                    assert ((rec & ~(SLOT_MASK | 1)) == 0).all()
                else:              # bits 0-1: it moves 0, 1 or 2 rows
                    d = rec & 3
                    assert ((rec & ~(SLOT_MASK | 3)) == 0).all() and (d <= 2).all() and (d == 0).any()
                assert (d[:, 0] == 0).all(), "the first record of a column carries no step"
                row0 = by + off[:, None] // pitch + np.cumsum(d, axis=1)
                col0 = bx + off % pitch
                want = s[y0 + wy * R:y0 + wy * R + R, ox + wx * 32:ox + wx * 32 + 32]  # [row][column]
                got_rows = np.empty((R, 32), np.int64); got_rows[:, col] = row0.T
                got_cols = np.empty(32, np.int64); got_cols[col] = col0
                got_field = np.empty((R, 32), np.int64); got_field[:, col] = field.T
                assert (got_rows == want[..., 1] >> 10).all() and (got_cols[None, :] == want[..., 0]).all()
                wimg.check(got_field.ravel(), (want[..., 1] & 1023).ravel())
                # the whole column's windows stay inside the box and the plane
                assert (col0 - bx + k <= pitch).all() and (row0[:, -1] - by + k <= bh).all()
                assert (col0 >= 0).all() and (col0 + k <= iw).all() and (row0 >= 0).all() and (row0 + k <= ih).all()
                rows_used = max(rows_used, int(row0.max()) - by + k)
            assert rows_used > lower, "the job names the lowest box that holds its windows"
            produced[y0:y0 + sh, ox:ox + 64] += 1
            continue
        # 32 x 32 jobs: class 0 (also one quadrant of a tile), class 1, seam.  Warp w, word j = one pixel of the 8 x 4
        # patch at rows 4w .., columns 8j ..
        if quad < 0:
            words = compact[rec_off * 4:rec_off * 4 + 8 * 128].reshape(8, 32, 4).astype(np.int64)  # [warp][lane][step]
            next_offset += 8 * 128 // 4
        else:
            # a quadrant job stores its four live warps x two live steps only; the kernel takes every other word as "skip"
            words = np.broadcast_to(((np.arange(32, dtype=np.int64) << 16) | SKIP)[None, :, None], (8, 32, 4)).copy()
            live_words = compact[rec_off * 4:rec_off * 4 + 4 * 64].reshape(4, 32, 2).astype(np.int64)
            words[4 * (quad >> 1):4 * (quad >> 1) + 4, :, 2 * (quad & 1):2 * (quad & 1) + 2] = live_words
            next_offset += 4 * 64 // 4
        off, pos, field, skip = words & 0x7FFF, (words >> 16) & 31, (words >> 17) & SLOT_MASK, (words & SKIP) != 0
        assert (np.sort(pos, axis=1) == np.arange(32)[None, :, None]).all(), "a patch holds every position once (also the skipped ones)"
        x = ox + 8 * np.arange(4)[None, None, :] + (pos & 7)
        y = y0 + 4 * np.arange(8)[:, None, None] + (pos >> 3)
        live = (x < mw) & (y < mh)
        if quad >= 0:
            live &= (x >= ox + 16 * (quad & 1)) & (x < ox + 16 * (quad & 1) + 16) & (y >= y0 + 16 * (quad >> 1)) & (y < y0 + 16 * (quad >> 1) + 16)
        assert (skip == ~live).all(), "exactly the pixels outside the plane / the quadrant are skipped"
        xs, ys, offs, fields = x[live], y[live], off[live], field[live]
        want = s[ys, xs]
        row0, col0 = by + offs // pitch, bx + offs % pitch
        assert (row0 == want[..., 1] >> 10).all()
        wimg.check(fields, want[..., 1] & 1023)
        assert (offs % pitch + k <= pitch).all() and (offs // pitch + k <= bh).all() and (row0 >= 0).all() and (row0 + k <= ih).all()
        assert int((offs // pitch).max()) + k > lower, "the job names the lowest box that holds its windows"
        if kind == SEAM:
            assert iw % 16 == 0 and bx < iw < bx + pitch, "the box of a seam tile wraps around the border"
            assert (col0 % iw == want[..., 0] % iw).all()
        else:
            assert (col0 == want[..., 0]).all() and (col0 >= 0).all() and (col0 + k <= iw).all(), "a staged tile never needs BORDER_WRAP"
        np.add.at(produced, (ys, xs), 1)
    assert (produced == 1).all(), "every output pixel belongs to exactly one job"
    assert compact is None or next_offset * 4 == compact.size


def test_job_counts_of_the_headline_plan():
    """cfg2 (8K equirect -> 3840x2560 cubemap, bicubic): the numbers DESIGN.md quotes."""
    for plane, want in ((0, dict(class0=4760, class1=0, seam=112, general=120, share=3120)),
                        (1, dict(class0=1176, class1=0, seam=56, general=32, share=760))):
        _, hp, _, _ = _plan(FULL["cfg2"], plane)
        assert hp.gather_plan()["counts"] == want


def _weight_load_wavefronts(field, live=None):
    """Bank model of one 128-bit weight load of a warp (tests and profiles/bank_sim.py): a quarter-warp of 8 lanes per
    pass, a pass costs as many wavefronts as its fullest 16-byte bank group holds DISTINCT addresses."""
    total = 0
    group = (field >> 4) & 7
    for q in range(4):
        lanes = np.arange(q * 8, q * 8 + 8)
        if live is not None:
            lanes = lanes[live[lanes]]
        if lanes.size:
            total += max(len(set(field[lanes][group[lanes] == b])) for b in range(8))
    return total


def test_weight_bank_balance_of_the_headline_plan():
    """The point of the second copy of the cubic table and of the host's lane dealing (GroupMatcher: even bank groups, none
    empty; PassDealer: exact deal to quarter-warps): modelled wavefronts of a 128-bit weight load over the share jobs and
    the 32 x 32 tile jobs of the cfg2 luma plan.  4.0 is the floor; a plain table costs 6.3, two copies with a greedy
    deal cost 4.46 / 4.98."""
    _, hp, _, _ = _plan(FULL["cfg2"], 0)
    g = hp.gather_plan()
    jobs, compact = g["jobs"], g["compact"]
    kinds = (jobs[:, 1] >> KIND_SHIFT) & 15
    share = jobs[np.isin(kinds, (SHARE, SHARE_STAY))][::16]
    R = share_rows(4)
    nwords = R // 8 * 128 + 32
    total = n = 0
    for ox, oy, boxxy, rec_off in share:
        words = compact[rec_off * 4:rec_off * 4 + 8 * nwords].reshape(8, nwords).astype(np.int64)
        for w in range(8):
            px = words[w, :R // 8 * 128].reshape(R // 8, 32, 4)
            for j in range(R):
                field = (px[j >> 3, :, (j >> 1) & 3] >> (16 * (j & 1))) & SLOT_MASK
                total += _weight_load_wavefronts(field)
                n += 1
    assert total / n < 4.55, f"share jobs: {total / n:.2f} wavefronts per weight load"
    tiles = jobs[(kinds == CLASS0) & ((jobs[:, 0] & 7) == 0)][::16]
    total = n = 0
    for ox, oy, boxxy, rec_off in tiles:
        words = compact[rec_off * 4:rec_off * 4 + 8 * 128].reshape(8, 32, 4).astype(np.int64)
        for w in range(8):
            for j in range(4):
                live = (words[w, :, j] & SKIP) == 0
                if live.any():
                    total += _weight_load_wavefronts((words[w, :, j] >> 17) & SLOT_MASK, live)
                    n += 1
    assert total / n < 4.65, f"tile jobs: {total / n:.2f} wavefronts per weight load"


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


def test_lane_dealing_is_a_valid_and_exact_deal():
    """T360B200_dealLanes (csrc/gather_plan.cpp: GroupMatcher + PassDealer) on random and on adversarial warp steps: every
    pixel gets its own lane and a copy that exists; the wavefronts it reports are what the bank model counts for the deal
    (or fewer, when two pixels share a slot: one broadcast read); never worse than leaving the pixels where they are; and
    the floor of 4 whenever the bank groups can be filled evenly."""
    rng = np.random.default_rng(7)

    def model(k, phases, lane, copy):
        img_field = np.empty(32, np.int64)
        for ph, ln, cp in zip(phases, lane, copy):
            slot = (int(ph) & ~31) | ((int(ph) & 1) << 4) | ((int(ph) & 31) >> 1)  # kernels.cuh weightSlotOf
            pos = (slot & ~7) | ((slot + int(cp)) & 7) if cp else slot               # weightSlotInCopy
            img_field[ln] = (pos << 4) | (int(cp) << 14)
        return _weight_load_wavefronts(img_field)

    cases = [rng.integers(0, 1024, 32) for _ in range(300)]
    cases += [np.full(32, 37), np.arange(32) * 32 + 5, np.arange(32), (np.arange(32) % 4) * 2 + 64]  # one slot; one fracX; one fracY; four bank groups
    for smooth in range(100):  # an 8 x 4 patch of a smooth map: fracX and fracY drift slowly
        fx0, fy0, dx, dy = rng.integers(0, 32), rng.integers(0, 32), rng.uniform(-3, 3, 2), rng.uniform(-3, 3, 2)
        xs, ys = np.meshgrid(np.arange(8), np.arange(4))
        cases.append((((fy0 + xs * dy[0] + ys * dy[1]).astype(int) & 31) << 5 | ((fx0 + xs * dx[0] + ys * dx[1]).astype(int) & 31)).ravel())
    for phases in cases:
        w, lane, copy = t360.deal_lanes(t360.CUBIC, phases)
        assert sorted(lane.tolist()) == list(range(32)) and set(copy.tolist()) <= {0, 1}
        got = model(4, phases, lane, copy)
        assert 4 <= got <= w, (got, w)
        assert w <= model(4, phases, np.arange(32), np.zeros(32, int)) or len(set(phases.tolist())) < 32
        groups = np.bincount((phases.astype(int) & 31) >> 1 & 7, minlength=8)
        if (groups == 4).all() and len(set(phases.tolist())) == 32:
            assert w == 4
    w, lane, copy = t360.deal_lanes(t360.CUBIC, np.arange(20))  # fewer than 32 pixels: identity
    assert w == 0 and (lane == np.arange(20)).all() and (copy == 0).all()
    w, lane, copy = t360.deal_lanes(t360.LANCZOS4, rng.integers(0, 1024, 32))  # one copy: lanes only
    assert sorted(lane.tolist()) == list(range(32)) and (copy == 0).all()
