#!/usr/bin/env python
"""Offline model of the shared-memory bank behaviour of the staged gather (gatherFrameKernel; no GPU needed).

Takes the host plan of a BASELINE config (through libTransform360.so's host-plan API) and counts, per warp-step,
the shared-memory wavefronts of (a) the aligned 32-bit window ("tap") loads for a given staging pitch and
lane->pixel mapping, and (b) the 128-bit weight loads for a given slot hash, with and without dealing the pixels of a
32-pixel row segment to lanes by bank group.  The model (a pass = 32 lanes for 32-bit loads, 8 lanes for 128-bit
loads; cost = max number of distinct addresses falling into one bank / bank group) reproduced ncu's per-instruction
"L1 Wavefronts Shared" to ~1 % (1.57 vs 1.59 per tap load, 9.90 vs 9.87 per weight load before lane ordering,
7.0 vs 6.8-6.9 after), which is what made it usable for choosing the pitch, the slot hash and the lane order.

    python profiles/bank_sim.py [cfg2|cfg4]
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import transform360_b200 as t360  # noqa: E402
from tests.golden.cases import FULL  # noqa: E402


def max_distinct_per_bank(addr, nbanks):
    bank = addr % nbanks
    out = np.zeros(addr.shape[0], np.int64)
    for b in range(nbanks):
        a = np.sort(np.where(bank == b, addr, -1), axis=1)
        out = np.maximum(out, ((a[:, 1:] != a[:, :-1]) & (a[:, 1:] >= 0)).sum(axis=1) + (a[:, 0] >= 0))
    return out


def weight_wavefronts(phase_lanes, slot):
    tot = np.zeros(phase_lanes.shape[0], np.int64)
    for q in range(4):
        p = phase_lanes[:, q * 8:(q + 1) * 8]
        s = slot(p)
        mx = np.zeros(p.shape[0], np.int64)
        for g in range(8):
            a = np.sort(np.where(s == g, p, -1), axis=1)
            mx = np.maximum(mx, ((a[:, 1:] != a[:, :-1]) & (a[:, 1:] >= 0)).sum(axis=1) + (a[:, 0] >= 0))
        tot += mx
    return tot


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    case = FULL[name]
    ctx = t360.make_context(**case["ov"])
    (iw, ih), (ow, oh) = case["inp"], case["out"]
    hp = t360.HostPlan(ctx, iw, ih, ow, oh)
    s, k = hp.samples, hp.kernel_size
    col0, row0, phase = s[..., 0].astype(np.int64), (s[..., 1] >> 10).astype(np.int64), (s[..., 1] & 1023).astype(np.int64)
    rng = np.random.default_rng(0)
    tiles = [(tx, ty) for ty in range(oh // 32) for tx in range(ow // 32)]
    sel = [tiles[i] for i in rng.choice(len(tiles), 1200, replace=False)]
    print(f"{name}: kernel size {k}; tap loads (wavefronts per 32-bit load, lanes along x):")
    for pitch in (144, 176, 192, 208, 240):
        tot = n = 0
        for tx, ty in sel:
            c, r = col0[ty * 32:ty * 32 + 32, tx * 32:tx * 32 + 32], row0[ty * 32:ty * 32 + 32, tx * 32:tx * 32 + 32]
            if c.min() < 0 or r.min() < 0:
                continue
            bx, by = c.min() // 16 * 16, r.min()
            if c.max() + k - bx > pitch or r.max() + k - by > 64:
                continue
            off = (r - by) * pitch + (c - bx)
            for rr in range(k):
                a = (off + rr * pitch) // 4
                tot += max_distinct_per_bank(a, 32).sum() + max_distinct_per_bank(a + 1, 32).sum()
            n += 32 * k * 2
        print(f"  pitch {pitch:3d} B ({pitch // 4 % 32:2d} words mod 32): {tot / max(n, 1):.2f}")
    seg = phase[rng.choice(oh, 300, replace=False)].reshape(-1, 32)
    lane_of_i = (np.arange(32) % 4) * 8 + np.arange(32) // 4
    print("weight loads (wavefronts per 128-bit load): identity order / dealt by bank group")
    for label, slot in (("fracX & 7", lambda a: a & 7), ("(fracX + fracY) & 7", lambda a: (a + (a >> 5)) & 7),
                        ("fracX >> 2   [used]", lambda a: (a & 31) >> 2), ("(fracX >> 1) & 7", lambda a: ((a & 31) >> 1) & 7)):
        order = np.argsort(slot(seg) * 1024 + seg, axis=1, kind="stable")
        dealt = np.empty_like(seg)
        dealt[:, lane_of_i] = np.take_along_axis(seg, order, axis=1)
        print(f"  {label:22s} {weight_wavefronts(seg, slot).mean():5.2f} / {weight_wavefronts(dealt, slot).mean():5.2f}")


if __name__ == "__main__":
    main()
