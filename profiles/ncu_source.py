#!/usr/bin/env python
"""Per-instruction summary of an ncu `--page source --csv` export (SASS view):
    ncu -i prof.ncu-rep --page source --csv > src.csv ; python profiles/ncu_source.py src.csv [top]
Prints totals, the opcode histogram (by executed warp instructions), shared-memory wavefronts per opcode class and
the instructions with the most stall samples."""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    hdr = rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    f = lambda r, n: float(r[col[n]] or 0)
    tot_inst = sum(f(r, "Instructions Executed") for r in body)
    tot_samp = sum(f(r, "# Samples") for r in body)
    tot_wf = sum(f(r, "L1 Wavefronts Shared") for r in body)
    print(f"warp instructions {tot_inst / 1e6:.2f} M, stall samples {tot_samp:.0f}, shared wavefronts {tot_wf / 1e6:.2f} M")
    by_op = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0])
    for r in body:
        toks = r[col["Source"]].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        b = by_op[op]
        b[0] += f(r, "Instructions Executed"); b[1] += f(r, "# Samples"); b[2] += f(r, "L1 Wavefronts Shared"); b[3] += f(r, "L1 Wavefronts Shared Ideal")
    print(f"{'opcode':28s} {'Minst':>8s} {'%inst':>6s} {'%samples':>8s} {'wavefronts/inst':>15s} {'ideal':>6s}")
    for op, b in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:top]:
        wf = f"{b[2] / b[0]:.2f}" if b[2] and b[0] else ""
        ideal = f"{b[3] / b[0]:.2f}" if b[3] and b[0] else ""
        print(f"{op:28s} {b[0] / 1e6:8.3f} {100 * b[0] / tot_inst:6.1f} {100 * b[1] / max(tot_samp, 1):8.1f} {wf:>15s} {ideal:>6s}")
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    tot = {n: sum(f(r, n) for r in body) for n in stall_cols}
    print("stall reasons (all samples):", ", ".join(f"{n[6:]} {100 * v / max(tot_samp, 1):.1f}%" for n, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
    print("instructions with most samples:")
    for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:top]:
        reasons = sorted(((f(r, n), n[6:]) for n in stall_cols), reverse=True)[:2]
        print(f"  {r[col['Address']][-5:]} {100 * f(r, '# Samples') / max(tot_samp, 1):5.2f}%  {f(r, 'Instructions Executed') / 1e3:9.1f}k  {r[col['Source']][:70]:70s} {reasons[0][1]}/{reasons[1][1]}")


if __name__ == "__main__":
    main()
