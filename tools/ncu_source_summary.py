"""Summarise an `ncu --page source --csv --print-source sass` export: stall-reason totals and the hottest SASS regions.
usage: python tools/ncu_source_summary.py <file.csv> [top_n]"""
import csv
import sys


def load(path):
    with open(path) as f:
        r = csv.reader(f)
        name = next(r)[1]
        hdr = next(r)
        rows = [dict(zip(hdr, x)) for x in r if len(x) >= len(hdr) - 1]
    return name, hdr, rows


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    name, hdr, rows = load(path)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {s: sum(int(r[s] or 0) for r in rows) for s in stalls}
    allsamp = sum(int(r["# Samples"] or 0) for r in rows)
    instr = sum(int(r["Instructions Executed"] or 0) for r in rows)
    print(f"kernel: {name}\nSASS lines: {len(rows)}  samples: {allsamp}  warp-instructions executed: {instr}")
    print("stall totals:", ", ".join(f"{k[6:]}={v} ({100.0 * v / max(allsamp, 1):.1f}%)" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    # windows of 16 SASS lines ranked by samples
    W = 16
    wins = []
    for i in range(0, len(rows), W):
        chunk = rows[i:i + W]
        wins.append((sum(int(r["# Samples"] or 0) for r in chunk), i, chunk))
    wins.sort(key=lambda w: -w[0])
    print(f"\nhottest {top} windows of {W} SASS instructions:")
    for s, i, chunk in wins[:top]:
        ops = {}
        for r in chunk:
            op = r["Source"].split()[0] if r["Source"].split() else "?"
            if op.startswith("@"):
                op = r["Source"].split()[1]
            ops[op] = ops.get(op, 0) + 1
        dom = {k: sum(int(r[k] or 0) for r in chunk) for k in stalls}
        d = sorted(dom.items(), key=lambda kv: -kv[1])[:3]
        print(f"  line {i:5d}: samples {s:6d} ({100.0 * s / max(allsamp, 1):4.1f}%) exec/line {int(chunk[0]['Instructions Executed'] or 0):8d}  ops {dict(sorted(ops.items(), key=lambda kv: -kv[1])[:4])}  stalls {[(k[6:], v) for k, v in d]}")


if __name__ == "__main__":
    main()
