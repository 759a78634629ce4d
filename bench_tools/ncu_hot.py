"""Top stall sites of an `ncu --page source --csv` export (bench_tools/ncu_capture.sh):
    python bench_tools/ncu_hot.py gpurun_out/<tag>_source.csv.gz [N]
Prints the N instructions with the most warp-stall samples, their dominant stall reason and execution count."""
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(gzip.open(path, "rt") if path.endswith(".gz") else open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    base = None
    tot = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        base = addr if base is None else base
        s = int(r[ix["# Samples"]] or 0)
        tot += s
        st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
        data.append((s, addr - base, r[1].strip(), int(r[ix["Instructions Executed"]] or 0), st))
    print(f"total samples {tot}")
    agg = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        for c in stall_cols:
            agg[c] = agg.get(c, 0) + int(r[ix[c]] or 0)
    print("by reason:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for s, off, txt, ex, st in sorted(data, reverse=True)[:n]:
        print(f"{s:6d} {100.0 * s / tot:5.1f}%  0x{off:05x}  x{ex:<7d} {txt[:64]:64s} {st[0][1][6:]}={st[0][0]} {st[1][1][6:]}={st[1][0]}")


if __name__ == "__main__":
    main()
