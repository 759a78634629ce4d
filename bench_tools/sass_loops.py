"""Instruction mix of the loops of one kernel in a cuobjdump -sass listing.

    cuobjdump -sass -fun <mangled> lib.so | python bench_tools/sass_loops.py [min_len]

A loop = a backward branch; for every loop longer than min_len instructions prints its address range, length and
per-class instruction counts (the numbers quoted in DESIGN.md / profiles/*_sass_*.txt)."""
import re
import sys
from collections import Counter

CLASSES = [
    ("FADD2/FFMA2/FMUL2", r"^(FADD2|FFMA2|FMUL2)"),
    ("FADD", r"^FADD\b"), ("FFMA", r"^FFMA\b"), ("FMUL", r"^FMUL\b"), ("MUFU", r"^MUFU"),
    ("SHFL", r"^SHFL"), ("LDS", r"^LDS"), ("STS", r"^STS"), ("LDG", r"^LDG"), ("STG", r"^STG"),
    ("LDL/STL", r"^(LDL|STL)"), ("MOV/SEL/PRMT", r"^(MOV|SEL|FSEL|PRMT|IMAD\.MOV|UMOV)"),
    ("BRA/BSSY/BSYNC/WARPSYNC", r"^(BRA|BSSY|BSYNC|WARPSYNC|CALL|RET|EXIT|BAR|NANOSLEEP|YIELD)"),
]


def main():
    min_len = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    ins = []
    for line in sys.stdin:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        txt = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
        ins.append((addr, txt))
    idx = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.match(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in idx:
                loops.append((idx[tgt], i))
    print(f"{len(ins)} instructions, {len(loops)} backward branches")
    for lo, hi in loops:
        n = hi - lo + 1
        if n < min_len:
            continue
        c = Counter()
        for _, t in ins[lo:hi + 1]:
            op = t.split()[0]
            for name, pat in CLASSES:
                if re.match(pat, op):
                    c[name] += 1
                    break
            else:
                c["other(int/addr/pred)"] += 1
        print(f"loop 0x{ins[lo][0]:x}-0x{ins[hi][0]:x}: {n} instr  " + "  ".join(f"{k}={v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
