"""Per-opcode executed-instruction mix and stall attribution from an ncu source page
(ncu -i X.ncu-rep --page source --csv > X_source.csv).  Usage: python profiles/opmix.py X_source.csv <units>
where <units> = walker-steps (or other work units) the profiled launch processed; prints warp-instructions per
32 units -- i.e. per warp-step for one-walker-per-lane kernels -- and the issue-slot model of DESIGN.md section 4
(IMAD.WIDE / IMAD.HI occupy the issue port for 4.1 cycles each, profiles/microbench/pipes_b200.txt)."""
import collections
import csv
import re
import sys


def main(path, units):
    rows = list(csv.reader(open(path)))
    kernels = []
    hdr = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernels.append([r[1], []])
        elif r and r[0] == "Address":
            hdr = r
        elif kernels and hdr and len(r) >= len(hdr):
            kernels[-1][1].append(r)
    for name, body in kernels:
        iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        stall_cols = {h: hdr.index(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
        ops, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
        for r in body:
            m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", r[iS])
            if not m:
                continue
            op = m.group(2)
            base = op.split(".")[0]
            key = base
            if base == "IMAD":
                key = "IMAD.WIDE" if "WIDE" in op else ("IMAD.HI" if ".HI" in op else "IMAD")
            elif base == "MUFU":
                key = op
            ops[key] += int(r[iE])
            samples[key] += int(r[iN])
            for h, i in stall_cols.items():
                stalls[h] += int(r[i])
        w = units / 32.0
        total = sum(ops.values())
        wide = ops["IMAD.WIDE"] + ops["IMAD.HI"]
        mufu = sum(v for k, v in ops.items() if k.startswith("MUFU"))
        print(f"kernel: {name[:120]}")
        print(f"  warp-instructions per 32 units: {total / w:.1f}  (IMAD.WIDE/HI {wide / w:.2f}, MUFU {mufu / w:.2f})")
        print(f"  issue-slot model: {(total - wide) / w + 4.1 * wide / w:.1f} cycles per 32 units per SMSP; "
              f"XU pipe {8 * mufu / w:.1f} cycles")
        for k, v in ops.most_common(16):
            print(f"    {k:12s} {v / w:8.2f}   samples {samples[k]}")
        print("  stall samples:", ", ".join(f"{k[6:]}={v}" for k, v in stalls.most_common(6)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]))
