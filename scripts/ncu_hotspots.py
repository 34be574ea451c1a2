"""Per-kernel stall-sample hot spots from an .ncu-rep captured with --set full (SASS level; runs without a GPU):
    python scripts/ncu_hotspots.py <rep> [top_n]
For every kernel: total warp-stall samples, the stall reasons ranked, and the top instructions by samples."""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": next(csv.reader([line]))[1], "lines": []}
        blocks.append(cur)
    elif cur is not None:
        cur["lines"].append(line)
for b in blocks:
    rows = list(csv.reader(io.StringIO("\n".join(b["lines"]))))
    if len(rows) < 2:
        continue
    hdr = rows[0]
    si = hdr.index("Warp Stall Sampling (All Samples)")
    ii = hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = 0
    reasons = collections.Counter()
    insts = []
    for k, r in enumerate(rows[1:]):
        if len(r) <= si or not r[si].isdigit():
            continue
        s = int(r[si])
        total += s
        for i, h in stall_cols:
            if r[i].isdigit():
                reasons[h] += int(r[i])
        insts.append((s, k, r[1].strip(), r[ii], [(h, int(r[i])) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0]))
    print("=== ", b["name"][:120])
    print(f"    samples {total}; instructions {len(insts)}")
    print("    stall reasons: " + ", ".join(f"{h[6:]} {100 * c / max(total, 1):.0f}%" for h, c in reasons.most_common(6)))
    for s, k, text, ex, why in sorted(insts, reverse=True)[:top_n]:
        w = ", ".join(f"{h[6:]} {c}" for h, c in sorted(why, key=lambda t: -t[1])[:3])
        print(f"    {100 * s / max(total, 1):5.1f}%  #{k:<5d} {text[:58]:58s} exec {ex:>8s}  [{w}]")
