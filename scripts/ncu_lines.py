"""Executed instructions and stall samples per source line of one captured kernel.
    python scripts/ncu_lines.py x.ncu-rep attempts_per_launch [top]"""
import csv, io, subprocess, sys
rep, att = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, agg, hdr = None, [], None
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci, cs = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        continue
    if hdr and r[0] not in ("", "Function Name") and len(r) > ci:
        try:
            agg.append((float(r[ci]) * 32 / att, float(r[cs] or 0), cur, r[0], r[1].strip()[:120]))
        except ValueError:
            pass
tot = sum(a[0] for a in agg); ts = sum(a[1] for a in agg)
print("total %.1f instr/attempt over %d lines" % (tot, len(agg)))
for n, st, f, ln, src in sorted(agg, reverse=True)[:top]:
    print("%6.1f %5.1f%%  %s:%s  %s" % (n, 100 * st / max(1, ts), f, ln, src))
