"""Dynamic opcode mix of one captured kernel: executed warp instructions per opcode, per attempt.
    python scripts/ncu_opmix.py x.ncu-rep attempts_per_launch [top]"""
import csv, io, subprocess, sys, collections
rep, att = sys.argv[1], float(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
i = next(k for k, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[i:]))))
hdr = rows[0]
ci, cs, cst = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
mix, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[1:]:
    if len(r) <= ci:
        continue
    op = r[cs].split()
    op = [o for o in op if not o.startswith("@")]
    name = op[0].rstrip(";") if op else "?"
    base = ".".join(name.split(".")[:2]) if name.startswith(("LDG", "STG", "MUFU", "IMAD", "LDS", "STS")) else name.split(".")[0]
    n = float(r[ci]) * 32 / att
    mix[base] += n
    stall[base] += float(r[cst] or 0)
    tot += n
print("total %.1f instr/attempt" % tot)
ts = sum(stall.values())
for k, v in mix.most_common(top):
    print("%-14s %7.2f   stall samples %4.1f%%" % (k, v, 100 * stall[k] / max(1, ts)))
