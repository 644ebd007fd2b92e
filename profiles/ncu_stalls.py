"""Per-SASS-instruction stall samples of one reason from an .ncu-rep (source page): where the warps wait.
    python profiles/ncu_stalls.py REP REASON [top]      REASON e.g. stall_no_inst, stall_barrier, stall_long_sb
Prints the totals of every reason, then the `top` instructions (address offset, samples of REASON, executed count, SASS)."""
import csv
import subprocess
import sys

rep, reason = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = None
data = []
for r in rows:
    if len(r) > 10 and r[0] == "Address":
        hdr = r
        data = []          # keep the LAST kernel of the report
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
col = {n: i for i, n in enumerate(hdr)}
reasons = [n for n in hdr if n.startswith("stall_") and "(" not in n]
tot = {n: sum(int(d[col[n]] or 0) for d in data) for n in reasons}
allsum = sum(tot.values())
print("samples by reason:", ", ".join("%s %.1f%%" % (n[6:], 100.0 * v / allsum) for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
base = int(data[0][col["Address"]], 16)
best = sorted(data, key=lambda d: -int(d[col[reason]] or 0))[:top]
for d in best:
    print("+%05x  %6s %s  exec %9s  %s" % (int(d[col["Address"]], 16) - base, d[col[reason]], reason[6:], d[col["Instructions Executed"]], d[col["Source"]].strip()[:90]))
