"""Hot regions of an ncu source page (SASS view): runs of consecutive instructions with the same
execution count, with their share of executed instructions and of stall samples.
    python profiles/ncu_hot.py report.ncu-rep [min_share]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
end_i = next((i for i, r in enumerate(rows) if i > hdr_i and r and r[0] == "Kernel Name"), len(rows))
print(" ".join(rows[hdr_i - 1][:2])[:110])
rows = rows[:end_i]   # first kernel of the report only
iS, iE, iSmp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(r[iS].strip(), int(r[iE] or 0), int(r[iSmp] or 0)) for r in rows[hdr_i + 1:] if len(r) > iE]
tot_e = sum(d[1] for d in data)
tot_s = sum(d[2] for d in data)
print("instructions executed %d, samples %d, static %d" % (tot_e, tot_s, len(data)))
# opcode histogram weighted by executions
ops = {}
for s, e, smp in data:
    op = s.split()[0] if not s.startswith("@") else s.split()[1]
    op = op.split(".")[0]
    o = ops.setdefault(op, [0, 0])
    o[0] += e
    o[1] += smp
print("top opcodes by executed: " + ", ".join("%s %.1f%%(st %.1f%%)" % (k, 100.0 * v[0] / tot_e, 100.0 * v[1] / max(tot_s, 1))
                                          for k, v in sorted(ops.items(), key=lambda kv: -kv[1][0])[:24]))
if len(sys.argv) > 2 and sys.argv[2] == "list":
    for k, (s, e, smp) in enumerate(data):
        if e > 0.0005 * tot_e:
            print("%5d %9d %6d  %s" % (k, e, smp, s[:100]))
