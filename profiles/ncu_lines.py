"""Per-source-line share of executed instructions and stall samples from an ncu report captured with
--import-source on (the CUDA + SASS source page).
    python profiles/ncu_lines.py report.ncu-rep [kernel-regex] [top_n] [stall]   (last word: sort by stall samples)
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
by_stall = len(sys.argv) > 4 and sys.argv[4] == "stall"
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, lines = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        iE, iSm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr is not None and r[0].isdigit() and len(r) > iE:
        num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
        lines.append((cur_file, int(r[0]), r[1].strip(), num(r[iE]), num(r[iSm])))
tot = sum(l[3] for l in lines)
ts = sum(l[4] for l in lines)
print("instructions executed %d, samples %d" % (tot, ts))
for f, n, src, e, sm in sorted(lines, key=lambda l: -(l[4] if by_stall else l[3]))[:top_n]:
    print("%5.1f%% instr %5.1f%% stall  %s:%d  %s" % (100.0 * e / tot, 100.0 * sm / max(ts, 1), f, n, src[:110]))
