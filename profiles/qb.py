"""Run bench.py with the given args and print the few numbers used while tuning."""
import json
import subprocess
import sys
out = subprocess.run([sys.executable, "bench.py", "--no-cpu", "--no-cloud"] + sys.argv[1:], capture_output=True, text=True).stdout
d = json.loads(out.strip().splitlines()[-1])
print(" ".join(sys.argv[1:]), "| ms/step %.4f kernel_ms %.4f frac %.3f Mpix/s %.0f e2e %.0f" % (
    d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["value"], d["e2e"]["value"]))
