"""one-line summary of a bench.py JSON line (file argument or stdin)"""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
det = d.get("detail", {})
print("n_gpus %s  value %.1f M/s  e2e %.1f M/s  ms/step %.2f  e2e ms/step %s  map_kernel_ms %s  em_loop_ms %s  host_wall %s  classes %s" % (
    d.get("n_gpus"), d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], d["e2e"].get("ms_per_step"),
    det.get("map_kernel_ms_per_step"), det.get("em_loop_ms_per_step"), det.get("host_wall_ms_per_step"), det.get("n_classes")))
