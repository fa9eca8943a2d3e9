"""Print the headline fields of a bench.py JSON line (development aid)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g  e2e %.4g  ms/frame %.4f (pair %.4f build %.4f)  roofline.frac %.3f  launches %s" % (
    d["value"], d["e2e"]["value"], d.get("ms_per_frame", 0), d.get("pair_kernel_ms_per_frame", 0),
    d.get("build_ms_per_frame", 0), d["roofline"]["frac"], d.get("gpu_launches")))
print("e2e steps", d["e2e"].get("step_ms"), "clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
