"""Print the headline fields of a bench.py JSON line.  usage: bench_summary.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
print("value %.1f %s | e2e %.1f | ms/step %.0f | launches %d" % (d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d.get("gpu_launches", 0)))
if "roofline" in d:
    r = d["roofline"]
    print("roofline %s: %.2f of %.2f %s = %.3f ; share of step %.3f" % (r["kernel"], r["achieved"], r["peak"], r["unit"], r["frac"], r["share_of_step"]))
if "desync" in d:
    print("desync %.1f" % d["desync"]["value"])
if "cpu_baseline" in d:
    print("cpu_baseline %.2f (%s)" % (d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"]))
print("parity", d.get("parity"), "stage_ms", d.get("stage_ms"), "clocks", d.get("clocks"))
