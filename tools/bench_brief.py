"""Prints the headline fields of a bench.py JSON line (development aid)."""
import json
import sys

for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.startswith("{")]
    if not lines:
        print(path, "no JSON line")
        continue
    d = json.loads(lines[-1])
    print(path)
    print("  value %.4g %s  ms/step %.2f  n_gpus %d  launches %s" % (d["value"], d["unit"], d["ms_per_step"], d["n_gpus"], d.get("gpu_launches")))
    print("  stages", {k: round(v, 3) for k, v in d.get("stages_ms_per_step", {}).items()})
    print("  knn", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.get("knn", {}).items()})
    if d.get("roofline"):
        print("  roofline %.3f (%s %.1f of %.1f %s)" % (d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["achieved"], d["roofline"]["peak"], d["roofline"]["unit"]))
    for k, v in (d.get("roofline_stages") or {}).items():
        print("  stage %-10s %.3f of HBM  (%.3f ms, %.0f GB/s)" % (k, v["frac"], v["ms"], v["achieved"]))
    if d.get("e2e"):
        print("  e2e %.4g  ms/step %.2f" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))
    if d.get("diffusion"):
        print("  diffusion", d["diffusion"])
    if d.get("per_rank"):
        for k, v in d["per_rank"].items():
            print("  per-rank %-12s %s" % (k, v))
    print("  verify", d.get("verify"))
    print("  clocks", d.get("clocks"))
    if d.get("cpu_baseline"):
        print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
