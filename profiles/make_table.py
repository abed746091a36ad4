#!/usr/bin/env python
"""Markdown table from a jsonl of bench.py lines:  python profiles/make_table.py profiles/r01_all_workloads.jsonl"""
import json, sys
rows = []
for line in open(sys.argv[1]):
    line = line.strip()
    if not line:
        continue
    try:
        d = json.loads(line)
    except Exception:
        continue
    if "roofline" not in d:
        continue
    r, c = d["roofline"], d.get("cpu_baseline") or {}
    rows.append((d["config"]["name"], d["config"]["envs_per_gpu"], d["value"], 1e3 * d["ms_per_step"], r["bytes_per_env_step"],
                 r["achieved"], r["frac"], d["e2e"]["value"], c.get("value"), c.get("cores")))
print("| workload | envs/GPU | env-steps/s (device) | µs / batch step | B / env-step (this layout) | achieved GB/s | frac of 6545 GB/s | e2e env-steps/s (host buffers) | CPU port env-steps/s (cores) | e2e ÷ CPU |")
print("|---|---|---|---|---|---|---|---|---|---|")
for n, B, v, us, by, gb, fr, e2e, cpu, cores in rows:
    print("| %s | %d | %.3g | %.1f | %.0f | %.0f | %.3f | %.3g | %s | %s |" % (
        n, B, v, us, by, gb, fr, e2e, ("%.3g (%s)" % (cpu, cores)) if cpu else "-", ("%.0fx" % (e2e / cpu)) if cpu else "-"))
