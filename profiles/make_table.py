#!/usr/bin/env python
"""Markdown tables from ONE bench.py line (round 2: the default run carries every workload):
    python profiles/make_table.py profiles/r02_bench_default.json [profiles/r02_scaling_1gpu.json ...2gpu.json ...]"""
import json, sys


def load(path):
    for line in open(path):
        line = line.strip()
        if line.startswith("{"):
            try:
                return json.loads(line)
            except Exception:
                pass
    raise SystemExit("no JSON line in " + path)


d = load(sys.argv[1])
print("| workload | envs/GPU | env-steps/s (device) | µs / batch step | protocol (R batches x C chains) | µs isolated (flush + event pair per step) | B / env-step (this layout) | measured DRAM B / env-step (ncu) | frac of %.0f GB/s (isolated) | e2e env-steps/s (host buffers) | e2e + obs node columns | CPU port env-steps/s (cores) | Python reference, 1 core / all cores | e2e ÷ CPU port |"
      % d["roofline"]["peak"])
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for w in d["workloads"]:
    r, c, py = w["roofline"], w.get("cpu_baseline") or {}, w.get("cpu_reference_python") or {}
    tr = r.get("traffic")
    pyref = "-"
    if py:
        a, b = py.get("one_core_step_only_steps_per_s"), py.get("all_cores_steps_per_s")
        pyref = "%s / %s (%s)" % ("%.3g" % a if a else "n/a", "%.3g" % b if b else "n/a", py.get("cores"))
    st = w.get("streaming")
    proto = ("streaming %d x %d" % (st["replicas"], st["chunks"])) if st else "isolated"
    iso = w.get("isolated") or {}
    eo = w.get("e2e_obs") or {}
    print("| %s | %d | %.3g | %.1f | %s | %s | %.0f | %s | %.3f (%s) | %.3g | %s | %s | %s | %s |" % (
        w["name"], w["envs_per_gpu"], w["value"], 1e3 * w["ms_per_step"], proto, ("%.1f" % (1e3 * iso["ms_per_step"])) if iso else "-",
        r["bytes_per_env_step"], ("%.0f" % (tr / w["envs_per_gpu"])) if tr else "-", r["frac"],
        ("%.3f" % r["frac_isolated"]) if r.get("frac_isolated") is not None else "-", w["e2e"]["value"],
        ("%.3g" % eo["value"]) if eo else "-",
        ("%.3g (%s)" % (c["value"], c["cores"])) if c else "-", pyref, ("%.0fx" % (w["e2e"]["value"] / c["value"])) if c else "-"))
if len(sys.argv) > 2:
    print()
    print("| GPUs | cfg2 env-steps/s (weak, 65,536 envs/GPU) | cfg2 e2e | cfg5 Multicast (524,288 envs total) | cfg5 DistributionCenter (524,288 total) | cfg5 all 1,048,576 envs, one step of both | envs per GPU and kind |")
    print("|---|---|---|---|---|---|---|")
    base = None
    for p in sys.argv[2:]:
        s = load(p)
        c5 = s["cfg5_strong_scaling"]
        row = (s["value"], s["e2e"]["value"], c5["cfg5_multicast"]["value"], c5["cfg5_distcenter"]["value"], c5["combined"]["value"])
        if base is None:
            base = row
        n = s["n_gpus"]
        print("| %d | %s | %d |" % (n, " | ".join("%.3g (%.2f)" % (v, v / (b * (n if i < 2 else n))) for i, (v, b) in enumerate(zip(row, base))),
                                   c5["cfg5_multicast"]["envs_per_gpu"]))
    print("\n(in parentheses: efficiency = value / (N x the 1-GPU value))")
