#!/usr/bin/env python
"""Renders README's "Results" section from the committed bench lines:
    python profiles/make_readme_results.py > /tmp/results.md     (then pasted between the markers of README.md)"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    for line in open(os.path.join(HERE, name)):
        line = line.strip()
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit("no JSON in " + name)


d = load("r02_bench_default.json")
ref = load("r02_bench_reference_arm.json")
scal = [p for p in ("r02_scaling_1gpu.json", "r02_scaling_2gpu.json", "r02_scaling_4gpu.json", "r02_scaling_8gpu.json") if os.path.exists(os.path.join(HERE, p))]
tab = subprocess.check_output([sys.executable, os.path.join(HERE, "make_table.py"), os.path.join(HERE, "r02_bench_default.json")] +
                              [os.path.join(HERE, p) for p in scal]).decode()
e = d["e2e"]
st = d.get("streaming") or {}
print("""Device numbers: everything resident, the uniform valid action is drawn in the step kernel (`ge_step_sampled`: sampler + step + new
mask + auto-reset).  **Streaming protocol** (the `value`): EXACTLY K steps in ONE timed region (one CUDA event pair, barrier + synchronize on
both sides), no flush inside; the steps rotate over R independent resident batches of B envs, R chosen so that one rotation touches more than
3 x L2 (every step finds its batch cold); one step of a batch = C sub-batch launches on C free-running stream chains, launched with
programmatic dependent launch.  **Isolated protocol** (round 1's, kept beside it): a 256 MiB write + read flushes L2 before every step, each
step between its own event pair (~6 µs of launch + event overhead inside every pair).  Workloads whose R batches do not fit in HBM (config 4) report
the isolated protocol only.  `e2e`: the step through `ge_step_host_compact` with pinned host buffers (actions in; reward / one flag byte /
float32 solution_cost / packed mask out; two slices, the write-back of one overlapping the kernel of the next; L2 flushed before every call); `e2e + obs`: the same call also rewrites
the observation's node columns `x[B, N, F]` on the device (`ge_batch.obs_x`).  CPU port: `oracle/graphenvs_oracle.c` with OpenMP over envs on
the GPU box's host threads.  Python reference: the unmodified `graph_envs` loop as recorded on the build container
(`profiles/r02_python_reference_cpu.json`).
""")
print("**Headline** (BASELINE config 2, LongestPath N=50 E=200 p=2, 65,536 envs, 1 B200): **%.3g env-steps/s** on the device (%.1f µs per batch "
      "step, streaming over %d batches x %d chains; %.2f of the measured %.0f GB/s on the %.0f bytes per env-step this layout must move, %.0f measured by "
      "ncu; isolated: %.1f µs, %.2f), **%.3g env-steps/s end to end** "
      "through the C ABI with host buffers (%.3g with the device sampler choosing the actions between calls), %.3g with the observation's node "
      "columns rewritten every step; reference arm (CPU port, %d threads): %.3g env-steps/s.\n" % (
          d["value"], 1e3 * d["ms_per_step"], st.get("replicas", 1), st.get("chunks", 1), d["roofline"]["frac"], d["roofline"]["peak"],
          d["roofline"]["bytes_per_env_step"], (d["roofline"]["traffic"] or 0) / d["config"]["envs_per_gpu"],
          1e3 * d["isolated"]["ms_per_step"], d["roofline"]["frac_isolated"], e["value"], e.get("value_with_device_policy_between_calls") or float("nan"),
          d["e2e_obs"]["value"], ref["cpu_baseline"]["cores"], ref["value"]))
print(tab)
f = d["feature_extraction_us_per_env"]
print("`feature_extraction.generate_features` on the device (µs per env): " + ", ".join(
    "%s N=%d E=%d: %.2f" % (k, v["n_nodes"], v["n_edges"], v["us_per_env"]) for k, v in f.items()) +
    "  (round 1: 190 at TSP N=200 dense, 34 at MIS N=200, 73 at N=500, 0.55 at cfg2).\n")
t = d["instance_turnover"]
for k, v in t.items():
    a, b = v["state_only_auto_reset_same_graph"], v["pool_turnover"]
    print("Instance turnover, %s: %.3g env-steps/s stepping the same graphs (auto-reset), %.3g env-steps/s when every finished env gets its next "
          "instance from a resident pool of %d x B prepared instances (copy + reset, 3 launches per step; mean episode %.1f steps; %d banks regenerated "
          "in the background during the %d timed steps)." % (k, a["value"], b["value"], b["pool"]["banks"], b["mean_episode_steps"],
                                                             b["pool"]["banks_regenerated_in_background"], b["steps"]))
c5 = d["cfg5_strong_scaling"]
print("\nBASELINE config 5 on ONE GPU (524,288 envs per kind resident): Multicast %.3g env-steps/s (%.1f GB), DistributionCenter %.3g env-steps/s "
      "(%.1f GB), one step of all 1,048,576 envs in %.2f ms (%.3g env-steps/s)." % (
          c5["cfg5_multicast"]["value"], c5["cfg5_multicast"]["memory_gb_per_gpu"], c5["cfg5_distcenter"]["value"],
          c5["cfg5_distcenter"]["memory_gb_per_gpu"], c5["combined"]["ms_per_step_pair"], c5["combined"]["value"]))
