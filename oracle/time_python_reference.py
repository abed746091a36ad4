"""TEST / MEASUREMENT INFRASTRUCTURE.  Times the UNMODIFIED reference (graph_envs from /root/reference,
imported through oracle/ref_loader.py behind the gymnasium stubs) in THIS container: the README.md:54-68
loop -- `env.reset(seed=s)` then `env.step(np.random.choice(mask.nonzero()[0]))` until done -- for every
BASELINE config, (i) one process on one core and (ii) the same loop fanned out over all cores with
multiprocessing, one env per process (SURVEY.md 8(d) "CPU baseline beside it").

The reference needs networkx and /root/reference, neither of which exists on the GPU box, so this cannot run
beside the GPU numbers; bench.py carries the committed result file as `cpu_reference_python` with its date,
host and core count stated.

    python oracle/time_python_reference.py [seconds per config, default 6] [config names ...] > profiles/r02_python_reference_cpu.json
"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# (name, env id, kwargs) -- BASELINE.json configs; sizes as stated (the reference is intractable at some of them,
# a fixed small number of steps is timed there and the entry says so)
CONFIGS = [
    ("cfg1_shortest_path", "ShortestPath-v0", dict(n_nodes=10, n_edges=20, weighted=True)),
    ("cfg2_longest_path", "LongestPath-v0", dict(n_nodes=50, n_edges=200, weighted=True, parenting=2)),
    ("cfg3_mst", "SteinerTree-v0", dict(n_nodes=100, n_edges=500, n_dests=99, weighted=True, is_eval_env=True)),
    ("cfg4_tsp_p1", "TSP-v0", dict(n_nodes=200, n_edges=19900, weighted=True, parenting=1)),
    ("cfg4_tsp_p2", "TSP-v0", dict(n_nodes=200, n_edges=19900, weighted=True, parenting=2)),
    ("cfg4_mis", "MaxIndependentSet-v0", dict(n_nodes=200, n_edges=5970, weighted=True)),
    ("cfg5_multicast", "MulticastRouting-v0", dict(n_nodes=500, n_edges=4000, n_dests=3, parenting=4)),
    ("cfg5_distcenter", "DistributionCenter-v0", dict(n_nodes=500, n_edges=4000, parenting=2, target_count=100, max_distance=1)),
    ("densest", "DensestSubgraph-v0", dict(n_nodes=500, n_edges=4000, parenting=1)),
    ("perishable", "PerishableProductDelivery-v0", dict(n_nodes=50, n_edges=200, n_products=3, parenting=1)),
]


def loop(env_id, kwargs, budget_s, seed0):
    """README loop for `budget_s` seconds: returns (steps, step seconds, resets, reset seconds)."""
    import ref_loader
    gym, _ = ref_loader.load()
    env = gym.make(env_id, **kwargs)
    steps = resets = 0
    t_step = t_reset = 0.0
    seed = seed0
    t_end = time.perf_counter() + budget_s
    while time.perf_counter() < t_end:
        c0 = time.perf_counter()
        obs, info = env.reset(seed=seed)
        t_reset += time.perf_counter() - c0
        resets += 1
        seed += 1
        done = False
        while not done and time.perf_counter() < t_end:
            valid = info["mask"].nonzero()[0]
            a = np.random.choice(valid)
            c0 = time.perf_counter()
            obs, reward, done, _, info = env.step(a)
            t_step += time.perf_counter() - c0
            steps += 1
    return steps, t_step, resets, t_reset


def _worker(args):
    return loop(*args)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
    cores = os.cpu_count()
    out = {"what": "unmodified reference graph_envs (pure Python + networkx), README.md:54-68 loop", "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
           "host": "build container (no GPU): %s, %d cores visible" % (platform.processor() or platform.machine(), cores),
           "python": platform.python_version(), "seconds_per_config": budget, "configs": {}}
    import networkx
    out["networkx"] = networkx.__version__
    out["numpy"] = np.__version__
    only = set(sys.argv[2:])                      # optional: time just these configs
    for name, env_id, kw in CONFIGS:
        if only and name not in only:
            continue
        s, ts, r, tr = loop(env_id, kw, budget, 0)
        with mp.Pool(cores) as pool:
            res = pool.map(_worker, [(env_id, kw, budget, 1000 * (i + 1)) for i in range(cores)])
        wall = budget
        entry = {
            "env_id": env_id, "kwargs": kw,
            "one_core": {"steps": s, "resets": r, "step_only_steps_per_s": s / ts if ts > 0 else None,
                         "incl_reset_steps_per_s": s / (ts + tr) if ts + tr > 0 else None,
                         "ms_per_step": 1e3 * ts / s if s else None, "ms_per_reset": 1e3 * tr / r if r else None},
            "all_cores": {"processes": cores, "steps": sum(x[0] for x in res),
                          "step_only_steps_per_s": sum(x[0] / x[1] for x in res if x[1] > 0),
                          "incl_reset_steps_per_s": sum(x[0] for x in res) / wall},
        }
        if s < 50:
            entry["note"] = "reference intractable at this size: only %d steps fit in %.0f s" % (s, budget)
        out["configs"][name] = entry
        print(name, json.dumps(entry["one_core"]), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
