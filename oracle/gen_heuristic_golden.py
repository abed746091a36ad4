"""TEST INFRASTRUCTURE. Records eval-heuristic values of the UNMODIFIED reference (reset(seed) with
is_eval_env=True) for instance sizes beyond tests/golden/*.npz -> tests/golden/heuristics.json.
    python oracle/gen_heuristic_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

gym, ge = ref_loader.load()
out = []
for env_id, kws, seeds in [
    ("MulticastRouting-v0", [dict(n_nodes=10, n_edges=20, n_dests=3), dict(n_nodes=30, n_edges=90, n_dests=5),
                             dict(n_nodes=60, n_edges=300, n_dests=4), dict(n_nodes=120, n_edges=600, n_dests=8),
                             dict(n_nodes=40, n_edges=-1, n_dests=3, weighted=False)], range(8)),
    ("ShortestPath-v0", [dict(n_nodes=80, n_edges=240)], range(4)),
    ("TSP-v0", [dict(n_nodes=12, n_edges=30, parenting=1), dict(n_nodes=20, n_edges=60, parenting=2),
                dict(n_nodes=14, n_edges=91, parenting=1), dict(n_nodes=30, n_edges=100, parenting=2, weighted=False)], range(4)),
    ("MaxIndependentSet-v0", [dict(n_nodes=20, n_edges=40, weighted=False), dict(n_nodes=40, n_edges=120, weighted=False)], range(5)),
    ("SteinerTree-v0", [dict(n_nodes=30, n_edges=80, n_dests=3), dict(n_nodes=60, n_edges=200, n_dests=5),
                        dict(n_nodes=40, n_edges=100, n_dests=10, weighted=False)], range(5)),
    ("SteinerTree-v0", [dict(n_nodes=60, n_edges=200, n_dests=59), dict(n_nodes=60, n_edges=200, n_dests=1)], range(4)),
]:
    for kw in kws:
        for s in seeds:
            k = dict(kw, is_eval_env=True)
            if env_id == "MulticastRouting-v0":
                k["parenting"] = 4
            env = gym.make(env_id, **k)
            env.reset(seed=s)
            h = float(getattr(env, "approx_solution", getattr(env, "optimal_solution", 0)))
            out.append({"env_id": env_id, "kwargs": k, "seed": s, "heuristic": h})
path = os.path.join(os.path.dirname(HERE), "tests", "golden", "heuristics.json")
json.dump(out, open(path, "w"), indent=0)
print(len(out), "values ->", path)
