"""Helpers shared by the CPU (oracle) and GPU (CUDA) parity tests: load the recorded runs of the
unmodified reference (tests/golden/*.npz, written by oracle/gen_golden.py) and turn a case into
the plain instance description both implementations are constructed from."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ENV_IDS = ["ShortestPath-v0", "LongestPath-v0", "SteinerTree-v0", "TSP-v0", "MaxIndependentSet-v0",
           "DensestSubgraph-v0", "MulticastRouting-v0", "DistributionCenter-v0", "PerishableProductDelivery-v0"]
DYN_COLS = {"ShortestPath-v0": 2, "LongestPath-v0": 2, "SteinerTree-v0": 2, "TSP-v0": 4,
            "MaxIndependentSet-v0": 2, "DensestSubgraph-v0": 1, "MulticastRouting-v0": 4,
            "DistributionCenter-v0": 5, "PerishableProductDelivery-v0": 16}


def load_cases(env_id):
    z = np.load(os.path.join(GOLDEN_DIR, env_id.replace("-v0", "") + ".npz"))
    meta = json.loads(str(z["meta"]))
    out = []
    for i, m in enumerate(meta["cases"]):
        pre = "c%d_" % i
        rec = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        out.append((m, rec))
    return out


def all_cases():
    for env_id in ENV_IDS:
        for m, r in load_cases(env_id):
            yield m, r


def case_id(m):
    kw = ",".join("%s=%s" % (k[:5], v) for k, v in m["kwargs"].items() if k not in ("n_nodes", "n_edges"))
    return "%s-N%d-M%d-%s-s%d-%s" % (m["env_id"][:-3], m["N"], m["M"], kw, m["seed"], m["policy"])


def instance_kwargs(m, r):
    """Instance description (what reset() decided) extracted from a golden case."""
    env_id, kw = m["env_id"], m["kwargs"]
    nd = DYN_COLS[env_id]
    nodes0 = r["nodes0"]
    d = dict(N=m["N"], links=r["edge_links"], w64=r["w64"], parenting=kw.get("parenting", -1),
             features=nodes0[:, nd:].astype(np.float32), heuristic=m.get("heuristic", 0.0))
    if env_id == "DistributionCenter-v0" and "parenting" not in kw:
        d["parenting"] = 2
    if env_id == "MulticastRouting-v0" and "parenting" not in kw:
        d["parenting"] = 4
    if env_id in ("ShortestPath-v0", "LongestPath-v0"):
        d.update(src=m["src"], dest=m["dest"])
    elif env_id == "SteinerTree-v0":
        d.update(src=m["src"], dests=r["dests"], n_dests=len(r["dests"]))
    elif env_id == "TSP-v0":
        if kw.get("spatial"):
            d["node_xy"] = nodes0[:, 2:4].astype(np.float64)
    elif env_id == "MaxIndependentSet-v0":
        d["node_cost"] = nodes0[:, 0].astype(np.float64)
    elif env_id == "DensestSubgraph-v0":
        d["n_choices"] = int(m["n_choices"])
    elif env_id == "MulticastRouting-v0":
        d.update(src=0, dests=r["dests"], n_dests=len(r["dests"]), max_distance=float(nodes0[0, 2]))
    elif env_id == "DistributionCenter-v0":
        d.update(dests=r["targets"], node_cost=nodes0[:, 0].astype(np.float64),
                 max_distance=float(kw.get("max_distance", 1)))
    elif env_id == "PerishableProductDelivery-v0":
        d.update(dests=np.concatenate([r["pickups"], r["dropoffs"]]).astype(np.int32), n_dests=len(r["pickups"]),
                 max_distance=float(m["delivery_time"]))
    return d
