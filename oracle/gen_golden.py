"""TEST INFRASTRUCTURE. Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/graph_envs, imported through oracle/ref_loader.py) in this container with
numpy 2.3.5 / networkx 3.6.1 / CPython 3.12.  The reference's own tests pin no results for the
hot path, so these recorded runs of the reference itself are the pin.

    python oracle/gen_golden.py            # rewrites tests/golden/

Each case = one `reset(seed)` + one full episode.  Recorded per case:
  instance : edge_links int32[M,2] (reference order), w64 float64[M] (the nx edge attribute),
             nodes0 float32[N,F] / edges0 float32[M,Fe] right after reset, terminals, params
  features : nx values float64[N,5] (before the reference's float32 rounding)
  trace    : actions, reward (f64), done, solved (-1 = key absent), has_mask, mask bool[T,A],
             solution_cost / heuristic (nan = key absent), dynamic node columns after every
             step, edge IS_TAKEN column after every step (edge-action envs), obs sha256 prefixes
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

gym, ge = ref_loader.load()
import networkx as nx  # noqa: E402
import graph_envs.feature_extraction as fe  # noqa: E402

_captured = {}
_orig_generate = fe.generate_features


def _capture_generate(G):
    _captured["G"] = G
    return _orig_generate(G)


fe.generate_features = _capture_generate  # harness-side hook; reference files untouched

N_STRUCT = 5


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


def weight_attr(env_id):
    return "weight" if env_id == "TSP-v0" else "delay"


def nx_features64(G):
    cl = nx.clustering(G)
    pr = nx.pagerank(G)
    bt = nx.betweenness_centrality(G)
    cs = nx.closeness_centrality(G)
    return np.array([[G.degree(v), bt[v], cs[v], pr[v], cl[v]] for v in G.nodes], dtype=np.float64)


def run_case(env_id, kwargs, seed, policy, max_steps=100000):
    env = gym.make(env_id, **kwargs)
    obs, info = env.reset(seed=seed)
    G = _captured["G"]
    g = env.graph
    N = g.nodes.shape[0]
    M = g.edge_links.shape[0]
    attr = weight_attr(env_id)
    w64 = np.array([G[u][v].get(attr, 1.0) for u, v in g.edge_links], dtype=np.float64)
    rec = {
        "edge_links": g.edge_links.astype(np.int32),
        "w64": w64,
        "nodes0": g.nodes.copy(),
        "edges0": g.edges.copy(),
        "features64": nx_features64(G),
        "mask0": info["mask"].copy(),
    }
    meta = {"env_id": env_id, "kwargs": kwargs, "seed": seed, "policy": policy, "N": int(N), "M": int(M),
            "obs0_sha": sha(obs), "obs_len": int(obs.shape[0])}
    for name in ("src", "dest", "start", "head"):
        if hasattr(env, name):
            meta[name] = int(getattr(env, name))
    if hasattr(env, "dests"):
        rec["dests"] = np.asarray(env.dests, dtype=np.int32)
    if hasattr(env, "optimal_solution"):
        meta["heuristic"] = float(env.optimal_solution)
    if hasattr(env, "approx_solution"):
        meta["heuristic"] = float(env.approx_solution)
    if hasattr(env, "n_choices"):
        meta["n_choices"] = float(env.n_choices)
    if hasattr(env, "pickups"):           # PerishableProductDelivery
        rec["pickups"] = np.asarray(env.pickups, dtype=np.int32)
        rec["dropoffs"] = np.asarray(env.dropoffs, dtype=np.int32)
        meta["delivery_time"] = float(env.delivery_time)
    if hasattr(env, "in_range_dict"):
        tg = sorted(int(t) for t in env.in_range_dict)
        rec["targets"] = np.array(tg, dtype=np.int32)
        tab = np.zeros((len(tg), N), dtype=np.uint8)
        for i, t in enumerate(tg):
            tab[i, env.in_range_dict[t]] = 1
        rec["in_range"] = tab
    n_dyn = g.nodes.shape[1] - N_STRUCT
    has_edge_taken = g.edges.shape[1] > 1

    actions, rewards, dones, solveds, has_masks, masks = [], [], [], [], [], []
    costs, heurs, dyn, etaken, obs_shas = [], [], [], [], []
    mask = info["mask"]
    rng = np.random  # README loop: actions continue the global numpy stream
    done = False
    t = 0
    while not done and t < max_steps:
        valid = mask.nonzero()[0]
        if policy == "lowest":
            a = int(valid[0])
        elif policy == "random":
            a = int(rng.choice(valid))
        elif policy == "highest":
            a = int(valid[-1])
        elif policy == "anynode":  # envs whose invalid moves have defined semantics
            a = int(rng.randint(0, mask.shape[0]))
        elif policy == "start_first":
            a = int(env.start)
        else:
            raise ValueError(policy)
        obs, reward, done, trunc, info = env.step(a)
        assert trunc is False
        actions.append(a)
        rewards.append(float(reward))
        dones.append(bool(done))
        solveds.append(-1 if "solved" not in info else int(bool(info["solved"])))
        has_masks.append("mask" in info)
        if "mask" in info:
            mask = info["mask"]
        masks.append(np.asarray(mask, dtype=bool).copy())
        costs.append(float(info["solution_cost"]) if "solution_cost" in info else np.nan)
        heurs.append(float(info["heuristic_solution"]) if "heuristic_solution" in info else np.nan)
        dyn.append(env.graph.nodes[:, :n_dyn].copy())
        if has_edge_taken:
            etaken.append((env.graph.edges[:, 1] > 0.5).astype(np.uint8))
        obs_shas.append(sha(obs))
        t += 1
    A = mask.shape[0]
    rec.update({
        "actions": np.array(actions, dtype=np.int32),
        "reward": np.array(rewards, dtype=np.float64),
        "done": np.array(dones, dtype=np.uint8),
        "solved": np.array(solveds, dtype=np.int8),
        "has_mask": np.array(has_masks, dtype=np.uint8),
        "mask": np.array(masks, dtype=bool).reshape(len(actions), A),
        "solution_cost": np.array(costs, dtype=np.float64),
        "heuristic": np.array(heurs, dtype=np.float64),
        "nodes_dyn": np.array(dyn, dtype=np.float32).reshape(len(actions), N, n_dyn),
    })
    if has_edge_taken:
        rec["edge_taken"] = np.array(etaken, dtype=np.uint8).reshape(len(actions), M)
    meta["obs_shas"] = obs_shas
    meta["final_obs_sha"] = obs_shas[-1] if obs_shas else meta["obs0_sha"]
    return meta, rec


def cases():
    out = []
    SP = "ShortestPath-v0"
    for s in range(10):  # BASELINE config 1
        out.append((SP, dict(n_nodes=10, n_edges=20, weighted=True, is_eval_env=True), s, "random"))
    out.append((SP, dict(n_nodes=10, n_edges=20, weighted=True, is_eval_env=True), 0, "lowest"))
    out.append((SP, dict(n_nodes=10, n_edges=20, weighted=False, is_eval_env=True), 1, "random"))
    for s in range(3):
        out.append((SP, dict(n_nodes=30, n_edges=60, weighted=True, is_eval_env=True), s, "random"))
    out.append((SP, dict(n_nodes=70, n_edges=160, weighted=True, is_eval_env=True), 0, "random"))

    LP = "LongestPath-v0"
    for p in (0, 1, 2, 3):
        pol = "anynode" if p == 0 else "random"
        for s in range(3):
            out.append((LP, dict(n_nodes=10, n_edges=20, weighted=True, is_eval_env=True, parenting=p), s, pol))
        out.append((LP, dict(n_nodes=10, n_edges=20, weighted=True, is_eval_env=True, parenting=p), 0, "lowest"))
    for s in range(4):  # BASELINE config 2 shape
        out.append((LP, dict(n_nodes=50, n_edges=200, weighted=True, is_eval_env=True, parenting=2), s, "random"))
    out.append((LP, dict(n_nodes=50, n_edges=200, weighted=True, parenting=2), 4, "highest"))
    out.append((LP, dict(n_nodes=64, n_edges=150, weighted=True, parenting=2), 0, "random"))
    out.append((LP, dict(n_nodes=70, n_edges=140, weighted=True, parenting=2), 0, "random"))
    out.append((LP, dict(n_nodes=70, n_edges=140, weighted=False, parenting=3), 1, "random"))
    out.append((LP, dict(n_nodes=40, n_edges=-1, weighted=True, parenting=1), 1, "random"))
    out.append((LP, dict(n_nodes=40, n_edges=60, weighted=True, parenting=3), 2, "lowest"))

    ST = "SteinerTree-v0"
    for nd in (1, 3, 9):
        for s in range(2):
            out.append((ST, dict(n_nodes=10, n_edges=20, n_dests=nd, weighted=True, is_eval_env=True), s, "random"))
    out.append((ST, dict(n_nodes=10, n_edges=20, n_dests=9, weighted=True, is_eval_env=True), 0, "lowest"))
    out.append((ST, dict(n_nodes=30, n_edges=80, n_dests=5, weighted=True, is_eval_env=True), 0, "random"))
    out.append((ST, dict(n_nodes=30, n_edges=80, n_dests=29, weighted=False, is_eval_env=True), 1, "random"))
    out.append((ST, dict(n_nodes=100, n_edges=500, n_dests=99, weighted=True, is_eval_env=True), 0, "random"))  # cfg 3

    TS = "TSP-v0"
    for p in (1, 2):
        for s in range(3):
            out.append((TS, dict(n_nodes=10, n_edges=30, weighted=True, parenting=p, is_eval_env=True), s, "random"))
        out.append((TS, dict(n_nodes=10, n_edges=30, weighted=True, parenting=p), 0, "lowest"))
        out.append((TS, dict(n_nodes=12, n_edges=66, weighted=True, parenting=p), 0, "random"))  # complete graph
        out.append((TS, dict(n_nodes=30, n_edges=90, weighted=True, parenting=p), 1, "random"))
        out.append((TS, dict(n_nodes=10, n_edges=30, weighted=True, parenting=p), 0, "start_first"))
    out.append((TS, dict(n_nodes=40, n_edges=100, weighted=False, parenting=2), 2, "random"))
    out.append((TS, dict(n_nodes=70, n_edges=200, weighted=True, parenting=2), 0, "random"))
    out.append((TS, dict(n_nodes=12, n_edges=30, weighted=True, parenting=2, spatial=True), 3, "random"))

    MI = "MaxIndependentSet-v0"
    out.append((MI, dict(n_nodes=10, n_edges=20, weighted=True, is_eval_env=True), 0, "lowest"))
    out.append((MI, dict(n_nodes=10, n_edges=20, weighted=False, is_eval_env=True), 1, "random"))
    out.append((MI, dict(n_nodes=40, n_edges=100, weighted=True), 2, "random"))
    out.append((MI, dict(n_nodes=70, n_edges=300, weighted=True), 3, "random"))

    DS = "DensestSubgraph-v0"
    for p in (0, 1):
        for s in range(3):
            out.append((DS, dict(n_nodes=10, n_edges=20, parenting=p), s, "random"))
        out.append((DS, dict(n_nodes=10, n_edges=20, parenting=p), 0, "lowest"))
        out.append((DS, dict(n_nodes=10, n_edges=20, parenting=p), 0, "highest"))
        out.append((DS, dict(n_nodes=30, n_edges=-1, parenting=p, n_choices=12), 1, "random"))
        out.append((DS, dict(n_nodes=70, n_edges=200, parenting=p), 2, "random"))

    MC = "MulticastRouting-v0"
    for p in (1, 2, 3, 4):
        for s in range(3):
            out.append((MC, dict(n_nodes=10, n_edges=20, n_dests=3, parenting=p, is_eval_env=True), s, "random"))
        out.append((MC, dict(n_nodes=10, n_edges=20, n_dests=3, parenting=p, is_eval_env=True), 0, "lowest"))
        out.append((MC, dict(n_nodes=30, n_edges=90, n_dests=5, parenting=p), 1, "random"))
    for s in range(3):
        out.append((MC, dict(n_nodes=60, n_edges=300, n_dests=4, parenting=4), s, "random"))
    out.append((MC, dict(n_nodes=40, n_edges=-1, n_dests=3, weighted=False, parenting=4), 0, "random"))

    DC = "DistributionCenter-v0"
    for p in (1, 2):
        for s in range(3):
            out.append((DC, dict(n_nodes=10, n_edges=20, parenting=p), s, "random"))
        out.append((DC, dict(n_nodes=10, n_edges=20, parenting=p), 0, "lowest"))
        out.append((DC, dict(n_nodes=40, n_edges=120, parenting=p, max_distance=1.2), 1, "random"))
        out.append((DC, dict(n_nodes=40, n_edges=120, parenting=p, max_distance=0.7, target_count=10), 2, "random"))
    out.append((DC, dict(n_nodes=70, n_edges=280, parenting=2, weighted=False, max_distance=2), 0, "random"))
    out.append((DC, dict(n_nodes=100, n_edges=500, parenting=2, max_distance=1), 0, "random"))
    PP = "PerishableProductDelivery-v0"
    for s in range(3):
        out.append((PP, dict(n_nodes=10, n_edges=20, n_products=3, parenting=1, is_eval_env=True), s, "random"))
    out.append((PP, dict(n_nodes=10, n_edges=20, n_products=3, parenting=1, is_eval_env=True), 0, "lowest"))
    out.append((PP, dict(n_nodes=30, n_edges=90, n_products=5, parenting=1, is_eval_env=True), 1, "random"))
    out.append((PP, dict(n_nodes=60, n_edges=200, n_products=2, parenting=1), 2, "random"))
    out.append((PP, dict(n_nodes=20, n_edges=-1, n_products=1, weighted=False, parenting=1, is_eval_env=True), 3, "random"))
    out.append((PP, dict(n_nodes=70, n_edges=160, n_products=4, parenting=1, is_eval_env=True), 0, "random"))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    by_env = {}
    only = set(sys.argv[1:])            # optional: regenerate just these env ids
    for env_id, kw, seed, pol in cases():
        if only and env_id not in only:
            continue
        meta, rec = run_case(env_id, kw, seed, pol)
        by_env.setdefault(env_id, []).append((meta, rec))
    versions = {"numpy": np.__version__, "networkx": nx.__version__, "python": sys.version.split()[0]}
    for env_id, lst in by_env.items():
        arrays = {}
        metas = []
        for i, (meta, rec) in enumerate(lst):
            metas.append(meta)
            for k, v in rec.items():
                arrays["c%d_%s" % (i, k)] = v
        arrays["meta"] = np.array(json.dumps({"versions": versions, "cases": metas}))
        path = os.path.join(OUT, env_id.replace("-v0", "") + ".npz")
        np.savez_compressed(path, **arrays)
        nsteps = sum(len(m["obs_shas"]) for m in metas)
        print("%-24s %3d cases %6d steps -> %s (%d KB)" % (env_id, len(lst), nsteps, path, os.path.getsize(path) // 1024))


if __name__ == "__main__":
    main()
