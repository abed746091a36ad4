"""CPU: the C oracle (oracle/graphenvs_oracle.c) replayed against recorded runs of the unmodified
reference (tests/golden).  This is what pins the oracle; the CUDA path is then checked against
the oracle (and against the same fixtures) in the -m gpu tests."""
import hashlib

import numpy as np
import pytest

import golden_util as gu
from oracle import oracle as orc

CASES = list(gu.all_cases())


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:12]


@pytest.mark.parametrize("case", CASES, ids=[gu.case_id(m) for m, _ in CASES])
def test_replay(case):
    m, r = case
    env_id = m["env_id"]
    env = orc.OracleEnv(orc.KINDS[env_id], **gu.instance_kwargs(m, r))
    nd = gu.DYN_COLS[env_id]
    # state right after reset: node/edge matrices, obs (utils.py:87-88), initial mask
    np.testing.assert_array_equal(env.nodes, r["nodes0"])
    np.testing.assert_array_equal(env.edges, r["edges0"])
    assert sha(env.obs()) == m["obs0_sha"]
    np.testing.assert_array_equal(env.mask(reset_patch=True), r["mask0"])
    T = len(r["actions"])
    for t in range(T):
        out = env.step(int(r["actions"][t]))
        assert out["status"] == 0, (t, out)
        assert out["done"] == bool(r["done"][t]), t
        assert out["solved"] == int(r["solved"][t]), t
        assert out["has_mask"] == bool(r["has_mask"][t]), t
        if out["has_mask"]:
            np.testing.assert_array_equal(out["mask"], r["mask"][t], err_msg="mask step %d" % t)
        exp_r = r["reward"][t]
        assert abs(out["reward"] - exp_r) <= 1e-5 * max(1.0, abs(exp_r)), (t, out["reward"], exp_r)
        for key in ("solution_cost", "heuristic"):
            e = r[key][t]
            if np.isnan(e):
                assert np.isnan(out[key]), (key, t)
            elif not (key == "heuristic"):
                assert abs(out[key] - e) <= 1e-5 * max(1.0, abs(e)), (key, t, out[key], e)
        np.testing.assert_array_equal(env.nodes[:, :nd], r["nodes_dyn"][t], err_msg="dyn cols step %d" % t)
        if "edge_taken" in r:
            np.testing.assert_array_equal((env.edges[:, 1] > 0.5).astype(np.uint8), r["edge_taken"][t])
        assert sha(env.obs()) == m["obs_shas"][t], "obs step %d" % t
    assert T == 0 or bool(r["done"][-1])


@pytest.mark.parametrize("case", CASES, ids=[gu.case_id(m) for m, _ in CASES])
def test_features(case):
    """feature_extraction.generate_features: float64 values vs networkx, and float32 rounding."""
    m, r = case
    env_id = m["env_id"]
    env = orc.OracleEnv(orc.KINDS[env_id], **gu.instance_kwargs(m, r))
    f64 = env.features64(weighted_pr=(env_id == "TSP-v0"))
    ref = r["features64"]
    np.testing.assert_allclose(f64, ref, rtol=1e-9, atol=1e-12)
    nd = gu.DYN_COLS[env_id]
    np.testing.assert_allclose(f64.astype(np.float32), r["nodes0"][:, nd:], rtol=1e-6, atol=1e-9)


def _heur_cases():
    for m, r in CASES:
        kw = m["kwargs"]
        if not kw.get("is_eval_env"):
            continue
        if m["env_id"] in ("ShortestPath-v0", "LongestPath-v0"):
            yield m, r
        if m["env_id"] == "SteinerTree-v0" and kw["n_dests"] in (1, m["N"] - 1):
            yield m, r


HC = list(_heur_cases())


@pytest.mark.parametrize("case", HC, ids=[gu.case_id(m) for m, _ in HC])
def test_heuristics(case):
    """Tie-independent eval heuristics: Dijkstra value (shortest_path.py:90, longest_path.py:105,
    steiner_tree.py:79) and Kruskal total weight (steiner_tree.py:81)."""
    m, r = case
    env_id = m["env_id"]
    kwargs = gu.instance_kwargs(m, r)
    env = orc.OracleEnv(orc.KINDS[env_id], **kwargs)
    if env_id == "ShortestPath-v0":
        got = env.sssp(m["src"])[m["dest"]]
    elif env_id == "LongestPath-v0":
        got = -env.sssp(m["src"])[m["dest"]]
    elif m["kwargs"]["n_dests"] == 1:
        got = env.sssp(m["src"])[int(r["dests"][0])]
    else:
        got = env.mst_weight()
    assert got == pytest.approx(m["heuristic"], rel=1e-12)


def test_dc_in_range_tables():
    for m, r in gu.load_cases("DistributionCenter-v0"):
        kw = gu.instance_kwargs(m, r)
        env = orc.OracleEnv(7, **kw)
        for i, t in enumerate(r["targets"]):
            d = env.sssp(int(t), cutoff=kw["max_distance"])
            np.testing.assert_array_equal((d <= kw["max_distance"]).astype(np.uint8), r["in_range"][i])
