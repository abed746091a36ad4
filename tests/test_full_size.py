"""GPU: BASELINE.json's full-size configurations.  The oracle only replays a slice of each batch
(trajectory checksum, bit-exact); the whole batch is checked through size-independent properties
of each problem (episode lengths, cost identities, mask invariants) and through rank-slice
equality of the per-env trajectory checksums ("checksum of checksums")."""
import numpy as np
import pytest
import torch

import cuda_util as cu
from graphenvs_b200 import BatchedGraphEnv
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
SEED = 4242


def _rollout(env, T, t0=0):
    for t in range(t0, t0 + T):
        env.sample_actions(SEED, t)
        env.step_async(env.actions_dev)
    torch.cuda.synchronize()


def _oracle_slice_matches(env, env_id, T, n=192, lo=0):
    inst = env.export_instances(lo, n)
    oenvs = [cu.oracle_from_instance(env_id, i, env.params) for i in inst]
    sr, ep, cs = orc.rollout(oenvs, T, SEED, env_id0=env.desc.env_id0 + lo)
    traj = env.t["traj"][lo:lo + n].cpu().numpy().view(np.uint64)
    acc = env.t["acc"][:, lo:lo + n].cpu().numpy()
    np.testing.assert_array_equal(traj, cs)
    np.testing.assert_array_equal(acc[0].astype(np.int64), ep)
    np.testing.assert_allclose(acc[2], sr, rtol=1e-5, atol=1e-4)
    masks = np.stack([oe.mask(reset_patch=False) for oe in oenvs])
    got = env.mask[lo:lo + n].cpu().numpy()
    np.testing.assert_array_equal(got, masks)


def test_cfg2_longest_path_65536():
    B, T = 65536, 48
    env = BatchedGraphEnv("LongestPath-v0", B, 50, 200, parenting=2, auto_reset=True)
    env.generate(seed=1)
    env.reset()
    _rollout(env, T)
    _oracle_slice_matches(env, "LongestPath-v0", T, n=256, lo=0)
    _oracle_slice_matches(env, "LongestPath-v0", T, n=64, lo=B - 64)
    nb, mb, head = env.t["node_bits"], env.t["mask_bits"], env.t["head"].long()
    assert not (nb & mb).any(), "a visited node is offered as an action"
    adj = env.adjacency_rows()
    rows = adj[torch.arange(B, device=adj.device), head]
    assert not (mb & ~rows).any(), "mask must be a subset of N(head)"
    acc = env.t["acc"]
    assert torch.equal(acc[0], acc[1]), "parenting=2 prunes dead ends: every finished episode is solved"
    assert float(acc[0].sum()) > B  # > 1 episode per env on average in 48 steps
    # rank-sliced batches reproduce the full batch bit for bit
    full = env.t["traj"].clone()
    parts = []
    for r in range(4):
        e = BatchedGraphEnv("LongestPath-v0", B // 4, 50, 200, parenting=2, auto_reset=True, env_id0=r * (B // 4))
        e.generate(seed=1)
        e.reset()
        _rollout(e, T)
        parts.append(e.t["traj"])
    assert torch.equal(full, torch.cat(parts))


def test_cfg3_mst_32768():
    B, N, E = 32768, 100, 500
    env = BatchedGraphEnv("SteinerTree-v0", B, N, E, n_dests=N - 1, is_eval_env=True, auto_reset=False)
    env.generate(seed=2)
    env.reset()
    for t in range(N - 1):
        env.sample_actions(SEED, t)
        env.step_async(env.actions_dev)
        if t == N - 3:
            torch.cuda.synchronize()
            assert not env.t["done"].any(), "a spanning tree needs exactly N-1 edge picks"
    torch.cuda.synchronize()
    flags = env.flags.cpu().numpy()
    assert (flags[:, 0] == 1).all() and (flags[:, 1] == 1).all() and (flags[:, 2] == 0).all()
    cost, mst = env.solution_cost.cpu().numpy(), env.t["heuristic"].cpu().numpy()
    assert (cost >= mst - 1e-4).all(), "a random spanning tree cannot beat the MST weight"
    assert (mst >= 0.3 * (N - 1) - 1e-9).all() and (mst <= 0.9 * (N - 1) + 1e-9).all()
    assert int(env.t["node_bits"].cpu().numpy().view(np.uint32).astype(np.uint64).sum()) > 0
    inst = env.export_instances(0, 32)
    for b, ins in enumerate(inst):
        oe = cu.oracle_from_instance("SteinerTree-v0", ins, env.params)
        assert oe.mst_weight() == pytest.approx(mst[b], rel=1e-9)


def test_cfg4_mis_and_tsp_16384():
    B, N = 16384, 200
    mis = BatchedGraphEnv("MaxIndependentSet-v0", B, N, 5970, auto_reset=False)
    mis.generate(seed=3)
    mis.reset()
    _rollout(mis, N - 1)
    assert not mis.t["done"].any()
    _rollout(mis, 1, t0=N - 1)
    assert mis.t["done"].all(), "MaxIndependentSet episodes take exactly N steps"
    total = mis.t["node_cost"].double().sum(1)
    torch.testing.assert_close(mis.solution_cost, total, rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(-mis.t["acc"][2], total, rtol=1e-5, atol=1e-4)

    tsp = BatchedGraphEnv("TSP-v0", B, N, N * (N - 1) // 2, parenting=1, auto_reset=False)
    tsp.generate(seed=3)
    tsp.reset()
    m0 = tsp.mask
    assert not m0[:, 0].any() and m0[:, 1:].all(), "complete graph: every node but the start is a first move"
    _rollout(tsp, N)
    fl = tsp.flags.cpu().numpy()
    assert (fl[:, 0] == 1).all() and (fl[:, 1] == 1).all(), "on a complete graph every random tour closes after N moves"
    c = tsp.solution_cost.cpu().numpy()
    assert (c >= 0.3 * N - 1e-9).all() and (c <= 0.9 * N + 1e-9).all()
    _oracle_slice_matches_tsp(tsp, N)


def _oracle_slice_matches_tsp(tsp, T):
    inst = tsp.export_instances(0, 24)
    oenvs = [cu.oracle_from_instance("TSP-v0", i, tsp.params) for i in inst]
    sr, ep, cs = orc.rollout(oenvs, T, SEED)
    np.testing.assert_array_equal(tsp.t["traj"][:24].cpu().numpy().view(np.uint64), cs)
    np.testing.assert_allclose(tsp.t["acc"][2, :24].cpu().numpy(), sr, rtol=1e-5)
    # masks mid-episode (after the closing move the oracle rollout has auto-reset its envs, the batch above has not)
    tsp.reset()
    _rollout(tsp, T // 2)
    oenvs = [cu.oracle_from_instance("TSP-v0", i, tsp.params) for i in inst]
    orc.rollout(oenvs, T // 2, SEED)
    masks = np.stack([oe.mask(reset_patch=False) for oe in oenvs])
    np.testing.assert_array_equal(tsp.mask[:24].cpu().numpy(), masks)


def test_cfg4_tsp_parenting2_slice():
    """parenting=2 at N=200 dense: the residual graph stays complete, so pruning removes nothing and
    the tour still closes after N moves; checked against the oracle on a slice."""
    B, N = 2048, 200
    tsp = BatchedGraphEnv("TSP-v0", B, N, N * (N - 1) // 2, parenting=2, auto_reset=False)
    tsp.generate(seed=5)
    tsp.reset()
    _rollout(tsp, N)
    fl = tsp.flags.cpu().numpy()
    assert (fl[:, 0] == 1).all() and (fl[:, 1] == 1).all()
    inst = tsp.export_instances(0, 4)
    oenvs = [cu.oracle_from_instance("TSP-v0", i, tsp.params) for i in inst]
    sr, ep, cs = orc.rollout(oenvs, N, SEED)
    np.testing.assert_array_equal(tsp.t["traj"][:4].cpu().numpy().view(np.uint64), cs)
    # masks mid-episode (the oracle rollout auto-resets after the closing move): 60 moves into a fresh episode
    tsp.reset()
    _rollout(tsp, 60)
    oenvs = [cu.oracle_from_instance("TSP-v0", i, tsp.params) for i in inst]
    orc.rollout(oenvs, 60, SEED)
    np.testing.assert_array_equal(tsp.mask[:4].cpu().numpy(), np.stack([oe.mask(reset_patch=False) for oe in oenvs]))


def test_cfg5_multicast_and_distcenter_large():
    N, E = 500, 4000
    B = 131072   # the batch bench.py steps per GPU; 256-env oracle slices from the start, the middle and the end
    mc = BatchedGraphEnv("MulticastRouting-v0", B, N, E, n_dests=3, parenting=4, auto_reset=True)
    mc.generate(seed=6)
    mc.reset()
    T = 24
    _rollout(mc, T)
    for lo in (0, B // 2 - 128, B - 256):
        _oracle_slice_matches(mc, "MulticastRouting-v0", T, n=256, lo=lo)
    # parenting 4 keeps exactly one candidate edge per frontier vertex
    mb = mc.t["mask_bits"][:64].cpu().numpy().view(np.uint32)
    col = mc.t["col"][:64].cpu().numpy()
    bits = np.unpackbits(mb.view(np.uint8), axis=1, bitorder="little")[:, :2 * E].astype(bool)
    for b in range(64):
        dst = col[b, :2 * E][bits[b]]
        assert len(set(dst.tolist())) == dst.size, "two candidate edges for one frontier vertex"

    dc = BatchedGraphEnv("DistributionCenter-v0", B, N, E, parenting=2, target_count=100, max_distance=1, auto_reset=True)
    dc.generate(seed=7)
    dc.reset()
    del mc
    torch.cuda.empty_cache()
    _rollout(dc, 30)
    for lo in (0, B // 2 - 128, B - 256):
        _oracle_slice_matches(dc, "DistributionCenter-v0", 30, n=256, lo=lo)
    taken, covered, mask = dc.t["node_bits"], dc.t["node_bits2"], dc.t["mask_bits"]
    assert not (taken & mask).any()
    assert ((taken & covered) == taken).all(), "a chosen centre covers itself"
