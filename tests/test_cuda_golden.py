"""GPU: the CUDA engine (through the C ABI) replays the recorded runs of the unmodified reference.
Bit-exact: masks, done/solved flags, dynamic node/edge columns and the full float32 observation
(sha256 of utils.vectorize_graph's layout).  1e-5 relative: reward, solution_cost."""
import numpy as np
import pytest
import torch

import cuda_util as cu

pytestmark = pytest.mark.gpu
GROUPS = cu.grouped_cases()


@pytest.mark.parametrize("path", ["auto", "warp"])  # auto = lane-per-env kernels where they apply (N <= 64)
@pytest.mark.parametrize("key", list(GROUPS), ids=["-".join(str(x) for x in k) for k in GROUPS])
def test_replay_group(key, path):
    cases = GROUPS[key]
    if path == "warp" and not cu.has_fast_path(key[0], key[1], key[3]):
        pytest.skip("warp-per-env is already the auto path here")
    env = cu.batch_from_cases(cases, force_warp=(path == "warp"))
    B = len(cases)
    info = env.reset()
    torch.cuda.synchronize()
    mask0 = info["mask"].cpu().numpy()
    obs0 = env.obs_flat().cpu().numpy()
    for b, (m, r) in enumerate(cases):
        np.testing.assert_array_equal(mask0[b], r["mask0"], err_msg="mask0 env %d" % b)
        assert cu.sha(obs0[b]) == m["obs0_sha"], "obs0 env %d" % b
    T = max(len(r["actions"]) for _, r in cases)
    for t in range(T):
        acts = np.zeros(B, dtype=np.int32)
        for b, (_, r) in enumerate(cases):
            if t < len(r["actions"]):
                acts[b] = r["actions"][t]
        reward, done, info = env.step(torch.from_numpy(acts).cuda())
        torch.cuda.synchronize()
        reward = reward.cpu().numpy(); done = done.cpu().numpy()
        solved = info["solved"].cpu().numpy(); status = info["status"].cpu().numpy()
        has_mask = info["has_mask"].cpu().numpy(); cost = info["solution_cost"].cpu().numpy()
        mask = info["mask"].cpu().numpy()
        obs = env.obs_flat().cpu().numpy()
        for b, (m, r) in enumerate(cases):
            if t >= len(r["actions"]):
                assert status[b] == 2, "stepping a finished env must report AFTER_DONE"
                continue
            tag = "env %d step %d" % (b, t)
            assert status[b] == 0, tag
            assert bool(done[b]) == bool(r["done"][t]), tag
            assert int(solved[b]) == int(r["solved"][t]), tag
            assert bool(has_mask[b]) == bool(r["has_mask"][t]), tag
            np.testing.assert_array_equal(mask[b], r["mask"][t], err_msg="mask " + tag)
            e = r["reward"][t]
            assert abs(reward[b] - e) <= 1e-5 * max(1.0, abs(e)), (tag, reward[b], e)
            e = r["solution_cost"][t]
            if np.isnan(e):
                assert np.isnan(cost[b]), tag
            else:
                assert abs(cost[b] - e) <= 1e-5 * max(1.0, abs(e)), (tag, cost[b], e)
            assert cu.sha(obs[b]) == m["obs_shas"][t], "obs " + tag


def test_invalid_actions_report_status():
    """Actions the reference rejects with AssertionError leave the state untouched (status 1)."""
    cases = GROUPS[[k for k in GROUPS if k[0] == "ShortestPath-v0" and k[1] == 10][0]]
    env = cu.batch_from_cases(cases)
    env.reset()
    before = env.obs_flat().clone()
    mask = env.mask.cpu().numpy()
    bad = np.array([int(np.flatnonzero(~mk)[0]) for mk in mask], dtype=np.int32)
    _, done, info = env.step(torch.from_numpy(bad).cuda())
    assert (info["status"].cpu().numpy() == 1).all() and not done.any()
    assert torch.equal(before, env.obs_flat())
    _, _, info = env.step(torch.full((len(cases),), 99, dtype=torch.int32).cuda())
    assert (info["status"].cpu().numpy() == 1).all()
