"""GPU: sub-batch stream chains (SliceStreams) and programmatic dependent launch (GE_FLAG_PDL) are scheduling choices only --
a rollout stepped as C free-running slices with PDL, eagerly and from a captured CUDA graph, ends in exactly the state and
trajectory checksums of the plain one-launch-per-step rollout (same seeds, same in-kernel sampler)."""
import pytest
import torch

from graphenvs_b200 import BatchedGraphEnv
from graphenvs_b200.batch import SliceStreams

pytestmark = pytest.mark.gpu

CFG = [
    ("LongestPath-v0", 50, 200, {"parenting": 2}, 4096 + 96),          # lane kernel, staged tiles, ragged last slice
    ("ShortestPath-v0", 10, 20, {}, 4096),
    ("TSP-v0", 70, 200, {"parenting": 2}, 1024),                        # group kernel
    ("SteinerTree-v0", 100, 500, {"n_dests": 99}, 1024),                # incremental tree kernel
    ("MaxIndependentSet-v0", 200, 600, {}, 1024),                       # incremental MIS kernel
    ("MulticastRouting-v0", 120, 500, {"parenting": 4, "n_dests": 3}, 512),
    ("DistributionCenter-v0", 120, 500, {"parenting": 2, "target_count": 20, "max_distance": 1}, 512),
    ("DensestSubgraph-v0", 120, 500, {"parenting": 1}, 512),
    ("PerishableProductDelivery-v0", 50, 200, {"n_products": 3, "parenting": 1}, 2048),
]
STATE = ("node_bits", "node_bits2", "edge_bits", "dist32", "bestkey", "head", "cost", "counters", "done", "mask_bits", "traj", "acc")
T = 24


def _make(cfg):
    env_id, N, E, kw, B = cfg
    env = BatchedGraphEnv(env_id, B, N, E, auto_reset=True, **kw)
    env.generate(seed=77)
    env.reset()
    env.enable_env_clock()
    return env


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "graph"])
@pytest.mark.parametrize("cfg", CFG, ids=[c[0][:-3] for c in CFG])
def test_slice_streams_with_pdl_equal_plain_stepping(cfg, graph):
    ref = _make(cfg)
    for _ in range(T):
        ref.step_sampled(5, 0)
    torch.cuda.synchronize()

    env = _make(cfg)
    env.enable_pdl()
    ss = SliceStreams(env, 3)
    assert sum(n for _, n in ss.bounds) == env.B and all(lo % 32 == 0 for lo, _ in ss.bounds)
    if graph:
        ss.fork(); ss.step_sampled(5, 0); ss.join()          # warm-up outside the capture (kernel attributes), counts as step 1
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ss.fork()
            ss.step_sampled(5, 0)
            ss.join()
        for _ in range(T - 1):
            g.replay()
    else:
        ss.fork()
        for _ in range(T):
            ss.step_sampled(5, 0)
        ss.join()
    torch.cuda.synchronize()
    for k in STATE:
        if k in ref.t:
            assert torch.equal(ref.t[k], env.t[k]), "%s differs between plain and sliced + PDL stepping" % k
    assert torch.equal(ref.reward, env.reward) and torch.equal(ref.flags, env.flags)
    assert torch.equal(ref.actions_dev, env.actions_dev)
    assert float(ref.stats()[0]) == float(env.stats()[0])


def test_pdl_chain_on_one_stream_equals_plain():
    cfg = CFG[0]
    ref, env = _make(cfg), _make(cfg)
    env.enable_pdl()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(T):
            env.step_sampled(9, 0)
    for _ in range(T):
        ref.step_sampled(9, 0)
    torch.cuda.synchronize()
    for k in STATE:
        if k in ref.t:
            assert torch.equal(ref.t[k], env.t[k]), k


@pytest.mark.parametrize("cfg", [("LongestPath-v0", 50, 200, {"parenting": 2}, 20 * 1024 - 200),
                                 ("DistributionCenter-v0", 120, 500, {"parenting": 2, "target_count": 20, "max_distance": 1}, 6 * 1024 - 40)],
                         ids=["lane", "distcenter"])
@pytest.mark.parametrize("compact", [False, True], ids=["full", "compact"])
def test_streamed_host_step_equals_sliced(cfg, compact):
    """ge_step_host_pipelined / _compact with chunks = 0 (one step kernel signalling ge_batch.progress + a concurrent write-back kernel
    that ships every 1024-env chunk as soon as it is complete) delivers exactly what the two-slice path delivers; many chunks, ragged tail."""
    import ctypes
    logs = []
    for chunks in (2, 0):
        e = _make(cfg)
        B = e.B
        assert e.lib.ge_progress_supported(ctypes.byref(e.desc)) == 1
        if compact:
            h_rew, h_flg, h_cost, h_bits = e.host_io_compact()
        else:
            _blk, h_rew, h_flg, h_cost, h_bits = e.host_io()
        h_act = torch.zeros(B, dtype=torch.int32).pin_memory()
        side = torch.cuda.Stream()
        torch.cuda.synchronize()
        stepper = e.host_stepper(h_act, h_rew, h_flg, h_cost, None, h_bits, stream=side, pipelined=True, chunks=chunks, compact=compact)
        log = []
        for t in range(T):
            h_act.copy_(e.sample_actions(9, t).cpu())
            torch.cuda.synchronize()
            stepper()
            assert torch.equal(h_bits, e.t["mask_bits"].cpu()) and torch.equal(h_rew, e.reward.cpu())
            log.append((h_rew.clone(), h_flg.clone(), torch.nan_to_num(h_cost.clone(), nan=-7.0), h_bits.clone()))
        logs.append((log, e.t["traj"].clone()))
        del stepper
    for a, b in zip(logs[0][0], logs[1][0]):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert torch.equal(logs[0][1], logs[1][1])
