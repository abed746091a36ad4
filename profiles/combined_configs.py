"""BASELINE configs 4 and 5 as literally stated: TWO env kinds resident and stepping together on one GPU
(config 4: TSP N=200 dense + MaxIndependentSet N=200, 16,384 envs each; config 5: MulticastRouting +
DistributionCenter N=500 E=4000, `per_kind` envs each).  The two step kernels are independent, so they are
launched on two forked streams inside one CUDA graph; value = env-steps of both kinds / time of the pair.
    python profiles/combined_configs.py [per_kind_cfg5]      -> one JSON line per config"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphenvs_b200 import BatchedGraphEnv

SEED = 20260101
per5 = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frd = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").view(torch.int64)
sink = torch.zeros((), dtype=torch.int64, device="cuda")


def run(name, specs, K=200, G=50):
    envs = []
    for env_id, B, N, E, kw in specs:
        e = BatchedGraphEnv(env_id, B, N, E, auto_reset=True, **kw)
        e.generate(seed=SEED)
        e.release_w64()
        e.reset()
        e.enable_env_clock()
        envs.append(e)
    side = [torch.cuda.Stream() for _ in envs]

    def step(ev=None):
        flush.fill_(1)
        torch.sum(frd, dim=(0,), out=sink)
        if ev:
            ev[0].record()
        cur = torch.cuda.current_stream()
        for e, s in zip(envs, side):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                e.step_sampled(SEED, 0)
        for s in side:
            cur.wait_stream(s)
        if ev:
            ev[1].record()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True, external=True) for _ in range(2)] for _ in range(G)]
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(G):
            step(evs[i])
    g.replay()
    torch.cuda.synchronize()
    ms = []
    for _ in range(K // G):
        g.replay()
        torch.cuda.synchronize()
        ms += [e[0].elapsed_time(e[1]) for e in evs]
    total_envs = sum(e.B for e in envs)
    out = {"config": name, "kinds": [[s[0], s[1]] for s in specs], "envs_total": total_envs, "steps": len(ms),
           "ms_per_pair_step": float(np.mean(ms)), "env_steps_per_s": total_envs / (float(np.mean(ms)) * 1e-3),
           "memory_gb": sum(e.memory_bytes() for e in envs) / 1e9,
           "episodes": [float(e.stats()[0].item()) for e in envs]}
    print(json.dumps(out), flush=True)
    del envs
    torch.cuda.empty_cache()


run("cfg4: TSP p1 + MaxIndependentSet, 16384 envs each", [("TSP-v0", 16384, 200, 19900, dict(parenting=1)), ("MaxIndependentSet-v0", 16384, 200, 5970, {})])
run("cfg4: TSP p2 + MaxIndependentSet, 16384 envs each", [("TSP-v0", 16384, 200, 19900, dict(parenting=2)), ("MaxIndependentSet-v0", 16384, 200, 5970, {})])
run("cfg5: MulticastRouting + DistributionCenter, %d envs each" % per5,
    [("MulticastRouting-v0", per5, 500, 4000, dict(n_dests=3, parenting=4)),
     ("DistributionCenter-v0", per5, 500, 4000, dict(parenting=2, target_count=100, max_distance=1))], K=100, G=25)
