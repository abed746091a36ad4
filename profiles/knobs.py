"""Diagnostic (needs a -DGE_KNOBS build): step-kernel time at cfg2 with parts of the kernel switched off.
    GE_KNOBS=1 python -c 'from graphenvs_b200 import _native; _native.build(force=True)'; python profiles/knobs.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphenvs_b200 import BatchedGraphEnv
B = int(os.environ.get("B", 65536))
print("GE_LANE_T", os.environ.get("GE_LANE_T"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
frd = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").view(torch.int64)
sink = torch.zeros((), dtype=torch.int64, device="cuda")
for name, bits in [("full", 0), ("no_bfs", 0x100), ("no_bfs_no_stage", 0x300), ("no_maskbytes", 0x400), ("no_wmat", 0x800),
                   ("nothing", 0xf00), ("stage_only", 0x1000), ("empty", 0x2200), ("full2", 0)]:
    env = BatchedGraphEnv("LongestPath-v0", B, 50, 200, parenting=2, auto_reset=True)
    env.generate(seed=1); env.reset(); env.enable_env_clock()
    env.desc.flags |= bits
    G = 50
    ev = [[torch.cuda.Event(enable_timing=True, external=True) for _ in range(3)] for _ in range(G)]
    def step(e=None):
        flush.fill_(1); torch.sum(frd, dim=(0,), out=sink)
        if e: e[0].record()
        if e: e[1].record()
        env.step_sampled(1, 0)
        if e: e[2].record()
    for _ in range(5): step()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(G): step(ev[i])
    ks, ss = [], []
    for _ in range(6):
        g.replay(); torch.cuda.synchronize()
        ks += [e[1].elapsed_time(e[2]) for e in ev]; ss += [e[0].elapsed_time(e[1]) for e in ev]
    print("%-18s step_kernel %.2f us (median %.2f)  sample %.2f us" % (name, 1e3 * np.mean(ks[G:]), 1e3 * np.median(ks[G:]), 1e3 * np.mean(ss[G:])))
