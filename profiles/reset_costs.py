"""Reset-time (per-instance) kernels timed with CUDA events: instance generation, derived arrays,
heuristics / in-range tables, structural features, state reset.  Prints one JSON line per workload.
    python profiles/reset_costs.py [B]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench
from graphenvs_b200 import BatchedGraphEnv, _native

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096


def timed(fn, reps=1):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for wl, (env_id, N, E, kw, _, _, desc) in bench.WORKLOADS.items():
    kw = dict(kw)
    kw["is_eval_env"] = True
    env = BatchedGraphEnv(env_id, B, N, E, structural_features=True, auto_reset=True, **kw)
    L, d = env.lib, env.desc
    t_gen = timed(lambda: env.generate(seed=3))         # includes derived arrays + prepare + features of generate()
    t_gen_only = timed(lambda: _native.check(L.ge_generate(C.byref(d), 3, env.t["row_ptr"].data_ptr(), env.t["col"].data_ptr(),
                                                           env.t["w64"].data_ptr() if "w64" in env.t else None,
                                                           env.t["w32"].data_ptr() if "w32" in env.t else None, env._stream()))
                        ) if "w64" in env.t else float("nan")
    t_feat = timed(lambda: _native.check(L.ge_features(C.byref(d), env._stream())))
    t_reset = timed(lambda: env.reset(), reps=3)
    out = {"workload": wl, "B": B, "us_per_env": {"generate+derive+prepare+features": 1e3 * t_gen / B, "generate": 1e3 * t_gen_only / B,
                                                    "features": 1e3 * t_feat / B, "reset": 1e3 * t_reset / B},
           "ms_total": {"generate_all": t_gen, "features": t_feat, "reset": t_reset}}
    print(json.dumps(out), flush=True)
    del env
    torch.cuda.empty_cache()
