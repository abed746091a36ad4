"""TEST INFRASTRUCTURE: ctypes front-end of oracle/libgraphenvs_oracle.so (the CPU checker).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (graphenvs_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgraphenvs_oracle.so")

KINDS = {
    "ShortestPath-v0": 0, "LongestPath-v0": 1, "SteinerTree-v0": 2, "TSP-v0": 3,
    "MaxIndependentSet-v0": 4, "DensestSubgraph-v0": 5, "MulticastRouting-v0": 6,
    "DistributionCenter-v0": 7, "PerishableProductDelivery-v0": 8,
}
EDGE_ACTION = {2, 6}


class OStep(C.Structure):
    _fields_ = [("reward", C.c_double), ("solution_cost", C.c_double), ("heuristic", C.c_double),
                ("done", C.c_int), ("solved", C.c_int), ("has_mask", C.c_int), ("status", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "graphenvs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libgraphenvs_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        P = C.c_void_p
        L.oenv_create.restype = P
        L.oenv_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, P, P, P, P, P, P, P, P]
        L.oenv_free.argtypes = [P]
        L.oenv_mask.argtypes = [P, P]
        L.oenv_mask_len.argtypes = [P]
        L.oenv_step.argtypes = [P, C.c_int, C.POINTER(OStep), P]
        L.oenv_obs.argtypes = [P, P]
        L.oenv_obs_len.argtypes = [P]
        L.oenv_nodes.restype = C.POINTER(C.c_float)
        L.oenv_nodes.argtypes = [P]
        L.oenv_edges.restype = C.POINTER(C.c_float)
        L.oenv_edges.argtypes = [P]
        L.oenv_F.argtypes = [P]
        L.oenv_Fe.argtypes = [P]
        L.oenv_reset_state.argtypes = [P]
        L.oenv_features.argtypes = [P, C.c_int, P]
        L.oenv_mst_weight.restype = C.c_double
        L.oenv_mst_weight.argtypes = [P]
        L.oenv_sssp_pair.restype = C.c_double
        L.oenv_sssp_pair.argtypes = [P, C.c_int, C.c_int]
        L.oenv_sssp_all.argtypes = [P, C.c_int, C.c_double, C.c_int, P]
        L.oenv_sample_action.argtypes = [P, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32]
        L.oenv_rollout.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, P, P, P]
        L.oenv_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """One environment instance in the oracle.  `links` int32[M,2] in reference order."""

    def __init__(self, kind, N, links, w64, parenting=-1, features=None, src=0, dest=0, n_dests=0,
                 n_choices=0, max_distance=0.0, heuristic=0.0, dests=None, node_cost=None, node_xy=None):
        L = lib()
        self.kind, self.N = int(kind), int(N)
        links = np.ascontiguousarray(links, dtype=np.int32).reshape(-1, 2)
        self.M = links.shape[0]
        w64 = np.ascontiguousarray(w64, dtype=np.float64)
        feats = None if features is None else np.ascontiguousarray(features, dtype=np.float32)
        dests = None if dests is None else np.ascontiguousarray(dests, dtype=np.int32)
        n_targets = 0
        if self.kind == 7:
            n_targets = 0 if dests is None else dests.shape[0]
        if self.kind == 8:      # dests = pickups then dropoffs; n_dests = n_products; max_distance = delivery_time
            n_dests = dests.shape[0] // 2
        ip = np.array([src, dest, n_dests, n_choices, n_targets, 0], dtype=np.int32)
        dp = np.array([max_distance, heuristic], dtype=np.float64)
        nc = None if node_cost is None else np.ascontiguousarray(node_cost, dtype=np.float64)
        xy = None if node_xy is None else np.ascontiguousarray(node_xy, dtype=np.float64)
        self._keep = (links, w64, feats, dests, nc, xy)
        self.h = L.oenv_create(self.kind, self.N, self.M, int(parenting), _ptr(links), _ptr(w64), _ptr(feats),
                               _ptr(ip), _ptr(dp), _ptr(dests), _ptr(nc), _ptr(xy))
        self.A = L.oenv_mask_len(self.h)
        self.F, self.Fe = L.oenv_F(self.h), L.oenv_Fe(self.h)

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.oenv_free(self.h)
            self.h = None

    def mask(self, reset_patch=False):
        m = np.zeros(self.A, dtype=np.uint8)
        lib().oenv_mask(self.h, _ptr(m))
        if reset_patch and self.kind == 3 and m.sum() == 0:  # tsp.py:154-155
            m[0] = 1
        return m.astype(bool)

    def step(self, a):
        r = OStep()
        m = np.zeros(self.A, dtype=np.uint8)
        lib().oenv_step(self.h, int(a), C.byref(r), _ptr(m))
        return {"reward": r.reward, "done": bool(r.done), "solved": r.solved, "has_mask": bool(r.has_mask),
                "status": r.status, "solution_cost": r.solution_cost, "heuristic": r.heuristic,
                "mask": m.astype(bool) if r.has_mask else None}

    def obs(self):
        out = np.zeros(lib().oenv_obs_len(self.h), dtype=np.float32)
        lib().oenv_obs(self.h, _ptr(out))
        return out

    @property
    def nodes(self):
        return np.ctypeslib.as_array(lib().oenv_nodes(self.h), shape=(self.N, self.F)).copy()

    @property
    def edges(self):
        return np.ctypeslib.as_array(lib().oenv_edges(self.h), shape=(self.M, self.Fe)).copy()

    def reset_state(self):
        lib().oenv_reset_state(self.h)

    def features64(self, weighted_pr=False):
        out = np.zeros((self.N, 5), dtype=np.float64)
        rc = lib().oenv_features(self.h, int(weighted_pr), _ptr(out))
        if rc != 0:
            raise RuntimeError("pagerank power iteration failed to converge")
        return out

    def mst_weight(self):
        return lib().oenv_mst_weight(self.h)

    def sssp(self, s, cutoff=None):
        d = np.zeros(self.N, dtype=np.float64)
        lib().oenv_sssp_all(self.h, int(s), 0.0 if cutoff is None else float(cutoff), int(cutoff is not None), _ptr(d))
        return d


def sample_action(mask, seed, env, t):
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    return lib().oenv_sample_action(_ptr(m), m.shape[0], int(seed), int(env), int(t))


def set_threads(n=0):
    """OpenMP team size of rollout(): n <= 0 -> every CPU this process may run on.  Returns the size in use."""
    if n <= 0:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return int(lib().oenv_set_threads(int(n)))


def rollout(envs, n_steps, seed, env_id0=0, t0=0):
    """Random-valid-policy rollout with auto-reset over a list of OracleEnv (OpenMP over envs)."""
    n = len(envs)
    arr = (C.c_void_p * n)(*[e.h for e in envs])
    sr = np.zeros(n, dtype=np.float64)
    ep = np.zeros(n, dtype=np.int64)
    cs = np.zeros(n, dtype=np.uint64)
    lib().oenv_rollout(arr, n, int(env_id0), int(n_steps), int(seed), int(t0), _ptr(sr), _ptr(ep), _ptr(cs))
    return sr, ep, cs
