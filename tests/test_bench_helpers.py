"""CPU: the parts of bench.py that do not need a GPU -- workload table, compulsory-bytes model, clock-sample
parsing, the host policy, and the reference arm's JSON contract (run for real on a tiny budget)."""
import ctypes as C
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from graphenvs_b200 import _native  # noqa: E402
from graphenvs_b200.spec import ENV_SPECS  # noqa: E402


def _fake_env(wl):
    env_id, N, E, kw, B, survey, desc = bench.WORKLOADS[wl]
    d = _native.GeBatch()
    d.kind, d.B, d.N, d.M = ENV_SPECS[env_id].kind, 4, N, 2 * E
    d.parenting = kw.get("parenting", -1)
    d.n_targets = kw.get("target_count", 0)
    d.n_dests = kw.get("n_dests", kw.get("n_products", 0))
    _native.lib().ge_fill_layout(C.byref(d))
    return types.SimpleNamespace(desc=d, N=N, M=2 * E, env_id=env_id, t={"mask_bytes": 1})


def test_workloads_cover_every_baseline_config_and_env():
    ids = {w[0] for w in bench.WORKLOADS.values()}
    assert ids == set(ENV_SPECS), "every env of north_star has a bench workload"
    assert bench.DEFAULT_WORKLOAD == "cfg2_longest_path"
    env_id, N, E, kw, B, _, _ = bench.WORKLOADS[bench.DEFAULT_WORKLOAD]
    assert (env_id, N, E, kw["parenting"], B) == ("LongestPath-v0", 50, 200, 2, 65536)   # BASELINE.json configs[1]


def test_compulsory_bytes_model_is_positive_and_below_the_survey_figure_for_incremental_kernels():
    for wl, spec in bench.WORKLOADS.items():
        b = bench.layout_bytes_per_step(_fake_env(wl))
        assert b > 50, wl
        if wl in ("cfg3_mst", "cfg5_multicast", "cfg4_tsp_p2", "cfg2_longest_path"):
            assert b < spec[5], (wl, b, spec[5])


def test_clock_sampler_parses_nvidia_smi_rows():
    s = bench.ClockSampler(0)
    s.proc = types.SimpleNamespace(terminate=lambda: None)
    rows = ["0, 1965, 1965, 380.5, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
            "0, 1950, 1965, 401.0, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
            "garbage"]
    s.rows = [(100.0 + i, r) for i, r in enumerate(rows)]
    out = s.stop(99.0, 110.0)
    assert out["sm_mhz"] == 1957.5 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 2


def test_host_policy_picks_valid_actions():
    rng = np.random.default_rng(0)
    mask = rng.random((500, 37)) < 0.2
    mask[:, 5] = True
    a = bench.host_policy(rng, mask)
    assert a.dtype == np.int32 and mask[np.arange(500), a].all()
    assert len(np.unique(a)) > 5


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="1")
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "5", "--warmup", "3",
                                   "--workload", "cfg1_shortest_path"], env=env).decode().strip().splitlines()[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["steps"] == 5 and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 stay silent
    env["RANK"] = "1"
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"], env=env)
    assert out.decode().strip() == ""
