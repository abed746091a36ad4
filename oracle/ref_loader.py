"""TEST INFRASTRUCTURE (oracle side). Imports the UNMODIFIED reference package from
/root/reference behind stub `gymnasium` / `matplotlib` / `torch_geometric` packages
(SURVEY.md Appendix B).  Only usable in the build container - /root/reference does not
exist on the GPU box, so nothing under tests -m gpu / bench.py / smoke() may import this.
It is used by oracle/gen_golden.py to produce the committed fixtures in tests/golden/.
"""
import importlib
import os
import sys
import warnings

REF_ROOT = os.environ.get("GRAPHENVS_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "graph_envs"))


def load():
    """Return (gymnasium-like module, graph_envs module) with the 9 reference envs registered."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    warnings.filterwarnings("ignore")
    for name in ("gymnasium", "matplotlib", "torch_geometric"):
        try:
            importlib.import_module(name)
        except Exception:
            if _STUBS not in sys.path:
                sys.path.insert(0, _STUBS)
    if REF_ROOT not in sys.path:
        sys.path.insert(1, REF_ROOT)
    gym = importlib.import_module("gymnasium")
    ge = importlib.import_module("graph_envs")
    return gym, ge
