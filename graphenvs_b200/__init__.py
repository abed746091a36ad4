"""graphenvs_b200 -- B200-native batched engine for the GraphEnvs hot path (step / mask / obs).

Public surface:
  make(id, **kwargs)                 gymnasium-shaped single env (reference ids and kwargs)
  make_batched(id, num_envs, ...)    batched vector env with the same per-env semantics
  BatchedGraphEnv                    the engine class behind both
  utils.vectorize_graph / devectorize_graph / get_env_info   (graph_envs/utils.py layout contract)
"""
from .spec import ENV_SPECS, get_env_info, get_num_features  # noqa: F401

name = "graphenvs_b200"
__version__ = "0.1.0"


def __getattr__(attr):  # lazy: keeps `import graphenvs_b200` cheap and torch-free until needed
    if attr in ("BatchedGraphEnv",):
        from .batch import BatchedGraphEnv
        return BatchedGraphEnv
    if attr in ("make", "make_batched", "register_with_gymnasium", "registry"):
        from . import registration
        return getattr(registration, attr)
    if attr in ("Instance", "generate_instance"):
        from . import instances
        return getattr(instances, attr)
    raise AttributeError(attr)
