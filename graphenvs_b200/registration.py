"""Env ids and factories (graph_envs/__init__.py:9-56).

make(id, **kw)                   -> GraphEnv        same ids / kwargs as gym.make on the reference
make_batched(id, num_envs, **kw) -> BatchedGraphEnv the batched vector-env entry point
register_with_gymnasium()        registers the 8 ids with gymnasium when it is importable, so
                                 `gym.make('<Env>-v0', ...)` resolves to this engine.
"""
from .spec import ENV_SPECS

registry = {env_id: "graphenvs_b200.registration:_entry_%d" % spec.kind for env_id, spec in ENV_SPECS.items()}


def make(env_id, **kwargs):
    if env_id not in ENV_SPECS:
        raise KeyError("No registered env with id: %s" % env_id)
    from .single import GraphEnv
    return GraphEnv(env_id, **kwargs)


def make_batched(env_id, num_envs, **kwargs):
    if env_id not in ENV_SPECS:
        raise KeyError("No registered env with id: %s" % env_id)
    from .batch import BatchedGraphEnv
    return BatchedGraphEnv(env_id, num_envs, **kwargs)


def _entry(env_id):
    def ctor(**kwargs):
        return make(env_id, **kwargs)
    ctor.__name__ = "make_" + env_id.replace("-", "_")
    return ctor


for _id, _spec in ENV_SPECS.items():
    globals()["_entry_%d" % _spec.kind] = _entry(_id)


def register_with_gymnasium():
    """Returns True when the ids were registered with an importable gymnasium."""
    try:
        from gymnasium.envs.registration import register, registry as gym_registry
    except Exception:
        return False
    for env_id, entry in registry.items():
        if env_id not in gym_registry:
            register(id=env_id, entry_point=entry, disable_env_checker=True, order_enforce=False)
    return True
