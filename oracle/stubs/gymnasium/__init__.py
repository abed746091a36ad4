"""Minimal stand-in for `gymnasium`, used ONLY by oracle/ref_loader.py to import the
unmodified reference from /root/reference when the real package is absent.
TEST INFRASTRUCTURE - never imported by the product package."""
import importlib
from . import spaces
from .envs import registration

class Env:
    def reset(self, seed=None, options=None):
        return None

def make(id, **kwargs):
    entry = registration.registry[id]
    mod, cls = entry.split(":")
    return getattr(importlib.import_module(mod), cls)(**kwargs)
