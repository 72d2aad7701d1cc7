"""`register` / `make` of the gym shim (test infrastructure only).

No wrappers are applied: gym 0.26 would wrap in PassiveEnvChecker ->
OrderEnforcing -> TimeLimit(100) (Environments/__init__.py:6), none of which
changes a value on the hot path (episodes end at step 80 < 100).
"""
import importlib

_REGISTRY = {}


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)


def make(id, **kwargs):
    entry_point, _ = _REGISTRY[id]
    mod_name, cls_name = entry_point.split(":")
    mod = importlib.import_module(mod_name)
    return getattr(mod, cls_name)(**kwargs)
