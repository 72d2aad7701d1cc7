"""Load the reference PPO classes (Model_PPO, Env_rollout, Algo_PPO) UNMODIFIED from
`Coop-MH-PPO-scalable.py` (build container only; TEST INFRASTRUCTURE).

The script is a notebook export: class definitions (PY:42-1001) followed by an interactive driver
(`input()` prompts, PY:1003+).  Only the definitions are executed; matplotlib (not installed, used
only by the analysis cells) is stubbed.  The classes read the globals `env`, `nb_lines` that the
driver cell would have set (PY:125, 111); `make_algo` provides them.
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _HERE)
import refdriver as rd  # noqa: E402

PY = os.path.join(rd.REF_ROOT, "Coop-MH-PPO-scalable.py")


NB_COOP = os.path.join(rd.REF_ROOT, "Coop-MH-PPO.ipynb")     # on the coop env class
NB_NAIF = os.path.join(rd.REF_ROOT, "MH-PPO.ipynb")          # on the naif env class (same PPO cell)


def _source_lines(path):
    """The PPO definitions: the script itself, or the first code cell of a notebook."""
    if path.endswith(".ipynb"):
        import json
        cell = next(c for c in json.load(open(path))["cells"] if c["cell_type"] == "code")
        return "".join(cell["source"]).split("\n")
    return open(path).read().split("\n")


def load_namespace(path=None):
    path = path or PY
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.rcParams = {}
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    if rd.REF_ROOT not in sys.path:
        sys.path.insert(0, rd.REF_ROOT)
    src = _source_lines(path)
    cut = next(i for i, l in enumerate(src) if l.startswith("#5) Computing part"))
    ns = {"__name__": "ref_ppo"}
    exec(compile("\n".join(src[:cut]), path, "exec"), ns)
    return ns


def make_algo(ns, variant, nb_car, nb_ped, nb_lines, seed=0):
    """Algo_PPO exactly as the driver cell builds it (PY:1026-1047)."""
    import torch
    env = rd.make_env(variant, nb_car, nb_ped, nb_lines)
    ns["env"] = env
    ns["nb_lines"] = nb_lines
    ns["nb_car"] = nb_car                                # global read by the older notebooks' Env_rollout.__init__
    ns["print"] = lambda *a, **k: None
    torch.manual_seed(seed)
    num_states_c = 2 + 9 + 2
    num_states_d = 2 + (6 * (2 * nb_lines - 1)) + 8 + 2
    if variant != "coop_scalable":                       # the older notebooks' driver cell: 5 columns per other car
        num_states_d = 2 + (5 * (nb_car - 1)) + 8 + 2
    mean = (rd.CAR_B[1, 0] + rd.CAR_B[0, 0]) / 2.0
    std = (rd.CAR_B[1, 0] - rd.CAR_B[0, 0]) / 2.0
    env._mh_rng.set_stream(seed, 0, 0)
    algo = ns["Algo_PPO"](ns["Model_PPO"], env, num_algo=100 * nb_ped + 10 * nb_car + nb_lines, num_states_c=num_states_c,
                          num_states_d=num_states_d, num_actions=1, mean=mean, std=std, nb_cars=nb_car, dt=0.3)
    return algo, env
