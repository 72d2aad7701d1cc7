"""The C-ABI library loads on a machine without a GPU and exports every symbol include/mhppo.h
declares; with no device every compute entry point refuses (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mhppo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mhppo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import mhppo_b200
    from mhppo_b200 import _lib
    L = mhppo_b200.lib()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(L, s), "libmhppo_b200.so does not export %s" % s
    assert sorted(_lib.SYMBOLS) == syms, "python binding and header disagree"
    assert L.mhppo_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mhppo_b200
    from mhppo_b200._lib import EnvCfg
    L = mhppo_b200.lib()
    assert L.mhppo_device_count() == 0
    cfg = EnvCfg(variant=5, nb_car=4, nb_ped=3, nb_lines=2, max_episode=80, sin_model=1, device=0, dt=0.3, seed=1, n_envs=8)
    h = C.c_void_p()
    assert L.mhppo_env_create(C.byref(cfg), C.byref(h)) == -2  # MHPPO_ENODEV
    assert b"no CPU fallback" in L.mhppo_last_error()
    with pytest.raises(Exception):
        mhppo_b200.make("Crosswalk_hybrid_multi_coop_scalable-v0", 8, nb_car=4, nb_ped=3, nb_lines=2)


def test_create_rejects_bad_configs():
    import mhppo_b200
    from mhppo_b200._lib import EnvCfg
    L = mhppo_b200.lib()
    h = C.c_void_p()
    bad = EnvCfg(variant=9, nb_car=1, nb_ped=1, nb_lines=1, max_episode=80, dt=0.3, n_envs=4)
    assert L.mhppo_env_create(C.byref(bad), C.byref(h)) == -1
    bad = EnvCfg(variant=5, nb_car=5, nb_ped=1, nb_lines=2, max_episode=80, dt=0.3, n_envs=4)  # nb_car > 2*nb_lines
    assert L.mhppo_env_create(C.byref(bad), C.byref(h)) == -1
    bad = EnvCfg(variant=2, nb_car=1, nb_ped=1, nb_lines=1, max_episode=80, dt=0.3, n_envs=0)
    assert L.mhppo_env_create(C.byref(bad), C.byref(h)) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mh-ppo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "mhppo_oracle" not in txt, f
