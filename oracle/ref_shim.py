"""Test-only import shim for the upstream reference (oracle infrastructure, NOT product code).

Lets the unmodified reference modules under ``/root/reference`` import in this container, where
``gym``/``gymnasium``/``gym3``/``matplotlib``/``procgen``/... are absent.  Used only by
``oracle/mint_golden.py`` (to mint ``tests/golden/*.npz``) and by ``tests/test_oracle_vs_reference.py``
(skipped when ``/root/reference`` is missing, i.e. on the GPU box).  Nothing here is importable from
the product package.

What it fakes: ``spaces.Box/Discrete``, ``gymnasium.Env``,
``gymnasium.utils.seeding.np_random`` (= ``Generator(PCG64(SeedSequence(seed)))``, gymnasium 0.29.1),
``gym3.env.Env``; aliases ``np.float_`` (removed in numpy 2, used at
``discrete_env/acrobot_pre_vec.py:523``); lets ``random.sample`` accept a ``set`` the way
Python <= 3.10 did (``boxworld/boxworld_gen_vec.py:9,19,21``).
"""
import importlib.abc
import importlib.machinery
import os
import random as _random
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TPP_REFERENCE_ROOT", "/root/reference")

_FAKE_TOPLEVEL = (
    "gym", "gymnasium", "gym3", "matplotlib", "moviepy", "procgen", "vector_quantize_pytorch",
    "torchinfo", "pygame", "Box2D", "mujoco", "optuna", "pysr", "imageio", "wandb", "mbrl", "omegaconf",
)


class _Anything:
    """Stand-in class for any symbol imported from a faked module."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _FakeModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[0].islower():
            # `from gymnasium import spaces` style: lower-case names are treated as sub-modules
            return importlib.import_module(f"{self.__name__}.{name}")
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


class _FakeFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _FAKE_TOPLEVEL and fullname not in sys.modules:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _FakeModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        _populate(module)


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=dtype)
            high = np.full(shape, high, dtype=dtype)
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class GymnasiumEnv:
    """gymnasium.Env essentials: the `np_random` property backed by `_np_random`."""
    _np_random = None
    metadata = {}

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, _ = np_random_seeding(None)
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value


class Gym3Env:
    """gym3.env.Env essentials (gym3 0.3.3): stores spaces and `num`."""

    def __init__(self, ob_space=None, ac_space=None, num=1):
        self.ob_space, self.ac_space, self.num = ob_space, ac_space, num


def np_random_seeding(seed=None):
    """gymnasium.utils.seeding.np_random (0.29.1): PCG64 over a SeedSequence."""
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    return np.random.Generator(np.random.PCG64(seed_seq)), np_seed


def _populate(module):
    name = module.__name__
    if name in ("gym.spaces", "gymnasium.spaces"):
        module.Box, module.Discrete = Box, Discrete
    if name in ("gym", "gymnasium"):
        module.Env = GymnasiumEnv
    if name == "gymnasium.utils.seeding":
        module.np_random = np_random_seeding
    if name == "gym3.env":
        module.Env = Gym3Env


_installed = False


def install():
    """Idempotently install the fakes and put the reference FIRST on sys.path (its top-level package
    names `agents`/`common`/`utils` collide with unrelated site-packages; ours are all under
    `tpp_b200`/`oracle`, so nothing of this repo is shadowed)."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.meta_path.insert(0, _FakeFinder())
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    _orig_sample = _random.sample

    def _sample_compat(population, k, **kw):
        if isinstance(population, (set, frozenset)):
            population = tuple(population)
        return _orig_sample(population, k, **kw)

    _random.sample = _sample_compat
    sys.path.insert(0, REFERENCE_ROOT)
    for _m in ("agents", "common", "utils"):
        sys.modules.pop(_m, None)
    _installed = True


def available():
    return os.path.isdir(REFERENCE_ROOT)


def load(modname):
    """Import a reference module by its dotted name, e.g. ``discrete_env.cartpole_pre_vec``."""
    install()
    return importlib.import_module(modname)
