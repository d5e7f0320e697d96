"""CPU: the C-ABI library loads, exports every symbol include/tpp_b200.h declares, validates arguments without
touching a GPU, and its host-side level generator equals the oracle (no device compute here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle.boxworld import world_gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tpp_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_symbols_are_exported_and_bound(lib):
    header = open(os.path.join(ROOT, "include", "tpp_b200.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(tpp_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 20
    cdll = lib.load()
    for name in declared:
        assert hasattr(cdll, name), f"{name} declared in the header but not exported"
    assert declared - {"tpp_error_string"} == set(lib.SIGNATURES), "ctypes binding out of sync with the header"
    assert cdll.tpp_version() == lib.ABI_VERSION


def test_struct_sizes_match_the_header(lib):
    assert ctypes.sizeof(lib.EnvCfg) == 4 * 4 + 8 + 16 * 4 * 2 + 8 * 4
    assert ctypes.sizeof(lib.BoxWorldState) == 4 * 4 + 8 + 4 * 4 + 11 * 8
    assert ctypes.sizeof(lib.LossCfg) == 8 * 4 + 8
    assert ctypes.sizeof(lib.AdamState) == 4 * 8 + 2 * 4 + 2 * 4 + 2 * 8


def test_null_and_bad_arguments_are_rejected_without_a_gpu(lib):
    cdll = lib.load()
    assert cdll.tpp_gae(None, None, None, None, None, None, 4, 4, 4, 0.99, 0.95, None) == 10001
    assert cdll.tpp_env_step(None, None, None, None, None, None, None, None, None, None, 0, 0, None) == 10001
    assert cdll.tpp_gemm_f32(None, 1, 1, None, 1, 1, None, 1, None, None, 1, 1, 1, 0, 1, None) == 10001
    with pytest.raises(lib.TppError):
        lib.call("tpp_adam_clip_step", None, None, None, None, None, 0, None)
    assert b"TPP_EINVAL" in cdll.tpp_error_string(10001)


@pytest.mark.parametrize("spec", [(6, 2, 1, 1), (12, 5, 3, 3), (12, 4, 2, 2), (9, 3, 2, 2), (12, 5, 0, 0)])
def test_host_level_generator_equals_oracle(lib, spec):
    from tpp_b200.boxworld.box_world_env_vec import generate_levels_host
    for seed0 in (0, 6033, 2 ** 32 - 3, 2 ** 40 + 1):
        w, d, p = generate_levels_host(*spec, seed0, 24)
        for i in range(24):
            ow, op, od = world_gen(*spec, seed0 + i)
            assert np.array_equal(w[i], ow) and np.array_equal(p[i], op) and np.array_equal(d[i], od.astype(np.int8))


def test_generator_rejects_unsupported_specs(lib):
    with pytest.raises(lib.TppError):
        lib.call("tpp_boxworld_gen_levels_host", 40, 2, 1, 1, 0, 1, 1, 1, 1)


def test_product_fails_loudly_without_cuda(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tpp_b200.common.storage import Storage
    with pytest.raises(lib.TppError):
        Storage((9,), 4, 8, 8, "cpu")


def test_host_randperm_equals_torch_randperm_and_leaves_the_same_generator_state(lib):
    """Minibatch indices must be bit-exact (reference common/storage.py:87: torch.randperm on the default CPU
    generator): the library's MT19937 + Fisher-Yates restatement against torch itself, including the generator state
    it leaves behind and draws interleaved with other torch CPU random numbers."""
    import torch
    from tpp_b200.common.storage import Storage
    for seed in (0, 1, 6033, 4321):
        sizes = (10, 700, 5, 1 << 16, 3, 1000, 64, 65, 2, 1, 625, 624)
        torch.manual_seed(seed)
        want = [torch.randperm(n) for n in sizes]
        s_want = torch.get_rng_state()
        torch.manual_seed(seed)
        got = [Storage.randperm(n) for n in sizes]
        assert all(torch.equal(a, b) for a, b in zip(want, got))
        assert torch.equal(s_want, torch.get_rng_state())
        torch.manual_seed(seed)
        a1, x1, b1 = torch.randperm(1000), torch.rand(5), torch.randperm(77)
        torch.manual_seed(seed)
        a2, x2, b2 = Storage.randperm(1000), torch.rand(5), Storage.randperm(77)
        assert torch.equal(a1, a2) and torch.equal(x1, x2) and torch.equal(b1, b2)
    out = torch.empty(4096, dtype=torch.int64)
    torch.manual_seed(7)
    w = torch.randperm(4096)
    torch.manual_seed(7)
    assert Storage.randperm(4096, out=out) is out and torch.equal(out, w)
    # int32 form (what PPO.train uploads: half the pinned bytes): same permutation, same generator state afterwards
    out32 = torch.empty(1 << 20, dtype=torch.int32)
    torch.manual_seed(8)
    w, after = torch.randperm(1 << 20), torch.rand(3)
    torch.manual_seed(8)
    Storage.randperm(1 << 20, out=out32)
    assert torch.equal(out32.long(), w) and torch.equal(torch.rand(3), after)


def test_weight_gradient_split_policy():
    """Host logic of the engine: k-splits of a weight-gradient GEMM keep <= ~1024 samples per accumulator, fill at
    least one wave of the 148 SMs and do not leave a nearly empty last wave."""
    from tpp_b200.common.engine import MLPEngineTC
    for M in (256, 2048, 8192, 32768, 131072, 1 << 20):
        for ctas in (1, 2, 4, 6, 10):
            s = MLPEngineTC._wgrad_split(M, ctas)
            assert 1 <= s <= max(1, -(-M // 32))
            if M >= 148 * 32:
                assert s * ctas >= 148                                    # at least one wave
                assert -(-M // s) <= 1024 + 32                            # accumulation length bound
                waves = -(-s * ctas // 148)
                assert s * ctas > (waves - 1) * 148 + 148 // 2 or waves == 1   # last wave at least half full


def test_accumulation_group_size_policy():
    """Host logic of PPO.optimize: how many minibatches of a gradient-accumulation window share one pass."""
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.engine import ImpalaEngineTC, MLPEngine
    agent = PPO.__new__(PPO)
    agent.fuse_accum, agent.x_entropy_coef, agent.max_group_rows = "auto", 0.0, 1 << 18
    mlp, impala = object.__new__(MLPEngine), object.__new__(ImpalaEngineTC)
    assert agent._group_size(16, 128, 8192, mlp) == 16            # the bench workload: the whole window
    assert agent._group_size(16, 128, 8192, impala) == 1          # convolution engine: per minibatch
    assert agent._group_size(1, 8, 4096, mlp) == 1                # no accumulation
    assert agent._group_size(16, 120, 8192, mlp) == 1             # windows do not tile the epoch
    assert agent._group_size(4, 8, 100, mlp) == 1                 # grouped loss kernel needs mb % 256 == 0
    agent.max_group_rows = 40000
    assert agent._group_size(16, 128, 8192, mlp) == 4             # row budget: largest divisor that fits
    agent.max_group_rows, agent.fuse_accum = 1 << 18, 2
    assert agent._group_size(16, 128, 8192, mlp) == 2
    agent.fuse_accum = 1
    assert agent._group_size(16, 128, 8192, mlp) == 1
    agent.fuse_accum, agent.x_entropy_coef = "auto", 0.05
    assert agent._group_size(16, 128, 8192, mlp) == 1             # cross-batch entropy keeps the ungrouped path


def test_shape_specialised_entries_refuse_other_shapes_without_a_gpu():
    """The FMA-pipe convolution entries check their shape table before anything touches the device: unsupported
    shapes answer TPP_ENOTSUP (``_lib.try_call`` -> False, the caller takes the tensor-core form), bad arguments
    TPP_EINVAL (TppError).  Dummy non-null pointers: no launch happens on either path."""
    from tpp_b200 import _lib
    one = ctypes.c_void_p(16)
    assert _lib.try_call("tpp_conv3x3_wgrad", one, 0, one, one, 2, 14, 14, 16, 16, None) is False
    assert _lib.try_call("tpp_conv3x3_fma", one, 0, one, 16, None, None, None, 0, one, None, None, None, 2, 14, 14, 16,
                         16, None) is False
    assert _lib.try_call("tpp_conv3x3_wgrad_first", one, 1, 1, 1, one, one, 2, 32, 32, 16, None) is False
    assert _lib.try_call("tpp_conv3x3_fwd_first", one, 1, 1, 1, one, one, one, 2, 32, 32, 16, None) is False
    with pytest.raises(_lib.TppError):
        _lib.try_call("tpp_conv3x3_wgrad", None, 0, one, one, 2, 32, 32, 16, 16, None)


def test_tile_and_kernel_selection_policies():
    """Host logic of the engines: which tile a dense GEMM gets (lean persistent pairs unless the contraction is too short
    to hide a lean epilogue) and which convolutions leave the tensor core for the FMA-pipe kernels."""
    from tpp_b200 import _lib
    from tpp_b200.common.engine import ImpalaEngineTC, MLPEngineTC
    e = object.__new__(MLPEngineTC)
    e.pair_min_n, e.wide_tile_rows, e.pair_block_n = 256, 8192, _lib.TC_TILE_PAIR_PERSISTENT
    e.lean_kinds, e.lean_min_k, e.precision, e.small_tile_elems = ("fwd", "dgrad", "wgrad"), 128, 3, 148 * 128 * 128 // 2
    assert e._bn(131072, 256, "fwd") == _lib.TC_TILE_PAIR_PERSISTENT_LEAN
    assert e._bn(131072, 256, "dgrad", K=256) == _lib.TC_TILE_PAIR_PERSISTENT_LEAN
    assert e._bn(131072, 256, "dgrad", K=64) == _lib.TC_TILE_PAIR_PERSISTENT          # 16 epilogue warps
    assert e._bn(131072, 64, "fwd") == _lib.TC_TILE_PAIR64_PERSISTENT
    assert e._bn(4096, 256, "fwd") == 64 and e._bn(256, 64, "fwd") == 0
    e.precision = 1
    assert e._bn(131072, 256, "fwd") == _lib.TC_TILE_PAIR_PERSISTENT                    # lean exists for 3xTF32 only
    i = object.__new__(ImpalaEngineTC)
    i.wgrad_cc = True
    first = dict(implicit=False, cin=3, cout=16)
    assert i._first_cc(first, 64, 64, (12288, 64, 1, 4096))
    assert not i._first_cc(first, 14, 14, (588, 14, 1, 196))                            # Box-World frames: col + GEMM
    assert not i._first_cc(first, 64, 64, (12288, 64, 2, 4096))                         # x stride must be 1
    assert not i._first_cc(dict(implicit=True, cin=16, cout=16), 64, 64, (1, 1, 1, 1))
    i.wgrad_cc = False
    assert not i._first_cc(first, 64, 64, (12288, 64, 1, 4096))
