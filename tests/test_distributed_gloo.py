"""CPU, world_size 2 over gloo: the host-side sharding logic (tpp_b200/parallel.py).  Checks the two claims the
multi-GPU design rests on: (1) all-reduced advantage moments give every shard the reference's GLOBAL
normalisation, (2) summed shard gradients x 1/world equal the gradient of the full minibatch, so clip+Adam after the
all-reduce keeps single-process semantics."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ppo as oppo


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from tpp_b200 import parallel
    r, _, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    T, N, A = 12, 16, 3
    g = torch.Generator().manual_seed(0)                      # identical global data on both ranks
    rew, value = torch.randn(T, N, generator=g), torch.randn(T + 1, N, generator=g)
    done = (torch.rand(T, N, generator=g) < 0.1).float()
    obs = torch.randn(T, N, 9, generator=g)
    act = torch.randint(0, A, (T, N), generator=g).float()
    old_logp = -torch.rand(T, N, generator=g) - 0.3
    lo, hi = parallel.shard_envs(N, rank, world)
    assert hi - lo == N // world
    # (1) global advantage normalisation from per-shard moments
    adv, ret = oppo.gae(rew[:, lo:hi], done[:, lo:hi], value[:, lo:hi])
    m = torch.tensor([adv.double().sum(), (adv.double() ** 2).sum(), float(adv.numel())], dtype=torch.float64)
    parallel.allreduce_moments_(m)
    mean, std = parallel.mean_std_from_moments(m)
    adv_n = (adv - mean) / (std + 1e-8)
    adv_full, _ = oppo.gae(rew, done, value)
    want = oppo.normalize_adv(adv_full)[:, lo:hi]
    np.testing.assert_allclose(adv_n.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)
    # (2) gradient of the union minibatch == mean of shard gradients
    torch.manual_seed(1)
    pol = oppo.OraclePolicy(oppo.OracleMLP(9, 4, 32, 16), A)

    def grad_of(sl):
        pol.zero_grad()
        d, v, _ = pol(obs[:, sl].reshape(-1, 9))
        loss, _ = oppo.ppo_loss(d, v, act[:, sl].reshape(-1), old_logp[:, sl].reshape(-1),
                                value[:-1, sl].reshape(-1), ret_full[:, sl].reshape(-1), adv_norm_full[:, sl].reshape(-1))
        loss.backward()
        return torch.cat([p.grad.reshape(-1) for p in pol.parameters()])
    ret_full = adv_full + value[:-1]
    adv_norm_full = oppo.normalize_adv(adv_full)
    flat = grad_of(slice(lo, hi)).clone()
    parallel.allreduce_gradients_(flat)
    flat *= 1.0 / world
    full = grad_of(slice(0, N))
    # pi/value terms are means over the minibatch -> shard mean == full gradient (entropy term likewise)
    np.testing.assert_allclose(flat.numpy(), full.numpy(), rtol=2e-4, atol=2e-7)
    hp = parallel.shard_hyperparameters(dict(n_envs=4096, mini_batch_size=8192, n_steps=256), world)
    assert hp["n_envs"] == 2048 and hp["mini_batch_size"] == 4096
    assert parallel.rank_seed(6033, rank, 500) == 6033 + 500 * rank
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_sharding_semantics(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]


def test_shard_envs_rejects_uneven_split():
    from tpp_b200 import parallel
    with pytest.raises(ValueError):
        parallel.shard_envs(10, 0, 4)
    assert parallel.shard_envs(4096, 3, 8) == (1536, 2048)
