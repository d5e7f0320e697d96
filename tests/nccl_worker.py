"""Worker of tests/test_distributed_nccl.py (launched by torchrun, one rank per GPU, NCCL): env-sharded
``PPO.optimize()`` -- global advantage moments, ONE gradient all-reduce per optimizer step, 1/world folded into the
clip+Adam kernel -- must equal the single-GPU ``optimize()`` on the UNION of the shards' rollouts when minibatch k of
the single-GPU run is the union of every rank's minibatch k.  Replicas must stay bit-identical."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_agent(T, N, A, device, use_graph, **kw):
    from tpp_b200.agents.ppo import PPO
    from tpp_b200.common.model import MLPModel
    from tpp_b200.common.policy import CategoricalPolicy
    from tpp_b200.common.storage import Storage
    torch.manual_seed(5)
    pol = CategoricalPolicy(MLPModel(9, 4, 64, 32), False, A).to(device).flatten_()
    st = Storage((9,), 32, T, N, device)
    agent = PPO(None, pol, None, st, device, 0, n_steps=T, n_envs=N, epoch=2, n_minibatch=4, learning_rate=1e-3,
                entropy_coef=0.01, use_cuda_graph=use_graph, **kw)
    return agent, st


def fill(st, data, lo, hi):
    n = hi - lo
    st.obs_fm[:, :, :n] = data["obs"][:, lo:hi].permute(0, 2, 1)
    st.act_i32[:, :n] = data["act"][:, lo:hi]
    st.logp[:, :n] = data["logp"][:, lo:hi]
    st.value[:, :n] = data["value"][:, lo:hi]
    st.rew[:, :n] = data["rew"][:, lo:hi]
    st.done_u8[:, :n] = data["done"][:, lo:hi]


def main():
    out_dir = sys.argv[1]
    mode = sys.argv[2] if len(sys.argv) > 2 else "eager"
    use_graph = mode in ("graph", "nccl")          # "nccl": ncclAllReduce captured in the epoch graph instead of the
    peer = mode != "nccl"                          # one-shot peer-memory all-reduce kernel
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    device = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(device))
    T, N, A = 16, 128, 3
    n_loc = N // world
    mb_loc, n_mb = 128 // world * 2, None
    g = torch.Generator().manual_seed(0)                                  # identical union data on every rank
    data = dict(obs=torch.randn(T + 1, N, 9, generator=g), act=torch.randint(0, A, (T, N), generator=g).int(),
                logp=-torch.rand(T, N, generator=g) * 1.2 - 0.3, value=torch.randn(T + 1, N, generator=g) * 0.4,
                rew=torch.randn(T, N, generator=g), done=(torch.rand(T, N, generator=g) < 0.1).to(torch.uint8))
    data = {k: v.to(device) for k, v in data.items()}
    # per-rank local minibatch indices (drawn identically everywhere so that rank 0 can build the union run)
    epochs = 2
    perms = [[torch.randperm(T * n_loc, generator=g) for _ in range(world)] for _ in range(epochs)]

    lo, hi = rank * n_loc, (rank + 1) * n_loc
    agent, st = make_agent(T, n_loc, A, device, use_graph, mini_batch_size=mb_loc, peer_reduce=peer)
    agent.shard(world)
    assert (agent.optimizer.peer is not None) == peer
    fill(st, data, lo, hi)
    calls = {"e": 0}

    def local_indices(mb):
        e = calls["e"]
        calls["e"] += 1
        return perms[e][rank][:(T * n_loc) // mb * mb].view(-1, mb).to(device)
    st.epoch_indices = local_indices
    st.compute_estimates(0.99, 0.95, True, True)
    summary = agent.optimize()
    flat = agent.policy.flat.clone()
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    in_sync = all(torch.equal(gathered[0], x) for x in gathered)
    if agent.optimizer.peer is not None:
        agent.optimizer.peer.check()
    result = {"rank": rank, "in_sync": bool(in_sync), "steps": agent.optimizer.step_count}
    if rank == 0:
        # the single-GPU run on the union: minibatch k = union over ranks of their minibatch k
        ref, sr = make_agent(T, N, A, device, use_graph, mini_batch_size=mb_loc * world)
        fill(sr, data, 0, N)
        calls_u = {"e": 0}

        def union_indices(mb):
            e = calls_u["e"]
            calls_u["e"] += 1
            rows = []
            for r in range(world):
                p = perms[e][r][:(T * n_loc) // mb_loc * mb_loc].view(-1, mb_loc)
                t, env = p // n_loc, p % n_loc
                rows.append(t * N + (r * n_loc + env))
            return torch.cat(rows, 1).to(device)
        sr.epoch_indices = union_indices
        sr.compute_estimates(0.99, 0.95, True, True)
        # (1) exact global advantage normalisation on the shard (3 doubles all-reduced per rollout)
        adv_err = float((sr.adv[:, lo:hi] - st.adv[:, :n_loc]).abs().max())
        s_ref = ref.optimize()
        a, b = flat.cpu().numpy(), ref.policy.flat.cpu().numpy()
        result.update(adv_err=adv_err, steps_ref=ref.optimizer.step_count,
                      max_abs=float(np.abs(a - b).max()), scale=float(np.abs(b).max()),
                      rel_ok=bool(np.allclose(a, b, rtol=2e-4, atol=2e-6)),
                      loss=[summary["Loss/total"], s_ref["Loss/total"]])
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump(result, f)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)       # (no destroy_process_group: CUDA graphs that captured the all-reduce are still alive)


if __name__ == "__main__":
    main()
