"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement (torch CPU fp32, the reference's own arithmetic library) of the storage / PPO part of the path:

* rollout buffer + GAE + advantage normalisation   -> common/storage.py:21-79
* minibatch index stream                             -> common/storage.py:81-92 (SubsetRandomSampler/BatchSampler)
* recurrent (env-permuting) minibatches + GRU policy -> common/storage.py:93-110, common/model.py:212-276,
                                                        common/policy.py:49-69; NOTE agents/ppo.py:116-121: optimize()
                                                        evaluates embedder + heads WITHOUT the GRU (its call through
                                                        the policy is commented out), so the GRU only acts in predict()
* MLP / IMPALA policies and heads                    -> common/model.py:134-208, 954-980; common/policy.py:36-87
* PPO loss                                           -> agents/ppo.py:131-169, common/misc_util.py:32-51
* clip_grad_norm_ + Adam(eps=1e-5) + LR anneal       -> agents/ppo.py:58,173-176; common/misc_util.py:92-96
* one full optimize() / train-iteration loop         -> agents/ppo.py:96-279

Pinned against the live reference in tests/test_oracle_vs_reference.py (container only) and through the
fixtures tests/golden/ppo_*.npz minted from it by oracle/mint_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# Storage
# ----------------------------------------------------------------------------------------------

def gae(rew, done, value, gamma=0.99, lmbda=0.95):
    """rew, done: [T, N]; value: [T+1, N] (torch fp32). Returns (adv_raw, returns). storage.py:56-77."""
    T = rew.shape[0]
    adv = torch.zeros_like(rew)
    A = 0
    for i in reversed(range(T)):
        delta = (rew[i] + gamma * value[i + 1] * (1 - done[i])) - value[i]
        adv[i] = A = gamma * lmbda * A * (1 - done[i]) + delta
    return adv, adv + value[:-1]


def normalize_adv(adv):
    """storage.py:78-79: global mean, unbiased std."""
    return (adv - torch.mean(adv)) / (torch.std(adv) + 1e-8)


def epoch_indices(batch_size, mini_batch_size):
    """The index lists one epoch of fetch_train_generator yields (default CPU generator, drop_last)."""
    perm = torch.randperm(batch_size).tolist()
    n = batch_size // mini_batch_size
    return [perm[i * mini_batch_size:(i + 1) * mini_batch_size] for i in range(n)]


def recurrent_epoch_indices(n_steps, n_envs, mini_batch_size):
    """common/storage.py:93-110: one ``torch.randperm(num_envs)``, whole trajectories of ``num_envs_per_batch`` envs per
    minibatch, rows flattened time-major (``batch[:, idxes].reshape(-1)``).  Returns (list of flat index lists into the
    [T*N] arrays, list of env index lists)."""
    per_epoch = (n_steps * n_envs) // mini_batch_size
    envs_per_batch = n_envs // per_epoch
    perm = torch.randperm(n_envs)
    flat, envs = [], []
    for start in range(0, n_envs, envs_per_batch):
        idxes = perm[start:start + envs_per_batch]
        envs.append(idxes.tolist())
        flat.append((torch.arange(n_steps)[:, None] * n_envs + idxes[None, :]).reshape(-1).tolist())
    return flat, envs


# ----------------------------------------------------------------------------------------------
# Policies
# ----------------------------------------------------------------------------------------------

def _xavier(m):
    if isinstance(m, (nn.Linear, nn.Conv2d)):
        nn.init.xavier_uniform_(m.weight.data, 1.0)
        nn.init.constant_(m.bias.data, 0)


def _orth(m, gain):
    nn.init.orthogonal_(m.weight.data, gain)
    nn.init.constant_(m.bias.data, 0)
    return m


class OracleMLP(nn.Module):
    def __init__(self, in_channels, depth=4, mid_weight=256, latent_size=64):
        super().__init__()
        mid = []
        for _ in range(depth - 2):
            mid += [nn.Linear(mid_weight, mid_weight), nn.ReLU()]
        self.model = nn.Sequential(nn.Linear(in_channels, mid_weight), nn.ReLU(), nn.Sequential(*mid),
                                   nn.Linear(mid_weight, latent_size))
        self.output_dim = latent_size
        self.apply(_xavier)

    def features(self, x):
        return self.model(x), None


class _Res(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv2d(c, c, 3, 1, 1), nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        return self.conv2(F.relu(self.conv1(F.relu(x)))) + x


class _Block(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv, self.res1, self.res2 = nn.Conv2d(cin, cout, 3, 1, 1), _Res(cout), _Res(cout)

    def forward(self, x):
        return self.res2(self.res1(F.max_pool2d(self.conv(x), 3, 2, 1)))


class OracleImpala(nn.Module):
    def __init__(self, in_channels, output_dim=256, latent_dim=32, input_hw=(64, 64)):
        super().__init__()
        self.block1, self.block2, self.block3 = _Block(in_channels, 16), _Block(16, 32), _Block(32, latent_dim)
        h, w = input_hw
        for _ in range(3):
            h, w = (h + 1) // 2, (w + 1) // 2
        self.fc = nn.Linear(latent_dim * h * w, output_dim)
        self.output_dim = output_dim
        self.apply(_xavier)

    def features(self, x):
        h = F.relu(self.block3(self.block2(self.block1(x)))).flatten(1)
        fs = torch.mean(torch.max(torch.tanh(torch.abs(h * 100)), 0)[0])    # model.py:207
        return F.relu(self.fc(h)), fs


class _GRU(nn.Module):
    """common/model.py:212-216: nn.GRU under the name ``gru`` (orthogonal_init leaves an nn.GRU untouched)."""

    def __init__(self, d):
        super().__init__()
        self.gru = nn.GRU(d, d)


class OraclePolicy(nn.Module):
    def __init__(self, embedder, action_size, recurrent=False):
        super().__init__()
        self.embedder = embedder
        self.fc_policy = _orth(nn.Linear(embedder.output_dim, action_size), 0.01)
        self.fc_value = _orth(nn.Linear(embedder.output_dim, 1), 1.0)
        self.recurrent = recurrent
        if recurrent:
            self.gru = _GRU(embedder.output_dim)

    def predict(self, x, hx, mask):
        """policy(obs, hidden, mask) with one row per hidden state (agents/ppo.py:72-81, common/model.py:219-226):
        -> (distribution, value, next hidden)."""
        feat, _ = self.embedder.features(x)
        if self.recurrent:
            out, hx = self.gru.gru(feat.unsqueeze(0), (hx * mask.unsqueeze(-1)).unsqueeze(0))
            feat, hx = out.squeeze(0), hx.squeeze(0)
        log_probs = F.log_softmax(self.fc_policy(feat), dim=1)
        return torch.distributions.Categorical(logits=log_probs), self.fc_value(feat).reshape(-1), hx

    def forward(self, x):
        feat, fs = self.embedder.features(x)
        log_probs = F.log_softmax(self.fc_policy(feat), dim=1)
        return torch.distributions.Categorical(logits=log_probs), self.fc_value(feat).reshape(-1), fs


# ----------------------------------------------------------------------------------------------
# Loss / optimiser
# ----------------------------------------------------------------------------------------------

def cross_batch_entropy(dist):
    """misc_util.py:32-51 (categorical branch) -> (marginal - conditional, conditional)."""
    cond = -(dist.probs * dist.logits).sum(-1).mean()
    p = dist.probs.mean(0)
    marg = -(p * torch.log(p)).sum()
    return marg - cond, cond


def ppo_loss(dist, value, act, old_logp, old_value, ret, adv, eps_clip=0.2, value_coef=0.5, entropy_coef=0.01,
             entropy_multiplier=1.0, x_entropy_coef=0.0, fs=None, fs_coef=0.0):
    """agents/ppo.py:131-169. Returns (loss, dict of the logged terms)."""
    logp = dist.log_prob(act)
    ratio = torch.exp(logp - old_logp)
    surr1 = ratio * adv
    surr2 = torch.clamp(ratio, 1.0 - eps_clip, 1.0 + eps_clip) * adv
    pi_loss = -torch.min(surr1, surr2).mean()
    clipped = old_value + (value - old_value).clamp(-eps_clip, eps_clip)
    v_loss = 0.5 * torch.max((value - ret).pow(2), (clipped - ret).pow(2)).mean()
    x_ent, ent = cross_batch_entropy(dist)
    loss = pi_loss + value_coef * v_loss - entropy_coef * ent * entropy_multiplier - x_entropy_coef * x_ent
    terms = dict(pi_loss=pi_loss, value_loss=v_loss, entropy=ent, x_entropy=x_ent)
    if fs is not None:
        loss = loss + fs_coef * fs
        terms["fs"] = fs
    terms["total"] = loss
    return loss, terms


def optimize(policy, optimizer, data, n_steps, n_envs, epoch=3, n_minibatch=8, mini_batch_size=8192,
             grad_clip_norm=0.5, recurrent=False, **loss_kw):
    """agents/ppo.py:96-208 on a dict of flat [T*N, ...] fp32 tensors
    (obs, act, old_logp, old_value, ret, adv).  Returns the per-minibatch term lists."""
    batch_size = n_steps * n_envs // n_minibatch
    mini_batch_size = min(mini_batch_size, batch_size)
    accum = batch_size / mini_batch_size
    cnt, logs = 1, []
    fs_coef = loss_kw.pop("fs_coef", 0.0)
    for _ in range(epoch):
        batches = recurrent_epoch_indices(n_steps, n_envs, mini_batch_size)[0] if recurrent \
            else epoch_indices(n_steps * n_envs, mini_batch_size)
        for idx in batches:
            dist, value, fs = policy(data["obs"][idx])        # (recurrent too: the GRU is not on optimize()'s path)
            loss, terms = ppo_loss(dist, value, data["act"][idx], data["old_logp"][idx], data["old_value"][idx],
                                   data["ret"][idx], data["adv"][idx], fs=fs, fs_coef=fs_coef, **loss_kw)
            loss.backward()
            if cnt % accum == 0:
                torch.nn.utils.clip_grad_norm_(policy.parameters(), grad_clip_norm)
                optimizer.step()
                optimizer.zero_grad()
            cnt += 1
            logs.append({k: float(v.detach()) for k, v in terms.items()})
    return logs


def adjust_lr(init_lr, timesteps, max_timesteps):
    return init_lr * (1 - (timesteps / max_timesteps))


def make_adam(policy, lr):
    return torch.optim.Adam(policy.parameters(), lr=lr, eps=1e-5)   # agents/ppo.py:58


# ----------------------------------------------------------------------------------------------
# One full CPU PPO iteration (rollout + GAE + update): the cpu_baseline leg of bench.py
# ----------------------------------------------------------------------------------------------

def ppo_iteration(env_step, obs0, policy, optimizer, n_steps, n_envs, gamma, lmbda, obs_transform=None, **opt_kw):
    """env_step(action ndarray[N]) -> (obs, reward, done).  obs0: current observation (ndarray).
    Follows agents/ppo.py:228-254 (train env only).  Returns (last obs, logs)."""
    tf = obs_transform or (lambda o: o)
    obs = obs0
    O, Act, Lp, V, R, D = [], [], [], [], [], []
    with torch.no_grad():
        for _ in range(n_steps):
            x = torch.FloatTensor(tf(obs))
            dist, value, _ = policy(x)
            act = dist.sample()
            O.append(x); Act.append(act.float()); Lp.append(dist.log_prob(act)); V.append(value)
            obs, rew, done = env_step(act.numpy())
            R.append(torch.from_numpy(np.asarray(rew, dtype=np.float32).copy()))
            D.append(torch.from_numpy(np.asarray(done, dtype=np.float32).copy()))
        _, last_v, _ = policy(torch.FloatTensor(tf(obs)))
    rew, done, value = torch.stack(R), torch.stack(D), torch.stack(V + [last_v])
    adv, ret = gae(rew, done, value, gamma, lmbda)
    adv = normalize_adv(adv)
    data = dict(obs=torch.stack(O).reshape(n_steps * n_envs, *O[0].shape[1:]), act=torch.stack(Act).reshape(-1),
                old_logp=torch.stack(Lp).reshape(-1), old_value=value[:-1].reshape(-1), ret=ret.reshape(-1),
                adv=adv.reshape(-1))
    logs = optimize(policy, optimizer, data, n_steps, n_envs, **opt_kw)
    return obs, logs
