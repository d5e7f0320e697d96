"""Measured numerical error of the floating-point kernels against float64 evaluations of the same operations
(round 2; feeds the tolerance table of DESIGN.md section 5).  err = max |x - x64| / max |x64| over the tensor."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpp_b200.common.engine import MLPEngineTC  # noqa: E402
from tpp_b200.common.model import MLPModel  # noqa: E402
from tpp_b200.common.policy import CategoricalPolicy  # noqa: E402


def rel(a, b):
    """(max-norm error, L2 error), both relative to the float64 reference.  The float64 network uses the ENGINE's ReLU
    masks: a pre-activation within rounding of zero is otherwise masked differently in fp32 and float64, which moves
    one sample's whole contribution to a weight row (with the random loss gradient used here ~1/sqrt(M) of the row's
    magnitude: 1e-3 at M = 16384, identical for every engine variant) -- a property of comparing a piecewise-linear
    network across precisions, not of the kernels."""
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def main():
    rows = []
    torch.manual_seed(0)
    for name, in_dim, raw, M in (("MLP 588-256-256-256-64 on pixel rows (Box-World), M=16384", 588, True, 16384),
                                 ("MLP 9-256-256-256-64 (cartpole), M=4096", 9, False, 4096)):
        A = 4
        pol = CategoricalPolicy(MLPModel(in_dim, 4, 256, 64), False, A).to("cuda").flatten_()
        with torch.no_grad():
            pol.flat.add_(0.02 * torch.randn_like(pol.flat))
        for prec, soc in ((3, False), (3, True), (1, False)):
            eng = MLPEngineTC(pol, A, precision=prec, raw_pixels=raw, split_on_chip=soc)
            torch.manual_seed(in_dim)          # the same inputs for every variant
            if raw:
                x = torch.zeros(M, eng.ld_in, device="cuda")
                x[:, :in_dim] = torch.randint(0, 256, (M, in_dim), device="cuda").float()
                x64 = (x[:, :in_dim].double() / 255.0)
            else:
                x = torch.randn(M, in_dim, device="cuda")
                x64 = x.double()
            head = eng.forward(x, M, raw=raw).clone()
            dhead = torch.randn(M, eng.ld_head, device="cuda") / M
            dhead[:, A + 1:] = 0
            pol.flat_grad.zero_()
            eng.backward(dhead, M)
            g = pol.flat_grad.clone()
            # float64 autograd of the same network
            p64 = {k: v.detach().double().clone().requires_grad_(True) for k, v in pol.state_dict().items()}
            names = sorted({k.rsplit(".", 1)[0] for k in p64 if k.startswith("embedder.")},
                           key=lambda s: [int(t) if t.isdigit() else t for t in s.split(".")])
            # (ReLU masks taken from the engine's own activations: a pre-activation within rounding of zero would
            # otherwise be masked differently in float64, moving one sample's whole contribution -- see rel())
            ws = eng._workspace(M)
            h = x64
            for i, n in enumerate(names):
                h = h @ p64[n + ".weight"].t() + p64[n + ".bias"]
                if eng.layers[i][4]:
                    h = h * (ws.h[i]["hi"][:, :h.shape[1]] > 0).double()
            out = torch.cat((h @ p64["fc_policy.weight"].t() + p64["fc_policy.bias"],
                             h @ p64["fc_value.weight"].t() + p64["fc_value.bias"]), 1)
            (out * dhead[:, :A + 1].double()).sum().backward()
            g64 = torch.cat([p64[k].grad.reshape(-1) for k in pol.flat_order()])
            tag = ("3xTF32, lo halves formed on chip" if soc else "3xTF32, (hi, lo) pairs in HBM") if prec == 3 else "1xTF32"
            rows.append((f"{name}, {tag}: logits + value", rel(head[:, :A + 1], out.detach())))
            rows.append((f"{name}, {tag}: flat parameter gradient", rel(g, g64)))
            if prec == 3 and not soc and eng.fused_rollout_ok(raw):
                N = 4096
                act = torch.zeros(N, dtype=torch.int32, device="cuda")
                lp, vv = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
                ho = torch.zeros(N, eng.ld_head, device="cuda")
                tick = torch.zeros(1, dtype=torch.int64, device="cuda")
                if raw:
                    eng.rollout_fused(x[:N], N, eng.ld_in, True, act, lp, vv, 0, tick, 0, head_out=ho)
                else:
                    xs = x[:N].t().contiguous()
                    eng.rollout_fused(xs, N, N, False, act, lp, vv, 0, tick, 0, head_out=ho)
                rows.append((f"{name}: fused rollout kernel logits + value (first 4096 rows)",
                             rel(ho[:, :A + 1], out.detach()[:N])))
    print("| quantity | max err / max |ref| vs float64 | L2 err / L2 |ref| |")
    print("|---|---:|---:|")
    for n, e in rows:
        print(f"| {n} | {e[0]:.2e} | {e[1]:.2e} |")


if __name__ == "__main__":
    main()
