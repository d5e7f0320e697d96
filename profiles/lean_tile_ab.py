"""A/B of the update pass with the LEAN (3-stage, 8 epilogue warps) persistent pair tile on subsets of the GEMMs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tpp_b200.common.engine import MLPEngineTC  # noqa: E402
from tpp_b200.common.model import MLPModel  # noqa: E402
from tpp_b200.common.policy import CategoricalPolicy  # noqa: E402


def main():
    M, in_dim, A = int(os.environ.get("ROWS", 131072)), 588, 4
    torch.manual_seed(0)
    pol = CategoricalPolicy(MLPModel(in_dim, 4, 256, 64), False, A).to("cuda").flatten_()
    ref = None
    for kinds in ((), ("fwd",), ("fwd", "dgrad"), ("fwd", "dgrad", "wgrad"), ("dgrad",), ("wgrad",)):
        eng = MLPEngineTC(pol, A, precision=3, raw_pixels=True)
        eng.lean_kinds = kinds
        x = torch.zeros(M, eng.ld_in, device="cuda")
        x[:, :in_dim] = torch.randint(0, 256, (M, in_dim), device="cuda", generator=torch.Generator("cuda").manual_seed(1)).float()
        dhead = torch.randn(M, eng.ld_head, device="cuda", generator=torch.Generator("cuda").manual_seed(2)) / M
        dhead[:, A + 1:] = 0
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for it in range(8):
            pol.flat_grad.zero_()
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            eng.forward(x, M, raw=True)
            e1.record()
            eng.backward(dhead, M)
            e2.record()
            torch.cuda.synchronize()
            ts.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        ts = ts[3:]
        f = sum(t[0] for t in ts) / len(ts) * 1e3
        b = sum(t[1] for t in ts) / len(ts) * 1e3
        g = pol.flat_grad.clone()
        if ref is None:
            ref = g
        d = float((g.double() - ref.double()).abs().max() / ref.double().abs().max())
        print(f"lean on {str(kinds):28s}: forward {f:.0f} us, backward {b:.0f} us, pass {f + b:.0f} us; gradient vs baseline {d:.1e}")


if __name__ == "__main__":
    main()
