"""SASS op-count summary of csrc/libtpp_b200.so: per kernel, the mnemonics that prove which hardware paths are in the
binary (tcgen05 / TMEM / TMA / cluster / mbarrier / cp.async / PDL / vector reductions, and FFMA for the FMA-pipe kernels).

    python profiles/sass_summary.py > profiles/sass_summary_r02.md          # needs cuobjdump + c++filt, no GPU
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "train-procgen-pytorch_b200", "csrc", "libtpp_b200.so")
COLS = [("UTCHMMA (tcgen05.mma kind::tf32)", lambda m: m.startswith("UTCHMMA") and ".2CTA" not in m),
        ("UTCHMMA.2CTA (cta_group::2)", lambda m: m.startswith("UTCHMMA") and ".2CTA" in m),
        ("LDTM (tcgen05.ld)", lambda m: m.startswith("LDTM")),
        ("UTMALDG (TMA tiled load)", lambda m: m.startswith("UTMALDG") and "IM2COL" not in m),
        ("UTMALDG ... IM2COL", lambda m: m.startswith("UTMALDG") and "IM2COL" in m),
        ("UTCBAR (tcgen05.commit)", lambda m: m.startswith("UTCBAR")),
        ("SYNCS (mbarrier)", lambda m: m.startswith("SYNCS")),
        ("UCGABAR (barrier.cluster)", lambda m: m.startswith("UCGABAR")),
        ("LDGSTS (cp.async)", lambda m: m.startswith("LDGSTS")),
        ("ACQBULK/PDL (griddepcontrol)", lambda m: m.startswith("ACQBULK") or m.startswith("PREEXIT")),
        ("RED v4 f32 (red.global.add.v4.f32)", lambda m: m.startswith("RED") and ".128" in m or (m.startswith("RED") and "F32" in m)),
        ("FFMA", lambda m: m.startswith("FFMA")),
        ("SHFL (warp shuffles)", lambda m: m.startswith("SHFL"))]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur].append(m.group(1))
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS op-count summary of `libtpp_b200.so` (round 2, final HEAD)\n")
    print("`python profiles/sass_summary.py` = `cuobjdump -sass train-procgen-pytorch_b200/csrc/libtpp_b200.so` (sm_100a), counted per")
    print("kernel: the mnemonics that prove the tcgen05 / TMEM / TMA / cluster / mbarrier paths are really in the binary")
    print("(B200_PROFILING.md), plus FFMA for the FMA-pipe convolution kernels.  Kernels with none of the first eleven and fewer")
    print("than 500 FFMA are omitted.\n")
    print("| kernel | SASS instr | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---:|" + "---:|" * len(COLS))
    for mangled, name in zip(kernels, names):
        ops = kernels[mangled]
        counts = [sum(1 for o in ops if f(o)) for _, f in COLS]
        if not any(counts[:11]) and counts[11] < 500:
            continue
        short = re.sub(r"\(.*", "", name).replace("tpp::", "").replace("void ", "")
        print(f"| `{short[:90]}` | {len(ops)} | " + " | ".join(str(c) for c in counts) + " |")


if __name__ == "__main__":
    sys.exit(main())
