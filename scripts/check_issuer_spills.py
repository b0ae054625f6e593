"""The one thread that issues tcgen05.mma must not touch local memory between MMAs: a register spill there (the compiler hoists ~150
loop-invariant descriptors out of the stage loop unless told otherwise) cost 1.5 ms per 1e6 points.  Disassembles the built library and
reports, per points-on-M kernel, the local loads / stores inside the span of its UTCHMMA instructions.
Usage: python scripts/check_issuer_spills.py [lib.so]   (exit code 1 if any)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def check_uniform(lib, max_r2ur=100):
    """The role branches of the points-on-M kernels must be uniform branches (warp index taken through a shuffle): when the compiler treats
    the epilogue as divergent code it routes every memory descriptor and barrier address through R2UR (270 of them) and spills -- 6 % of the
    kernel time.  Returns {kernel: (R2UR count, local-memory instructions)} of the offenders."""
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    bad = {}
    for block in out.split("Function : ")[1:]:
        name = block.split("\n", 1)[0].strip()
        if "nsf_pm_jet_kernel" not in name or "ELb1ELb1EEEv" in name:      # (the cycle-counter instantiations <..., true, true> are diagnostics)
            continue
        r2ur = len(re.findall(r"\bR2UR\b", block))
        local = len(re.findall(r"\b(?:LDL|STL)\b", block))
        if r2ur > max_r2ur or local:
            bad[name] = (r2ur, local)
    return bad


def check(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    bad, seen = {}, 0
    for block in out.split("Function : ")[1:]:
        name = block.split("\n", 1)[0].strip()
        if "nsf_pm_jet_kernel" not in name:
            continue
        lines = block.split("\n")
        mma = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
        if not mma:
            continue
        seen += 1
        spills = [i for i in range(mma[0], mma[-1] + 1) if re.search(r"\b(LDL|STL)\b", lines[i])]
        if spills:
            bad[name] = len(spills)
    return seen, bad


if __name__ == "__main__":
    seen, bad = check(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nsfnet_b200", "libnsf_b200.so"))
    print(f"{seen} tcgen05 kernels checked;", "no local-memory traffic inside the MMA issue spans" if not bad else f"SPILLS inside the issue span: {bad}")
    bad_u = check_uniform(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nsfnet_b200", "libnsf_b200.so"))
    print("no spills, uniform role branches" if not bad_u else f"divergent-code symptoms (R2UR count, local-memory instructions): {bad_u}")
    sys.exit(1 if bad or bad_u else 0)
