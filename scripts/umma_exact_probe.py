"""Raw tcgen05 kind::tf32 results for chosen inputs -> gpurun_out/umma_exact_probe.npz (fitted offline by scripts/emu_tc_numerics.py:
which rounding model reproduces the tensor core bit for bit?).  Plain tf32 MMAs (no split), K/8 chained accumulations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nsfnet_b200 import _capi
lib = _capi.load()
rng = np.random.default_rng(7)


def tf32(x):
    x = np.ascontiguousarray(x, np.float32)
    return ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xffffe000)).view(np.float32)


out = {}
n = 32
for k in (8, 16, 40, 80, 120):
    for dist in ("normal", "wide", "pos"):
        if dist == "normal":
            A = rng.standard_normal((128, k)); B = rng.standard_normal((n, k))
        elif dist == "wide":
            A = rng.standard_normal((128, k)) * np.exp2(rng.integers(-12, 4, (128, k))); B = rng.standard_normal((n, k)) * np.exp2(rng.integers(-6, 3, (n, k)))
        else:
            A = rng.random((128, k)) + 0.5; B = rng.random((n, k)) + 0.5
        A = tf32(A.astype(np.float32)); B = tf32(B.astype(np.float32))
        a = torch.as_tensor(A).cuda(); b = torch.as_tensor(B).cuda()
        d = torch.empty((128, n), device="cuda")
        _capi.check(lib, lib.nsf_selftest_umma(0, 3, a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, None))
        torch.cuda.synchronize()
        out[f"A_{k}_{dist}"] = A; out[f"B_{k}_{dist}"] = B; out[f"D_{k}_{dist}"] = d.cpu().numpy()
# single MMA, products far below one dominant product: how many bits below the largest term survive the alignment?
k = 8
for sh in range(0, 40, 2):
    A = np.zeros((128, k), np.float32); B = np.zeros((n, k), np.float32)
    A[:, 0] = tf32((1.0 + rng.random(128)).astype(np.float32)); B[:, 0] = 1.0
    A[:, 1:] = tf32(((1.0 + rng.random((128, k - 1))) * 2.0 ** -sh).astype(np.float32)); B[:, 1:] = tf32((1.0 + rng.random((n, k - 1))).astype(np.float32)) * np.where(rng.random((n, k - 1)) < 0.5, -1, 1).astype(np.float32)
    a = torch.as_tensor(A).cuda(); b = torch.as_tensor(B).cuda()
    d = torch.empty((128, n), device="cuda")
    _capi.check(lib, lib.nsf_selftest_umma(0, 3, a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, None))
    torch.cuda.synchronize()
    out[f"A_sh{sh}"] = A; out[f"B_sh{sh}"] = B; out[f"D_sh{sh}"] = d.cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/umma_exact_probe.npz", **out)
print("saved", len(out))
