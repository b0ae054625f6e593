"""Exploratory: run nsf_selftest_umma for every variant / flag combination and print the error against a
float64 product (used once to pin down the descriptor semantics; the regression test is tests/test_gpu_umma.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nsfnet_b200 import _capi
lib = _capi.load()
rng = np.random.default_rng(0)
for (n, k) in [(48, 80), (64, 80), (128, 120), (80, 48), (80, 64), (16, 8), (256, 16)]:
    A = rng.standard_normal((128, k)).astype(np.float32)
    B = rng.standard_normal((n, k)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    a = torch.as_tensor(A).cuda(); b = torch.as_tensor(B).cuda()
    for variant in (0, 1, 2):
        for flags in (0, 16, 32, 64, 96):
            d = torch.full((128, n), float('nan'), device='cuda')
            rc = lib.nsf_selftest_umma(0, variant | flags, a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, None)
            if rc != 0:
                print(n, k, variant, flags, 'rc', rc, lib.nsf_last_error().decode()); continue
            torch.cuda.synchronize()
            out = d.cpu().numpy().astype(np.float64)
            err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
            print(f"n={n:3d} k={k:3d} variant={variant} flags={flags:3d} rel_err={err:.3e}", flush=True)
