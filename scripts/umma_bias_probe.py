"""Is the tcgen05 fp32 accumulation biased (truncation)?  Signed relative error of positive dot products."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nsfnet_b200 import _capi
lib = _capi.load()
rng = np.random.default_rng(0)
for k in (8, 16, 40, 80, 160):
    n = 32
    A = rng.random((128, k)).astype(np.float32) + 0.5
    B = rng.random((n, k)).astype(np.float32) + 0.5
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    ref32 = (A @ B.T).astype(np.float64)
    a = torch.as_tensor(A).cuda(); b = torch.as_tensor(B).cuda()
    for variant in (3 | 16, 2 | 16):
        d = torch.empty((128, n), device='cuda')
        _capi.check(lib, lib.nsf_selftest_umma(0, variant, a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, None))
        torch.cuda.synchronize()
        out = d.cpu().numpy().astype(np.float64)
        rel = (out - ref) / ref
        print(f"k={k:3d} variant={variant:2d} mean signed rel err {rel.mean():+.3e}  rms {np.sqrt((rel**2).mean()):.3e}   (numpy fp32 matmul: mean {((ref32-ref)/ref).mean():+.3e} rms {np.sqrt((((ref32-ref)/ref)**2).mean()):.3e})")
