"""tcgen05 descriptor encodings: nsf_selftest_umma runs kind::tf32 MMAs with the K-major no-swizzle
shared-memory layouts of the jet kernel; compare against a float64 product.  1xTF32 must show the
~1e-3 truncation error, the 3xTF32 split must bring it to fp32 level."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(32, 80), (80, 32), (48, 80), (16, 8), (256, 16), (64, 120)])
def test_umma_kmajor_layouts(n, k):
    import torch
    from nsfnet_b200 import _capi
    lib = _capi.load()
    rng = np.random.default_rng(n * 1000 + k)
    A = rng.standard_normal((128, k)).astype(np.float32)
    B = rng.standard_normal((n, k)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    a, b = torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda()
    for variant in (2, 3, 4):       # 4: A operand from tensor memory; 2: "n-contiguous" operand images (wgrad), 3: padded row images (forward / dgrad)
        errs = {}
        for flags in (0, 16):
            d = torch.full((128, n), float("nan"), device="cuda")
            _capi.check(lib, lib.nsf_selftest_umma(0, variant | flags, a.data_ptr(), b.data_ptr(), d.data_ptr(), n, k, None))
            torch.cuda.synchronize()
            out = d.cpu().numpy().astype(np.float64)
            errs[flags] = np.linalg.norm(out - ref) / np.linalg.norm(ref)
        assert 1e-4 < errs[0] < 3e-3, errs       # tensor cores really ran in tf32
        assert errs[16] < 2e-6, errs             # 3xTF32 recovers fp32-level accuracy
