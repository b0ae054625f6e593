"""Host-side pieces that need no GPU: the C-ABI library loads and exports every symbol the header
declares, FCNet keeps the reference's module tree / state_dict layout on one flat buffer, the data
layer reproduces the reference's point sets."""
import ctypes
import os
import re

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from nsfnet_b200 import _capi, build
    path = build.build_cuda()
    lib = ctypes.CDLL(path)            # loads without a GPU (libcudart is lazily initialised)
    hdr = open(os.path.join(ROOT, "include", "nsf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(nsf_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_capi.EXPORTS) == declared
    _capi.bind(lib)
    assert lib.nsf_abi_version() == 1
    assert lib.nsf_last_error() is not None


def test_fcnet_layout_matches_reference_module():
    from nsfnet_b200.net import FCNet
    from oracle.autograd_port import RefNet
    torch.manual_seed(0)
    net = FCNet(num_ins=2, num_outs=3, num_layers=4, hidden_size=120)
    ref = RefNet(2, 3, 4, 120)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    assert net.depth == 5
    ref.load_state_dict(net.state_dict())
    flat = net.flat_params()
    assert flat.numel() == 44283
    assert np.array_equal(flat.numpy(), ref.flat())
    # parameters are views of the flat buffer: an in-place optimizer update is visible in it
    with torch.no_grad():
        net.layers.layer_0.weight.add_(1.0)
    assert np.allclose(net.flat_params()[:240].numpy(), ref.flat()[:240] + 1.0)
    # load_state_dict copies in place, the flat view survives
    net.load_state_dict(ref.state_dict())
    assert net._is_flat() and np.array_equal(net.flat_params().numpy(), ref.flat())
    x = torch.rand(7, 2)
    assert torch.allclose(net(x), ref(x))
    d = FCNet()   # reference defaults (net.py:23-27)
    assert (d.num_ins, d.num_outs, d.num_layers, d.hidden_size) == (3, 3, 10, 50)


def test_cavity_data_matches_reference_sets():
    from nsfnet_b200.cavity_data import DataLoader, cavity_boundary, lhs_sample, sdf_weights
    from oracle.jet_numpy import cavity_boundary as oracle_boundary
    xb, yb, ub, vb = cavity_boundary(513)
    ox, oy, ou, ov = oracle_boundary(513)
    assert xb.shape == (2052, 1) and np.array_equal(xb.ravel(), ox) and np.array_equal(ub.ravel(), ou)
    assert abs(ub.max() - 0.98652) < 1e-5         # lid maximum
    s = lhs_sample(2, [[0, 1], [0, 1]], 1000, np.random.default_rng(0))
    for d in range(2):                            # exactly one sample per 1/N stratum
        assert np.array_equal(np.sort((s[:, d] * 1000).astype(int)), np.arange(1000))

    class Cfg:
        enabled, min_weight, decay = True, 0.2, 5.0
    dl = DataLoader(N_f=2000, sdf_weighting=Cfg(), seed=1)
    dl.loading_boundary_data()
    x, y = dl.loading_training_data()
    w = dl.get_sdf_weights()
    assert x.shape == (2000, 1) and w.shape == (2000,) and w.dtype == np.float32 and abs(w.mean() - 1) < 1e-5
    dist = np.minimum(np.minimum(x, 1 - x), np.minimum(y, 1 - y)).ravel()
    assert np.all(np.diff(dist) > -2e-3)          # sorted by wall distance (to the discrete boundary set)
    assert w[0] > w[-1]


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """The ctypes mirrors in nsfnet_b200/_capi.py against include/nsf_b200.h as the C compiler lays it out (sizes and
    field offsets): the boundary is plain C, so any binding -- cgo, JNI, cffi -- sees exactly these numbers."""
    import subprocess
    from nsfnet_b200 import _capi
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "nsf_b200.h"
#define F(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("NsfNetDesc %zu\nNsfPhysics %zu\nNsfDataBlock %zu\nNsfAdamDev %zu\n", sizeof(NsfNetDesc), sizeof(NsfPhysics), sizeof(NsfDataBlock), sizeof(NsfAdamDev));
  F(NsfPhysics, inv_Re); F(NsfPhysics, eq4_weight); F(NsfPhysics, flags); F(NsfPhysics, alpha_evm_init); F(NsfPhysics, n_f_norm);
  F(NsfDataBlock, p); F(NsfDataBlock, n); F(NsfDataBlock, cu); F(NsfDataBlock, cp);
  F(NsfAdamDev, lr); F(NsfAdamDev, grad_scale); F(NsfAdamDev, step);
  printf("NSF_LOSS_SLOTS %d\nNSF_MAX_BLOCKS %d\nNSF_VTM_FROM_E %u\n", NSF_LOSS_SLOTS, NSF_MAX_BLOCKS, NSF_VTM_FROM_E);
  return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.rsplit(" ", 1) for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines())
    for name, cls in (("NsfNetDesc", _capi.NsfNetDesc), ("NsfPhysics", _capi.NsfPhysics), ("NsfDataBlock", _capi.NsfDataBlock),
                      ("NsfAdamDev", _capi.NsfAdamDev)):
        assert int(out[name]) == ctypes.sizeof(cls), name
    for key, val in out.items():
        if "." in key:
            struct, field = key.split(".")
            assert int(val) == getattr(getattr(_capi, struct), field).offset, key
    assert int(out["NSF_LOSS_SLOTS"]) == _capi.NSF_LOSS_SLOTS and int(out["NSF_MAX_BLOCKS"]) == _capi.NSF_MAX_BLOCKS
    assert int(out["NSF_VTM_FROM_E"]) == _capi.NSF_VTM_FROM_E


def test_mma_issue_loop_has_no_register_spills():
    """scripts/check_issuer_spills.py: no local-memory load / store between the tcgen05.mma instructions of the points-on-M kernels
    (a spill in the single issuing thread's loop cost 20 % of the step; the issuer's descriptors are kept loop-variant for that reason)."""
    import importlib.util
    import shutil
    from nsfnet_b200 import _capi
    if shutil.which("cuobjdump") is None or not os.path.exists(_capi.LIB_PATH):
        pytest.skip("needs cuobjdump and the built library")
    spec = importlib.util.spec_from_file_location("check_issuer_spills", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "check_issuer_spills.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    seen, bad = mod.check(_capi.LIB_PATH)
    assert seen >= 8 and not bad, bad
    # and the kernels are spill-free with uniform role branches (warp index through a shuffle; otherwise hundreds of R2UR)
    assert not mod.check_uniform(_capi.LIB_PATH)
