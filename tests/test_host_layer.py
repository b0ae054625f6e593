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
