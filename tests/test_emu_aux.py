"""Step-level kernels (device-resident Adam, Latin-hypercube points, wall distance, SDF weights) checked WITHOUT a GPU
through the host emulation build of nsfnet_b200/csrc/nsf_aux.cu (tests/emu); the CUDA build of the same per-element
code is checked on the B200 by tests/test_gpu_fused.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from nsfnet_b200 import _capi
from tests.emu import emu


@pytest.fixture(scope="module")
def lib():
    return emu.load()


def P(a):
    return a.ctypes.data_as(C.c_void_p)


def test_adam_dev_matches_torch_adam(lib):
    rng = np.random.default_rng(0)
    n = 1000
    p0 = rng.standard_normal(n).astype(np.float32)
    p = p0.copy(); m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    st = _capi.NsfAdamDev(1e-3, 0.9, 0.999, 1e-8, 1.0, 0, (C.c_int32 * 2)(0, 0))
    tp = torch.nn.Parameter(torch.tensor(p0))
    opt = torch.optim.Adam([tp], lr=1e-3, weight_decay=0.0)
    for k in range(7):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 2, n)).astype(np.float32)
        _capi.check(lib, lib.nsf_adam_dev(P(p), P(g), P(m), P(v), n, C.byref(st), None))
        _capi.check(lib, lib.nsf_adam_tick(C.byref(st), None))
        tp.grad = torch.tensor(g)
        opt.step()
        if k == 3:      # a new stage learning rate is a write to the state
            st.lr = 2e-4
            opt.param_groups[0]["lr"] = 2e-4
    assert st.step == 7
    assert np.max(np.abs(p - tp.detach().numpy())) < 2e-7 * max(1.0, np.abs(p0).max())
    # a fresh optimizer (freeze / defreeze, ev :489-511) = zero moments, step 0: the first update is a sign step of size lr
    m[:] = 0; v[:] = 0; st.step = 0; st.lr = 1e-3
    before = p.copy()
    g = rng.standard_normal(n).astype(np.float32)
    _capi.check(lib, lib.nsf_adam_dev(P(p), P(g), P(m), P(v), n, C.byref(st), None))
    assert np.allclose(before - p, 1e-3 * np.sign(g), rtol=1e-4, atol=0)


@pytest.mark.parametrize("n", [1, 2, 7, 1000, 4096, 100003])
def test_lhs_points_is_a_latin_hypercube(lib, n):
    x = np.empty(n, np.float32); y = np.empty(n, np.float32)
    _capi.check(lib, lib.nsf_lhs_points(n, 0, n, 5, 0.0, 1.0, 0.0, 1.0, P(x), P(y), None))
    for c in (x, y):   # exactly one sample per 1/N stratum (tools.py:30-57)
        cells = np.floor(c.astype(np.float64) * n).astype(np.int64)
        if n <= 2 ** 22:
            assert np.array_equal(np.sort(cells), np.arange(n))
        assert c.min() >= 0.0 and c.max() < 1.0
    if n >= 1000:
        assert abs(np.corrcoef(x, y)[0, 1]) < 0.1
        assert abs(np.corrcoef(x[:-1], x[1:])[0, 1]) < 0.1     # neighbouring rows are not neighbouring strata
    # any row range of the same design, from any rank
    if n >= 7:
        a, b = n // 3, n - n // 4
        xs = np.empty(b - a, np.float32); ys = np.empty(b - a, np.float32)
        _capi.check(lib, lib.nsf_lhs_points(n, a, b - a, 5, 0.0, 1.0, 0.0, 1.0, P(xs), P(ys), None))
        assert np.array_equal(xs, x[a:b]) and np.array_equal(ys, y[a:b])
        x2 = np.empty(n, np.float32); y2 = np.empty(n, np.float32)
        _capi.check(lib, lib.nsf_lhs_points(n, 0, n, 6, -1.0, 1.0, 2.0, 3.0, P(x2), P(y2), None))
        assert not np.array_equal(x2, x) and x2.min() >= -1.0 and y2.min() >= 2.0 and y2.max() <= 3.0
    assert lib.nsf_lhs_points(n, 1, n, 5, 0.0, 1.0, 0.0, 1.0, P(x), P(y), None) == -1      # NSF_E_ARG: rows past the design


def test_wall_distance_and_sdf_weights_match_ckdtree(lib):
    from nsfnet_b200.cavity_data import cavity_boundary, sdf_weights, wall_distance
    xb, yb, _, _ = cavity_boundary(513)
    xb32, yb32 = xb.ravel().astype(np.float32), yb.ravel().astype(np.float32)
    rng = np.random.default_rng(1)
    n = 3000
    x, y = rng.random(n).astype(np.float32), rng.random(n).astype(np.float32)
    x[:3] = [0.0, 1.0, 0.5]; y[:3] = [0.0, 0.3, 0.5]
    d = np.empty(n, np.float32)
    _capi.check(lib, lib.nsf_wall_distance(P(x), P(y), n, P(xb32), P(yb32), xb32.size, P(d), None))
    pts, bc = np.stack([x, y], 1).astype(np.float64), np.stack([xb32, yb32], 1).astype(np.float64)
    d_ref = wall_distance(pts, bc)
    assert np.max(np.abs(d - d_ref)) < 2e-7
    w = np.empty(n, np.float32); acc = np.zeros(1, np.float64)
    _capi.check(lib, lib.nsf_sdf_weights(P(x), P(y), n, P(xb32), P(yb32), xb32.size, 0.2, 5.0, P(w), P(acc), None))
    assert abs(acc[0] - w.astype(np.float64).sum()) < 1e-9 * n
    w_ref = sdf_weights(pts, bc, 0.2, 5.0)
    assert np.max(np.abs(w / (acc[0] / n) - w_ref)) < 5e-7
    # the reference's clamps (cavity_data.py:123-126)
    _capi.check(lib, lib.nsf_sdf_weights(P(x), P(y), n, P(xb32), P(yb32), xb32.size, 7.0, -3.0, P(w), None, None))
    assert np.allclose(w, 1.0)


def test_vtm_from_e_equals_explicit_init(lib):
    """NSF_VTM_FROM_E: `init_vis_t` (ev :138-140) fused into the first loss evaluation == passing alpha_init*|e| explicitly."""
    from oracle import jet_numpy as J
    rng = np.random.default_rng(3)
    md, ed = J.NetDesc(2, 3, 3, 16), J.NetDesc(2, 1, 4, 40)
    pm, pe = J.init_params(md, 1) * 1.5, J.init_params(ed, 2) * 3.0
    n, nb = 70, 9
    x, y = rng.random(n), rng.random(n)
    xb, yb, ub, vb = rng.random(nb), rng.random(nb), rng.random(nb), rng.random(nb)
    blocks = [(xb, yb, ub, vb, None, 10. / nb, 10. / nb, 0.)]
    a_init, a_now = 0.05, 0.03
    base = emu.run_step(lib, (2, 3, 3, 16), pm, _capi.physics(50., alpha_evm=a_now, has_evm=True), x, y, blocks=blocks,
                        evm_desc=(2, 1, 4, 40), params_evm=pe)                     # gives e
    vtm0 = (np.float32(a_init) * np.abs(base["e"])).astype(np.float32)
    assert (vtm0 < 20. / 50.).any()                                             # the cap is not active everywhere
    want = emu.run_step(lib, (2, 3, 3, 16), pm, _capi.physics(50., alpha_evm=a_now, has_evm=True), x, y, blocks=blocks,
                        evm_desc=(2, 1, 4, 40), params_evm=pe, vtm_in=vtm0)
    got = emu.run_step(lib, (2, 3, 3, 16), pm, _capi.physics(50., alpha_evm=a_now, has_evm=True, vtm_from_e_alpha=a_init), x, y,
                       blocks=blocks, evm_desc=(2, 1, 4, 40), params_evm=pe)
    for k in ("grad_main", "loss_parts", "resid", "vis_t", "vtm_out", "e"):
        assert np.array_equal(got[k], want[k]), k
    assert not np.array_equal(got["vis_t"], base["vis_t"])
