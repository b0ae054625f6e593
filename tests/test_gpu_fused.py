"""SURVEY 8f rows 1-2 on the GPU: the fused iteration (nsf_step + device-resident Adam, replayed as one CUDA graph)
against the reference-style loop (loss.backward() + torch.optim.Adam) from identical states, and the on-device point
layer (Latin hypercube, wall distance, SDF weights) against the host data layer / scipy's cKDTree."""
import os

import numpy as np
import pytest

from oracle import jet_numpy as J

pytestmark = pytest.mark.gpu


def _ev(n_f=3000, seed=1):
    import torch
    from nsfnet_b200.ev_nsfnet import PysicsInformedNeuralNetwork
    torch.manual_seed(seed)
    P = PysicsInformedNeuralNetwork(Re=2000, layers=6, hidden_size=80, layers_1=4, hidden_size_1=40, N_f=n_f, alpha_evm=0.05,
                                    bc_weight=10, eq_weight=1, supervised_data_weight=0.0)
    rng = np.random.default_rng(0)
    xb, yb, ub, vb = J.cavity_boundary(65)
    P.set_boundary_data((xb, yb, ub, vb))
    P.set_eq_training_data((rng.random(n_f).astype(np.float32), rng.random(n_f).astype(np.float32)))
    P.log_interval = 10 ** 9
    P.checkpoints = False
    P.verbose = False
    return P


def _state(P):
    return (P.net.flat_params().detach().clone(), P.net_1.flat_params().detach().clone() if P.net_1 is not None else None,
            P.vis_t_minus.clone() if getattr(P, "vis_t_minus", None) is not None else None)


def _restore(P, st):
    import torch
    with torch.no_grad():
        P.net.flat_params().copy_(st[0])
        if st[1] is not None:
            P.net_1.flat_params().copy_(st[1])
    if st[2] is not None:
        P.vis_t_minus = st[2].clone()


def test_fused_ev_loop_matches_reference_style_loop():
    import torch
    P = _ev()
    st0 = _state(P)
    steps = 12
    # reference-style loop (ev :440-487): freeze at 0, fresh Adam again at epoch 1
    P.train(num_epoch=steps, lr=1e-3)
    ref = _state(P)
    ref_loss = float(P.loss)
    for graph in (False, True):
        _restore(P, st0)
        P.enable_fused_step(True, graph=graph)
        P.train(num_epoch=steps, lr=1e-3)
        got = _state(P)
        assert P.adam_step_count() == steps - 1            # the optimizer was re-created at epoch 1 (ev :461-462)
        d = (got[0] - ref[0]).abs().max().item()
        assert d < 2e-6, (graph, d)                         # 12 steps of lr = 1e-3: identical up to Adam's rounding order
        assert torch.equal(got[1], ref[1])                  # net_1 frozen
        assert abs(float(P.loss) - ref_loss) < 1e-5 * abs(ref_loss)
        assert (got[2] - ref[2]).abs().max().item() <= 1e-6 * ref[2].abs().max().item() + 1e-12
        if graph:
            assert any(isinstance(v, tuple) for v in P._graphs.values())      # the iteration really ran from a captured graph
            fused_eager = eager_params
            assert (got[0] - fused_eager).abs().max().item() < 5e-7           # graph replay == the same kernels launched eagerly
        eager_params = got[0]
    # an unfrozen step (epoch % 10000 == 0): fresh Adam over both nets, every parameter moves by lr (SURVEY a8 probe)
    _restore(P, st0)
    P.enable_fused_step(True, graph=True)
    P.opt.param_groups[0]["lr"] = 1e-3
    P.defreeze_evm_net(10000)
    before = _state(P)
    P._fused_step_replayable()
    after = _state(P)
    moved = (after[1] - before[1]).abs()
    assert abs(moved.median().item() - 1e-3) < 1e-5 and moved.max().item() < 1.01e-3
    # ... and against torch.optim.Adam on the same state
    P.enable_fused_step(False)
    _restore(P, st0)
    P.defreeze_evm_net(10000)
    loss, _ = P.fwd_computing_loss_2d(); P.opt.zero_grad(); loss.backward(); P.opt.step()
    t = _state(P)
    assert (t[0] - after[0]).abs().max().item() < 2e-6 and (t[1] - after[1]).abs().max().item() < 2e-6


def test_fused_ns_loop_keeps_one_adam_across_stages():
    import torch
    from nsfnet_b200.nsfnet import PysicsInformedNeuralNetwork
    torch.manual_seed(3)
    P = PysicsInformedNeuralNetwork(Re=100, layers=4, hidden_size=120, N_f=2000, bc_weight=10, eq_weight=1)
    rng = np.random.default_rng(0)
    xb, yb, ub, vb = J.cavity_boundary(65)
    P.set_boundary_data((xb, yb, ub, vb))
    P.set_eq_training_data((rng.random(2000).astype(np.float32), rng.random(2000).astype(np.float32)))
    P.log_interval = 10 ** 9; P.checkpoints = False
    p0 = P.net.flat_params().detach().clone()
    P.train(num_epoch=6, lr=1e-3); P.train(num_epoch=6, lr=2e-4)     # two stages, one Adam (NSFnet :76-79, :234-235)
    ref = P.net.flat_params().detach().clone()
    with torch.no_grad():
        P.net.flat_params().copy_(p0)
    P.opt = torch.optim.Adam(P.net.parameters(), lr=1e-3, weight_decay=0.0)
    P.enable_fused_step(True)
    P.train(num_epoch=6, lr=1e-3); P.train(num_epoch=6, lr=2e-4)
    assert P.adam_step_count() == 12
    assert (P.net.flat_params() - ref).abs().max().item() < 2e-6


def test_device_point_layer_matches_host_layer():
    import torch
    from nsfnet_b200.cavity_data import (DataLoader, DeviceDataLoader, cavity_boundary, lhs_sample_device, sdf_weights,
                                         sdf_weights_device, wall_distance, wall_distance_device)
    dev = torch.device("cuda:0")
    n = 1_000_000
    x, y = lhs_sample_device(n, dev, seed=7)
    for c in (x, y):      # exactly one point per 1/N stratum (tools.py:30-57)
        cells = torch.floor(c.double() * n).long()
        assert torch.equal(torch.sort(cells).values, torch.arange(n, device=dev))
    assert abs(torch.corrcoef(torch.stack([x, y]))[0, 1].item()) < 0.01
    xs, ys = lhs_sample_device(n, dev, seed=7, first=250_000, count=1000)          # a rank's row range of the same design
    assert torch.equal(xs, x[250_000:251_000]) and torch.equal(ys, y[250_000:251_000])
    xb, yb, _, _ = cavity_boundary(513)
    xbd = torch.as_tensor(xb.ravel(), dtype=torch.float32, device=dev); ybd = torch.as_tensor(yb.ravel(), dtype=torch.float32, device=dev)
    m = 200_000
    d = wall_distance_device(x[:m].contiguous(), y[:m].contiguous(), xbd, ybd).cpu().numpy()
    pts = np.stack([x[:m].cpu().numpy(), y[:m].cpu().numpy()], 1).astype(np.float64)
    bc = np.stack([xbd.cpu().numpy(), ybd.cpu().numpy()], 1).astype(np.float64)
    assert np.max(np.abs(d - wall_distance(pts, bc))) < 2e-7
    w = sdf_weights_device(x[:m].contiguous(), y[:m].contiguous(), xbd, ybd, 0.2, 5.0).cpu().numpy()
    assert np.max(np.abs(w - sdf_weights(pts, bc, 0.2, 5.0))) < 2e-6 and abs(w.mean() - 1) < 1e-5

    class Cfg:
        enabled, min_weight, decay = True, 0.2, 5.0
    dl = DeviceDataLoader(dev, N_f=50_000, sdf_weighting=Cfg(), sort_training_points=True, seed=1)
    dl.loading_boundary_data()
    xd, yd = dl.loading_training_data()
    wd = dl.get_sdf_weights()
    assert xd.is_cuda and xd.shape == (50_000,) and wd.shape == (50_000,) and abs(wd.mean().item() - 1) < 1e-5
    dist = wall_distance_device(xd, yd, xbd, ybd)
    assert torch.all(dist[1:] >= dist[:-1])                 # sorted by wall distance (tools.py:68-83)
    # the solver takes the shard as is
    P = _ev(n_f=1000)
    P.set_eq_training_shard((xd, yd), weights=wd)
    assert P.x_f.data_ptr() == xd.data_ptr() and P._n_f_global == 50_000
    loss, _ = P.fwd_computing_loss_2d()
    assert np.isfinite(float(loss))


def test_deferred_init_vis_t_equals_the_setter_time_forward():
    """`init_vis_t` (ev :138-140) is fused into the first loss evaluation on new points (NSF_VTM_FROM_E); reading
    `vis_t_minus` earlier, or changing net_1 in between, must give exactly what the reference's eager call gives."""
    import torch

    def run(mode):
        P = _ev(seed=4)
        with torch.no_grad():
            P.net_1.flat_params().mul_(4.0)           # |e| large enough that min(vis_t0, alpha |e|) is not the cap everywhere
        P.set_eq_training_data((P.x_f.cpu().numpy(), P.y_f.cpu().numpy()))      # (re)initialises the lag state with alpha = 0.05
        if mode == "eager":
            assert P._vtm_pending is not None
            _ = P.vis_t_minus                          # evaluated on the spot
            assert P._vtm_pending is None
        if mode in ("changed", "changed_eager"):
            if mode == "changed_eager":
                _ = P.vis_t_minus
            with torch.no_grad():
                P.net_1.layers.layer_0.weight.mul_(1.25)     # through the module: the deferred forward must use the old weights
        P.set_alpha_evm(0.03)
        P.freeze_evm_net(0)
        loss, _ = P.fwd_computing_loss_2d(); P.opt.zero_grad(); loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in P.net.parameters()]).clone()
        return float(loss), g, P.vis_t.clone(), P.vis_t_minus.clone()
    le, ge, ve, me = run("eager")
    ll, gl, vl, ml = run("lazy")
    assert ll == le and torch.equal(gl, ge) and torch.equal(vl, ve) and torch.equal(ml, me)
    assert (ve < 20.0 / 2000.0).float().mean().item() > 0.005      # the cap is not active everywhere
    lc, gc, vc, mc = run("changed")
    lce, gce, vce, mce = run("changed_eager")
    assert lc == lce and torch.equal(gc, gce) and torch.equal(vc, vce) and torch.equal(mc, mce)
    assert torch.equal(vc, ve) and not torch.equal(mc, me)     # vis_t from the OLD weights' e, the new lag state from the new ones
