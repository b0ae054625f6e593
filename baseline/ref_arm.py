"""Reference arm: the UNMODIFIED reference solvers (a verbatim copy of /root/reference/{NSFnet,ev-NSFnet} under
baseline/_ref/, git-ignored, shipped to the GPU box by gpurun) driven through their own public API.

Used only by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` / ``reference_cuda_eager`` legs).  Nothing of this
repo's package runs on these paths: the nets are the reference's ``FCNet``, the derivatives its seven
``torch.autograd.grad(create_graph=True)`` sweeps, the gradient its ``loss.backward()``, the update ``torch.optim.Adam``.

The only adaptation is outside the reference's code: ``matplotlib`` (imported and never used by tools.py) is stubbed, and
for a CPU run of the ev solver -- whose constructor hard-selects ``cuda:{LOCAL_RANK}`` (ev-NSFnet/pinn_solver.py:57-63) --
the constructed object's nets are moved to the CPU and its ``device`` attribute is set accordingly (on a box without any GPU
the object is allocated with ``object.__new__`` and the same attributes are set by hand, as tests/golden/make_golden.py does).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "ev-NSFnet", "pinn_solver.py")) and os.path.isfile(os.path.join(REF, "NSFnet", "pinn_solver.py"))


def _load(variant):
    mpl = types.ModuleType("matplotlib"); mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl); sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    d = os.path.join(REF, variant)
    sys.path.insert(0, d)
    try:
        for m in ("net", "pinn_solver", "tools", "cavity_data"):
            sys.modules.pop(m, None)
        spec = importlib.util.spec_from_file_location("ref_" + variant.replace("-", "_"), os.path.join(d, "pinn_solver.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        spec2 = importlib.util.spec_from_file_location("ref_data_" + variant.replace("-", "_"), os.path.join(d, "cavity_data.py"))
        data = importlib.util.module_from_spec(spec2)
        spec2.loader.exec_module(data)
    finally:
        sys.path.remove(d)
    return mod, data


def _quiet(P):
    P.log_interval = 10 ** 9
    P.print_log = lambda *a, **k: None
    P.save = lambda *a, **k: None


def build(workload: str, device: str, n_f: int, seed: int = 0):
    """The reference solver object of `workload` ("ev": ev-NSFnet Re=2000, 6x80 + 4x40; "ns": NSFnet Re=1000, 4x120) on `device`,
    with the reference's own 2052 boundary points and n_f uniform collocation points."""
    variant = "ev-NSFnet" if workload == "ev" else "NSFnet"
    mod, data = _load(variant)
    dev = torch.device(device)
    torch.manual_seed(seed)
    os.environ.setdefault("RANK", "0"); os.environ.setdefault("LOCAL_RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
    if workload == "ev":
        kw = dict(Re=2000, layers=6, layers_1=4, hidden_size=80, hidden_size_1=40, N_f=n_f, alpha_evm=0.05, bc_weight=10, eq_weight=1,
                  supervised_data_weight=0.0)
        saved_ws = os.environ.get("WORLD_SIZE")
        os.environ["WORLD_SIZE"] = "1"; os.environ["RANK"] = "0"         # the arm is a single-process run even under torchrun
        try:
            if torch.cuda.is_available():
                import contextlib
                import io
                with contextlib.redirect_stdout(io.StringIO()):
                    P = mod.PysicsInformedNeuralNetwork(**kw)
                if dev.type == "cpu":
                    P.net = P.net.to(dev); P.net_1 = P.net_1.to(dev); P.device = dev
                    P.opt = torch.optim.Adam(list(P.net.parameters()) + list(P.net_1.parameters()), lr=1e-3, weight_decay=0.0)
            else:
                P = object.__new__(mod.PysicsInformedNeuralNetwork)
                P.rank = 0; P.local_rank = 0; P.world_size = 1; P.device = dev; P.is_distributed = False
                P.Re = 2000; P.vis_t0 = 20.0 / 2000; P.alpha_evm = 0.05; P.alpha_b = 10; P.alpha_e = 1; P.alpha_s = 0.0
                P.evm = None; P.vis_t = None; P.vis_t_minus = None; P.eq_weights = None; P.coord_scale = 1.0; P.coord_scale_sq = 1.0
                P.x_s = P.y_s = P.u_s = P.v_s = P.p_s = None; P._p_mask = None
                P.supervision_enabled = False; P.supervision_point_count = 0; P.supervision_total_points = 0; P.supervision_has_data = False
                P.loss_s = 0.0; P.N_f = n_f; P.current_stage = ' '; P.layers = 6; P.hidden_size = 80; P.layers_1 = 4; P.hidden_size_1 = 40
                P.net = P.initialize_NN(num_ins=2, num_outs=3, num_layers=6, hidden_size=80)
                P.net_1 = P.initialize_NN(num_ins=2, num_outs=1, num_layers=4, hidden_size=40)
                P.opt = torch.optim.Adam(list(P.net.parameters()) + list(P.net_1.parameters()), lr=1e-3, weight_decay=0.0)
        finally:
            if saved_ws is not None:
                os.environ["WORLD_SIZE"] = saved_ws
        P.world_size = 1; P.rank = 0
    else:
        mod.device = dev                  # module-level global read at call time (NSFnet/pinn_solver.py:24)
        P = mod.PysicsInformedNeuralNetwork(Re=1000, layers=4, hidden_size=120, N_f=n_f, bc_weight=10, eq_weight=1)
        P.net = P.net.to(dev)
        P.opt = torch.optim.Adam(P.net.parameters(), lr=1e-3, weight_decay=0)
    _quiet(P)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        dl = data.DataLoader(N_f=n_f, N_b=1000)
        xb, yb, ub, vb = dl.loading_boundary_data()
        rng = np.random.default_rng(seed)
        P.set_boundary_data(X=(xb, yb, ub, vb))
        P.set_eq_training_data(X=(rng.random((n_f, 1)), rng.random((n_f, 1))))
        if workload == "ev":
            P.freeze_evm_net(0)
    return P


def loop_body(P, workload):
    """One iteration of the reference's solve_Adam loop (ev-NSFnet/pinn_solver.py:465-472, NSFnet/pinn_solver.py:250-254)."""
    if workload == "ev":
        loss, _ = P.fwd_computing_loss_2d()
        P.opt.zero_grad()
        loss.backward()
        P.opt.step()
    else:
        loss, _ = P.fwd_computing_loss_2d()
        loss.backward()
        P.opt.step()
        P.opt.zero_grad()
    return loss


def time_steps(workload: str, device: str, n_f: int, steps: int, warmup: int, threads: int | None = None, seed: int = 0):
    """Seconds per full reference iteration at n_f collocation points.  Returns (sec_per_step, threads_used, last_loss)."""
    threads = threads or os.cpu_count()
    if device == "cpu":
        torch.set_num_threads(threads)
    P = build(workload, device, n_f, seed)
    cuda = device != "cpu"
    for _ in range(warmup):
        loop_body(P, workload)
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    loss = None
    for _ in range(steps):
        loss = loop_body(P, workload)
    if cuda:
        torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / max(steps, 1)
    out = float(loss.detach()) if loss is not None else float("nan")
    del P
    if cuda:
        torch.cuda.empty_cache()
    return sec, threads, out
