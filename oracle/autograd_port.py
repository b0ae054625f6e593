"""CPU oracle (PyTorch autograd) -- a port of the reference's own evaluation order.

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py CPU-baseline legs).

Where ``jet_numpy`` restates the *math*, this file restates the reference's *algorithm*: the
tanh MLP evaluated by ``torch.nn`` and the derivatives obtained by seven
``torch.autograd.grad(create_graph=True)`` sweeps followed by ``loss.backward()`` --
ev-NSFnet/pinn_solver.py:290-342,372-428,456-472 and NSFnet/pinn_solver.py:132-163,197-226,
240-254.  It is what ``bench.py`` times as the CPU baseline (``cpu_baseline.kind == "port"``)
because /root/reference itself does not travel to the GPU box.  Parity with the real
reference is pinned by tests/golden (see tests/test_oracle_golden.py).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import numpy as np
import torch


class RefNet(torch.nn.Module):
    """Same module tree / state_dict keys as FCNet (net.py:22-54): layers.layer_i.{weight,bias}."""

    def __init__(self, num_ins=2, num_outs=3, num_layers=4, hidden_size=120):
        super().__init__()
        dims = [num_ins] + [hidden_size] * num_layers + [num_outs]
        mods = []
        for i in range(len(dims) - 2):
            mods.append(("layer_%d" % i, torch.nn.Linear(dims[i], dims[i + 1])))
            mods.append(("activation_%d" % i, torch.nn.Tanh()))
        mods.append(("layer_%d" % (len(dims) - 2), torch.nn.Linear(dims[-2], dims[-1])))
        self.layers = torch.nn.Sequential(OrderedDict(mods))

    def forward(self, x):
        return self.layers(x)

    def flat(self) -> np.ndarray:
        return np.concatenate([p.detach().cpu().numpy().reshape(-1) for p in self.parameters()])

    def load_flat(self, flat):
        flat = torch.as_tensor(np.asarray(flat))
        off = 0
        with torch.no_grad():
            for p in self.parameters():
                n = p.numel()
                p.copy_(flat[off:off + n].reshape(p.shape).to(p.dtype))
                off += n
        assert off == flat.numel()

    def flat_grad(self) -> np.ndarray:
        return np.concatenate([(p.grad if p.grad is not None else torch.zeros_like(p)).cpu().numpy().reshape(-1)
                               for p in self.parameters()])


def _grad(y, xs):
    g = torch.autograd.grad([y], xs, grad_outputs=[torch.ones_like(y)], create_graph=True, allow_unused=True)
    return [gi if gi is not None else torch.zeros_like(x) for gi, x in zip(g, xs)]


class RefSolver:
    """Minimal re-statement of PysicsInformedNeuralNetwork's hot path (both variants)."""

    def __init__(self, Re, layers=4, hidden=120, evm_layers: Optional[int] = None, evm_hidden: int = 40,
                 alpha_b=10.0, alpha_e=1.0, alpha_evm=0.03, dtype=torch.float32, lr=1e-3):
        self.Re, self.alpha_b, self.alpha_e, self.alpha_evm = Re, alpha_b, alpha_e, alpha_evm
        self.dtype = dtype
        self.net = RefNet(2, 3, layers, hidden).to(dtype)
        self.net_1 = RefNet(2, 1, evm_layers, evm_hidden).to(dtype) if evm_layers else None
        self.vis_t0 = 20.0 / Re
        self.vis_t_minus = None
        self.eq_weights = None
        self.coord_scale = 1.0
        self.opt = torch.optim.Adam(self.net.parameters(), lr=lr, weight_decay=0.0)

    # -- setters (ev :142-184) ------------------------------------------------------
    def set_boundary_data(self, X):
        self.x_b, self.y_b, self.u_b, self.v_b = [torch.tensor(np.asarray(a)).to(self.dtype).reshape(-1, 1) for a in X]

    def set_eq_training_data(self, X, weights=None):
        self.x_f = torch.tensor(np.asarray(X[0])).to(self.dtype).reshape(-1, 1).requires_grad_(True)
        self.y_f = torch.tensor(np.asarray(X[1])).to(self.dtype).reshape(-1, 1).requires_grad_(True)
        self.eq_weights = None if weights is None else torch.tensor(np.asarray(weights)).to(self.dtype)
        if self.net_1 is not None:      # init_vis_t (ev :138-140)
            with torch.no_grad():
                e = self.net_1(torch.cat((self.x_f, self.y_f), 1))[:, 0:1]
            self.vis_t_minus = self.alpha_evm * e.abs().numpy()

    # -- hot path ---------------------------------------------------------------------
    def equations(self, x, y):
        X = torch.cat((x, y), 1)
        uvp = self.net(X)
        u, v, p = uvp[:, 0:1], uvp[:, 1:2], uvp[:, 2:3]
        u_x, u_y = _grad(u, [x, y]); u_xx = _grad(u_x, [x])[0]; u_yy = _grad(u_y, [y])[0]
        v_x, v_y = _grad(v, [x, y]); v_xx = _grad(v_x, [x])[0]; v_yy = _grad(v_y, [y])[0]
        p_x, p_y = _grad(p, [x, y])
        s, s2 = self.coord_scale, self.coord_scale ** 2
        u_x, u_y, v_x, v_y, p_x, p_y = [s * t for t in (u_x, u_y, v_x, v_y, p_x, p_y)]
        u_xx, u_yy, v_xx, v_yy = [s2 * t for t in (u_xx, u_yy, v_xx, v_yy)]
        if self.net_1 is None:
            nu = 1.0 / self.Re
            e = None
        else:
            e = self.net_1(X)[:, 0:1]
            if self.vis_t_minus is not None:
                self.vis_t = torch.tensor(np.minimum(self.vis_t0, self.vis_t_minus)).float().to(self.dtype)
            else:
                self.vis_t = torch.tensor(self.vis_t0).float().to(self.dtype)
            self.vis_t_minus = self.alpha_evm * e.detach().abs().numpy()
            nu = 1.0 / self.Re + self.vis_t
        eq1 = (u * u_x + v * u_y) + p_x - nu * (u_xx + u_yy)
        eq2 = (u * v_x + v * v_y) + p_y - nu * (v_xx + v_yy)
        eq3 = u_x + v_y
        if e is None:
            return eq1, eq2, eq3
        return eq1, eq2, eq3, (eq1 * (u - 0.5) + eq2 * (v - 0.5)) - e

    def loss_fn(self):
        uvp_b = self.net(torch.cat((self.x_b, self.y_b), 1))
        self.loss_b = torch.mean((self.u_b.reshape(-1) - uvp_b[:, 0]) ** 2) + \
            torch.mean((self.v_b.reshape(-1) - uvp_b[:, 1]) ** 2)
        self.eqs = self.equations(self.x_f, self.y_f)

        def wmse(r):
            r = r.reshape(-1)
            if self.eq_weights is not None and r.numel() == self.eq_weights.numel():
                r = r * torch.sqrt(self.eq_weights.reshape(-1))
            return torch.mean(r * r)
        self.loss_eq = [wmse(r) for r in self.eqs]
        self.loss_e = self.loss_eq[0] + self.loss_eq[1] + self.loss_eq[2]
        if len(self.loss_eq) == 4:
            self.loss_e = self.loss_e + 0.1 * self.loss_eq[3]
        self.loss = self.alpha_b * self.loss_b + self.alpha_e * self.loss_e
        return self.loss

    def adam_step(self):
        """One solve_Adam iteration (ev :465-472)."""
        loss = self.loss_fn()
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return float(loss.detach())


def time_reference_step(kind: str, n_f: int, n_b: int = 2052, steps: int = 3, warmup: int = 1, threads: Optional[int] = None,
                        seed: int = 0):
    """Time full Adam steps of the autograd port on the host cores.  Returns (seconds/step, threads)."""
    import os
    import time
    from .jet_numpy import cavity_boundary
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    if kind == "ev":
        s = RefSolver(2000, 6, 80, 4, 40, alpha_evm=0.05)
    else:
        s = RefSolver(1000, 4, 120)
    rng = np.random.default_rng(seed)
    xb, yb, ub, vb = cavity_boundary(n_b // 4)
    s.set_boundary_data((xb, yb, ub, vb))
    s.set_eq_training_data((rng.random(n_f), rng.random(n_f)))
    for _ in range(warmup):
        s.adam_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.adam_step()
    return (time.perf_counter() - t0) / steps, threads
