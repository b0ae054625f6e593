"""CPU oracle (numpy) for the NSFnet / ev-NSFnet training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``nsfnet_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs do.

It restates, in closed form, what the reference computes with eight ``torch.autograd``
sweeps per step (all citations relative to /root/reference):

* tanh MLP ``FCNet``                                   -- ``NSFnet/net.py:22-54`` (= ``ev-NSFnet/net.py``)
* ``neural_net_equations`` (u,v,p,e + first/second x/y
  derivatives, eq1..eq3 and the ev residual eq4)       -- ``ev-NSFnet/pinn_solver.py:290-342``,
                                                          ``NSFnet/pinn_solver.py:132-163``
* lagged entropy viscosity ``vis_t``                   -- ``ev-NSFnet/pinn_solver.py:67,138-140,327-334``
* ``fwd_computing_loss_2d`` (SDF-weighted MSE,
  boundary MSE, optional supervised MSE)               -- ``ev-NSFnet/pinn_solver.py:372-428``,
                                                          ``NSFnet/pinn_solver.py:197-226``
* ``loss.backward()`` (parameter gradients)            -- ``ev-NSFnet/pinn_solver.py:468-469``

Instead of autograd it propagates a second-order Taylor jet (streams 0,x,y,xx,yy) through
the MLP and runs the hand-derived adjoint, which is exactly the algorithm the CUDA kernels
implement.  Parity of this restatement is *pinned* against outputs of the reference's own
code (``tests/golden/*.npz``, produced by ``tests/golden/make_golden.py`` which imports
/root/reference) by ``tests/test_oracle_golden.py``.

Flat parameter order = ``state_dict`` order of ``FCNet``: W0 [out,in] row-major, b0, W1, b1, ...
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# network description / parameter packing
# --------------------------------------------------------------------------------------
@dataclass(frozen=True)
class NetDesc:
    """Shape of one FCNet: n_in -> hidden x n_hidden_layers -> n_out (net.py:30)."""
    n_in: int
    n_out: int
    n_hidden_layers: int
    hidden: int

    @property
    def dims(self) -> List[int]:
        return [self.n_in] + [self.hidden] * self.n_hidden_layers + [self.n_out]

    @property
    def n_params(self) -> int:
        d = self.dims
        return sum(d[i] * d[i + 1] + d[i + 1] for i in range(len(d) - 1))

    def offsets(self) -> List[Tuple[int, int, int, int]]:
        """[(w_off, b_off, out, in)] per linear layer in the flat buffer."""
        res, off, d = [], 0, self.dims
        for i in range(len(d) - 1):
            fi, fo = d[i], d[i + 1]
            res.append((off, off + fi * fo, fo, fi))
            off += fi * fo + fo
        return res


def unpack(flat: np.ndarray, desc: NetDesc, dtype=np.float64):
    flat = np.asarray(flat)
    assert flat.size == desc.n_params, (flat.size, desc.n_params)
    out = []
    for (wo, bo, fo, fi) in desc.offsets():
        out.append((flat[wo:wo + fo * fi].reshape(fo, fi).astype(dtype),
                    flat[bo:bo + fo].astype(dtype)))
    return out


def pack(layers: Sequence[Tuple[np.ndarray, np.ndarray]], dtype=np.float32) -> np.ndarray:
    return np.concatenate([np.concatenate([W.reshape(-1), b.reshape(-1)]) for W, b in layers]).astype(dtype)


def init_params(desc: NetDesc, seed: int) -> np.ndarray:
    """torch.nn.Linear-style U(+-1/sqrt(fan_in)) init (net.py:40,45), numpy RNG (not bit-equal to torch)."""
    rng = np.random.default_rng(seed)
    layers = []
    d = desc.dims
    for i in range(len(d) - 1):
        k = 1.0 / np.sqrt(d[i])
        layers.append((rng.uniform(-k, k, size=(d[i + 1], d[i])), rng.uniform(-k, k, size=(d[i + 1],))))
    return pack(layers)


# --------------------------------------------------------------------------------------
# value-only forward / reverse (neural_net_u; boundary + supervised terms)
# --------------------------------------------------------------------------------------
def mlp_forward(layers, X):
    """FCNet.forward (net.py:52-54).  X [N,n_in] -> [N,n_out]; also returns the tanh outputs."""
    a = X
    acts = [a]
    for (W, b) in layers[:-1]:
        a = np.tanh(a @ W.T + b)
        acts.append(a)
    W, b = layers[-1]
    return a @ W.T + b, acts


def mlp_backward(layers, acts, out_bar):
    """Adjoint of mlp_forward for a given d(loss)/d(out) [N,n_out] -> [(dW, db)]."""
    grads = [None] * len(layers)
    W, b = layers[-1]
    grads[-1] = (out_bar.T @ acts[-1], out_bar.sum(0))
    a_bar = out_bar @ W
    for l in range(len(layers) - 2, -1, -1):
        t = acts[l + 1]
        z_bar = a_bar * (1.0 - t * t)
        grads[l] = (z_bar.T @ acts[l], z_bar.sum(0))
        if l > 0:
            a_bar = z_bar @ layers[l][0]
    return grads


# --------------------------------------------------------------------------------------
# second-order jet forward / reverse (neural_net_equations + backward)
# --------------------------------------------------------------------------------------
STREAMS = ("0", "x", "y", "xx", "yy")


def mlp_jet_forward(layers, x, y):
    """Propagate (value, d/dx, d/dy, d2/dx2, d2/dy2) through the tanh MLP.

    Equivalent to the reference's u/v/p + 7 autograd.grad sweeps
    (ev-NSFnet/pinn_solver.py:291-309) but in one pass.  Returns out[s] -> [N,n_out] and the
    per-layer stash (t, zx, zy, zxx, zyy) and post-activation streams needed by the adjoint.
    """
    N = x.shape[0]
    dt = layers[0][0].dtype
    a = {"0": np.stack([x, y], 1).astype(dt),
         "x": np.tile(np.array([1.0, 0.0], dt), (N, 1)),
         "y": np.tile(np.array([0.0, 1.0], dt), (N, 1)),
         "xx": np.zeros((N, 2), dt), "yy": np.zeros((N, 2), dt)}
    ins, stash = [a], []
    for (W, b) in layers[:-1]:
        z = {s: a[s] @ W.T for s in STREAMS}
        z["0"] = z["0"] + b
        t = np.tanh(z["0"])
        d1 = 1.0 - t * t
        d2 = -2.0 * t * d1
        a = {"0": t, "x": d1 * z["x"], "y": d1 * z["y"],
             "xx": d2 * z["x"] ** 2 + d1 * z["xx"], "yy": d2 * z["y"] ** 2 + d1 * z["yy"]}
        stash.append((t, z["x"], z["y"], z["xx"], z["yy"]))
        ins.append(a)
    W, b = layers[-1]
    out = {s: a[s] @ W.T for s in STREAMS}
    out["0"] = out["0"] + b
    return out, ins, stash


def mlp_jet_backward(layers, ins, stash, out_bar):
    """Adjoint of mlp_jet_forward.  out_bar[s] = d(loss)/d(out[s]) [N,n_out] -> [(dW, db)]."""
    grads = [None] * len(layers)
    W, b = layers[-1]
    dW = sum(out_bar[s].T @ ins[-1][s] for s in STREAMS)
    grads[-1] = (dW, out_bar["0"].sum(0))
    a_bar = {s: out_bar[s] @ W for s in STREAMS}
    for l in range(len(layers) - 2, -1, -1):
        t, zx, zy, zxx, zyy = stash[l]
        d1 = 1.0 - t * t
        d2 = -2.0 * t * d1
        d3 = -2.0 * d1 * (1.0 - 3.0 * t * t)
        zb = {
            "x": a_bar["x"] * d1 + 2.0 * a_bar["xx"] * d2 * zx,
            "y": a_bar["y"] * d1 + 2.0 * a_bar["yy"] * d2 * zy,
            "xx": a_bar["xx"] * d1,
            "yy": a_bar["yy"] * d1,
            "0": (a_bar["0"] * d1 + a_bar["x"] * d2 * zx + a_bar["y"] * d2 * zy
                  + a_bar["xx"] * (d3 * zx * zx + d2 * zxx) + a_bar["yy"] * (d3 * zy * zy + d2 * zyy)),
        }
        dW = sum(zb[s].T @ ins[l][s] for s in STREAMS)
        grads[l] = (dW, zb["0"].sum(0))
        if l > 0:
            Wl = layers[l][0]
            a_bar = {s: zb[s] @ Wl for s in STREAMS}
    return grads


# --------------------------------------------------------------------------------------
# physics
# --------------------------------------------------------------------------------------
@dataclass
class Physics:
    Re: float
    alpha_b: float = 10.0          # bc_weight   (ev :40, NSFnet/train.py:26)
    alpha_e: float = 1.0           # eq_weight
    alpha_s: float = 0.0           # supervised_data_weight
    alpha_evm: float = 0.03        # ev only (ev :39)
    vis_t0: Optional[float] = None  # ev: 20/Re (ev :67)
    coord_scale: float = 1.0       # ev :311-324
    eq4_weight: float = 0.1        # ev :397
    has_evm: bool = False
    evm_trainable: bool = False    # net_1 requires_grad (ev :489-511)

    def __post_init__(self):
        if self.vis_t0 is None:
            self.vis_t0 = 20.0 / self.Re


@dataclass
class StepResult:
    loss: float
    loss_e: float
    loss_b: float
    loss_s: float
    loss_eq: List[float]
    eq: List[np.ndarray]                 # eq1, eq2, eq3[, eq4]   each [N]
    grad_main: np.ndarray                # flat, state_dict order
    grad_evm: Optional[np.ndarray]
    e: Optional[np.ndarray]              # net_1 output at the collocation points [N]
    vis_t: Optional[np.ndarray]          # viscosity actually used this step [N]
    vis_t_minus_next: Optional[np.ndarray]  # alpha_evm*|e| handed to the next step (ev :334)
    extras: dict = field(default_factory=dict)


def residuals(out, nu, e=None, coord_scale=1.0):
    """eq1..eq3 (and eq4 if e is given) from the output jet (ev :311-342, NSFnet :159-163)."""
    s1, s2 = coord_scale, coord_scale * coord_scale
    u, v = out["0"][:, 0], out["0"][:, 1]
    ux, uy, uxx, uyy = s1 * out["x"][:, 0], s1 * out["y"][:, 0], s2 * out["xx"][:, 0], s2 * out["yy"][:, 0]
    vx, vy, vxx, vyy = s1 * out["x"][:, 1], s1 * out["y"][:, 1], s2 * out["xx"][:, 1], s2 * out["yy"][:, 1]
    px, py = s1 * out["x"][:, 2], s1 * out["y"][:, 2]
    eq1 = (u * ux + v * uy) + px - nu * (uxx + uyy)
    eq2 = (u * vx + v * vy) + py - nu * (vxx + vyy)
    eq3 = ux + vy
    eqs = [eq1, eq2, eq3]
    if e is not None:
        eqs.append((eq1 * (u - 0.5) + eq2 * (v - 0.5)) - e)
    d = dict(u=u, v=v, ux=ux, uy=uy, vx=vx, vy=vy)
    return eqs, d


def step(main_flat, main_desc: NetDesc, phys: Physics, x, y, xb, yb, ub, vb,
         evm_flat=None, evm_desc: Optional[NetDesc] = None, w=None, vis_t_minus=None,
         sup=None, n_f_norm=None, n_b_norm=None, dtype=np.float64) -> StepResult:
    """One ``fwd_computing_loss_2d()`` + ``loss.backward()`` (ev :372-428,:468-469).

    ``vis_t_minus``: alpha_evm*|e| of the previous evaluation (None -> constant vis_t0, ev :327-331).
    ``w``: per-point SDF weights (ev :387-392) or None.
    ``sup``: optional (x,y,u,v,p) supervised block with NaN-able p (ev :399-411).
    ``n_f_norm``/``n_b_norm``: mean denominators (global counts under data parallelism).
    """
    x = np.asarray(x, dtype).reshape(-1)
    y = np.asarray(y, dtype).reshape(-1)
    N = x.shape[0]
    nf = float(n_f_norm if n_f_norm is not None else N)
    layers = unpack(main_flat, main_desc, dtype)
    out, ins, stash = mlp_jet_forward(layers, x, y)

    e = None
    if phys.has_evm:
        evm_layers = unpack(evm_flat, evm_desc, dtype)
        e_out, e_acts = mlp_forward(evm_layers, np.stack([x, y], 1))
        e = e_out[:, 0]
        if vis_t_minus is not None:
            vis_t = np.minimum(dtype(phys.vis_t0), np.asarray(vis_t_minus, dtype).reshape(-1))
        else:
            vis_t = np.full(N, phys.vis_t0, dtype)
        # the reference rounds vis_t to fp32 (torch.tensor(...).float(), ev :328-331)
        vis_t = vis_t.astype(np.float32).astype(dtype)
        nu = 1.0 / phys.Re + vis_t
    else:
        vis_t = None
        nu = np.full(N, 1.0 / phys.Re, dtype)

    eqs, d = residuals(out, nu, e, phys.coord_scale)
    wv = np.ones(N, dtype) if w is None else np.asarray(w, dtype).reshape(-1)
    loss_eq = [float(np.sum(wv * r * r) / nf) for r in eqs]
    loss_e = loss_eq[0] + loss_eq[1] + loss_eq[2] + (phys.eq4_weight * loss_eq[3] if phys.has_evm else 0.0)

    # ---- adjoint seeds (SURVEY.md 8a "math contract") ----
    c = phys.alpha_e / nf
    u, v = d["u"], d["v"]
    eq1, eq2, eq3 = eqs[:3]
    eq4 = eqs[3] if phys.has_evm else np.zeros(N, dtype)
    k4 = 2.0 * phys.eq4_weight if phys.has_evm else 0.0
    g1 = c * wv * (2.0 * eq1 + k4 * eq4 * (u - 0.5))
    g2 = c * wv * (2.0 * eq2 + k4 * eq4 * (v - 0.5))
    g3 = 2.0 * c * wv * eq3
    g4 = k4 * c * wv * eq4
    s1, s2 = phys.coord_scale, phys.coord_scale ** 2
    zero = np.zeros(N, dtype)
    ob = {
        "0": np.stack([g1 * d["ux"] + g2 * d["vx"] + g4 * eq1, g1 * d["uy"] + g2 * d["vy"] + g4 * eq2, zero], 1),
        "x": s1 * np.stack([g1 * u + g3, g2 * u, g1], 1),
        "y": s1 * np.stack([g1 * v, g2 * v + g3, g2], 1),
        "xx": s2 * np.stack([-nu * g1, -nu * g2, zero], 1),
        "yy": s2 * np.stack([-nu * g1, -nu * g2, zero], 1),
    }
    if main_desc.n_out > 3:
        for s in STREAMS:
            ob[s] = np.concatenate([ob[s], np.zeros((N, main_desc.n_out - 3), dtype)], 1)
    grads = mlp_jet_backward(layers, ins, stash, ob)

    # ---- boundary term (value stream only) ----
    xb = np.asarray(xb, dtype).reshape(-1); yb = np.asarray(yb, dtype).reshape(-1)
    ub = np.asarray(ub, dtype).reshape(-1); vb = np.asarray(vb, dtype).reshape(-1)
    nb = float(n_b_norm if n_b_norm is not None else xb.shape[0])
    ob_out, b_acts = mlp_forward(layers, np.stack([xb, yb], 1))
    du, dv = ob_out[:, 0] - ub, ob_out[:, 1] - vb
    loss_b = float(np.sum(du * du) / nb + np.sum(dv * dv) / nb)
    bbar = np.zeros_like(ob_out)
    bbar[:, 0] = 2.0 * phys.alpha_b / nb * du
    bbar[:, 1] = 2.0 * phys.alpha_b / nb * dv
    gb = mlp_backward(layers, b_acts, bbar)
    grads = [(g[0] + h[0], g[1] + h[1]) for g, h in zip(grads, gb)]

    # ---- optional supervised term (ev :399-411) ----
    loss_s = 0.0
    if sup is not None and phys.alpha_s != 0.0:
        xs, ys, us, vs, ps = [None if a is None else np.asarray(a, dtype).reshape(-1) for a in sup]
        ns = float(xs.shape[0])
        so, s_acts = mlp_forward(layers, np.stack([xs, ys], 1))
        sbar = np.zeros_like(so)
        du, dv = so[:, 0] - us, so[:, 1] - vs
        loss_s = float(np.mean(du * du) + np.mean(dv * dv))
        sbar[:, 0] = 2.0 * phys.alpha_s / ns * du
        sbar[:, 1] = 2.0 * phys.alpha_s / ns * dv
        if ps is not None:
            m = np.isfinite(ps)
            if m.any():
                dp = np.where(m, so[:, 2] - np.where(m, ps, 0.0), 0.0)
                nm = float(m.sum())
                loss_s += float(np.sum(dp * dp) / nm)
                sbar[:, 2] = 2.0 * phys.alpha_s / nm * dp
        gs = mlp_backward(layers, s_acts, sbar)
        grads = [(g[0] + h[0], g[1] + h[1]) for g, h in zip(grads, gs)]

    grad_evm = None
    if phys.has_evm and phys.evm_trainable:
        grad_evm = pack(mlp_backward(evm_layers, e_acts, (-g4).reshape(-1, 1)), dtype)

    loss = phys.alpha_b * loss_b + phys.alpha_e * loss_e + phys.alpha_s * loss_s
    return StepResult(
        loss=loss, loss_e=loss_e, loss_b=loss_b, loss_s=loss_s, loss_eq=loss_eq, eq=eqs,
        grad_main=pack(grads, dtype), grad_evm=grad_evm, e=e, vis_t=vis_t,
        vis_t_minus_next=(phys.alpha_evm * np.abs(e) if e is not None else None),
        extras=dict(out=out))


# --------------------------------------------------------------------------------------
# data helpers used by tests / bench (host side, numpy)
# --------------------------------------------------------------------------------------
def cavity_boundary(n_side: int = 513):
    """The reference's deterministic boundary set (ev-NSFnet/cavity_data.py:47-72): lower, upper
    (regularised lid u = 1 - cosh(10(x-.5))/cosh(5)), left, right; 4*n_side points."""
    s = np.linspace(0.0, 1.0, n_side)
    lid = 1.0 - np.cosh(10.0 * (s - 0.5)) / np.cosh(5.0)
    xb = np.concatenate([s, s, np.zeros(n_side), np.ones(n_side)])
    yb = np.concatenate([np.zeros(n_side), np.ones(n_side), s, s])
    ub = np.concatenate([np.zeros(n_side), lid, np.zeros(n_side), np.zeros(n_side)])
    vb = np.zeros(4 * n_side)
    return xb, yb, ub, vb
