"""Shared implementation of the two ``PysicsInformedNeuralNetwork`` solvers (NSFnet / ev-NSFnet).

Mirrors the reference's solver API (NSFnet/pinn_solver.py:26-389, ev-NSFnet/pinn_solver.py:27-765):
same constructor keywords, setters, ``neural_net_u``, ``neural_net_equations``,
``fwd_computing_loss_2d`` -> ``(loss, [loss_e, loss_b])``, ``train`` / ``solve_Adam``,
``freeze_evm_net`` / ``defreeze_evm_net``, ``evaluate`` / ``test`` / ``save``.  Underneath, one
loss + gradient evaluation is ONE call into libnsf_b200.so (``nsf_step``, include/nsf_b200.h)
instead of two module forwards, seven ``torch.autograd.grad`` sweeps and a backward pass; the
returned ``loss`` is a tensor whose ``.backward()`` fills ``p.grad`` for every parameter with
``requires_grad`` (so the reference's own loops, or an L-BFGS closure, work unchanged).  Adam
stays ``torch.optim.Adam``.

Data parallelism (ev-NSFnet/pinn_solver.py:103-106,142-184,414-424): points are sharded in
contiguous blocks exactly like the reference; instead of DDP buckets + three scalar all-reduces
there is a single ``all_reduce(SUM)`` of ``[grad_main | grad_evm | loss partial sums]`` and the
means are taken over the GLOBAL point counts, so the W-rank gradient equals the 1-rank gradient
on the union of the shards.  ``ddp_compat=True`` reproduces the reference's extra 1/W factor.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _capi
from .net import FCNet


def _dev_f32(a, device) -> torch.Tensor:
    """numpy / tensor -> contiguous 1-D fp32 device tensor (the reference's .float().to(device))."""
    if isinstance(a, torch.Tensor):
        t = a.detach()
    else:
        t = torch.as_tensor(np.asarray(a))
    # a PINNED fp32 host tensor is copied asynchronously on the current stream (the kernels that read it are ordered behind the
    # copy; the caller must not overwrite the buffer before the stream gets there); everything else is the reference's blocking copy
    pinned = t.device.type == "cpu" and t.dtype == torch.float32 and t.is_pinned()
    out = t.to(device=device, dtype=torch.float32, non_blocking=pinned).reshape(-1).contiguous()
    if out.data_ptr() % 16:      # an odd-offset view of a device tensor (a shard x[s:e]): the C ABI wants 16-byte aligned point arrays
        out = out.clone()
    return out


def shard_bounds(total: int, rank: int, world_size: int):
    """Contiguous block split of the reference (ev :144-147, :165-168): N // W points per rank, the last
    rank takes the remainder."""
    per = total // world_size
    s = rank * per
    e = s + per if rank < world_size - 1 else total
    return s, e


class _StepFn(torch.autograd.Function):
    """loss = nsf_step(...); backward hands out the gradient the kernel already produced."""

    @staticmethod
    def forward(ctx, solver, *params):
        loss = solver._launch_step()
        ctx.solver = solver
        ctx.n_in = len(params)
        ctx.eval_id = solver._eval_id          # the flat gradient of THIS evaluation lives in solver._gbuf until the next one
        ctx.slices = solver._param_slices
        return loss

    @staticmethod
    def backward(ctx, gout):
        s = ctx.solver
        if ctx.eval_id != s._eval_id:
            raise RuntimeError("loss.backward() after a later loss evaluation: the kernel keeps ONE gradient buffer per solver "
                               "(call backward() before the next fwd_computing_loss_2d / neural_net_equations on the training set)")
        scaled = s._gbuf * gout
        outs = []
        for (net_id, off, numel, shape, req) in ctx.slices:
            if req:
                base = 0 if net_id == 0 else s._n_main
                outs.append(scaled[base + off: base + off + numel].view(shape))
        assert len(outs) == ctx.n_in
        return (None, *outs)


class SolverBase:
    tb_writer = None
    global_step = 0
    HAS_EVM = False
    verbose = True
    # `init_vis_t` (ev :138-140, called by set_eq_training_data :184) is deferred into the first loss evaluation on the new
    # points, which computes the same net_1 forward anyway (NSF_VTM_FROM_E): identical lag state, one forward instead of two.
    # Any read of `vis_t_minus` before that evaluates it on the spot; set to False to always evaluate it in the setter.
    lazy_init_vis_t = True
    _vtm = None
    _vtm_pending = None

    @property
    def vis_t_minus(self):
        if self._vtm_pending is not None:
            self._materialise_vis_t()
        return self._vtm

    @vis_t_minus.setter
    def vis_t_minus(self, value):
        self._vtm = value
        self._vtm_pending = None

    # ------------------------------------------------------------------------------------
    def _init_common(self, Re, layers, hidden_size, N_f, bc_weight, eq_weight, num_ins, num_outs, learning_rate,
                     net_params, opt, layers_1=None, hidden_size_1=None, num_outs_1=1, net_params_1=None,
                     alpha_evm=0.0, supervised_data_weight=0.0):
        if not torch.cuda.is_available():
            raise RuntimeError("nsfnet_b200 needs a CUDA device (sm_100); there is no CPU path")
        self.rank = int(os.environ.get("RANK", 0))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.world_size = int(os.environ.get("WORLD_SIZE", 1))
        self.device = torch.device(f"cuda:{self.local_rank}")
        torch.cuda.set_device(self.local_rank)
        self._lib = _capi.load()

        self.evm = None
        self.Re = Re
        self.layers, self.hidden_size, self.N_f = layers, hidden_size, N_f
        self.layers_1, self.hidden_size_1 = layers_1, hidden_size_1
        self.current_stage = " "
        self.alpha_evm = alpha_evm
        self.alpha_b, self.alpha_e, self.alpha_s = bc_weight, eq_weight, supervised_data_weight
        self.loss_b = self.loss_e = self.loss_s = 0.0
        self.x_s = self.y_s = self.u_s = self.v_s = self.p_s = None
        self.supervision_point_count = 0
        self.supervision_total_points = 0
        self.supervision_has_data = False
        self.supervision_enabled = False
        self._n_s_global = 0
        self._n_ps_global = 0
        self.eq_weights = None
        self.coord_scale = 1.0
        self.coord_scale_sq = 1.0
        self.vis_t = None
        self.vis_t_minus = None
        self.ddp_compat = False
        self.store_residuals = True
        self.x_b = self.x_f = None
        self.graph_replays = 0       # iterations executed as a CUDA-graph replay (introspection)
        self._fused = False          # enable_fused_step(): device-resident Adam + CUDA-graph replay of the iteration
        self._fused_graph = True
        self._graphs = {}
        self._adam = None
        self._eval_id = 0            # bumped by every nsf_step launch (guards _StepFn.backward against a stale gradient buffer)
        self._resid = self._e = self._vis = None

        self.net = self.initialize_NN(num_ins=num_ins, num_outs=num_outs, num_layers=layers, hidden_size=hidden_size).to(self.device)
        self.net_1 = None
        if self.HAS_EVM:
            self.net_1 = self.initialize_NN(num_ins=num_ins, num_outs=num_outs_1, num_layers=layers_1,
                                            hidden_size=hidden_size_1).to(self.device)
        self.is_distributed = dist.is_available() and dist.is_initialized() and self.world_size > 1
        if net_params:
            if self.rank == 0:
                print(f"Loading net params from {net_params}")
            self.net.load_state_dict(torch.load(net_params, map_location=self.device))
        if net_params_1 and self.HAS_EVM:
            if self.rank == 0:
                print(f"Loading net_1 params from {net_params_1}")
            self.net_1.load_state_dict(torch.load(net_params_1, map_location=self.device))
        self.net.flatten_()
        if self.net_1 is not None:
            self.net_1.flatten_()
        if self.is_distributed:  # what DDP's constructor does (ev :103-106): rank 0's weights everywhere
            dist.broadcast(self.net.flat_params(), src=0)
            if self.net_1 is not None:
                dist.broadcast(self.net_1.flat_params(), src=0)

        self._ctx = _capi.Context(self._lib, self.local_rank, self.net.desc, self.net_1.desc if self.net_1 is not None else None)
        self._n_main = self.net.flat_params().numel()
        self._n_evm = self.net_1.flat_params().numel() if self.net_1 is not None else 0
        # [grad_main | grad_evm | loss_parts]: one buffer => one all-reduce
        self._buf = torch.zeros(self._n_main + self._n_evm + _capi.NSF_LOSS_SLOTS, dtype=torch.float32, device=self.device)
        self._gbuf = self._buf[: self._n_main + self._n_evm]
        self._parts = self._buf[self._n_main + self._n_evm:]
        self._param_slices = []

        params = list(self.net.parameters()) + (list(self.net_1.parameters()) if self.net_1 is not None else [])
        self.opt = torch.optim.Adam(params, lr=learning_rate, weight_decay=0.0) if not opt else opt

    def initialize_NN(self, num_ins=3, num_outs=3, num_layers=10, hidden_size=50):
        return FCNet(num_ins=num_ins, num_outs=num_outs, num_layers=num_layers, hidden_size=hidden_size,
                     activation=torch.nn.Tanh)

    # ---- setters (ev :142-262, NSFnet :82-110) -------------------------------------------------
    def _shard(self, total):
        return shard_bounds(total, self.rank, self.world_size)

    def set_boundary_data(self, X=None, time=False):
        total = np.asarray(X[0]).shape[0]
        s, e = self._shard(total)
        self.x_b, self.y_b, self.u_b, self.v_b = [_dev_f32(np.asarray(a)[s:e], self.device) for a in X[:4]]
        self._n_b_global = total
        self._graphs = {}            # captured iterations bake in the old buffer addresses
        if self.rank == 0 and self.verbose:
            print(f"GPU {self.rank}: Processing {e - s} boundary points out of {total} total")

    def set_eq_training_data(self, X=None, time=False, weights=None):
        total = np.asarray(X[0]).shape[0] if not isinstance(X[0], torch.Tensor) else X[0].shape[0]
        s, e = self._shard(total)
        if isinstance(X[0], torch.Tensor):
            self.x_f = _dev_f32(X[0].reshape(-1)[s:e], self.device)
            self.y_f = _dev_f32(X[1].reshape(-1)[s:e], self.device)
        else:
            self.x_f = _dev_f32(np.asarray(X[0])[s:e], self.device)
            self.y_f = _dev_f32(np.asarray(X[1])[s:e], self.device)
        if weights is not None:
            w = weights if isinstance(weights, torch.Tensor) else np.asarray(weights)
            self.eq_weights = _dev_f32(w.reshape(-1)[s:e], self.device)
        else:
            self.eq_weights = None
        self._n_f_global = total
        n = self.x_f.numel()
        if self._resid is None or self._resid.numel() != 4 * n:      # output buffers: once per shard size, not per call
            self._resid = torch.empty(4 * n, dtype=torch.float32, device=self.device)
            self._e = torch.empty(n, dtype=torch.float32, device=self.device)
            self._vis = torch.empty(n, dtype=torch.float32, device=self.device)
        self._graphs = {}            # captured iterations bake in the old point / lag-state addresses
        if self.rank == 0 and self.verbose:
            print(f"GPU {self.rank}: Processing {e - s} equation points out of {total} total")
        if self.HAS_EVM:
            self.init_vis_t()

    def set_eq_training_shard(self, X, weights=None, n_global=None):
        """Extension (no reference counterpart): THIS rank's collocation shard, already on the device or the host, plus
        the global point count.  What every rank of a data-parallel run calls when its shard is generated in place
        (nsfnet_b200.cavity_data.DeviceDataLoader) instead of slicing one W*N host array per rank (ev :165-177)."""
        ws, rk = self.world_size, self.rank
        self.world_size, self.rank = 1, 0
        try:
            self.set_eq_training_data(X, weights=weights)
        finally:
            self.world_size, self.rank = ws, rk
        n_loc = self.x_f.numel()
        if n_global is None:
            n_global = n_loc
            if self.is_distributed:
                t = torch.tensor([n_loc], dtype=torch.int64, device=self.device)
                dist.all_reduce(t)
                n_global = int(t.item())
        self._n_f_global = int(n_global)

    def set_coordinate_transform(self, scale):
        if scale is None or scale <= 0:
            self.coord_scale, self.coord_scale_sq = 1.0, 1.0
        else:
            self.coord_scale = float(scale)
            self.coord_scale_sq = self.coord_scale ** 2

    def clear_supervised_data(self):
        self.x_s = self.y_s = self.u_s = self.v_s = self.p_s = None
        self.supervision_point_count = 0
        self.supervision_total_points = 0
        self.supervision_has_data = False
        self.supervision_enabled = False
        self._n_s_global = self._n_ps_global = 0
        self._graphs = {}

    def set_supervised_data(self, data):
        if data is None:
            return self.clear_supervised_data()
        x, y, u, v, p = [None if a is None else np.asarray(a).reshape(-1) for a in data]
        total = x.shape[0]
        self.supervision_total_points = int(total)
        if total == 0:
            return self.clear_supervised_data()
        idx = np.array_split(np.arange(total), self.world_size)[self.rank] if self.world_size > 1 else slice(None)
        self.x_s, self.y_s, self.u_s, self.v_s = [_dev_f32(a[idx], self.device) for a in (x, y, u, v)]
        self.p_s = _dev_f32(p[idx], self.device) if p is not None else None
        self.supervision_point_count = int(self.x_s.numel())
        self._n_s_global = total
        self._n_ps_global = int(np.isfinite(p).sum()) if p is not None else 0
        self.supervision_has_data = True
        self.supervision_enabled = self.alpha_s != 0.0
        self._graphs = {}

    def set_supervised_loss_weight(self, weight):
        self.alpha_s = float(weight)
        self.supervision_enabled = self.supervision_has_data and self.alpha_s != 0.0

    def set_optimizers(self, opt):
        self.opt = opt

    def set_stage(self, stage):
        self.stage = stage
        self.current_stage = stage

    def set_alpha_evm(self, alpha):
        self.alpha_evm = alpha

    def set_eq_training_func(self, train_data_func):
        self.train_data_func = train_data_func

    # ---- native calls ---------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _phys(self, n_f_norm=0.0, vtm_from_e_alpha=None):
        trainable = self.HAS_EVM and any(p.requires_grad for p in self.net_1.parameters())
        return _capi.physics(float(self.Re), vis_t0=(20.0 / self.Re) if self.HAS_EVM else 0.0, alpha_evm=float(self.alpha_evm),
                             alpha_e=float(self.alpha_e), coord_scale=float(self.coord_scale), eq4_weight=0.1,
                             has_evm=self.HAS_EVM, evm_trainable=trainable, n_f_norm=float(n_f_norm),
                             vtm_from_e_alpha=vtm_from_e_alpha)

    def _forward_net(self, which, x, y, params=None):
        net = self.net if which == 0 else self.net_1
        x = _dev_f32(x, self.device); y = _dev_f32(y, self.device)
        n = x.numel()
        out = torch.empty((n, net.num_outs), dtype=torch.float32, device=self.device)
        flat = net.flat_params() if params is None else params
        self._ctx.forward(which, flat.data_ptr(), x.data_ptr(), y.data_ptr(), n, out.data_ptr(), self._stream())
        return out

    def _evm_version(self):
        """Changes whenever net_1's weights are written through the module (optimizer step, load_state_dict, copy_) or
        through the flat buffer."""
        return (self.net_1.flat_params()._version,) + tuple(p._version for p in self.net_1.parameters())

    def init_vis_t(self):
        """vis_t_minus = alpha_evm * |e(x_f, y_f)| with the current weights (ev :138-140)."""
        if not self.lazy_init_vis_t:
            e = self._forward_net(1, self.x_f, self.y_f)
            self.vis_t_minus = (float(self.alpha_evm) * e.abs()).reshape(-1).contiguous()
            return
        self._vtm = None
        self._vtm_pending = dict(alpha=float(self.alpha_evm), version=self._evm_version(), n=self.x_f.numel(),
                                 snapshot=self.net_1.flat_params().detach().clone(), x=self.x_f, y=self.y_f)

    def _materialise_vis_t(self):
        p, self._vtm_pending = self._vtm_pending, None
        params = None if self._evm_version() == p["version"] else p["snapshot"]     # the weights of the time of the call
        e = self._forward_net(1, p["x"], p["y"], params=params)
        self._vtm = (p["alpha"] * e.abs()).reshape(-1).contiguous()

    def _launch_step(self):
        self._eval_id += 1
        net, net1 = self.net, self.net_1
        pm = net.flat_params()
        pe = net1.flat_params() if net1 is not None else None
        n_f = self.x_f.numel()
        W = self.world_size if self.is_distributed else 1
        n_f_glob = self._n_f_global if W > 1 else n_f
        n_b = self.x_b.numel() if self.x_b is not None else 0
        n_b_glob = self._n_b_global if W > 1 else n_b
        blocks = []
        if n_b > 0:
            c = float(self.alpha_b) / float(n_b_glob)
            blocks.append(_capi.NsfDataBlock(self.x_b.data_ptr(), self.y_b.data_ptr(), self.u_b.data_ptr(), self.v_b.data_ptr(),
                                             None, n_b, c, c, 0.0, 0))
        sup = self.supervision_enabled and self.x_s is not None and self.supervision_point_count > 0
        if sup:
            n_s_glob = self._n_s_global if W > 1 else self.supervision_point_count
            cs = float(self.alpha_s) / float(n_s_glob)
            cp = float(self.alpha_s) / float(self._n_ps_global) if (self.p_s is not None and self._n_ps_global > 0) else 0.0
            blocks.append(_capi.NsfDataBlock(self.x_s.data_ptr(), self.y_s.data_ptr(), self.u_s.data_ptr(), self.v_s.data_ptr(),
                                             self.p_s.data_ptr() if self.p_s is not None else None,
                                             self.supervision_point_count, cs, cs, cp, 0))
        w = self.eq_weights if (self.eq_weights is not None and self.eq_weights.numel() == n_f) else None
        lazy_alpha = None
        pend = self._vtm_pending
        if self.HAS_EVM and pend is not None:
            if pend["n"] == n_f and pend["x"] is self.x_f and self._evm_version() == pend["version"]:
                lazy_alpha = pend["alpha"]          # init_vis_t fused into this evaluation (same weights, same points)
                self._vtm_pending = None
                self._vtm = torch.empty(n_f, dtype=torch.float32, device=self.device)
            else:
                self._materialise_vis_t()
        vtm = self._vtm if (self.HAS_EVM and lazy_alpha is None and self._vtm is not None and self._vtm.numel() == n_f) else None
        if self.HAS_EVM and vtm is None and lazy_alpha is None:
            self.vis_t_minus = torch.empty(n_f, dtype=torch.float32, device=self.device)
        phys = self._phys(n_f_glob, vtm_from_e_alpha=lazy_alpha)
        gm = self._buf
        ge = self._buf[self._n_main:] if self._n_evm else None
        keep = self.store_residuals
        self._ctx.step(pm.data_ptr(), pe.data_ptr() if pe is not None else None, self.x_f.data_ptr(), self.y_f.data_ptr(),
                       w.data_ptr() if w is not None else None, vtm.data_ptr() if vtm is not None else None,
                       self._vtm.data_ptr() if self.HAS_EVM else None, n_f, blocks, phys,
                       gm.data_ptr(), ge.data_ptr() if ge is not None else None, self._parts.data_ptr(),
                       self._resid.data_ptr() if keep else None, self._e.data_ptr() if self.HAS_EVM else None,
                       self._vis.data_ptr() if self.HAS_EVM else None, self._stream())
        if self.HAS_EVM and not (phys.flags & _capi.NSF_EVM_TRAINABLE):
            self._buf[self._n_main:self._n_main + self._n_evm].zero_()
        if W > 1:
            dist.all_reduce(self._buf, op=dist.ReduceOp.SUM)
            if self.ddp_compat:  # the reference's tracked `loss /= world_size` on top of DDP averaging
                self._gbuf.div_(W)

        P = self._parts
        nf = float(n_f_glob)
        self.loss_eq1, self.loss_eq2, self.loss_eq3 = P[0] / nf, P[1] / nf, P[2] / nf
        loss_e = self.loss_eq1 + self.loss_eq2 + self.loss_eq3
        if self.HAS_EVM:
            self.loss_eq4 = P[3] / nf
            loss_e = loss_e + 0.1 * self.loss_eq4
        self.loss_e = loss_e
        self.loss_b = (P[6] + P[7]) / float(max(n_b_glob, 1))
        loss = self.alpha_b * self.loss_b + self.alpha_e * self.loss_e
        if sup:
            k = 10 if n_b > 0 else 6
            n_s_glob = self._n_s_global if W > 1 else self.supervision_point_count
            self.loss_s = (P[k] + P[k + 1]) / float(n_s_glob) + (P[k + 2] / float(self._n_ps_global) if self._n_ps_global > 0 else 0.0)
            loss = loss + self.alpha_s * self.loss_s
        else:
            self.loss_s = torch.zeros((), device=self.device)
        n = n_f
        if keep:
            self.eq1_pred, self.eq2_pred, self.eq3_pred = [self._resid[k * n:(k + 1) * n].view(n, 1) for k in range(3)]
            if self.HAS_EVM:
                self.eq4_pred = self._resid[3 * n:4 * n].view(n, 1)
        if self.HAS_EVM:
            self.evm = self._e.view(n, 1)
            self.vis_t = self._vis.view(n, 1)
        # bookkeeping for backward(): which slices of the flat gradient go to which parameter
        slices = []
        for net_id, nn_ in enumerate([net] + ([net1] if net1 is not None else [])):
            off = 0
            for p in nn_.parameters():
                slices.append((net_id, off, p.numel(), p.shape, p.requires_grad))
                off += p.numel()
        self._param_slices = slices
        return loss

    def _trainable(self) -> List[torch.nn.Parameter]:
        ps = list(self.net.parameters()) + (list(self.net_1.parameters()) if self.net_1 is not None else [])
        return [p for p in ps if p.requires_grad]

    # ---- the reference's hot-path methods -----------------------------------------------------
    def fwd_computing_loss_2d(self, loss_mode="MSE"):
        if loss_mode != "MSE":
            raise NotImplementedError("only the 'MSE' loss mode (the one the reference trains with) is implemented")
        assert self.x_f is not None and self.y_f is not None
        params = self._trainable()
        if torch.is_grad_enabled() and params:
            self.loss = _StepFn.apply(self, *params)
        else:
            self.loss = self._launch_step()
        return self.loss, [self.loss_e, self.loss_b]

    def _equations(self, x, y):
        x = _dev_f32(x, self.device); y = _dev_f32(y, self.device)
        n = x.numel()
        res = torch.empty(4 * n, dtype=torch.float32, device=self.device)
        e = torch.empty(n, dtype=torch.float32, device=self.device) if self.HAS_EVM else None
        vis = torch.empty(n, dtype=torch.float32, device=self.device) if self.HAS_EVM else None
        vtm_in = None
        if self.HAS_EVM:
            if self.vis_t_minus is not None and self.vis_t_minus.numel() == n:
                vtm_in = self.vis_t_minus
            vtm_out = torch.empty(n, dtype=torch.float32, device=self.device)
        self._ctx.residuals(self.net.flat_params().data_ptr(), self.net_1.flat_params().data_ptr() if self.HAS_EVM else None,
                            x.data_ptr(), y.data_ptr(), vtm_in.data_ptr() if vtm_in is not None else None,
                            vtm_out.data_ptr() if self.HAS_EVM else None, n, self._phys(n), res.data_ptr(),
                            e.data_ptr() if e is not None else None, vis.data_ptr() if vis is not None else None, self._stream())
        if self.HAS_EVM:
            self.vis_t_minus = vtm_out  # the reference overwrites the lag state on every call (ev :334)
            self.evm = e.view(n, 1)
            self.vis_t = vis.view(n, 1)
        return [res[k * n:(k + 1) * n].view(n, 1) for k in range(4)]

    def predict(self, net_params, X):
        x, y = X
        return self.neural_net_u(x, y)

    # ---- training loops ----------------------------------------------------------------------
    def train(self, num_epoch=1, lr=1e-4, optimizer=None, scheduler=None, batchsize=None, start_epoch=0):
        if self.opt is not None:
            self.opt.param_groups[0]["lr"] = lr
        else:
            self.opt = torch.optim.Adam(params=self.net.parameters(), lr=lr)
        if start_epoch:
            return self.solve_Adam(self.fwd_computing_loss_2d, num_epoch, batchsize, scheduler, start_epoch=start_epoch)
        return self.solve_Adam(self.fwd_computing_loss_2d, num_epoch, batchsize, scheduler)

    def _sync_fused_lr(self):
        """A scheduler (or the caller) may have changed ``self.opt``'s learning rate: the fused iteration reads it from a
        device scalar, so write it there (a fill on the same buffer -- captured graphs stay valid)."""
        lr = float(self.opt.param_groups[0]["lr"])
        if self._adam is not None and lr != getattr(self, "_fused_lr", None):
            self._adam_set_lr(lr)

    # ---- fused iteration ---------------------------------------------------
    def enable_fused_step(self, enabled=True, graph=True):
        """Extension: run ``solve_Adam`` as ``nsf_step`` + device-resident Adam (``nsf_adam_dev``; same update rule as
        ``torch.optim.Adam``, including the fresh-optimizer resets of freeze / defreeze, ev :489-511) and replay the whole
        iteration as ONE captured CUDA graph.  Semantics of the reference loop (ev :440-487, NSFnet :240-278) are kept:
        same iteration order, logging and checkpoint cadence; ``self.opt`` still carries the learning rate."""
        self._fused, self._fused_graph = bool(enabled), bool(graph)
        self._graphs = {}

    def release_graphs(self):
        """Drop the captured iterations.  Under data parallelism call this before ``dist.destroy_process_group()``: a live
        CUDA graph that holds NCCL kernels makes the communicator teardown wait forever."""
        import gc
        self._graphs = {}
        gc.collect()
        torch.cuda.synchronize(self.device)

    def _adam_reset(self):
        """A new ``torch.optim.Adam``: zero moments, step 0 (betas (0.9, 0.999), eps 1e-8, ev :126-129)."""
        n = self._n_main + self._n_evm
        if self._adam is None:
            self._adam = dict(m=torch.zeros(n, dtype=torch.float32, device=self.device),
                              v=torch.zeros(n, dtype=torch.float32, device=self.device),
                              state=torch.zeros(8, dtype=torch.float32, device=self.device))
        self._adam["m"].zero_(); self._adam["v"].zero_()
        st = self._adam["state"]
        st.copy_(torch.tensor([0.0, 0.9, 0.999, 1e-8, 1.0, 0.0, 0.0, 0.0], dtype=torch.float32))
        st.view(torch.int32)[5] = 0

    def _adam_set_lr(self, lr):
        self._adam["state"][0:1].fill_(float(lr))
        self._fused_lr = float(lr)

    def adam_step_count(self) -> int:
        return int(self._adam["state"].view(torch.int32)[5].item()) if self._adam is not None else 0

    def _fused_iteration(self):
        """One training iteration on the current stream: loss + gradient, Adam on every trainable flat buffer, tick."""
        loss = self._launch_step()
        lib, A = self._lib, self._adam
        st, strm = A["state"].data_ptr(), self._stream()
        nm = self._n_main
        _capi.check(lib, lib.nsf_adam_dev(self.net.flat_params().data_ptr(), self._buf.data_ptr(), A["m"].data_ptr(),
                                          A["v"].data_ptr(), nm, st, strm))
        if self.HAS_EVM and any(p.requires_grad for p in self.net_1.parameters()):
            _capi.check(lib, lib.nsf_adam_dev(self.net_1.flat_params().data_ptr(), self._buf[nm:].data_ptr(), A["m"][nm:].data_ptr(),
                                              A["v"][nm:].data_ptr(), self._n_evm, st, strm))
        _capi.check(lib, lib.nsf_adam_tick(st, strm))
        self.loss = loss
        return loss

    def _graph_key(self):
        ptr = lambda t: 0 if t is None else t.data_ptr()
        trainable = self.HAS_EVM and any(p.requires_grad for p in self.net_1.parameters())
        return (ptr(self.x_f), ptr(self.y_f), self.x_f.numel(), ptr(self.eq_weights), ptr(self._vtm), self._vtm_pending is None,
                ptr(self._resid), ptr(self._e), ptr(self._vis), ptr(self._buf),
                ptr(self.x_b), ptr(self.y_b), ptr(self.u_b), ptr(self.v_b),
                ptr(self.x_s), ptr(self.y_s), ptr(self.u_s), ptr(self.v_s), ptr(self.p_s), self.supervision_enabled,
                float(self.alpha_evm), float(self.alpha_b), float(self.alpha_e), float(self.alpha_s), float(self.coord_scale), float(self.Re),
                trainable, self.store_residuals, self._n_f_global, getattr(self, "_n_b_global", 0), self._n_s_global, self._n_ps_global,
                bool(self.ddp_compat), self.world_size if self.is_distributed else 1,
                ptr(self.net.flat_params()), ptr(self.net_1.flat_params()) if self.net_1 is not None else 0)

    def _fused_step_replayable(self):
        """Iteration through a captured graph when allowed (single process, timing hooks off), else eagerly.  The first call for
        a configuration runs eagerly (it creates the library's workspaces) and captures afterwards."""
        # under data parallelism the NCCL all-reduce sits in the middle of the iteration and is captured with it (NCCL >= 2.9.6;
        # the collective has run eagerly in the warm-up iteration); NSF_FUSED_GRAPH_DDP=0 launches the iteration eagerly instead
        use_graph = self._fused_graph and (not self.is_distributed or os.environ.get("NSF_FUSED_GRAPH_DDP", "1") == "1")
        if not use_graph:
            return self._fused_iteration()
        key = self._graph_key()
        ent = self._graphs.get(key)
        if ent is None:
            loss = self._fused_iteration()          # a real iteration, also the warm-up of everything the capture may not allocate
            key = self._graph_key()                  # (vis_t_minus may have been allocated by it)
            self._graphs[key] = "armed"
            return loss
        if ent == "armed":
            if len(self._graphs) > 8:
                self._graphs = {}
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g):
                self._fused_iteration()
            names = ["loss", "loss_e", "loss_b", "loss_s", "loss_eq1", "loss_eq2", "loss_eq3"] + (["loss_eq4"] if self.HAS_EVM else [])
            ent = (g, {k: getattr(self, k) for k in names})
            self._graphs[key] = ent
        g, outs = ent
        for k, v in outs.items():
            setattr(self, k, v)
        g.replay()
        self.graph_replays += 1
        return self.loss

    def get_runtime_stats(self, epoch_id, num_epoch):
        now = time.time()
        if not hasattr(self, "_epoch_start_wall"):
            return {}
        elapsed = now - self._epoch_start_wall
        avg = (epoch_id + 1) / elapsed if elapsed > 0 else 0.0
        remain = num_epoch - (epoch_id + 1)
        if self.vis_t is not None:
            vm = float(self.vis_t.mean())
            re_eff = 1.0 / (1.0 / self.Re + vm)
        else:
            vm = re_eff = float("nan")
        return dict(avg_it_s=avg, eta_seconds=remain / avg if avg > 0 else float("inf"), vis_t_mean=vm, Re_eff=re_eff)

    def print_log(self, loss, losses, epoch_id, num_epoch):
        lr = self.opt.param_groups[0]["lr"]
        now = time.time()
        start = getattr(self, "_epoch_start_wall", now)
        it_s = (epoch_id + 1) / max(now - start, 1e-9)
        n_pts = (self.x_f.numel() if self.x_f is not None else 0) + (self.x_b.numel() if self.x_b is not None else 0)
        msg = (f"[{self.current_stage}] epoch {epoch_id + 1}/{num_epoch} lr={lr:.2e} loss={float(loss.detach()):.4e} "
               f"eq1={float(self.loss_eq1):.3e} eq2={float(self.loss_eq2):.3e} eq3={float(self.loss_eq3):.3e}")
        if self.HAS_EVM:
            st = self.get_runtime_stats(epoch_id, num_epoch)
            msg += f" eq4={float(self.loss_eq4):.3e} Re_eff={st.get('Re_eff', float('nan')):.1f} alpha_evm={self.alpha_evm}"
        msg += f" bc={float(self.loss_b):.3e} it/s={it_s:.2f} throughput={it_s * n_pts:.1f} pts/s"
        print(msg)
        if self.tb_writer is not None:
            gs = self.global_step
            self.tb_writer.add_scalar("loss/total", float(loss), gs)
            self.tb_writer.add_scalar("loss/eq", float(self.loss_e), gs)
            self.tb_writer.add_scalar("loss/bc", float(self.loss_b), gs)
            self.tb_writer.add_scalar("lr", lr, gs)

    # ---- evaluation / checkpoint ---------------------------------------
    def _errors(self, x, y, u, v, p):
        """Relative L2 errors (percent) of u, v and the NaN-masked p against reference fields (ev :684-688), reduced on the
        device by ``nsf_error_norms``; returns the errors and the predictions (device tensors, shaped [N,1])."""
        x_t, y_t, u_t, v_t, p_t = [torch.as_tensor(np.asarray(a, dtype=np.float64).reshape(-1)) if not isinstance(a, torch.Tensor)
                                   else a.detach().reshape(-1) for a in (x, y, u, v, p)]
        xd, yd = _dev_f32(x_t, self.device), _dev_f32(y_t, self.device)
        n = xd.numel()
        uvp = self._forward_net(0, xd, yd)                                  # [n, 3]
        e_p = self._forward_net(1, xd, yd) if self.HAS_EVM else None
        ud, vd, pd = [_dev_f32(a, self.device) for a in (u_t, v_t, p_t)]
        sums = torch.empty(8, dtype=torch.float64, device=self.device)
        _capi.check(self._lib, self._lib.nsf_error_norms(uvp.data_ptr(), ud.data_ptr(), vd.data_ptr(), pd.data_ptr(), n,
                                                         sums.data_ptr(), self._stream()))
        s = sums.cpu().numpy()
        eu = 100.0 * np.sqrt(s[0] / s[1]); ev = 100.0 * np.sqrt(s[2] / s[3])
        ep = 100.0 * np.sqrt(s[4] / s[5]) if s[6] > 0 else float("nan")
        self.last_error_sums = s
        return (float(eu), float(ev), float(ep)), (uvp[:, 0:1], uvp[:, 1:2], uvp[:, 2:3], e_p)

    def evaluate(self, x, y, u, v, p):
        (eu, ev, ep), _ = self._errors(x, y, u, v, p)
        if self.rank == 0:
            print("------------------------")
            print("Error u: %.2f %%" % eu)
            print("Error v: %.2f %%" % ev)
            print("Error p: %.2f %%" % ep)
        return eu, ev, ep

    def test(self, x, y, u, v, p, loop=None, save_dir=None):
        import scipy.io
        (eu, ev, ep), outs = self._errors(x, y, u, v, p)
        u_p, v_p, p_p, e_p = [None if o is None else o.detach().cpu().numpy().reshape(-1, 1) for o in outs]
        if self.rank == 0:
            print("------------------------")
            print("Error u: %.3f %%" % eu)
            print("Error v: %.3f %%" % ev)
            print("Error p: %.3f %%" % ep)
            print("------------------------")
            side = int(round(np.sqrt(u_p.size)))  # 257 for the 256 files, 385 for cavity_Re4000_384 (the reference hard-codes 257)
            shp = (side, side) if side * side == u_p.size else (-1, 1)
            d = {"U_pred": u_p.reshape(shp), "V_pred": v_p.reshape(shp), "P_pred": p_p.reshape(shp), "error_u": eu,
                 "error_v": ev, "error_p": ep, "lam_bcs": self.alpha_b, "lam_equ": self.alpha_e}
            if e_p is not None:
                d["E_pred"] = e_p.reshape(shp)
            save_dir = save_dir or f"./results/Re{self.Re}/test_result"
            os.makedirs(save_dir, exist_ok=True)
            scipy.io.savemat(os.path.join(save_dir, f"cavity_result_loop_{loop}.mat"), d)
        return eu, ev, ep

    def _save_dir(self, directory, N_HLayer, N_neu, N_f):
        raise NotImplementedError

    def save(self, filename, directory=None, N_HLayer=None, N_neu=None, N_f=None):
        out = self._save_dir(directory or os.getcwd(), N_HLayer, N_neu, N_f)
        os.makedirs(out, exist_ok=True)
        # clones: the parameters are views into one flat buffer; save plain per-tensor storages
        torch.save({k: v.detach().clone() for k, v in self.net.state_dict().items()}, out + filename)
        if self.net_1 is not None:
            torch.save({k: v.detach().clone() for k, v in self.net_1.state_dict().items()}, out + filename + "_evm")
        return out

    # ---- full training state ------------------------------------------------
    def save_checkpoint(self, path):
        """Extension: everything a bit-exact resume needs.  ``save`` (above) writes the reference's weights-only ``.pth``
        files (ev :742-759); the reference forgets the optimizer moments, the step counters and the lagged viscosity
        ``vis_t_minus``, so a run resumed from them restarts Adam and takes its first step with ``vis_t = vis_t0``."""
        ck = {"format": "nsfnet_b200.checkpoint.v1",
              "net": {k: v.detach().clone() for k, v in self.net.state_dict().items()},
              "net_1": {k: v.detach().clone() for k, v in self.net_1.state_dict().items()} if self.net_1 is not None else None,
              "evm_trainable": [bool(p.requires_grad) for p in self.net_1.parameters()] if self.net_1 is not None else None,
              "opt": self.opt.state_dict() if self.opt is not None else None,
              "fused_adam": {k: v.detach().clone() for k, v in self._adam.items()} if self._adam is not None else None,
              "vis_t_minus": self.vis_t_minus.detach().clone() if self.vis_t_minus is not None else None,
              "global_step": int(self.global_step), "alpha_evm": float(self.alpha_evm), "current_stage": self.current_stage,
              "Re": float(self.Re), "n_f_local": int(self.x_f.numel()) if self.x_f is not None else 0}
        d = os.path.dirname(os.path.abspath(path))
        os.makedirs(d, exist_ok=True)
        torch.save(ck, path)
        return path

    def load_checkpoint(self, path):
        """Restore ``save_checkpoint`` state into this solver (same network shapes; the point sets are set by the caller as
        usual, BEFORE this call, so that the saved lag state replaces the one ``set_eq_training_data`` initialises).

        The restored Adam moments survive only when training continues WITHOUT the stage-start reset: the reference's
        ``solve_Adam`` creates a fresh optimizer at epoch 0 of every call (ev :452, ``freeze_evm_net(0)``), so pass
        ``start_epoch=k`` (the epoch the checkpoint was written at) to ``solve_Adam`` / ``train`` to continue the stage
        bit-exactly; a plain ``train()`` after ``load_checkpoint`` restarts Adam like the reference does."""
        ck = torch.load(path, map_location=self.device, weights_only=True)     # tensors, primitives and an optimizer state_dict only
        if ck.get("format") != "nsfnet_b200.checkpoint.v1":
            raise ValueError(f"{path} is not a nsfnet_b200 checkpoint")
        self.net.load_state_dict(ck["net"])
        if self.net_1 is not None and ck["net_1"] is not None:
            self.net_1.load_state_dict(ck["net_1"])
            for p, r in zip(self.net_1.parameters(), ck["evm_trainable"]):
                p.requires_grad = r
        if ck["opt"] is not None:
            params = self._trainable()
            self.opt = torch.optim.Adam(params, lr=ck["opt"]["param_groups"][0]["lr"], weight_decay=0.0)
            self.opt.load_state_dict(ck["opt"])
        if ck["fused_adam"] is not None:
            if self._adam is None:
                self._adam_reset()
            for k, v in ck["fused_adam"].items():
                self._adam[k].copy_(v)
        if ck["vis_t_minus"] is not None and self.x_f is not None and ck["vis_t_minus"].numel() == self.x_f.numel():
            self.vis_t_minus = ck["vis_t_minus"].to(self.device).clone()
        self.global_step = ck["global_step"]
        self.alpha_evm = ck["alpha_evm"]
        self.current_stage = ck["current_stage"]
        self._graphs = {}
        return ck
