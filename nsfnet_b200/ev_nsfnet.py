"""ev-NSFnet solver -- drop-in for ev-NSFnet/pinn_solver.py:27-765 (entropy-viscosity PINN).

Two networks (main u,v,p + EVM e), lagged per-point entropy viscosity ``vis_t = min(20/Re,
alpha_evm*|e_prev|)`` kept on the device (the reference round-trips it through numpy every step,
ev :327-334), residual eq4, SDF-weighted MSE, optional supervised term, the 1-in-10 000
freeze/unfreeze schedule with its fresh ``torch.optim.Adam`` objects (ev :456-511).
"""
from __future__ import annotations

import time

import torch

from .solver_core import SolverBase


class PysicsInformedNeuralNetwork(SolverBase):
    HAS_EVM = True

    def __init__(self, opt=None, Re=1000, layers=6, layers_1=6, hidden_size=80, hidden_size_1=20, N_f=100000,
                 alpha_evm=0.03, learning_rate=0.001, weight_decay=0.9, outlet_weight=1, bc_weight=10, eq_weight=1,
                 ic_weight=0.1, num_ins=2, num_outs=3, num_outs_1=1, supervised_data_weight=1, net_params=None,
                 net_params_1=None, checkpoint_freq=2000, checkpoint_path="./checkpoint/"):
        self.checkpoint_freq, self.checkpoint_path = checkpoint_freq, checkpoint_path
        self.alpha_i, self.alpha_o = ic_weight, outlet_weight
        self.vis_t0 = 20.0 / Re
        self._init_common(Re, layers, hidden_size, N_f, bc_weight, eq_weight, num_ins, num_outs, learning_rate, net_params,
                          opt, layers_1=layers_1, hidden_size_1=hidden_size_1, num_outs_1=num_outs_1,
                          net_params_1=net_params_1, alpha_evm=alpha_evm, supervised_data_weight=supervised_data_weight)
        if self.rank == 0:
            print("Distributed training setup:")
            print(f"  World size: {self.world_size}\n  Rank: {self.rank}\n  Local rank: {self.local_rank}\n  Device: {self.device}")

    def neural_net_u(self, x, y):
        """u [N], v [N], p [N,1], e [N,1] -- the shapes of ev :280-288."""
        uvp = self._forward_net(0, x, y)
        e = self._forward_net(1, x, y)
        return uvp[:, 0], uvp[:, 1], uvp[:, 2:3], e[:, 0:1]

    def neural_net_equations(self, x, y):
        eq1, eq2, eq3, eq4 = self._equations(x, y)
        return eq1, eq2, eq3, eq4

    def divergence(self, x_star, y_star):
        self.eq1_pred, self.eq2_pred, self.eq3_pred, self.eq4_pred = self.neural_net_equations(x_star, y_star)
        return self.eq3_pred

    # ev :440-487
    def solve_Adam(self, loss_func, num_epoch=1000, batchsize=None, scheduler=None, start_epoch=0):
        """``start_epoch`` (extension): continue a stage from a ``load_checkpoint`` state -- the epoch-0 optimizer reset is
        skipped and the loop runs epochs ``start_epoch .. num_epoch-1``."""
        if not hasattr(self, "cumulative_start_time"):
            self.cumulative_start_time = time.time()
        self._epoch_start_wall = time.time()
        if not hasattr(self, "log_interval"):
            self.log_interval = 100
        if not start_epoch:
            self.freeze_evm_net(0)
        fused = self._fused and loss_func == self.fwd_computing_loss_2d
        if fused and start_epoch and self._adam is None:
            self._adam_reset()
        for epoch_id in range(start_epoch, num_epoch):
            self.global_step += 1
            if epoch_id != 0 and epoch_id % 10000 == 0:
                self.defreeze_evm_net(epoch_id)
            if (epoch_id - 1) % 10000 == 0:
                self.freeze_evm_net(epoch_id)
            if fused:
                self._sync_fused_lr()
                loss = self._fused_step_replayable()
                losses = [self.loss_e, self.loss_b]
            else:
                loss, losses = loss_func()
                self.opt.zero_grad()
                loss.backward()
                self.opt.step()
            if scheduler:
                scheduler.step()
            interval = self.log_interval if self.log_interval > 0 else 100
            if self.rank == 0 and (epoch_id == 0 or (epoch_id + 1) % interval == 0 or epoch_id == num_epoch - 1):
                self.print_log(loss, losses, epoch_id, num_epoch)
            if self.rank == 0 and getattr(self, "checkpoints", True) and (epoch_id == 0 or epoch_id % 10000 == 0):
                self.save("model_cavity_loop%d.pth" % epoch_id, N_HLayer=self.layers, N_neu=self.hidden_size, N_f=self.N_f)
        if fused and self.is_distributed:
            self.release_graphs()      # no captured NCCL kernels outlive the loop (a later destroy_process_group would wait for them)

    def freeze_evm_net(self, epoch_id):
        for p in self.net_1.parameters():
            p.requires_grad = False
        self.opt = torch.optim.Adam([p for p in self.net.parameters() if p.requires_grad],
                                    lr=self.opt.param_groups[0]["lr"], weight_decay=0.0)
        if self._fused:
            self._adam_reset(); self._adam_set_lr(self.opt.param_groups[0]["lr"])

    def defreeze_evm_net(self, epoch_id):
        for p in self.net_1.parameters():
            p.requires_grad = True
        self.opt = torch.optim.Adam(list(self.net.parameters()) + list(self.net_1.parameters()),
                                    lr=self.opt.param_groups[0]["lr"], weight_decay=0.0)
        if self._fused:
            self._adam_reset(); self._adam_set_lr(self.opt.param_groups[0]["lr"])

    def _save_dir(self, directory, N_HLayer, N_neu, N_f):
        import numpy as np
        nn = f"{N_HLayer}x{N_neu}_Nf{np.int32(N_f / 1000)}k"
        lam = f"lamB{self.alpha_b}_alpha{self.alpha_evm}{self.current_stage}"
        return f"{directory}/results/Re{self.Re}/{nn}_{lam}/"
