"""NSFnet solver -- drop-in for NSFnet/pinn_solver.py:26-389 (plain Navier-Stokes PINN).

One network (u,v,p), constant viscosity 1/Re, eq1..eq3, boundary MSE; ``solve_Adam`` keeps the
reference's order ``backward -> step -> zero_grad`` and its single Adam object (NSFnet :240-278).
"""
from __future__ import annotations

import numpy as np
import torch

from .solver_core import SolverBase


class PysicsInformedNeuralNetwork(SolverBase):
    HAS_EVM = False

    def __init__(self, opt=None, Re=1000, layers=4, hidden_size=120, N_f=40000, stage=0, learning_rate=0.001,
                 weight_decay=0.9, outlet_weight=1, bc_weight=1, eq_weight=1, ic_weight=1, num_ins=2, num_outs=3,
                 supervised_data_weight=1, training_type="unsupervised", net_params=None, checkpoint_freq=10000,
                 checkpoint_path="./checkpoint/"):
        self.checkpoint_freq, self.checkpoint_path = checkpoint_freq, checkpoint_path
        self.alpha_i, self.alpha_o = ic_weight, outlet_weight
        self.stage = stage
        self.training_type = training_type
        self.vis_t0 = 5.0 / Re  # unused by the reference too (NSFnet :51)
        self.loss_history = []
        self._init_common(Re, layers, hidden_size, N_f, bc_weight, eq_weight, num_ins, num_outs, learning_rate, net_params,
                          opt, supervised_data_weight=0.0)

    def neural_net_u(self, x, y):
        """u, v, p as [N,1] columns (NSFnet :124-130)."""
        uvp = self._forward_net(0, x, y)
        return uvp[:, 0:1], uvp[:, 1:2], uvp[:, 2:3]

    def neural_net_equations(self, x, y):
        eq1, eq2, eq3, _ = self._equations(x, y)
        return eq1, eq2, eq3

    def divergence(self, x_star, y_star):
        # the reference's version is broken (NSFnet :382-389 unpacks 4 of 3); this is what it meant
        self.eq1_pred, self.eq2_pred, self.eq3_pred = self.neural_net_equations(x_star, y_star)
        return self.eq3_pred

    # NSFnet :240-278
    def solve_Adam(self, loss_func, num_epoch=1000, batchsize=None, scheduler=None, start_epoch=0):
        import time
        self._epoch_start_wall = time.time()
        epoch_id = start_epoch
        fused = self._fused and loss_func == self.fwd_computing_loss_2d
        if fused:                      # ONE Adam for the whole run (NSFnet :76-79); only the learning rate follows the stage
            if self._adam is None:
                self._adam_reset()
            self._adam_set_lr(self.opt.param_groups[0]["lr"])
        while epoch_id < num_epoch:
            if fused:
                self._sync_fused_lr()
                loss = self._fused_step_replayable()
                losses = [self.loss_e, self.loss_b]
            else:
                loss, losses = loss_func()
                loss.backward()
                self.opt.step()
                self.opt.zero_grad()
            if scheduler:
                scheduler.step()
            if self.rank == 0 and epoch_id % getattr(self, "log_interval", 1000) == 0:
                self.print_log(loss, losses, epoch_id, num_epoch)
            if self.rank == 0 and getattr(self, "checkpoints", True) and epoch_id % 10000 == 0:
                self.save("model_cavity_loop_%d.pth" % epoch_id, N_HLayer=self.layers, N_neu=self.hidden_size, N_f=self.N_f)
            epoch_id += 1
        if fused and self.is_distributed:
            self.release_graphs()      # no captured NCCL kernels outlive the loop (a later destroy_process_group would wait for them)

    def _save_dir(self, directory, N_HLayer, N_neu, N_f):
        nn = f"{N_HLayer}x{N_neu}_Nf{np.int32(N_f / 1000)}k"
        return f"{directory}/results/Re{self.Re}/{nn}_lamB{self.alpha_b}_{self.stage}/"
