"""ev-NSFnet trainer -- the reference's entry point (ev-NSFnet/train.py:15-224) on the B200 hot path.

    python -m nsfnet_b200.train --config configs/production.yaml [--dry-run] [--eval-file data.mat]
    torchrun --nproc_per_node=8 -m nsfnet_b200.train --config ...

Same flow: YAML config -> (optional) NCCL process group from the torchrun environment -> solver -> boundary /
collocation / SDF data -> per stage `set_alpha_evm(alpha); train(epochs, lr); evaluate(...)`.
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .cavity_data import DataLoader, DeviceDataLoader
from .config import ConfigManager


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="ev-NSFnet PINN trainer (B200 hot path)")
    p.add_argument("--config", type=str, default="configs/production.yaml")
    p.add_argument("--dry-run", action="store_true")
    p.add_argument("--eval-file", type=str, default=None, help="DNS .mat (X_ref,Y_ref,U_ref,V_ref,P_ref); default ./data/cavity_Re{Re}_256_Uniform.mat")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--no-fused", action="store_true", help="keep the reference's per-iteration Python (loss.backward() + torch.optim.Adam) "
                   "instead of the fused iteration (nsf_step + device-resident Adam replayed as one CUDA graph)")
    p.add_argument("--device-data", action="store_true", help="generate the collocation set (Latin hypercube, wall-distance sort, SDF "
                   "weights) on the GPU, every rank its own rows, instead of the host data layer")
    return p.parse_args(argv)


def setup_distributed() -> bool:
    """torchrun contract of the reference (train.py:22-43): RANK / LOCAL_RANK / WORLD_SIZE -> NCCL."""
    if not all(k in os.environ for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE")) or int(os.environ["WORLD_SIZE"]) <= 1:
        return False
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", device_id=torch.device(f"cuda:{local}"))
    return True


def build_pinn(cfg):
    from .ev_nsfnet import PysicsInformedNeuralNetwork
    return PysicsInformedNeuralNetwork(
        Re=cfg.physics.Re, layers=cfg.network.layers, layers_1=cfg.network.layers_1, hidden_size=cfg.network.hidden_size,
        hidden_size_1=cfg.network.hidden_size_1, N_f=cfg.training.N_f, alpha_evm=cfg.physics.alpha_evm,
        bc_weight=cfg.physics.bc_weight, eq_weight=cfg.physics.eq_weight,
        supervised_data_weight=cfg.supervision.loss_weight if cfg.supervision.enabled else 0.0)


def main(argv=None):
    args = parse_args(argv)
    cm = ConfigManager.from_file(args.config) if os.path.exists(args.config) else ConfigManager()
    cfg = cm.config
    distributed = setup_distributed()
    if not distributed:
        os.environ.update(RANK="0", LOCAL_RANK="0", WORLD_SIZE="1")
    rank = int(os.environ["RANK"])
    if rank == 0:
        cm.print_config()
    if args.dry_run:
        if rank == 0:
            print("dry-run: configuration only, no training")
        return 0
    if args.seed is not None:
        torch.manual_seed(args.seed)
    try:
        PINN = build_pinn(cfg)
        PINN.log_interval = cfg.training.log_interval
        PINN.enable_fused_step(not args.no_fused)
        if rank == 0 and cfg.training.enable_tensorboard:
            try:
                from torch.utils.tensorboard import SummaryWriter
                PINN.tb_writer = SummaryWriter(log_dir=os.path.join(cfg.training.tb_log_dir, f"{cfg.experiment_name}_{time.strftime('%Y%m%d_%H%M%S')}"))
            except Exception as e:      # tensorboard is optional
                print(f"TensorBoard disabled: {e}")
        kw = dict(N_f=cfg.training.N_f, N_b=1000, sort_training_points=cfg.training.sort_training_points,
                  sdf_weighting=cfg.training.sdf_weighting, coord_transform=cfg.training.coordinate_transform, seed=args.seed)
        loader = DeviceDataLoader(PINN.device, rank=PINN.rank, world_size=PINN.world_size, **kw) if args.device_data else DataLoader(**kw)
        PINN.set_boundary_data(X=loader.loading_boundary_data())
        train_pts = loader.loading_training_data()
        PINN.set_coordinate_transform(loader.get_coord_scale())
        if args.device_data:
            PINN.set_eq_training_shard(train_pts, weights=loader.get_sdf_weights(), n_global=cfg.training.N_f)
        else:
            PINN.set_eq_training_data(X=train_pts, weights=loader.get_sdf_weights())
        eval_file = args.eval_file or f"./data/cavity_Re{cfg.physics.Re}_256_Uniform.mat"
        eval_data = loader.loading_evaluate_data(eval_file) if os.path.exists(eval_file) else None
        sup = cfg.supervision
        if sup.enabled and sup.num_samples > 0 and eval_data is not None:
            x_s, y_s, u_s, v_s, p_s = eval_data
            n = min(int(sup.num_samples), x_s.shape[0])
            if distributed:      # every rank must use rank 0's sample (train.py:165-172)
                idx_t = torch.as_tensor(np.random.default_rng(args.seed).choice(x_s.shape[0], size=n, replace=False), device=PINN.device)
                dist.broadcast(idx_t, src=0)
                idx = idx_t.cpu().numpy()
            else:
                idx = np.random.default_rng(args.seed).choice(x_s.shape[0], size=n, replace=False)
            PINN.set_supervised_data((x_s[idx], y_s[idx], u_s[idx], v_s[idx], p_s[idx]))
            PINN.set_supervised_loss_weight(sup.loss_weight)
        else:
            PINN.clear_supervised_data()
            PINN.set_supervised_loss_weight(0.0)
        for st in cfg.training.training_stages:
            if rank == 0:
                print(f"=== {st.name}: alpha_evm={st.alpha} epochs={st.epochs} lr={st.lr}")
            PINN.current_stage = st.name
            PINN.set_alpha_evm(st.alpha)
            PINN.train(num_epoch=st.epochs, lr=st.lr)
            if rank == 0 and eval_data is not None:
                PINN.evaluate(*eval_data)
        if PINN.tb_writer is not None:
            PINN.tb_writer.close()
        main.last_solver = PINN          # for callers that drive main() in-process (tests)
    finally:
        if "PINN" in locals():
            PINN.release_graphs()
        if distributed and dist.is_initialized():
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
