"""Configuration tree of the ev-NSFnet trainer -- same YAML schema as the reference
(ev-NSFnet/config.py:9-142, configs/production.yaml): physics / network / training (+ stages, SDF weighting) /
supervision."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import yaml


@dataclass
class PhysicsConfig:
    Re: int = 5000
    alpha_evm: float = 0.05
    bc_weight: float = 10.0
    eq_weight: float = 1.0


@dataclass
class NetworkConfig:
    layers: int = 6
    layers_1: int = 4
    hidden_size: int = 80
    hidden_size_1: int = 40


@dataclass
class TrainingStage:
    alpha: float
    epochs: int
    lr: float
    name: str


@dataclass
class SupervisionConfig:
    enabled: bool = False
    num_samples: int = 0
    loss_weight: float = 1.0


@dataclass
class SDFWeightConfig:
    enabled: bool = False
    min_weight: float = 0.2
    decay: float = 5.0


def _default_stages():
    return [TrainingStage(0.05, 500000, 1e-3, "Stage 1"), TrainingStage(0.03, 500000, 2e-4, "Stage 2"),
            TrainingStage(0.01, 500000, 4e-5, "Stage 3"), TrainingStage(0.005, 500000, 1e-5, "Stage 4"),
            TrainingStage(0.002, 500000, 2e-6, "Stage 5"), TrainingStage(0.002, 500000, 2e-6, "Stage 6")]


@dataclass
class TrainingConfig:
    N_f: int = 120000
    log_interval: int = 1000
    enable_tensorboard: bool = True
    tb_log_dir: str = "runs"
    sort_training_points: bool = True
    sdf_weighting: SDFWeightConfig = field(default_factory=SDFWeightConfig)
    coordinate_transform: bool = False
    training_stages: List[TrainingStage] = field(default_factory=_default_stages)


@dataclass
class AppConfig:
    physics: PhysicsConfig = field(default_factory=PhysicsConfig)
    network: NetworkConfig = field(default_factory=NetworkConfig)
    training: TrainingConfig = field(default_factory=TrainingConfig)
    supervision: SupervisionConfig = field(default_factory=SupervisionConfig)
    experiment_name: str = "NSFnet_Restore"
    description: str = "Restored baseline with modern logging"


class ConfigManager:
    def __init__(self, config: AppConfig | None = None):
        self.config = config or AppConfig()

    @classmethod
    def from_file(cls, path: str) -> "ConfigManager":
        with open(path, "r", encoding="utf-8") as f:
            raw = yaml.safe_load(f) or {}
        cfg = AppConfig()
        cfg.experiment_name = raw.get("experiment_name", cfg.experiment_name)
        cfg.description = raw.get("description", cfg.description)
        for k, v in (raw.get("physics") or {}).items():
            setattr(cfg.physics, k, v)
        for k, v in (raw.get("network") or {}).items():
            setattr(cfg.network, k, v)
        tr = raw.get("training") or {}
        for k, v in tr.items():
            if k == "training_stages":
                cfg.training.training_stages = [TrainingStage(float(s["alpha"]), int(s["epochs"]), float(s["lr"]), str(s.get("name", f"Stage {i + 1}")))
                                                for i, s in enumerate(v)]
            elif k == "sdf_weighting":
                cfg.training.sdf_weighting = SDFWeightConfig(**(v or {}))
            else:
                setattr(cfg.training, k, v)
        if raw.get("supervision"):
            cfg.supervision = SupervisionConfig(**raw["supervision"])
        return cls(cfg)

    def print_config(self):
        c = self.config
        print(f"experiment: {c.experiment_name} -- {c.description}")
        print(f"physics   : {c.physics}")
        print(f"network   : {c.network}")
        t = c.training
        print(f"training  : N_f={t.N_f} log_interval={t.log_interval} sort={t.sort_training_points} sdf={t.sdf_weighting} "
              f"coordinate_transform={t.coordinate_transform}")
        for s in t.training_stages:
            print(f"  {s.name:10s} alpha={s.alpha:<7g} epochs={s.epochs:<8d} lr={s.lr:g}")
        print(f"supervision: {c.supervision}")
