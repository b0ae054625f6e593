"""The trainer entry point (ev-NSFnet/train.py:15-224 flow: YAML config -> solver -> boundary / collocation / SDF data ->
per stage set_alpha_evm + train + evaluate) end to end on the GPU, with the reference's loop body and with the fused
iteration, host data layer and on-device data layer."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CFG = """
experiment_name: tiny
description: two short stages
physics: {Re: 2000, alpha_evm: 0.05, bc_weight: 10, eq_weight: 1}
network: {layers: 6, layers_1: 4, hidden_size: 80, hidden_size_1: 40}
training:
  N_f: 6000
  sort_training_points: true
  log_interval: 20
  enable_tensorboard: false
  sdf_weighting: {enabled: true, min_weight: 0.2, decay: 5.0}
  coordinate_transform: false
  training_stages:
    - {alpha: 0.05, epochs: 25, lr: 1.0e-3, name: "Stage 1"}
    - {alpha: 0.03, epochs: 25, lr: 2.0e-4, name: "Stage 2"}
supervision: {enabled: false, num_samples: 0, loss_weight: 1.0}
"""


def _run(tmp_path, monkeypatch, extra):
    import torch
    from nsfnet_b200 import train
    monkeypatch.chdir(tmp_path)
    cfg = tmp_path / "tiny.yaml"
    cfg.write_text(CFG)
    # a DNS-like file so that the per-stage evaluate() runs (keys of the reference's .mat files)
    import scipy.io
    s = np.linspace(0, 1, 33)
    X, Y = np.meshgrid(s, s)
    os.makedirs(tmp_path / "data", exist_ok=True)
    scipy.io.savemat(str(tmp_path / "data" / "cavity_Re2000_256_Uniform.mat"),
                     {"X_ref": X, "Y_ref": Y, "U_ref": np.sin(X), "V_ref": np.cos(Y), "P_ref": X * Y})
    torch.manual_seed(7)
    assert train.main(["--config", str(cfg), "--seed", "3"] + extra) == 0
    P = train.main.last_solver
    return P.net.flat_params().detach().clone(), float(P.loss.detach()), P


def test_trainer_flow_fused_equals_reference_loop(tmp_path, monkeypatch):
    w_ref, l_ref, P = _run(tmp_path, monkeypatch, ["--no-fused"])
    assert P.current_stage == "Stage 2" and P.alpha_evm == 0.03 and P.opt.param_groups[0]["lr"] == 2.0e-4
    assert P.eq_weights is not None and abs(float(P.eq_weights.mean()) - 1.0) < 1e-5        # SDF weights in force
    out = [r for r, _, f in os.walk(tmp_path / "results") for x in f if x.endswith(".pth")]
    assert out, "the epoch-0 checkpoint of every stage (ev :484-487) was not written"
    w_fus, l_fus, Pf = _run(tmp_path, monkeypatch, [])
    assert Pf._fused and Pf.graph_replays >= 40          # 2 x 25 epochs, of which the first two of a stage run eagerly
    assert (w_fus - w_ref).abs().max().item() < 5e-6 and abs(l_fus - l_ref) < 1e-4 * abs(l_ref)


def test_trainer_flow_with_the_on_device_point_layer(tmp_path, monkeypatch):
    w, loss, P = _run(tmp_path, monkeypatch, ["--device-data"])
    assert np.isfinite(loss) and P.x_f.is_cuda and P.x_f.numel() == 6000 and P._n_f_global == 6000
    assert P.eq_weights is not None and abs(float(P.eq_weights.mean()) - 1.0) < 1e-5
