"""torchrun worker of tests/test_gpu_multi.py: the solver under NCCL data parallelism (one process per GPU).

Checks on every rank, then rank 0 prints "DDP_OK":
  * W-rank loss / gradient == the single-process evaluation on the union of the shards (<= 1e-6);
  * the reference-style loop (loss.backward() + torch.optim.Adam), the fused eager loop and the fused loop replayed as a
    CUDA graph WITH the NCCL all-reduce captured leave identical weights on all ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(150, exit=True)      # a hung collective must not hang the test run
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from nsfnet_b200.cavity_data import DeviceDataLoader, cavity_boundary
    from nsfnet_b200.ev_nsfnet import PysicsInformedNeuralNetwork
    from nsfnet_b200 import _capi
    from nsfnet_b200.solver_core import shard_bounds

    n_f = 20_001                      # not divisible by the world size
    torch.manual_seed(11)             # same weights everywhere (and rank 0's are broadcast anyway, ev :105-106)
    P = PysicsInformedNeuralNetwork(Re=2000, layers=6, hidden_size=80, layers_1=4, hidden_size_1=40, N_f=n_f, alpha_evm=0.05,
                                    bc_weight=10, eq_weight=1, supervised_data_weight=0.0)
    assert P.is_distributed and P.world_size == world
    P.log_interval = 10 ** 9; P.checkpoints = False; P.verbose = False
    xb, yb, ub, vb = cavity_boundary(513)
    P.set_boundary_data((xb, yb, ub, vb))                      # sharded like ev :144-147
    rng = np.random.default_rng(5)
    x, y = rng.random((n_f, 1)), rng.random((n_f, 1))
    w = (0.5 + rng.random(n_f)).astype(np.float32)
    P.set_eq_training_data((x, y), weights=w)                  # contiguous blocks, last rank takes the remainder (ev :165-177)
    s, e = shard_bounds(n_f, rank, world)
    assert P.x_f.numel() == e - s
    st0 = (P.net.flat_params().clone(), P.net_1.flat_params().clone(), P.vis_t_minus.clone())

    # ---- one evaluation: every rank holds the loss / gradient of the union ----------------------------------------
    P.freeze_evm_net(0)
    loss, _ = P.fwd_computing_loss_2d(); P.opt.zero_grad(); loss.backward()
    g = torch.cat([p.grad.reshape(-1) for p in P.net.parameters()])
    ctx = _capi.Context(P._lib, local, P.net.desc, P.net_1.desc)          # the same evaluation, un-sharded, by this process alone
    dv = lambda a: torch.as_tensor(np.asarray(a, np.float32).reshape(-1)).cuda()
    X, Y, Wt, XB, YB, UB, VB = dv(x), dv(y), dv(w), dv(xb), dv(yb), dv(ub), dv(vb)
    e_all = torch.empty((n_f, 1), device="cuda")
    ctx.forward(1, st0[1].data_ptr(), X.data_ptr(), Y.data_ptr(), n_f, e_all.data_ptr(), torch.cuda.current_stream().cuda_stream)
    vtm = (0.05 * e_all.abs()).reshape(-1).contiguous()
    gm = torch.empty(P._n_main, device="cuda"); lp = torch.empty(16, device="cuda")
    c = 10.0 / XB.numel()
    blk = _capi.NsfDataBlock(XB.data_ptr(), YB.data_ptr(), UB.data_ptr(), VB.data_ptr(), None, XB.numel(), c, c, 0.0, 0)
    ctx.step(st0[0].data_ptr(), st0[1].data_ptr(), X.data_ptr(), Y.data_ptr(), Wt.data_ptr(), vtm.data_ptr(), vtm.data_ptr(), n_f, [blk],
             _capi.physics(2000., alpha_evm=0.05, has_evm=True), gm.data_ptr(), None, lp.data_ptr(), None, None, None,
             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    loss1 = (lp[0] + lp[1] + lp[2] + 0.1 * lp[3]) / n_f + 10.0 * (lp[6] + lp[7]) / XB.numel()
    rel_g = ((g - gm).norm() / gm.norm()).item()
    rel_l = abs(float(loss) - float(loss1)) / abs(float(loss1))
    assert rel_g < 1e-6 and rel_l < 1e-6, (rel_g, rel_l)

    # ---- three loops from the same state ---------------------------------------------------------------------------
    def restore():
        with torch.no_grad():
            P.net.flat_params().copy_(st0[0]); P.net_1.flat_params().copy_(st0[1])
        P.vis_t_minus = st0[2].clone()
    results = {}
    for mode in ("torch", "fused_eager", "fused_graph"):
        restore()
        os.environ["NSF_FUSED_GRAPH_DDP"] = "1" if mode == "fused_graph" else "0"
        P.enable_fused_step(mode != "torch")
        P.train(num_epoch=10, lr=1e-3)
        results[mode] = P.net.flat_params().clone()
        if mode == "fused_graph":
            assert P.graph_replays >= 8                                    # 10 epochs, the first two of the stage run eagerly
        both = [torch.empty_like(results[mode]) for _ in range(world)]
        dist.all_gather(both, results[mode])
        assert all(torch.equal(both[0], b) for b in both), mode            # replicas stay bit-identical
    d1 = (results["fused_eager"] - results["torch"]).abs().max().item()
    d2 = (results["fused_graph"] - results["fused_eager"]).abs().max().item()
    assert d1 < 2e-6 and d2 < 5e-7, (d1, d2)

    # ---- every rank generates only its rows of ONE global Latin-hypercube design ------------------------------------
    dl = DeviceDataLoader(P.device, rank=rank, world_size=world, N_f=n_f, sort_training_points=False, seed=3)
    dl.loading_boundary_data()
    xs, ys = dl.loading_training_data()
    P.set_eq_training_shard((xs, ys))
    assert P._n_f_global == n_f
    cells = torch.floor(xs.double() * n_f).long()
    longest = max(shard_bounds(n_f, r, world)[1] - shard_bounds(n_f, r, world)[0] for r in range(world))
    pad = torch.full((longest,), -1, dtype=torch.long, device="cuda")
    pad[:cells.numel()] = cells
    got = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(got, pad)
    cat = torch.cat(got)
    assert torch.equal(torch.sort(cat[cat >= 0]).values, torch.arange(n_f, device="cuda"))      # one point per stratum, globally
    loss, _ = P.fwd_computing_loss_2d()
    assert np.isfinite(float(loss))
    assert not P._graphs          # solve_Adam released its captured NCCL kernels (a later destroy_process_group would hang on them)
    dist.barrier()
    if rank == 0:
        print(f"DDP_OK world={world} grad_rel={rel_g:.2e} loss_rel={rel_l:.2e} torch_vs_fused={d1:.2e} eager_vs_graph={d2:.2e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
