"""N > 1 path on CPU (gloo, world_size 2): the data-parallel design of SURVEY 8e -- contiguous shards,
kernels normalised by the GLOBAL point counts, ONE all-reduce(SUM) of [grad_main | grad_evm | loss sums] --
must give every rank the gradient and loss of the single-rank evaluation on the union of the shards.
The per-rank evaluation runs on the host emulation of the kernels (tests/emu); the collective is real."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nsfnet_b200 import _capi
from nsfnet_b200.solver_core import shard_bounds
from oracle import jet_numpy as J

MD, ED = (2, 3, 3, 24), (2, 1, 2, 12)
N_F, N_B = 203, 37           # deliberately not divisible by the world size


def _problem():
    rng = np.random.default_rng(7)
    pm = J.init_params(J.NetDesc(*MD), 3) * 1.5
    pe = J.init_params(J.NetDesc(*ED), 4)
    x, y = rng.random(N_F).astype(np.float32), rng.random(N_F).astype(np.float32)
    xb, yb, ub, vb = [rng.random(N_B).astype(np.float32) for _ in range(4)]
    vtm = (rng.random(N_F) * 0.02).astype(np.float32)
    w = (0.5 + rng.random(N_F)).astype(np.float32)
    return pm, pe, x, y, xb, yb, ub, vb, vtm, w


def _eval(lib, rank, world):
    from tests.emu import emu
    pm, pe, x, y, xb, yb, ub, vb, vtm, w = _problem()
    s, e = shard_bounds(N_F, rank, world)
    sb, eb = shard_bounds(N_B, rank, world)
    phys = _capi.physics(2000., alpha_evm=0.05, has_evm=True, evm_trainable=True, n_f_norm=N_F)
    c = 10. / N_B
    o = emu.run_step(lib, MD, pm, phys, x[s:e], y[s:e], blocks=[(xb[sb:eb], yb[sb:eb], ub[sb:eb], vb[sb:eb], None, c, c, 0.)],
                     evm_desc=ED, params_evm=pe, w=w[s:e], vtm_in=vtm[s:e])
    return np.concatenate([o["grad_main"], o["grad_evm"], o["loss_parts"]])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.emu import emu
        buf = torch.from_numpy(_eval(emu.load(), rank, world))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        if rank == 0:
            np.save(out, buf.numpy())
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one_rank_on_union(tmp_path):
    from tests.emu import emu
    lib = emu.load()
    single = _eval(lib, 0, 1)
    out = str(tmp_path / "w2.npy")
    mp.spawn(_worker, args=(2, 29571, out), nprocs=2, join=True)
    multi = np.load(out)
    n_main = J.NetDesc(*MD).n_params
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(multi[:n_main], single[:n_main]) < 2e-6          # grad_main
    assert rel(multi[n_main:-16], single[n_main:-16]) < 2e-6    # grad_evm
    assert np.allclose(multi[-16:], single[-16:], rtol=2e-6)    # loss partial sums, point counts
    assert multi[-16 + 5] == N_F


def test_shard_bounds_match_reference_split():
    assert shard_bounds(2052, 0, 8) == (0, 256) and shard_bounds(2052, 7, 8) == (1792, 2052)
    assert shard_bounds(10, 0, 1) == (0, 10)
    covered = []
    for r in range(3):
        s, e = shard_bounds(100, r, 3)
        covered += list(range(s, e))
    assert covered == list(range(100))
