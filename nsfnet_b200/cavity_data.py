"""Point sets of the lid-driven cavity -- the host-side data layer of the reference
(ev-NSFnet/cavity_data.py:25-160, tools.py:30-83), vectorised.

Same public surface (``DataLoader.loading_boundary_data / loading_training_data /
loading_evaluate_data / get_sdf_weights / get_coord_scale``), same point sets: the deterministic
4 x 513 boundary set with the regularised lid ``u = 1 - cosh(10 (x - .5)) / cosh(5)``, Latin-hypercube
collocation points (one stratified draw per 1/N cell and a shuffle per dimension, tools.py:30-57, done
with array ops instead of a per-sample Python loop), optional sort by distance to the wall
(tools.py:68-83) and the SDF weights ``(w_min + (1 - w_min) exp(-decay d)) / mean`` with d = distance
to the nearest discrete boundary point (cavity_data.py:118-130).
"""
from __future__ import annotations

import numpy as np


def cavity_boundary(n_side: int = 513, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0):
    """lower, upper (lid), left, right; returns x_b, y_b, u_b, v_b as [4*n_side, 1] float64."""
    sx = np.linspace(x_min, x_max, n_side)
    sy = np.linspace(y_min, y_max, n_side)
    lid = 1.0 - np.cosh(10.0 * (sx - 0.5)) / np.cosh(5.0)
    x_b = np.concatenate([sx, sx, x_min * np.ones(n_side), x_max * np.ones(n_side)]).reshape(-1, 1)
    y_b = np.concatenate([y_min * np.ones(n_side), y_max * np.ones(n_side), sy, sy]).reshape(-1, 1)
    u_b = np.concatenate([np.zeros(n_side), lid, np.zeros(n_side), np.zeros(n_side)]).reshape(-1, 1)
    v_b = np.zeros_like(u_b)
    return x_b, y_b, u_b, v_b


def lhs_sample(D, bounds, N, rng=None):
    """Latin hypercube sample, [N, D] (tools.py:30-57)."""
    rng = np.random.default_rng() if rng is None else rng
    b = np.asarray(bounds, dtype=np.float64)
    if np.any(b[:, 0] > b[:, 1]):
        raise ValueError("Wrong value bound")
    out = np.empty((N, D))
    for i in range(D):
        cell = (np.arange(N) + rng.random(N)) / N
        rng.shuffle(cell)
        out[:, i] = cell * (b[i, 1] - b[i, 0]) + b[i, 0]
    return out


def wall_distance(pts, pts_bc):
    from scipy.spatial import cKDTree
    d, _ = cKDTree(pts_bc).query(pts)
    return d


def sort_pts(pts, pts_bc):
    """Sort by distance to the nearest boundary point, ascending (tools.py:68-83)."""
    d = wall_distance(pts, pts_bc)
    idx = np.argsort(d, kind="stable")
    return pts[idx], d[idx]


def sdf_weights(pts, pts_bc, min_weight=0.2, decay=5.0):
    d = wall_distance(pts, pts_bc)
    min_w = max(1e-6, min(float(min_weight), 1.0))
    w = min_w + (1.0 - min_w) * np.exp(-max(0.0, float(decay)) * d)
    m = w.mean()
    return (w / m if m > 0 else w).astype(np.float32)


class DataLoader:
    def __init__(self, path=None, N_f=20000, N_b=1000, sort_training_points=True, sdf_weighting=None, coord_transform=False, seed=None):
        self.N_b, self.N_f = N_b, N_f
        self.x_min, self.x_max, self.y_min, self.y_max = 0.0, 1.0, 0.0, 1.0
        self.pts_bc = None
        self.sort_training_points = sort_training_points
        self.sdf_config = sdf_weighting
        self.sdf_enabled = bool(getattr(sdf_weighting, "enabled", False)) if sdf_weighting is not None else False
        self.sdf_weights = None
        self.coord_transform = coord_transform
        self.coord_scale = 2.0 if coord_transform else 1.0
        self.rng = np.random.default_rng(seed)

    def loading_boundary_data(self):
        x_b, y_b, u_b, v_b = cavity_boundary(513, self.x_min, self.x_max, self.y_min, self.y_max)
        pts = np.hstack((x_b, y_b))
        if self.coord_transform:
            pts = pts * 2.0 - 1.0
            x_b, y_b = pts[:, 0:1], pts[:, 1:2]
            self.x_min, self.x_max, self.y_min, self.y_max = -1.0, 1.0, -1.0, 1.0
        self.pts_bc = pts
        return x_b, y_b, u_b, v_b

    def loading_training_data(self):
        if self.pts_bc is None:
            raise RuntimeError("need to load boundary data first!")
        xye = lhs_sample(2, [[self.x_min, self.x_max], [self.y_min, self.y_max]], self.N_f, self.rng)
        if self.coord_transform:
            # the reference maps the already [-1,1]-bounded sample once more (cavity_data.py:83-84,100-102;
            # SURVEY 8c "known defects"); kept, because its checkpoints were trained on exactly this set
            xye = xye * 2.0 - 1.0
        if self.sort_training_points:
            xye, _ = sort_pts(xye, self.pts_bc)
        self.sdf_weights = None
        if self.sdf_enabled:
            self.sdf_weights = sdf_weights(xye, self.pts_bc, getattr(self.sdf_config, "min_weight", 0.2),
                                           getattr(self.sdf_config, "decay", 5.0))
        return xye[:, 0:1], xye[:, 1:2]

    def get_sdf_weights(self):
        return self.sdf_weights

    def get_coord_scale(self):
        return self.coord_scale

    def loading_evaluate_data(self, filename):
        import scipy.io
        data = scipy.io.loadmat(filename)
        x, y, u, v, p = data["X_ref"], data["Y_ref"], data["U_ref"], data["V_ref"], data["P_ref"]
        if self.coord_transform:
            x, y = x * 2.0 - 1.0, y * 2.0 - 1.0
        return tuple(a.reshape(-1, 1) for a in (x, y, u, v, p))


# ---- on-device point layer -------------------------------------------------------------------
# The same point sets produced by libnsf_b200 on the GPU (include/nsf_b200.h: nsf_lhs_points / nsf_wall_distance /
# nsf_sdf_weights): 1e6 points take milliseconds instead of the hours of the reference's per-sample Python loops, and
# under data parallelism every rank generates only its own rows of the one global Latin-hypercube design.

def _lib_stream(device):
    import torch
    from . import _capi
    return _capi.load(), torch.cuda.current_stream(device).cuda_stream


def lhs_sample_device(n_total, device, seed=0, bounds=((0.0, 1.0), (0.0, 1.0)), first=0, count=None):
    """Rows [first, first+count) of a Latin-hypercube design of ``n_total`` points -> (x, y) fp32 device tensors."""
    import torch
    from . import _capi
    count = n_total - first if count is None else count
    lib, st = _lib_stream(device)
    x = torch.empty(count, dtype=torch.float32, device=device)
    y = torch.empty(count, dtype=torch.float32, device=device)
    (x0, x1), (y0, y1) = bounds
    _capi.check(lib, lib.nsf_lhs_points(int(n_total), int(first), int(count), int(seed) & 0xFFFFFFFF, x0, x1, y0, y1,
                                        x.data_ptr(), y.data_ptr(), st))
    return x, y


def wall_distance_device(x, y, xb, yb):
    """Distance to the nearest discrete boundary point (cKDTree.query of cavity_data.py:118-121) for device tensors."""
    import torch
    from . import _capi
    lib, st = _lib_stream(x.device)
    d = torch.empty_like(x)
    _capi.check(lib, lib.nsf_wall_distance(x.data_ptr(), y.data_ptr(), x.numel(), xb.data_ptr(), yb.data_ptr(), xb.numel(), d.data_ptr(), st))
    return d


def sdf_weights_device(x, y, xb, yb, min_weight=0.2, decay=5.0, distributed=False):
    """SDF weights normalised to mean 1 over the GLOBAL point set (cavity_data.py:118-130; one all-reduce of
    (sum, count) when ``distributed``)."""
    import torch
    import torch.distributed as dist
    from . import _capi
    lib, st = _lib_stream(x.device)
    w = torch.empty_like(x)
    acc = torch.zeros(2, dtype=torch.float64, device=x.device)
    acc[1] = float(x.numel())
    _capi.check(lib, lib.nsf_sdf_weights(x.data_ptr(), y.data_ptr(), x.numel(), xb.data_ptr(), yb.data_ptr(), xb.numel(),
                                         float(min_weight), float(decay), w.data_ptr(), acc.data_ptr(), st))
    if distributed:
        dist.all_reduce(acc)
    mean = acc[0] / acc[1]
    return w * torch.where(mean > 0, 1.0 / mean, torch.ones_like(mean)).to(torch.float32)


class DeviceDataLoader(DataLoader):
    """``DataLoader`` whose collocation set lives on the GPU from the start.  ``loading_training_data`` returns THIS rank's
    shard (rows ``shard_bounds`` of the global design, or the whole sorted set when ``sort_training_points``) as device
    tensors for ``solver.set_eq_training_shard``."""

    def __init__(self, device, rank=0, world_size=1, seed=0, **kw):
        super().__init__(seed=seed, **kw)
        self.device, self.rank, self.world_size, self.seed = device, rank, world_size, 0 if seed is None else int(seed)

    def loading_training_data(self):
        import torch
        if self.pts_bc is None:
            raise RuntimeError("need to load boundary data first!")
        per = self.N_f // self.world_size
        first = self.rank * per
        count = per if self.rank < self.world_size - 1 else self.N_f - first
        bounds = ((self.x_min, self.x_max), (self.y_min, self.y_max))
        xb = torch.as_tensor(self.pts_bc[:, 0], dtype=torch.float32, device=self.device).contiguous()
        yb = torch.as_tensor(self.pts_bc[:, 1], dtype=torch.float32, device=self.device).contiguous()
        if self.sort_training_points:
            # the reference sorts the GLOBAL set by wall distance and shards it in contiguous blocks (tools.py:68-83, ev :165-177)
            x, y = lhs_sample_device(self.N_f, self.device, self.seed, bounds)
            if self.coord_transform:
                x, y = x * 2.0 - 1.0, y * 2.0 - 1.0
            idx = torch.argsort(wall_distance_device(x, y, xb, yb), stable=True)[first:first + count]
            x, y = x[idx].contiguous(), y[idx].contiguous()
        else:
            x, y = lhs_sample_device(self.N_f, self.device, self.seed, bounds, first, count)
            if self.coord_transform:
                x, y = x * 2.0 - 1.0, y * 2.0 - 1.0
        self.sdf_weights = None
        if self.sdf_enabled:
            self.sdf_weights = sdf_weights_device(x, y, xb, yb, getattr(self.sdf_config, "min_weight", 0.2),
                                                  getattr(self.sdf_config, "decay", 5.0), distributed=self.world_size > 1)
        return x, y
