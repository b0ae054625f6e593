"""FCNet -- the reference's tanh MLP module (net.py:22-54), same constructor, same module tree and
``state_dict`` keys (``layers.layer_{i}.{weight,bias}``), so reference checkpoints load unchanged.

What is different underneath: all parameters are views into ONE contiguous fp32 buffer laid out
in ``state_dict`` order (W0 [out,in] row-major, b0, W1, b1, ...), and all gradients are views into
one flat gradient buffer.  That flat layout is what libnsf_b200.so consumes (include/nsf_b200.h),
and it lets the optimizer / all-reduce work on a single tensor.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Tuple

import torch


class FCNet(torch.nn.Module):
    def __init__(self, num_ins=3, num_outs=3, num_layers=10, hidden_size=50, activation=torch.nn.Tanh):
        super().__init__()
        layers = [num_ins] + [hidden_size] * num_layers + [num_outs]
        self.depth = len(layers) - 1
        self.activation = activation
        self.num_ins, self.num_outs, self.num_layers, self.hidden_size = num_ins, num_outs, num_layers, hidden_size
        layer_list = []
        for i in range(self.depth - 1):
            layer_list.append(("layer_%d" % i, torch.nn.Linear(layers[i], layers[i + 1])))
            layer_list.append(("activation_%d" % i, self.activation()))
        layer_list.append(("layer_%d" % (self.depth - 1), torch.nn.Linear(layers[-2], layers[-1])))
        self.layers = torch.nn.Sequential(OrderedDict(layer_list))
        self._flat = None
        self._flat_grad = None

    # ---- flat parameter / gradient storage ---------------------------------------------------
    @property
    def desc(self) -> Tuple[int, int, int, int]:
        """(n_in, n_out, n_hidden_layers, hidden) -- NsfNetDesc."""
        return (self.num_ins, self.num_outs, self.num_layers, self.hidden_size)

    def _ordered_params(self) -> List[torch.nn.Parameter]:
        return list(self.parameters())

    def _is_flat(self) -> bool:
        f = self._flat
        if f is None:
            return False
        off = 0
        base = f.data_ptr()
        for p in self._ordered_params():
            if p.data_ptr() != base + 4 * off or p.device != f.device or p.dtype != torch.float32 or not p.is_contiguous():
                return False
            off += p.numel()
        return off == f.numel()

    def flatten_(self) -> "FCNet":
        """(Re)build the flat buffers and point every parameter (and its .grad) into them."""
        ps = self._ordered_params()
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        flat = torch.empty(n, dtype=torch.float32, device=dev)
        gflat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in ps:
                k = p.numel()
                flat[off:off + k].copy_(p.detach().reshape(-1).to(torch.float32))
                p.data = flat[off:off + k].view(p.shape)
                off += k
        self._flat, self._flat_grad = flat, gflat
        return self

    def flat_params(self) -> torch.Tensor:
        if not self._is_flat():
            self.flatten_()
        return self._flat

    def flat_grad(self) -> torch.Tensor:
        self.flat_params()
        return self._flat_grad

    def grad_views(self) -> List[torch.Tensor]:
        g = self.flat_grad()
        out, off = [], 0
        for p in self._ordered_params():
            k = p.numel()
            out.append(g[off:off + k].view(p.shape))
            off += k
        return out

    # ---- forward -----------------------------------------------------------------------------
    def forward(self, x):
        """Differentiable PyTorch evaluation, kept for API compatibility (net.py:52-54).  The
        solver's hot path does not come through here: it calls libnsf_b200 on the flat buffers."""
        return self.layers(x)
