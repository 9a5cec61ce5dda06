"""PyTorch owns every device buffer; this module hands raw pointers + the current stream to the
C ABI (seismicmesh_b200/_lib.py).  Plumbing only -- no arithmetic happens in torch."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DmPlan, DmSizeFn, check, lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError(
            "seismicmesh_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback."
        )


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def to_dev(a, dtype):
    """numpy / torch -> contiguous torch cuda tensor of `dtype` (no copy if already there)."""
    dev = device()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=dtype).contiguous()
    a = np.ascontiguousarray(a, dtype={torch.float64: np.float64, torch.int32: np.int32, torch.uint8: np.uint8}[dtype])
    return torch.from_numpy(a).to(dev)


def points_dev(x, dim=None):
    x = to_dev(x, torch.float64)
    if x.ndim == 1:
        x = x.reshape(1, -1)
    if dim is not None and x.shape[1] != dim:
        raise ValueError(f"expected points of shape (M,{dim}), got {tuple(x.shape)}")
    return x


class Plan:
    """DmPlan over a torch-owned workspace, sized for (N vertices, T cells)."""

    def __init__(self, N, T, dim):
        require_cuda()
        self.N, self.T, self.dim = int(N), int(T), int(dim)
        nbytes = lib.dm_plan_bytes(self.N, self.T, self.dim)
        if nbytes == 0:
            raise ValueError("bad plan dimensions")
        self.ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device())
        base = self.ws.data_ptr()
        aligned = (base + 255) & ~255
        self.c = DmPlan()
        check(lib.dm_plan_init(C.byref(self.c), self.N, self.T, self.dim, C.c_void_p(aligned), nbytes), "dm_plan_init")
        self._off = aligned - base

    def fits(self, N, T):
        return N == self.N and T <= self.T

    def _view(self, addr, nbytes, dtype):
        off = addr - self.ws.data_ptr()
        return self.ws[off : off + nbytes].view(dtype)

    # typed views into the workspace (for tests / API results)
    def keep(self):
        return self._view(self.c.keep, self.T, torch.uint8)

    def hbar(self, E):
        return self._view(self.c.hbar, E * 8, torch.float64)

    def degs(self):
        """(N,2) int32: {degree, number of lower neighbours} per vertex."""
        return self._view(self.c.degs, self.N * 8, torch.int32).view(self.N, 2)

    def cnt(self):
        """(N,) int32: kept cells incident to each vertex."""
        return self._view(self.c.cnt, self.N * 4, torch.int32)

    def scalars(self):
        return self._view(self.c.scalars, 8 * 8, torch.float64)

    def counters(self):
        return self._view(self.c.counters, 8 * 4, torch.int32)

    def num_bars(self):
        return int(self.counters()[0].item())


def size_fn_struct(kind, dim, hconst=0.0, axes=None, grid=None, cells=None):
    f = DmSizeFn()
    f.kind = kind
    f.dim = dim
    f.hconst = float(hconst)
    if kind == _lib.SIZE_GRID:
        for k in range(dim):
            f.n[k] = int(axes[k].numel())
            f.axis[k] = axes[k].data_ptr()
        f.grid = grid.data_ptr()
        f.cells = cells.data_ptr() if cells is not None else None
    return f


def prog_array(progs):
    """list of device program tensors -> (ctypes void* array, keep-alive list)."""
    arr = (C.c_void_p * max(1, len(progs)))()
    for i, p in enumerate(progs):
        arr[i] = p.data_ptr()
    return arr
