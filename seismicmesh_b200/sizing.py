"""Mesh-size functions with the reference's interface (SeismicMesh/sizing/size_function.py:1-12).

``SizeFunction(bbox, cell_size, hmin)`` keeps the reference constructor and ``.eval``; when
``cell_size`` is a gridded interpolant (our :class:`GridInterpolant`, or a SciPy
``RegularGridInterpolator`` as built by the reference at sizing/mesh_size_function.py:391-408)
it is lowered to a device descriptor and evaluated by the CUDA interpolation kernel with SciPy's
exact semantics (float32-rounded axes, linear extrapolation outside the grid).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import device as D
from ._lib import check, lib

__all__ = ["SizeFunction", "GridInterpolant", "grid_axes"]


def grid_axes(bbox, shape):
    """Axis vectors exactly as the reference builds them (mesh_size_function.py:514-523):
    float32 linspace, which SciPy up-casts to float64."""
    return [
        np.linspace(bbox[2 * k], bbox[2 * k + 1], int(n), dtype=np.float32).astype(np.float64)
        for k, n in enumerate(shape)
    ]


class GridInterpolant:
    """Bilinear / trilinear interpolant on a rectilinear grid, device resident.

    Semantics of ``scipy.interpolate.RegularGridInterpolator(axes, values, method="linear",
    bounds_error=False, fill_value=None)`` -- including its accumulation order, so results are
    bit-identical to SciPy's.
    """

    def __init__(self, axes, values):
        self.grid = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in axes)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.dim = len(self.grid)
        if self.dim not in (2, 3) or self.values.ndim != self.dim:
            raise ValueError("Dimension not supported")
        for a, n in zip(self.grid, self.values.shape):
            if a.ndim != 1 or len(a) != n or n < 2 or not np.all(np.diff(a) > 0):
                raise ValueError("grid axes must be strictly ascending and match the values' shape")
        self._dev = None

    def device_arrays(self):
        dev = D.device()
        if self._dev is None or self._dev[1].device != dev:
            self._dev = ([torch.from_numpy(a).to(dev) for a in self.grid], torch.from_numpy(self.values).to(dev))
        return self._dev

    def struct(self):
        axes, grid = self.device_arrays()
        return D.size_fn_struct(_lib.SIZE_GRID, self.dim, axes=axes, grid=grid)

    def __call__(self, x):
        as_torch = isinstance(x, torch.Tensor)
        xd = D.points_dev(x if as_torch else np.asarray(x, dtype=np.float64), self.dim)
        out = torch.empty(xd.shape[0], dtype=torch.float64, device=xd.device)
        f = self.struct()
        check(lib.dm_size_eval(C.byref(f), D.ptr(xd), xd.shape[0], D.ptr(out), D.stream_ptr()), "dm_size_eval")
        return out if as_torch else out.cpu().numpy()


def _lower_callable(cell_size):
    """GridInterpolant for a lowerable callable, else None (opaque Python callable)."""
    if isinstance(cell_size, GridInterpolant):
        return cell_size
    # a SciPy RegularGridInterpolator configured the way the reference configures it
    if (
        type(cell_size).__name__ == "RegularGridInterpolator"
        and getattr(cell_size, "method", None) == "linear"
        and getattr(cell_size, "bounds_error", True) is False
        and getattr(cell_size, "fill_value", 0) is None
        and len(cell_size.grid) in (2, 3)
    ):
        return GridInterpolant(cell_size.grid, np.asarray(cell_size.values, dtype=np.float64))
    return None


class SizeFunction:
    def __init__(self, bbox, cell_size, hmin):
        if not isinstance(bbox, tuple):
            raise ValueError("`bbox` must be a tuple")
        self.bbox = bbox
        if not callable(cell_size):
            raise ValueError("`cell_size` must be callable")
        self.cell_size = cell_size
        self.hmin = hmin
        self._interp = _lower_callable(cell_size)

    @classmethod
    def from_grid(cls, bbox, values, hmin):
        """Gridded size function over `bbox` with the reference's float32-linspace axes."""
        values = np.asarray(values, dtype=np.float64)
        return cls(tuple(bbox), GridInterpolant(grid_axes(bbox, values.shape), values), hmin)

    def interpolant(self):
        """The device-lowerable interpolant, or None for an opaque Python callable."""
        return self._interp

    def eval(self, x):
        if self._interp is not None:
            return self._interp(x)
        return self.cell_size(x)
