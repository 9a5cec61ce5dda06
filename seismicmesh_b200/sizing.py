"""Mesh-size functions with the reference's interface (SeismicMesh/sizing/size_function.py:1-12).

``SizeFunction(bbox, cell_size, hmin)`` keeps the reference constructor and ``.eval``; when
``cell_size`` is a gridded interpolant (our :class:`GridInterpolant`, or a SciPy
``RegularGridInterpolator`` as built by the reference at sizing/mesh_size_function.py:391-408)
it is lowered to a device descriptor and evaluated by the CUDA interpolation kernel with SciPy's
exact semantics (float32-rounded axes, linear extrapolation outside the grid).
"""
import ctypes as C
import warnings

import numpy as np
import torch

from . import _lib
from . import device as D
from ._lib import check, lib

__all__ = ["SizeFunction", "GridInterpolant", "grid_axes", "get_sizing_function_from_segy", "limgrad", "read_velocity_model"]


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

    def __init__(self, axes, values, cell_records=True):
        """values: host array, or a float64 CUDA tensor (the grid get_sizing_function_from_segy leaves on the
        device: it is used where it is and only comes to the host if somebody reads ``.values``)."""
        self.cell_records = bool(cell_records)  # 3-D: keep per-cell 64-B corner records on the device (8x the grid)
        self.grid = tuple(np.ascontiguousarray(a, dtype=np.float64) for a in axes)
        self._values_dev = None
        if isinstance(values, torch.Tensor) and values.is_cuda:
            self._values_dev = values.to(torch.float64).contiguous()
            self._values = None
            self.shape = tuple(int(n) for n in values.shape)
        else:
            self._values = np.ascontiguousarray(values, dtype=np.float64)
            self.shape = self._values.shape
        self.dim = len(self.grid)
        if self.dim not in (2, 3) or len(self.shape) != self.dim:
            raise ValueError("Dimension not supported")
        for a, n in zip(self.grid, self.shape):
            if a.ndim != 1 or len(a) != n or n < 2 or not np.all(np.diff(a) > 0):
                raise ValueError("grid axes must be strictly ascending and match the values' shape")
        self._dev = None

    @property
    def values(self):
        if self._values is None:
            self._values = self._values_dev.cpu().numpy()
        return self._values

    def device_arrays(self):
        dev = D.device()
        if self._dev is None or self._dev[1].device != dev:
            axes = [torch.from_numpy(a).to(dev) for a in self.grid]
            if self._values_dev is not None and self._values_dev.device == dev:
                grid = self._values_dev
            else:
                grid = torch.from_numpy(self.values).to(dev)
            cells = None
            if self.dim == 3 and self.cell_records:
                # a trilinear lookup then touches one aligned 64-B record instead of 4 sectors in 4 DRAM
                # pages; skipped when it would take more than a quarter of the free device memory
                nc = int(np.prod([n - 1 for n in self.shape]))
                if 64 * nc <= torch.cuda.mem_get_info(dev)[0] // 4:
                    cells = torch.empty(nc * 8, dtype=torch.float64, device=dev)
                    f = D.size_fn_struct(_lib.SIZE_GRID, 3, axes=axes, grid=grid)
                    check(lib.dm_size_build_cells(C.byref(f), D.ptr(cells), D.stream_ptr()), "dm_size_build_cells")
            self._dev = (axes, grid, cells)
        return self._dev

    def struct(self):
        axes, grid, cells = self.device_arrays()
        return D.size_fn_struct(_lib.SIZE_GRID, self.dim, axes=axes, grid=grid, cells=cells)

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


# ----------------------------------------------------------------------------------------------
# Sizing preprocessing (SURVEY section 8f "next #3"): velocity model -> gridded size function
# with the reference's interface (sizing/mesh_size_function.py:27-232).  The elementwise chain
# (wavelength sizing, clamp, CFL bound) is ONE fused kernel over the velocity grid and the gradient
# limiter -- the reference's only native code on this path (sizing/cpp/FastHJ.cpp) -- runs on the grid
# where that kernel left it, and so does the domain padding (np.pad's algorithm, dm_pad): the grid never leaves
# the device; the optional windowed-variance term likewise (dm_uniform_filter: SciPy's recurrence, bit-identical).
# ----------------------------------------------------------------------------------------------
_SIZING_DEFAULTS = {  # mesh_size_function.py:103-126
    "velocity_data": None, "vp_water": 1500.0, "hmin": 150.0, "hmax": 10000.0, "wl": 0, "freq": 2.0, "grad": 0.0,
    "grade": 0.0, "stencil_size": 10.0, "space_order": 1, "dt": 0.0, "cr_max": 1.0, "pad_style": "edge",
    "domain_pad": 0.0, "units": "m-s", "nz": None, "nx": None, "ny": None, "byte_order": "byte_order",
    "axes_order": (0, 1, 2), "axes_order_sort": "F", "dtype": "float32",
}


def _limgrad_device(f, grade, elen, max_sweeps=None):
    """dm_limgrad on a device tensor, in place (2-D or 3-D grid)."""
    shp = tuple(f.shape) if f.ndim == 3 else (f.shape[0], f.shape[1], 1)
    flag = torch.zeros(1, dtype=torch.int32, device=f.device)
    ftol = float(f.min().item()) * np.sqrt(1e-9)  # FastHJ.cpp:72 (EPS = 1e-9)
    sweeps = C.c_int(0)
    cap = int(max_sweeps) if max_sweeps is not None else 4 * int(sum(shp)) + 64
    tmp = torch.empty_like(f)  # the sweeps are Jacobi steps between two buffers (deterministic result)
    check(lib.dm_limgrad(D.ptr(f), D.ptr(tmp), shp[0], shp[1], shp[2], float(elen) * float(grade), ftol, cap, D.ptr(flag),
                         C.byref(sweeps), D.stream_ptr()), "dm_limgrad")
    limgrad.last_sweeps = sweeps.value
    return f


def limgrad(cell_size, grade, elen, max_sweeps=None):
    """Gradient-limit a gridded size function on the device (replaces _FastHJ.limgrad as called by
    _enforce_gradation_sizing, mesh_size_function.py:471-496): afterwards neighbouring nodes of the
    2*dim-edge stencil differ by at most ``elen*grade`` (+ the reference's ftol = min*sqrt(1e-9))."""
    a = np.ascontiguousarray(cell_size, dtype=np.float64)
    if a.ndim not in (2, 3):
        raise ValueError("Dimension not supported")
    shp = a.shape if a.ndim == 3 else (a.shape[0], a.shape[1], 1)
    D.require_cuda()
    f = torch.from_numpy(a).to(D.device())
    _limgrad_device(f, grade, elen, max_sweeps)
    return f.cpu().numpy().reshape(a.shape)


def _read_bin(filename, nz, nx, ny, byte_order, axes_order, axes_order_sort, dtype):
    """Binary velocity model -> (z, x, y) array, z flipped (mesh_size_function.py:609-630)."""
    if (nz is None) or (nx is None) or (ny is None):
        raise ValueError("Please specify the number of grid points in each dimension (e.g., `nz`, `nx`, `ny`)...")
    axes = [nz, nx, ny]
    axes = [axes[o] for o in np.argsort(axes_order)]
    if byte_order not in ("big", "little"):
        raise ValueError("Please specify byte_order as either: little or big.")
    vp = np.fromfile(filename, dtype=np.dtype(dtype).newbyteorder(">" if byte_order == "big" else "<"))
    vp = vp.reshape(*axes, order=axes_order_sort)
    return np.flipud(vp.transpose((*axes_order,))), nz, nx, ny


_PAD_MODES = {"edge": 0, "constant": 1, "linear_ramp": 2}


def _pad(array, padding, style, extra):  # mesh_size_function.py:575-587
    """np.pad(array, padding, style, ...) with the reference's three styles.  A CUDA tensor is padded on the
    device (dm_pad: NumPy's axis-by-axis algorithm and arithmetic, bit-identical) and stays there."""
    if style not in _PAD_MODES:
        raise ValueError("pad style currently not supported. Try `linear_ramp`, `edge`, or `constant`")
    if isinstance(array, torch.Tensor):
        a = array.to(torch.float64).contiguous()
        dim = a.ndim
        shape = (C.c_int64 * dim)(*a.shape)
        before = (C.c_int64 * dim)(*[int(w[0]) for w in padding])
        after = (C.c_int64 * dim)(*[int(w[1]) for w in padding])
        out = torch.empty(tuple(int(n + w[0] + w[1]) for n, w in zip(a.shape, padding)), dtype=torch.float64, device=a.device)
        flags = torch.zeros(2, dtype=torch.int32, device=a.device)
        check(lib.dm_pad(D.ptr(a), D.ptr(out), dim, shape, before, after, _PAD_MODES[style], float(extra[0]), float(extra[1]),
                         D.ptr(flags), D.stream_ptr()), "dm_pad")
        return out
    if style == "edge":
        return np.pad(array, padding, "edge")
    if style == "constant":
        return np.pad(array, padding, "constant", constant_values=tuple(extra))
    return np.pad(array, padding, "linear_ramp", end_values=tuple(extra))


_SEGY_FORMATS = {  # data sample format code (binary header bytes 3225-3226) -> big-endian dtype
    2: ">i4", 3: ">i2", 5: ">f4", 6: ">f8", 8: "i1", 9: ">i8", 10: ">u4", 11: ">u2", 12: ">u8", 16: "u1",
}


def _ibm32_to_float64(u):
    """IBM System/360 single precision (SEG-Y format code 1): sign, 7-bit excess-64 exponent of 16,
    24-bit fraction.  Every such value is exactly representable in float64."""
    u = u.astype(np.uint32)
    frac = (u & np.uint32(0x00FFFFFF)).astype(np.float64)
    expo = ((u >> np.uint32(24)) & np.uint32(0x7F)).astype(np.int64) - 64
    val = np.ldexp(frac, (4 * expo - 24).astype(np.int32))
    return np.where((u >> np.uint32(31)) != 0, -val, val)


def _read_segy(filename):
    """Velocity model from a SEG-Y file: what the reference gets out of ``segyio`` with
    ``ignore_geometry=True`` (sizing/mesh_size_function.py:633-646: one column per trace, samples
    down the column, then flipped so that row 0 is the deepest sample), read here with NumPy.
    Layout (SEG-Y rev 1): 3200-byte textual header, 400-byte binary header (samples per trace at
    bytes 3221-3222, format code at 3225-3226, extended textual headers at 3505-3506), then
    240-byte trace header + samples per trace, all traces the same length."""
    raw = np.fromfile(filename, dtype=np.uint8)
    if raw.size < 3600:
        raise ValueError(f"{filename}: not a SEG-Y file (shorter than its 3600-byte headers)")
    hdr = raw[3200:3600]

    def u16(off):
        return int(hdr[off]) << 8 | int(hdr[off + 1])

    ns, fmt = u16(20), u16(24)
    next_ = u16(304)
    ext = 0 if next_ >= 0x8000 else next_  # -1 = variable number of extended headers: not supported
    if fmt != 1 and fmt not in _SEGY_FORMATS:
        raise ValueError(f"{filename}: unsupported SEG-Y data sample format code {fmt}")
    width = 4 if fmt == 1 else np.dtype(_SEGY_FORMATS[fmt]).itemsize
    start = 3600 + 3200 * ext
    stride = 240 + ns * width
    if ns == 0 or (raw.size - start) % stride != 0 or raw.size <= start:
        raise ValueError(f"{filename}: size does not match {ns} samples per trace of {width} bytes")
    ntr = (raw.size - start) // stride
    body = raw[start:].reshape(ntr, stride)[:, 240:]
    if fmt == 1:
        traces = _ibm32_to_float64(np.ascontiguousarray(body).view(">u4"))
    else:
        traces = np.ascontiguousarray(body).view(_SEGY_FORMATS[fmt]).astype(np.float64)
    vp = np.ascontiguousarray(traces.T)  # (nz, nx): vp[:, index] = trace
    if np.amin(vp) < 1000.0:
        warnings.warn("Velocity appear to be in km/s. Maybe pass `units` km-s key pair?")
    return np.flipud(vp), int(ns), int(ntr), 0


def read_velocity_model(filename, nz=None, nx=None, ny=None, byte_order=None, axes_order=None, axes_order_sort=None,
                        dtype=None):
    """Read a velocity model: (vp, nz, nx, ny), z flipped (mesh_size_function.py:590-646).  SEG-Y through
    our own reader (IBM / IEEE / integer sample formats), anything else as a raw binary."""
    if str(filename).endswith(".segy"):
        return _read_segy(filename)
    return _read_bin(filename, nz, nx, ny, byte_order, axes_order, axes_order_sort, dtype)


def uniform_filter(a_dev, size, square_input=False):
    """scipy.ndimage.uniform_filter(a, size) (mode "reflect") of a 2-D / 3-D float64 CUDA tensor on the device
    (dm_uniform_filter: SciPy's running-sum recurrence per line, axis after axis: bit-identical)."""
    a = a_dev.contiguous()
    shp = tuple(a.shape) + (1,) * (3 - a.ndim)
    sz = (C.c_int * a.ndim)(*[int(s) for s in size])
    out, tmp = torch.empty_like(a), torch.empty_like(a)
    check(lib.dm_uniform_filter(D.ptr(a), D.ptr(out), D.ptr(tmp), shp[0], shp[1], shp[2], sz, a.ndim, 1 if square_input else 0,
                                D.stream_ptr()), "dm_uniform_filter")
    return out


def _gradient_sizing(vp, vp_dev, grad, stencil):
    """h_gr = grad / (normalised windowed variance of vp + 0.10)  (mesh_size_function.py:428-448) -> CUDA tensor."""
    window = [stencil] * vp.ndim if np.isscalar(stencil) else list(stencil)
    window = [int(w) for w in window]
    if any(w < 1 or w > n for w, n in zip(window, vp.shape)):  # windows longer than the grid: SciPy on the host
        from scipy import ndimage

        vp = vp_dev.cpu().numpy()  # (with the water layer filled in)
        win_mean = ndimage.uniform_filter(vp, tuple(window))
        win_var = ndimage.uniform_filter(vp**2, tuple(window)) - win_mean**2
        win_var = np.divide(win_var, np.amax(win_var))
        win_var -= np.amin(win_var)
        return torch.from_numpy(np.ascontiguousarray(grad / (win_var + 0.10))).to(vp_dev.device)
    mean = uniform_filter(vp_dev, window)
    sqr_mean = uniform_filter(vp_dev, window, square_input=True)
    var = torch.empty_like(vp_dev)
    st = D.stream_ptr()
    check(lib.dm_variance_size(D.ptr(mean), D.ptr(sqr_mean), var.numel(), 0, 0.0, 0.0, 0.0, D.ptr(var), st), "dm_variance_size")
    vmax, vmin = float(var.max().item()), float(var.min().item())
    # min(var / vmax) == min(var) / vmax: a correctly rounded division by a positive number is monotone
    check(lib.dm_variance_size(None, None, var.numel(), 1, vmax, vmin / vmax, float(grad), D.ptr(var), st), "dm_variance_size")
    return var


def get_sizing_function_from_segy(filename, bbox, comm=None, **kwargs):
    """Build a mesh-size function from a seismic velocity model: same name, arguments, defaults,
    errors and step order as the reference (mesh_size_function.py:27-232).  ``velocity_data=`` arrays
    binary files and SEG-Y files (own reader, no ``segyio``) are supported."""
    opts = dict(_SIZING_DEFAULTS)
    opts.update(kwargs)
    # The reference builds the grid on rank 0 only, hands the other ranks a placeholder (:128,:220) and
    # later ships every rank its resampled slab of rank 0's grid (migration.localize_sizing_function).
    # Here the grid is REPLICATED per GPU, so every rank builds the same size function
    # from the same input: same bbox (padded), same hmin, same grid everywhere, no placeholder to leak
    # into generate_mesh.  `comm` is accepted for call compatibility.
    vp, nz, nx, ny = opts["velocity_data"], opts["nz"], opts["nx"], opts["ny"]
    if vp is None:
        if str(filename).endswith(".segy"):
            vp, nz, nx, ny = _read_segy(filename)
        else:
            vp, nz, nx, ny = _read_bin(filename, nz, nx, ny, opts["byte_order"], opts["axes_order"],
                                       opts["axes_order_sort"], opts["dtype"])
    if opts["units"] == "km-s":
        vp *= 1000.0
    elif opts["units"] == "ft-s":
        vp *= 0.30
    if len(bbox) not in (4, 6):
        raise ValueError("Dimension not supported")
    dim = len(bbox) // 2
    if nz is None and opts["velocity_data"] is not None:  # lenient: grid shape of the array handed in
        nz, nx, ny = (tuple(vp.shape) + (None,))[:3]
    for key in kwargs:
        if key not in _SIZING_DEFAULTS:
            raise ValueError("Option %s with parameter %s not recognized " % (key, kwargs[key]))
    # ---- wavelength / gradient sizing (:411-450), clamp (:180-181), CFL bound (:453-468): argument checks in
    #      the reference's order, then ONE fused kernel over the velocity grid on the device
    want_grad = False
    if opts["wl"] > 0 or opts["grad"] > 0:
        if opts["wl"] < 0:
            raise ValueError("Parameter `wl` must be set > 0")
        if opts["freq"] < 0.0:
            raise ValueError("Parameter `freq` must be set > 0.0")
        if opts["grad"] < 0:
            raise ValueError("Parameter grad must be > 0")
        want_grad = opts["grad"] != 0.0
    cr_max, dt, so = opts["cr_max"], opts["dt"], opts["space_order"]
    if not ((cr_max == 0.0) or (dt == 0.0) or (so == 0.0)):
        if cr_max < 0:
            raise ValueError("Parameter `cr_max` must be > 0.0")
        if dt < 0:
            raise ValueError("Parameter `dt` must be > 0.0")
        if so < 1:
            raise ValueError("Parameter `space_order` must be >= 1 ")
    grade = opts["grade"]
    if grade < 0:
        raise ValueError("Parameter `grade` must be > 0.0")
    D.require_cuda()
    vp = np.ascontiguousarray(vp, dtype=np.float64)
    vp_dev = torch.from_numpy(vp).to(D.device())
    # water positions in shear-velocity data (:148-159: vp[vp < 1e-3] = vp_water), counted and replaced on the device
    cnt = torch.zeros(1, dtype=torch.int64, device=vp_dev.device)
    water = opts["vp_water"]
    ok = water is not None and 1300 <= water <= 1800
    check(lib.dm_replace_below(D.ptr(vp_dev), vp_dev.numel(), 1e-3, float(water) if ok else 0.0, 1 if ok else 0, D.ptr(cnt),
                               D.stream_ptr()), "dm_replace_below")
    if int(cnt.item()) > 0 and not ok:
        raise ValueError("vp_water is None or out of bounds. It should be >1300 and <1800 m/s")
    gr_dev = _gradient_sizing(vp, vp_dev, opts["grad"], opts["stencil_size"]) if want_grad else None
    cs_dev = torch.empty_like(vp_dev)
    check(lib.dm_size_from_velocity(D.ptr(vp_dev), D.ptr(gr_dev), vp_dev.numel(), dim, float(opts["freq"]), float(opts["wl"]),
                                    float(opts["hmin"]), float(opts["hmax"]), float(dt), float(cr_max), float(so), D.ptr(cs_dev),
                                    D.stream_ptr()), "dm_size_from_velocity")
    del vp_dev, gr_dev
    # gradation (:471-496): the limiter runs on the grid where it is
    if grade == 0.0:
        warnings.warn("Mesh size gradient is deactiavted. This may compromise mesh quality")
    else:
        if grade > 1.0:
            warnings.warn("Parameter `grade` is set pretty high (> 1.0)!")
        _limgrad_device(cs_dev, grade, (bbox[1] - bbox[0]) / nz)
    # domain extension (:526-572): np.pad's algorithm on the device, on the grid where the limiter left it.  (The
    # reference also pads vp, which nothing reads afterwards.)
    pad = opts["domain_pad"]
    if pad < 0:
        raise ValueError("Domain extension must be >= 0")
    if pad > 0:
        n = vp.shape
        d = [(bbox[2 * k + 1] - bbox[2 * k]) / n[k] for k in range(dim)]
        nn = [int(pad / dk) for dk in d]
        bbox = tuple(v for k in range(dim) for v in ((bbox[2 * k] - pad, bbox[2 * k + 1] + (pad if k > 0 else 0.0))))
        padding = tuple((nn[k], 0) if k == 0 else (nn[k], nn[k]) for k in range(dim))
        mx_h = float(cs_dev.max().item())
        cs_dev = _pad(cs_dev, padding, opts["pad_style"], [mx_h] * 2)
    cell_size = cs_dev
    # gridded interpolant with the reference's float32-linspace axes (:391-408, :514-523)
    return SizeFunction(tuple(bbox), GridInterpolant(grid_axes(bbox, cell_size.shape), cell_size), opts["hmin"])
