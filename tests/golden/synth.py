"""Synthetic velocity models and sizing cases shared by the golden-vector generator (which runs them
through the REFERENCE) and by the tests (which run them through the oracle and the CUDA path)."""
import numpy as np


def synth_vp_2d(nz, nx, bbox):
    z = np.linspace(bbox[0], bbox[1], nz)[:, None]
    x = np.linspace(bbox[2], bbox[3], nx)[None, :]
    vp = 1500 + (-z / (bbox[1] - bbox[0]) * 1.0) * 3000 + 150 * np.sin(x / 900.0) * np.cos(z / 400.0)
    return np.ascontiguousarray(vp)


def synth_vp_3d(nz, nx, ny, bbox):
    z = np.linspace(bbox[0], bbox[1], nz)[:, None, None]
    x = np.linspace(bbox[2], bbox[3], nx)[None, :, None]
    y = np.linspace(bbox[4], bbox[5], ny)[None, None, :]
    vp = 1500 + (-z / (bbox[1] - bbox[0])) * 3000 + 100 * np.sin(x / 700.0) * np.cos(y / 500.0)
    return np.ascontiguousarray(vp)


def salt_vp_2d(nz, nx, bbox):
    """Layered background + water layer + a fast salt ellipse: sharp interfaces, so the gradient
    limiter has real work to do (same family as the BP2004-shaped bench workload)."""
    z = np.linspace(bbox[0], bbox[1], nz)[:, None] * np.ones((1, nx))
    x = np.linspace(bbox[2], bbox[3], nx)[None, :] * np.ones((nz, 1))
    Lz, Lx = bbox[1] - bbox[0], bbox[3] - bbox[2]
    vp = 1500 + (-(z - bbox[1]) / Lz) * 3000 + 150 * np.sin(x / (0.045 * Lx)) * np.cos(z / (0.125 * Lz))
    vp[z > bbox[1] - 0.083 * Lz - 0.025 * Lz * np.sin(x / (0.12 * Lx))] = 1486.0
    vp[((x - (bbox[2] + 0.45 * Lx)) / (0.134 * Lx)) ** 2 + ((z - (bbox[0] + 0.5 * Lz)) / (0.21 * Lz)) ** 2 < 1] = 4790.0
    return np.ascontiguousarray(vp)


def salt_vp_3d(nz, nx, ny, bbox):
    z = np.linspace(bbox[0], bbox[1], nz)[:, None, None] * np.ones((1, nx, ny))
    x = np.linspace(bbox[2], bbox[3], nx)[None, :, None] * np.ones((nz, 1, ny))
    y = np.linspace(bbox[4], bbox[5], ny)[None, None, :] * np.ones((nz, nx, 1))
    Lz, Lx, Ly = bbox[1] - bbox[0], bbox[3] - bbox[2], bbox[5] - bbox[4]
    vp = 1500 + (-(z - bbox[1]) / Lz) * 2800 + 120 * np.sin(x / (0.05 * Lx)) * np.cos(y / (0.04 * Ly))
    r2 = ((x - (bbox[2] + 0.5 * Lx)) / (0.22 * Lx)) ** 2 + ((y - (bbox[4] + 0.45 * Ly)) / (0.2 * Ly)) ** 2 \
        + ((z - (bbox[0] + 0.55 * Lz)) / (0.25 * Lz)) ** 2
    vp[r2 < 1] = 4480.0
    return np.ascontiguousarray(vp)


def sizing_cases():
    """name -> (vp, bbox, kwargs) for get_sizing_function_from_segy(None, bbox, velocity_data=vp, **kwargs)."""
    b2 = (-3000.0, 0.0, 0.0, 8000.0)
    b3 = (-2000.0, 0.0, 0.0, 4000.0, 0.0, 3000.0)
    return {
        "salt2d_edge": (salt_vp_2d(73, 181, b2), b2, dict(hmin=40.0, wl=10, freq=2.0, grade=0.15, dt=0.001,
                                                          domain_pad=400.0, pad_style="edge", nz=73, nx=181)),
        "salt2d_grad_const": (salt_vp_2d(64, 150, b2), b2, dict(hmin=30.0, hmax=400.0, wl=8, freq=3.0, grad=60.0,
                                                                stencil_size=7, grade=0.25, cr_max=0.5, dt=0.002,
                                                                space_order=2, domain_pad=300.0,
                                                                pad_style="constant", nz=64, nx=150)),
        "salt3d_ramp": (salt_vp_3d(24, 45, 37, b3), b3, dict(hmin=100.0, wl=5, freq=2.0, grade=0.15, hmax=5e3,
                                                            domain_pad=250.0, pad_style="linear_ramp", nz=24,
                                                            nx=45, ny=37)),
        "smooth3d_nograde": (synth_vp_3d(12, 20, 16, b3), b3, dict(hmin=150.0, wl=5, freq=2.0, grade=0.0, nz=12,
                                                                 nx=20, ny=16)),
    }
