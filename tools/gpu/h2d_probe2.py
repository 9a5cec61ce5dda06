"""Which way of getting pinned host memory gives the fast host->device path?  Outcome (profiles/
r2r_h2d_probe_pinned_allocators.json): blocks from cudaHostAlloc / cudaHostRegister copy at 55 GB/s at every
size, UNTOUCHED `torch.empty(n).pin_memory()` blocks at 20-46 GB/s.  With real data in the buffers (as in
bench.py's e2e leg and generate_mesh) the allocator makes no difference: an A/B of the e2e leg on one box
gave 1.81 ms with cudaHostAlloc staging buffers against 1.74 ms with Tensor.pin_memory(), because that leg is
a serial chain (upload 65 MB -> stages B-D -> download 13 MB), not a copy at link rate.  Kept as evidence."""
import ctypes
import json
import mmap
import sys

import numpy as np
import torch

MiB = 1024 * 1024
out = {}
rt = ctypes.CDLL("libcudart.so.12")


def rate(fn, nbytes, reps=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9, 1)


d = torch.empty(256 * MiB, dtype=torch.uint8, device="cuda")
for mb in (13, 52, 64, 100, 128):
    n = mb * MiB
    r = []
    keep = []
    for rep in range(3):
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        keep.append(h)  # (fresh block every time)
        r.append((rate(lambda: d[:n].copy_(h, non_blocking=True), n), h.data_ptr() % (2 * MiB) // 4096))
    out[f"pin_memory_{mb}MiB"] = r
    del keep, h
    r = []
    for rep in range(3):
        m = mmap.mmap(-1, n + 2 * MiB)
        arr = np.frombuffer(m, dtype=np.uint8)
        arr[:] = 1
        rc = rt.cudaHostRegister(ctypes.c_void_p(arr.ctypes.data), ctypes.c_size_t(n + 2 * MiB), 0)
        t = torch.from_numpy(arr)[:n]
        r.append((rate(lambda: d[:n].copy_(t, non_blocking=True), n), rc))
        rt.cudaHostUnregister(ctypes.c_void_p(arr.ctypes.data))
        del t, arr
    out[f"hostregister_{mb}MiB"] = r
    r = []
    for flags in (0, 1, 4):  # default, portable, write-combined
        ptr = ctypes.c_void_p()
        rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(n), flags)
        buf = (ctypes.c_uint8 * n).from_address(ptr.value)
        arr = np.frombuffer(buf, dtype=np.uint8)
        t = torch.from_numpy(arr)
        r.append((rate(lambda: d[:n].copy_(t, non_blocking=True), n), flags, rc))
        del t, arr, buf
        rt.cudaFreeHost(ptr)
    out[f"cudaHostAlloc_flags_0_1_4_{mb}MiB"] = r
arena = torch.empty(256 * MiB, dtype=torch.uint8).pin_memory()
for mb in (13, 52, 64):
    n = mb * MiB
    out[f"arena256_slice_{mb}MiB"] = [rate(lambda: d[:n].copy_(arena[o * MiB: o * MiB + n], non_blocking=True), n) for o in (0, 7, 64)]
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
