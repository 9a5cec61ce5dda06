"""Why does a 64 MiB pinned host->device copy run at 36 GB/s and a 256 MiB one at 53 GB/s on this box?
Tries: order of the sizes, a 64 MiB slice of a large pinned block, chunked copies, cudaHostRegister'ed
numpy memory with and without transparent huge pages."""
import ctypes
import json
import mmap
import sys

import numpy as np
import torch

MiB = 1024 * 1024
out = {}


def rate(fn, nbytes, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return round(reps * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9, 2)


d = torch.empty(512 * MiB, dtype=torch.uint8, device="cuda")
# 1. order reversed
for mb in (256, 64, 8, 64, 256):
    h = torch.empty(mb * MiB, dtype=torch.uint8).pin_memory()
    out.setdefault("order_256_64_8_64_256", []).append(rate(lambda: d[: mb * MiB].copy_(h, non_blocking=True), mb * MiB))
    del h
# 2. slices of one large pinned block
big = torch.empty(512 * MiB, dtype=torch.uint8).pin_memory()
for mb in (8, 64, 256):
    out[f"slice_of_512MiB_block_{mb}MiB"] = rate(lambda: d[: mb * MiB].copy_(big[: mb * MiB], non_blocking=True), mb * MiB)
out["slice_64MiB_at_offset_100MiB"] = rate(lambda: d[: 64 * MiB].copy_(big[100 * MiB: 164 * MiB], non_blocking=True), 64 * MiB)
# 3. 64 MiB as 8 chunks of 8 MiB (what iterate_host does with the cell list)
def chunks():
    for k in range(8):
        d[k * 8 * MiB: (k + 1) * 8 * MiB].copy_(big[k * 8 * MiB: (k + 1) * 8 * MiB], non_blocking=True)
out["64MiB_in_8_chunks"] = rate(chunks, 64 * MiB)
# 4. touched-before vs fresh pages: write the host block first
big.fill_(1)
out["slice_64MiB_after_host_fill"] = rate(lambda: d[: 64 * MiB].copy_(big[: 64 * MiB], non_blocking=True), 64 * MiB)
# 5. cudaHostRegister on mmap'd memory, with MADV_HUGEPAGE
rt = ctypes.CDLL("libcudart.so.12") if True else None
try:
    for huge in (0, 1):
        m = mmap.mmap(-1, 128 * MiB)
        if huge:
            m.madvise(mmap.MADV_HUGEPAGE)
        arr = np.frombuffer(m, dtype=np.uint8)
        arr[:] = 3
        addr = arr.ctypes.data
        rc = rt.cudaHostRegister(ctypes.c_void_p(addr), ctypes.c_size_t(128 * MiB), 0)
        t = torch.from_numpy(arr)
        out[f"hostregister_mmap_hugepage{huge}_rc{rc}_64MiB"] = rate(lambda: d[: 64 * MiB].copy_(t[: 64 * MiB], non_blocking=True), 64 * MiB)
        rt.cudaHostUnregister(ctypes.c_void_p(addr))
        del t, arr
except Exception as e:  # noqa: BLE001
    out["hostregister_error"] = repr(e)
try:
    out["thp_enabled"] = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
except Exception:  # noqa: BLE001
    pass
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
