"""Test helper: write a minimal SEG-Y rev-1 file (3200-byte textual header, 400-byte binary header,
240-byte trace headers) from a (ns, ntraces) array, as IBM floats (format 1) or IEEE floats (5)."""
import math
import struct

import numpy as np


def ibm32(v):
    """IBM System/360 single-precision word of v (round to nearest on the 24-bit fraction)."""
    if v == 0.0:
        return 0
    sign = 0x80000000 if v < 0 else 0
    a = abs(float(v))
    e = int(math.floor(math.log(a, 16))) + 1
    frac = a / 16.0**e
    while frac >= 1.0:
        e, frac = e + 1, frac / 16.0
    while frac < 1.0 / 16.0:
        e, frac = e - 1, frac * 16.0
    m = int(round(frac * (1 << 24)))
    if m == 1 << 24:
        e, m = e + 1, m >> 4
    return sign | ((e + 64) << 24) | m


def write_segy(path, traces, fmt=1, dt_us=1000):
    traces = np.asarray(traces, dtype=np.float64)
    ns, ntr = traces.shape
    binary = bytearray(400)
    struct.pack_into(">H", binary, 16, dt_us)
    struct.pack_into(">H", binary, 20, ns)
    struct.pack_into(">H", binary, 24, fmt)
    with open(path, "wb") as f:
        f.write(b" " * 3200)
        f.write(bytes(binary))
        for k in range(ntr):
            hdr = bytearray(240)
            struct.pack_into(">H", hdr, 114, ns)
            f.write(bytes(hdr))
            if fmt == 1:
                f.write(struct.pack(f">{ns}I", *[ibm32(v) for v in traces[:, k]]))
            elif fmt == 5:
                f.write(struct.pack(f">{ns}f", *traces[:, k]))
            else:
                raise ValueError(fmt)
