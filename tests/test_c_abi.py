"""The drop-in boundary: libdistmesh_b200.so exports every function include/distmesh_b200.h declares,
the ctypes mirror binds exactly that set, and the struct layouts of the header and of the mirror
agree (checked with a C program compiled against the header).  No compute calls: runs without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "distmesh_b200.h")
LIB = os.path.join(ROOT, "seismicmesh_b200", "libdistmesh_b200.so")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(LIB)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in the header but not exported"
    lib.dm_version.restype = C.c_char_p
    assert b"sm_100a" in lib.dm_version()


def test_ctypes_mirror_binds_exactly_the_header():
    from seismicmesh_b200 import _lib

    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared()


def test_struct_layouts_match_the_header(tmp_path):
    from seismicmesh_b200 import _lib

    def fields(cls):
        return [f[0] for f in cls._fields_]

    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "distmesh_b200.h"', "int main(void) {"]
    for name, cls in (("DmSizeFn", _lib.DmSizeFn), ("DmPlan", _lib.DmPlan)):
        prog.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for fld in fields(cls):
            prog.append(f'  printf("{name}.{fld} %zu\\n", offsetof({name}, {fld}));')
    prog += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, cls in (("DmSizeFn", _lib.DmSizeFn), ("DmPlan", _lib.DmPlan)):
        assert int(out[name]) == C.sizeof(cls), name
        for fld in fields(cls):
            assert int(out[f"{name}.{fld}"]) == getattr(cls, fld).offset, f"{name}.{fld}"
    assert _lib.DM_MAX_LEVELS == 8 and _lib.DM_SDF_WORDS == 24


def test_missing_library_fails_loudly(tmp_path):
    """No CPU fallback: importing the package without the CUDA library raises."""
    code = (
        "import os, sys; sys.path.insert(0, %r); os.environ['DM_LIB_PATH'] = %r\n"
        "try:\n    import seismicmesh_b200\nexcept ImportError as e:\n    print('IMPORT_ERROR', e); sys.exit(0)\nsys.exit(1)\n"
    ) % (ROOT, str(tmp_path / "nope.so"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "no CPU fallback" in r.stdout
