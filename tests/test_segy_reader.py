"""SEG-Y reading for get_sizing_function_from_segy (reference: sizing/mesh_size_function.py:633-646
through segyio, which is not installed): our NumPy reader against the velocity model of the
reference's own fixture tests/testing.segy (decoded value by value by oracle/ref_harness.py's
stand-in and committed as tests/golden/segy_testing.npz by make_golden.py)."""
import os
import warnings

import numpy as np
import pytest
from conftest import load_golden
from segy_util import write_segy

REF_FIXTURE = "/root/reference/tests/testing.segy"


def _read(path):
    from seismicmesh_b200.sizing import _read_segy

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return _read_segy(str(path))


def test_ibm_float_file_round_trip(tmp_path):
    g = load_golden("segy_testing.npz")
    f = tmp_path / "model.segy"
    write_segy(f, g["traces"], fmt=1)
    vp, nz, nx, ny = _read(f)
    assert (nz, nx, ny) == (10, 10, 0)
    assert np.array_equal(vp, g["vp"])  # bit-exact: row 0 is the deepest sample (np.flipud, :646)
    assert vp[0, 0] == 7000.0 and vp[-1, 0] == 1500.0


@pytest.mark.skipif(not os.path.exists(REF_FIXTURE), reason="reference tree only exists in the build container")
def test_reference_fixture_decodes_to_the_golden_model():
    vp, nz, nx, ny = _read(REF_FIXTURE)
    assert (nz, nx, ny) == (10, 10, 0)
    assert np.array_equal(vp, load_golden("segy_testing.npz")["vp"])


def test_ibm_and_ieee_values(tmp_path):
    from seismicmesh_b200.sizing import _ibm32_to_float64

    # known IBM words: 1.0 = 0x41100000, -118.625 = 0xC276A000, 0 = 0
    w = np.array([0x41100000, 0xC276A000, 0x00000000, 0x42640000], dtype=np.uint32)
    assert _ibm32_to_float64(w).tolist() == [1.0, -118.625, 0.0, 100.0]
    rng = np.random.default_rng(3)
    tr = np.round(rng.uniform(1400, 6000, (37, 5)), 1)
    f5 = tmp_path / "ieee.segy"
    write_segy(f5, tr, fmt=5)
    vp, nz, nx, _ = _read(f5)
    assert (nz, nx) == (37, 5) and np.array_equal(vp, np.flipud(tr.astype(np.float32).astype(np.float64)))
    f1 = tmp_path / "ibm.segy"
    write_segy(f1, tr, fmt=1)
    vp1, _, _, _ = _read(f1)
    assert np.abs(vp1 - np.flipud(tr)).max() <= 6000 * 2.0**-20  # 24-bit fraction, up to 3 leading zero bits


def test_units_warning_and_errors(tmp_path):
    from seismicmesh_b200.sizing import _read_segy

    f = tmp_path / "kms.segy"
    write_segy(f, np.full((4, 3), 1.5), fmt=5)
    with pytest.warns(UserWarning, match="km/s"):
        _read_segy(str(f))
    short = tmp_path / "short.segy"
    short.write_bytes(b"\0" * 100)
    with pytest.raises(ValueError, match="SEG-Y"):
        _read_segy(str(short))
    bad = tmp_path / "bad.segy"
    write_segy(bad, np.full((4, 3), 2000.0), fmt=5)
    with open(bad, "ab") as fh:
        fh.write(b"\0" * 7)  # truncated trailing trace
    with pytest.raises(ValueError, match="size does not match"):
        _read_segy(str(bad))
