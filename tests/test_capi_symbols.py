"""The C-ABI library loads without a GPU and exports every symbol include/fbus_ekf.h declares; the entry points that
need a device fail loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fbus_ekf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fbus_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    from fbus_ekf_b200 import capi
    lib = capi.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libfbus_ekf.so does not export {n}"
    assert sorted(capi.EXPORTED_SYMBOLS) == names
    assert lib.fbus_abi_version() == capi.FBUS_ABI_VERSION == 3


def test_struct_sizes_match_header(built, tmp_path):
    """ctypes mirrors == C structs (compiled with gcc from the header)"""
    import subprocess
    from fbus_ekf_b200 import capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu\\n", sizeof(fbus_config), '
                   'sizeof(fbus_imu_stream), sizeof(fbus_det_frames), sizeof(fbus_state_soa), sizeof(fbus_synth_spec));return 0;}\n'
                   % os.path.join(ROOT, "include", "fbus_ekf.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(capi.FbusConfig), C.sizeof(capi.ImuStream), C.sizeof(capi.DetFrames), C.sizeof(capi.StateSoa),
                     C.sizeof(capi.SynthSpec)]


def test_helpers_without_device(built):
    from fbus_ekf_b200 import capi
    lib = capi.lib()
    cfg = capi.config_default()
    assert cfg.n_markers == 12 and cfg.marker_id[9] == 16
    assert abs(cfg.tsc_left[3] - 0.059967) < 1e-15 and cfg.n_water == 1.32 and cfg.reset_gap == 0.1
    R = np.array([1, 0, 0, 0, 0, -1, 0, 1, 0], dtype=np.float64)  # marker 1 of markersetup.yml
    q = np.zeros(4)
    lib.fbus_quat_from_rotmat(capi.dptr(R), capi.dptr(q))
    assert np.allclose(q, [np.sqrt(0.5), np.sqrt(0.5), 0, 0], atol=1e-15)
    R = np.array([1, 0, 0, 0, -1, 0, 0, 0, -1], dtype=np.float64)  # trace < 0 branch
    lib.fbus_quat_from_rotmat(capi.dptr(R), capi.dptr(q))
    assert np.allclose(q, [0, 1, 0, 0], atol=1e-15)
    # the workload generator's own NumPy copy of the rule (synth.quat_from_rotmat) and the product-free default config of the
    # CPU arm (orc.config_default) agree with the library bit for bit
    import orc
    from fbus_ekf_b200 import synth
    rng = np.random.default_rng(0)
    for _ in range(200):
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        A *= np.sign(np.linalg.det(A))
        lib.fbus_quat_from_rotmat(capi.dptr(np.ascontiguousarray(A.ravel())), capi.dptr(q))
        assert np.array_equal(q, synth.quat_from_rotmat(A))
    oc = orc.config_default()
    assert bytes(oc) == bytes(cfg)


def test_no_cpu_fallback(built):
    """without a CUDA device fbus_create must fail with FBUS_E_CUDA and say so"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from fbus_ekf_b200 import BatchFilter, FbusError
    with pytest.raises(FbusError) as e:
        BatchFilter(batch=4)
    assert "no usable CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)


def test_missing_library_fails_loudly(built, monkeypatch):
    from fbus_ekf_b200 import capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libfbus_ekf.so")
    with pytest.raises(RuntimeError) as e:
        capi.lib()
    assert "no CPU fallback" in str(e.value)
