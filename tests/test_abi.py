"""The C-ABI library loads and exports every symbol include/roft_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "roft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(roftb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    from roft_b200 import api
    lib = api.load_library()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in roft_b200.h but not exported"
    assert sorted(api.EXPORTED_SYMBOLS) == names


def test_config_defaults_are_the_fast_ycb_config(lib_built):
    from roft_b200 import api
    c = api.default_config()
    # config/config_fast_ycb.cfg
    assert (c.width, c.height) == (1280, 720) and abs(c.fx - 1229.4285612615463) < 1e-12 and c.cx == 640.0 and c.cy == 360.0
    assert c.subsampling_radius == 35 and c.weight_flow == 1 and c.depth_maximum == 2.0 and list(c.cov_flow) == [1.0, 1.0]
    assert list(c.v_sigma) == [0.1] * 6 and list(c.v_cov0) == [1e-3] * 6 and list(c.p_cov0) == [1e-3] * 12
    assert list(c.cov_v) == [0.1] * 3 and list(c.cov_w) == [1e-4] * 3 and list(c.cov_x) == [1e-3] * 3 and list(c.cov_q) == [1e-4] * 3
    assert (c.ut_alpha, c.ut_beta, c.ut_kappa) == (1.0, 2.0, 0.0)
    assert (c.use_pose, c.use_pose_resync, c.use_velocity, c.flow_aided) == (1, 1, 1, 1)
    assert c.segm_delay == 6 and c.pose_delay == 6
    assert ctypes.sizeof(api.RoftbConfig) % 8 == 0


def test_no_cpu_fallback(lib_built):
    """Without a CUDA device creating a context must fail loudly; bad arguments are rejected everywhere."""
    import torch
    from roft_b200 import api
    lib = api.load_library()
    bad = api.default_config(width=1281)
    h = ctypes.c_void_p()
    assert lib.roftb_create(ctypes.byref(bad), ctypes.byref(h)) < 0
    assert b"width" in lib.roftb_last_error(None)
    if not torch.cuda.is_available():
        with pytest.raises(api.RoftbError, match="no CUDA device"):
            api.Tracker(api.default_config(n_tracks=1, width=64, height=48))
    assert lib.roftb_filter_step(None, None) < 0 and lib.roftb_sync(None) < 0


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "roft_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "roft_oracle" not in txt and "cpu_ref" not in txt, f"{f} references the oracle"
