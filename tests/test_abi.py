"""CPU: the C-ABI library loads, exports every symbol include/msckf_b200.h declares, its
structs have the layout the ctypes mirror assumes, and it fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_functions():
    src = open(os.path.join(ROOT, "include", "msckf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mskf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from msckf_stereo_c_b200 import engine

    names = _header_functions()
    assert len(names) >= 24
    assert set(names) == set(engine.ABI_SYMBOLS)
    L = engine.lib()
    for n in names:
        assert hasattr(L, n), n


def test_struct_layout_matches_ctypes(tmp_path):
    from msckf_stereo_c_b200 import abi

    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-o", str(exe), os.path.join(ROOT, "tools", "check_abi_sizes.c")])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(t) for t in (abi.Config, abi.Feature, abi.TrackingInfo, abi.GridFeature, abi.State, abi.CamState)]
    assert got == want


def test_presets():
    from msckf_stereo_c_b200 import engine

    ref = engine.default_config("ref")
    # what the reference code hard-codes / reads
    assert (ref.pyramid_levels, ref.klt_win, ref.klt_max_iters) == (4, 15, 30)
    assert (ref.grid_row, ref.grid_col, ref.grid_min_feature_num, ref.grid_max_feature_num) == (4, 5, 3, 4)
    assert (ref.det_rows, ref.det_cols, ref.fast_threshold) == (30, 47, 10)
    assert ref.max_cam_state_size == 20 and ref.compat_stale_features == 1 and ref.use_ransac == 0
    bench = engine.default_config("bench")
    assert (bench.klt_win, bench.max_cam_state_size, bench.grid_max_feature_num) == (21, 30, 15)
    stress = engine.default_config("stress")
    assert (stress.img_rows, stress.img_cols, stress.pyramid_levels) == (1024, 1280, 6)
    with pytest.raises(ValueError):
        engine.default_config("nope")


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from msckf_stereo_c_b200 import engine

    with pytest.raises(engine.EngineError):
        engine.Engine(engine.default_config("ref"), 1)
