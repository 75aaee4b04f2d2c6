"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/revo_b200.h declares, struct layouts agree between the header and the ctypes mirror, and the
product fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from revo_b200 import build

    build.build()
    from revo_b200 import api

    return api.load_library()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "revo_b200.h")).read()
    return sorted(set(re.findall(r"REVO_API\s+[\w\s\*]+?\b(revo_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from revo_b200 import api

    syms = _declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/revo_b200.h but not exported"
    assert sorted(api.EXPORTED_SYMBOLS) == syms


def test_struct_layouts_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    from revo_b200 import api

    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "revo_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(revo_camera), sizeof(revo_pyr_config), sizeof(revo_opt_config),
         sizeof(revo_tracker_config), sizeof(revo_residual_info), sizeof(revo_track_result), sizeof(revo_trace_entry));
  printf("%zu %zu %zu %zu\n", offsetof(revo_track_result, error), offsetof(revo_track_result, res),
         offsetof(revo_track_result, n_evals), offsetof(revo_opt_config, huber_edge));
  return 0; }
'''
    src = tmp_path / "t.c"
    src.write_text(prog)
    exe = tmp_path / "t"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    sizes = [C.sizeof(x) for x in (api.revo_camera, api.revo_pyr_config, api.revo_opt_config, api.revo_tracker_config,
                                   api.revo_residual_info, api.revo_track_result, api.revo_trace_entry)]
    assert [int(v) for v in out[:7]] == sizes
    offs = [api.revo_track_result.error.offset, api.revo_track_result.res.offset, api.revo_track_result.n_evals.offset,
            api.revo_opt_config.huber_edge.offset]
    assert [int(v) for v in out[7:]] == offs


def test_defaults_match_reference(lib):
    """Defaults of ImgPyramidSettings / OptimizerSettings / TrackerSettings (camerapyr.h:40-64, optimizer.h:46-85, tracker.h:43-46)."""
    from revo_b200 import api

    pc = api.revo_pyr_config()
    lib.revo_pyr_config_default(C.byref(pc))
    assert (pc.n_levels, pc.canny_threshold1, pc.canny_threshold2, pc.use_edge_hist, pc.patch0) == (3, 150, 100, 1, 20)
    assert abs(pc.depth_min - 0.1) < 1e-7 and abs(pc.depth_max - 5.2) < 1e-6 and abs(pc.n_percentage - 0.3) < 1e-7
    tc = api.revo_tracker_config()
    lib.revo_tracker_config_default(C.byref(tc))
    assert tc.check_init_values == 1 and tc.pyr_min_lvl == 2 and tc.pyr_max_lvl == 0
    o = tc.opt
    assert o.lambda_success_fac == 0.5 and o.lambda_fail_fac == 2.0 and o.use_edge_filter == 1 and o.max_lm_tries == 0
    assert list(o.edge_distance_lvl) == [30, 20, 10, 5, 5, 5] and list(o.max_its_per_lvl) == [100] * 6
    assert all(abs(v - 0.999) < 1e-6 for v in o.convergence_eps) and abs(o.huber_edge - 0.3) < 1e-7
    # and the python mirror agrees with the C defaults
    py = api.TrackerSettings().optimizerSettings._c()
    assert bytes(py) == bytes(o)


def test_no_cpu_fallback(lib):
    """Without a CUDA device the context cannot be created and nothing else is callable."""
    import torch

    from revo_b200 import api

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.RevoError) as ei:
        api.Context(0)
    assert ei.value.code == 2
    assert b"no CUDA device" in lib.revo_strerror(2)


def test_camera_pyramid_rule():
    """fx,fy,cx,cy * 2^-l and w,h = floor(w * 2^-l) (camerapyr.h:98-103,139-144); nLevels+1 cameras."""
    from revo_b200 import api

    st = api.ImgPyramidSettings(PYR_MIN_LVL=3, width=640, height=480, fx=517.3, fy=516.5, cx=318.6, cy=255.3)
    cp = api.CameraPyr(st)
    assert cp.size() == st.nLevels() + 1 == 5
    assert (cp.at(3).width, cp.at(3).height) == (80, 60)
    assert abs(cp.at(2).fx - 517.3 / 4) < 1e-4 and abs(cp.at(1).cy - 255.3 / 2) < 1e-4


def test_pose_helpers_quaternion_round_trip():
    """revo_quat_to_R9 / revo_R9_to_quat (host arithmetic in the C ABI, the Sophus::SE3f side of the adapter): round trip,
    Eigen's branch for trace <= 0, and the tracker's own not-a-rotation error instead of Sophus' abort()."""
    from revo_b200 import api, synth, tum_io

    rng = np.random.default_rng(5)
    for _ in range(20):
        R = synth.se3_exp(np.r_[np.zeros(3), rng.normal(0, 1.5, 3)])[:3, :3]
        q = api.R_to_quat(R)
        assert abs(np.linalg.norm(q) - 1) < 1e-6
        assert np.allclose(q, tum_io.quaternion_from_R(R), atol=1e-6)
        assert np.allclose(api.quat_to_R(q), R, atol=2e-6)
        assert np.allclose(api.quat_to_R(3.0 * q), R, atol=2e-6)          # normalised by the library
    q = api.R_to_quat(np.diag([-1.0, -1.0, 1.0]))                            # trace <= 0 branch
    assert np.allclose(np.abs(q), [0, 0, 1, 0], atol=1e-7)
    with pytest.raises(api.RevoError) as e:
        api.R_to_quat(np.diag([1.0, 1.0, 1.01]))
    assert e.value.code == 5                                                 # REVO_ERR_NOT_ORTHOGONAL
    with pytest.raises(api.RevoError):
        api.R_to_quat(np.diag([1.0, 1.0, -1.0]))                             # det < 0
    with pytest.raises(api.RevoError):
        api.quat_to_R(np.zeros(4))
