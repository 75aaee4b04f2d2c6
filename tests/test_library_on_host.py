"""The `-m gpu` tests on the CPU: the whole library -- C ABI host code (slab layout, descriptor tables, launch sequencing, batch /
keyframe / tracking / quality entry points) + kernels -- is compiled against the fake CUDA runtime and the kernel emulation
layer (``tests/_cuda_emu_lib.py``), loaded through ``revo_b200.api`` in place of the CUDA library, and the GPU test FUNCTIONS
of ``test_gpu_pyramid.py`` / ``test_gpu_track.py`` are called with it at reduced image sizes.  Slow (minutes: one OS thread per
CUDA thread), therefore only run with REVO_RUN_EMULATED_LIBRARY=1; the last run is recorded in
``profiles/r1_emulated_library_tests.txt``.  The product never loads this library: there is no CPU path outside the tests."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.skipif(os.environ.get("REVO_RUN_EMULATED_LIBRARY") != "1",
                                 reason="slow emulation run of the GPU tests on the CPU (set REVO_RUN_EMULATED_LIBRARY=1)"),
              pytest.mark.timeout(3600, method="thread")]

SMALL = (160, 120)


@pytest.fixture(scope="module")
def emu_ctx(tmp_path_factory):
    import _cuda_emu_lib
    import conftest
    import test_gpu_pyramid
    import test_gpu_track
    from revo_b200 import api

    lib = _cuda_emu_lib.build(str(tmp_path_factory.mktemp("revo_b200_emu")), with_lean=os.environ.get("REVO_EMU_WITH_LEAN") == "1")
    saved = (api._LIB_PATH, api._lib, test_gpu_pyramid.synth_pair, test_gpu_track.synth_pair)
    api._LIB_PATH, api._lib = lib, None

    def small_pair(seed, w=SMALL[0], h=SMALL[1], xi=None):          # the GPU tests default to VGA: far too slow here
        w, h = (SMALL if (w, h) == (640, 480) else (w, h))
        return conftest.synth_pair(seed, w, h, xi)

    test_gpu_pyramid.synth_pair = test_gpu_track.synth_pair = small_pair
    ctx = api.Context(0)
    ctx.set_track_engine(1, 0)
    ctx.set_track_shape(2, 128)                                      # 2 CTAs x 128 threads per pair keep the OS-thread count sane
    yield ctx
    ctx.close()
    api._LIB_PATH, api._lib, test_gpu_pyramid.synth_pair, test_gpu_track.synth_pair = saved


def test_pyramids_through_the_c_abi(emu_ctx, orc32):
    import test_gpu_pyramid as G

    G.test_pyramid_bit_exact(emu_ctx, orc32, 3, 160, 120, 3)
    G.test_pyramid_bit_exact(emu_ctx, orc32, 2, 160, 120, 4)
    G.test_pyramid_noise_and_empty(emu_ctx, orc32)
    G.test_bgra_input_and_errors(emu_ctx, orc32)


def test_batched_builds_and_wire_format(emu_ctx, orc32):
    import test_gpu_pyramid as G

    G.test_pyramid_batch_matches_single(emu_ctx, orc32)
    G.test_uint16_depth_wire_format(emu_ctx, orc32)


def test_fill_in_and_viewer_cloud(emu_ctx, orc32):
    import test_gpu_pyramid as G

    G.test_pyramid_fill_in_path(emu_ctx, orc32)
    G.test_colored_point_cloud_from_device_arrays.__wrapped__(emu_ctx, orc32) if hasattr(
        G.test_colored_point_cloud_from_device_arrays, "__wrapped__") else G.test_colored_point_cloud_from_device_arrays(emu_ctx, orc32)


def test_tracking_through_the_c_abi(emu_ctx, orc32, orc64):
    import test_gpu_track as T

    T.test_eval_record_matches_oracle(emu_ctx, orc32, orc64, 1)
    T.test_track_level_fixed_iterations(emu_ctx, orc64, 1, 8)
    T.test_track_frames_default_rules(emu_ctx, orc32, orc64, 1)
    # at this image size the identity-vs-initial-pose check of the bad pair sits on the float32 / float64 floor knife edge
    # (tests/test_kernel_on_host.py): the device follows the float32 reference, so that oracle is the one to compare with
    T.test_track_batch_matches_single_and_check_init(emu_ctx, orc32)
    T.test_track_error_codes(emu_ctx, orc64)


def test_vote_and_main_loops(emu_ctx, orc32, orc64):
    import test_gpu_track as T

    T.test_end_to_end_gpu_pyramids_track_to_ground_truth(emu_ctx, orc64)
    T.test_tracking_quality_vote(emu_ctx, orc64)
    T.test_revo_main_loop_on_gpu(emu_ctx, orc32, "cluster")
    T.test_multi_stream_main_loop_on_gpu(emu_ctx, "cluster")


@pytest.mark.skipif(os.environ.get("REVO_EMU_WITH_LEAN") != "1", reason="the experimental lean engine is only built with REVO_EMU_WITH_LEAN=1")
@pytest.mark.parametrize("pack", ["0", "1"])
def test_lean_engine_through_the_c_abi(emu_ctx, orc32, orc64, pack):
    """scratch/experiments/track_lean.cu wired in as engine 4 (what enable_lean_engine.patch does): the tracking tests through
    the launcher, the C ABI and the Python API, scalar and packed accumulation."""
    import test_gpu_track as T

    os.environ["REVO_LEAN_PACK"] = pack
    emu_ctx.set_track_engine(4, 0)
    try:
        T.test_eval_record_matches_oracle(emu_ctx, orc32, orc64, 1)
        T.test_track_level_fixed_iterations(emu_ctx, orc64, 1, 8)
        T.test_track_frames_default_rules(emu_ctx, orc32, orc64, 1)
        T.test_track_batch_matches_single_and_check_init(emu_ctx, orc32)
    finally:
        emu_ctx.set_track_engine(1, 0)
        os.environ.pop("REVO_LEAN_PACK", None)
