#!/usr/bin/env python
"""Tracking-level check of the lean experiment's arithmetic on the CPU (no GPU needed): the level loop of k_track on one
host thread (tests/test_device_math_on_host.py: device source text + host shims), once with the library's per-point
functions (project_b / finish_point_b) and once with the experiment's (project_l / finish_point_p, packed accumulators),
against the float64 oracle after the same number of LM tries, and with the reference's own termination rules.

  python scratch/experiments/check_lean_tracking.py
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_device_math_on_host as H  # noqa: E402
from oracle import oracle as O  # noqa: E402
from revo_b200 import api, synth  # noqa: E402

LEAN_DRIVER = r'''
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 fmul2(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
@LEAN_PARTS@
static bool g_lean = false;
extern "C" void host_select_lean(int on) { g_lean = on != 0; }
'''


def main():
    lean = open(os.path.join(ROOT, "scratch", "experiments", "track_lean.cu")).read()
    g = H._grab
    parts = H.device_parts()
    lean_parts = [g(lean, r"^__device__ __forceinline__ ProjB project_l"), g(lean, r"^struct PackedAcc \{"),
                  g(lean, r"^__device__ __forceinline__ void finish_point_p")]
    driver = H.DRIVER.replace(
        '''        const ProjB P = project_b(true, p, L, R9, t3);''',
        '''        const ProjB P = g_lean ? project_l(p.x, p.y, p.z, L, R9, t3) : project_b(true, p, L, R9, t3);''').replace(
        '''        finish_point_b(P, r0, r1, L, ed, use_filter != 0, huber, acc);''',
        '''        if (g_lean) {
            PackedAcc S;
            S.clear();
            finish_point_p(P, r0, r1, fx * (1.0f / 32764.0f), fy * (1.0f / 32764.0f), use_filter ? ed : INFINITY, huber, S);
            S.unpack(acc, 1.0f);
        } else {
            finish_point_b(P, r0, r1, L, ed, use_filter != 0, huber, acc);
        }''')
    assert "g_lean ?" in driver and "S.unpack" in driver
    shim = H.SHIM.replace("struct float4 { float x, y, z, w; };", "struct float4 { float x, y, z, w; };\nstruct float2 { float x, y; };")
    src = shim + "\n".join(parts) + LEAN_DRIVER.replace("@LEAN_PARTS@", "\n".join(lean_parts)) + driver
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "scratch")) as d:
        cpp, so = os.path.join(d, "t.cpp"), os.path.join(d, "t.so")
        open(cpp, "w").write(src)
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                        cpp, "-o", so], check=True)
        lib = C.CDLL(so)
        orc = O.Oracle("f64")
        f = C.c_float

        def run(lean_on, cur, kf, lvl, R, T, oc):
            lib.host_select_lean(int(lean_on))
            cam = cur.cams[lvl]
            pts4 = np.ascontiguousarray(cur.edges3d[lvl], np.float32)
            dt = np.ascontiguousarray(kf.dt[lvl], np.float32)
            R9, t3 = np.ascontiguousarray(R.T.reshape(-1)), np.ascontiguousarray(T)
            err, n_evals, rec = C.c_float(0), C.c_int(0), np.zeros(32, np.float64)
            lib.host_track_level(pts4.ctypes.data_as(C.c_void_p), C.c_int(len(pts4)), dt.ctypes.data_as(C.c_void_p), C.c_int(cam.w),
                                 C.c_int(cam.h), f(cam.fx), f(cam.fy), f(cam.cx), f(cam.cy), R9.ctypes.data_as(C.c_void_p),
                                 t3.ctypes.data_as(C.c_void_p), C.byref(oc), C.c_int(lvl), C.byref(err), C.byref(n_evals),
                                 rec.ctypes.data_as(C.c_void_p))
            return R9.reshape(3, 3).T.copy(), t3.copy(), n_evals.value, err.value, int(rec[29]), int(rec[30])

        def rot(Ra, Rb):
            D = np.asarray(Ra, np.float64) @ np.asarray(Rb, np.float64).T
            return 0.5 * np.linalg.norm([D[2, 1] - D[1, 2], D[0, 2] - D[2, 0], D[1, 0] - D[0, 1]])

        worst = dict(fixed_rot=0.0, fixed_t=0.0, free_rot=0.0, free_t=0.0)
        same_evals_free = total_free = 0
        for seed in (1, 22, 5, 9):
            p = synth.make_pair(seed, 320, 240)
            cfg = O.PyrCfg(n_levels=3)
            kf = O.build_pyramid(orc, cfg, p["cam"], *p["key"])
            O.make_keyframe(orc, kf)
            cur = O.build_pyramid(orc, cfg, p["cam"], *p["cur"])
            T0 = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
            # (a) fixed number of LM tries: lean vs the float64 oracle, <= 1e-4 rad / m, same evaluation and point counts
            R, T = np.asarray(T0[:3, :3], np.float32), np.asarray(T0[:3, 3], np.float32)
            Ro, To = R.copy(), T.copy()
            oc = api.OptimizerSettings(USE_EDGE_FILTER=True, max_lm_tries=6, convergenceEps=[2.0] * 6)._c()
            for lvl in (2, 1, 0):
                ocfg = orc.default_cfg()
                for l in range(6):
                    ocfg.convergence_eps[l] = 2.0
                r = orc.track_level(cur.edges3d[lvl], kf.opt[lvl], cur.cams[lvl], Ro, To, ocfg, lvl, max_tries=6)
                Ro, To = r["R"].astype(np.float32), r["T"].astype(np.float32)
                R, T, ne, err, good, bad = run(True, cur, kf, lvl, R, T, oc)
                assert ne == r["n_evals"] and good == r["good"] and bad == r["bad"], (seed, lvl, ne, r["n_evals"], good, r["good"])
                worst["fixed_rot"] = max(worst["fixed_rot"], rot(R, Ro))
                worst["fixed_t"] = max(worst["fixed_t"], float(np.linalg.norm(T - To)))
            # (b) the reference's own termination rules: lean vs the library's functions on the same host loop
            oc = api.OptimizerSettings(USE_EDGE_FILTER=True)._c()
            Rb, Tb = np.asarray(T0[:3, :3], np.float32), np.asarray(T0[:3, 3], np.float32)
            Rl, Tl = Rb.copy(), Tb.copy()
            for lvl in (2, 1, 0):
                Rb, Tb, nb, _, _, _ = run(False, cur, kf, lvl, Rb, Tb, oc)
                Rl, Tl, nl, _, _, _ = run(True, cur, kf, lvl, Rl, Tl, oc)
                total_free += 1
                same_evals_free += int(nb == nl)
            worst["free_rot"] = max(worst["free_rot"], rot(Rl, Rb))
            worst["free_t"] = max(worst["free_t"], float(np.linalg.norm(Tl - Tb)))
        print(f"lean vs float64 oracle after 6 LM tries per level: max rot {worst['fixed_rot']:.2e} rad, max |dt| {worst['fixed_t']:.2e} m "
              f"(bar 1e-4), evaluation / good / bad counts equal")
        print(f"lean vs library functions, reference termination rules: max rot {worst['free_rot']:.2e} rad, max |dt| {worst['free_t']:.2e} m, "
              f"same evaluation count on {same_evals_free}/{total_free} levels")
        ok = worst["fixed_rot"] <= 1e-4 and worst["fixed_t"] <= 1e-4
        sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
