#!/usr/bin/env python
"""The lean tracking kernel (scratch/experiments/track_lean.cu: k_track_lean, scalar and packed accumulation) on the CPU through
the kernel emulation layer of tests/_cuda_emu.py, against the library kernel k_track on the same layer and against the
float64 oracle after the same number of LM tries.  Covers what the static checks cannot: the two pipelined segments
(cached points / uncached tail, pcap 18 and 2), the per-thread trip counts, the shared-memory addressing, the reduction
the level loop and the one-sided exchange of multi-CTA clusters (emulated distributed shared memory + transaction barriers).

  python scratch/experiments/check_lean_kernel.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _cuda_emu  # noqa: E402
import test_kernel_on_host as TK  # noqa: E402
from conftest import rot_angle  # noqa: E402
from oracle import oracle as O  # noqa: E402
from revo_b200 import synth  # noqa: E402


def main():
    orc = O.Oracle("f64")
    T0 = synth.se3_exp([0.002, -0.001, 0.0015, 0.001, -0.0005, 0.0007])
    R0, t0 = np.asarray(T0[:3, :3], np.float32), np.asarray(T0[:3, 3], np.float32)
    pairs = [TK.build_pair(orc, seed) + (R0, t0) for seed in (1, 22)]
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "scratch")) as d:
        lib = _cuda_emu.build(d, with_lean=True)
        worst = 0.0
        for n_tries in (5,):
            ref = {}
            for variant, name in ((0, "k_track"), (1, "k_track_lean"), (2, "k_track_lean packed")):
                for n_ctas, pcap, cpp in ((1, 18, 1), (2, 2, 1), (1, 0, 1), (1, 3, 2), (1, 18, 4)):
                    out, _ = TK.run_kernel(lib, variant, pairs, TK.tracker_cfg(n_tries), n_ctas=n_ctas, pcap=pcap, ctas_per_pair=cpp)
                    for i, (kf, cur, _, _) in enumerate(pairs):
                        Ro, To, evals, last = TK.oracle_chain(orc, kf, cur, R0, t0, n_tries)
                        R = out["R"][i].reshape(3, 3).T
                        assert out["rc"][i] == 0 and list(out["n_evals"][i][:3]) == evals, (name, out["n_evals"][i], evals)
                        assert out["good"][i] == last["good"] and out["bad"][i] == last["bad"], (name, out["good"][i], last["good"])
                        dr, dt = rot_angle(R, Ro), float(np.linalg.norm(out["t"][i] - To))
                        assert dr <= 1e-4 and dt <= 1e-4, (name, dr, dt)
                        worst = max(worst, dr, dt)
                        if variant == 0:
                            ref[(n_ctas, pcap, cpp, i)] = out[i].copy()
                        else:
                            b = ref[(n_ctas, pcap, cpp, i)]
                            assert np.abs(out["R"][i] - b["R"]).max() < 1e-6 and np.abs(out["t"][i] - b["t"]).max() < 1e-6
                print(f"{name:22s} ok (pcap 18 / 2 / 0, clusters of 1, 2 and 4 CTAs, {n_tries} LM tries per level)")
        # default termination rules and the init check: same decisions as the library kernel
        cfg = TK.tracker_cfg(0, check_init=1)
        for l in range(6):
            cfg.opt.convergence_eps[l] = 0.999
        base, _ = TK.run_kernel(lib, 0, pairs, cfg)
        for variant in (1, 2):
            out, _ = TK.run_kernel(lib, variant, pairs, cfg)
            assert np.array_equal(out["n_evals"], base["n_evals"]) and np.array_equal(out["used_identity_init"], base["used_identity_init"])
            assert np.abs(out["R"] - base["R"]).max() < 1e-6 and np.abs(out["t"] - base["t"]).max() < 1e-6 and np.array_equal(out["status"], base["status"])
        print(f"reference termination rules + init check: same evaluation counts {base['n_evals'][:, :3].tolist()} and poses as k_track")
        print(f"worst deviation from the float64 oracle after the same number of tries: {worst:.2e} (bar 1e-4)")


if __name__ == "__main__":
    main()
