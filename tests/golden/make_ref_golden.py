#!/usr/bin/env python
"""Golden vectors from the REFERENCE ITSELF: runs oracle/_ref/librevo_ref.so (the reference's own ImgPyramidRGBD / Optimizer /
TrackerNew sources compiled against the API shims of oracle/shim/, recipe `make -C oracle ref`, only possible where
/root/reference exists) on small seeded inputs and writes tests/golden/ref_golden.npz: the inputs, SHA-256 digests of every
pyramid array, and the optimizer / tracker / vote results.  tests/test_oracle.py::test_oracle_matches_reference_golden checks the
oracle against this file wherever the tests run (the GPU box has no /root/reference).

    python tests/golden/make_ref_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import oracle as O
    from oracle import ref as RF
    from revo_b200 import synth

    assert RF.build(), "oracle/_ref cannot be built here (no /root/reference)"
    out = {}
    meta = {"cases": []}
    orc = O.Oracle("f32")
    ocfg = orc.default_cfg()
    for seed in (11, 12):
        w, h, nl = 160, 120, 3
        p = synth.make_pair(seed, w, h)
        (kb, kd), (cb, cd) = p["key"], p["cur"]
        kd = kd.copy()
        kd[4:9, 10:30] = 0.0
        k16, c16 = np.round(kd.astype(np.float64) * 5000).astype(np.uint16), np.round(cd.astype(np.float64) * 5000).astype(np.uint16)
        scale = np.float32(1.0) / np.float32(5000.0)
        kd, cd = k16.astype(np.float32) * scale, c16.astype(np.float32) * scale            # exactly what the test reconstructs
        rk = RF.RefPyramid(p["cam"], nl, kb, kd)
        rk.make_keyframe()
        rc = RF.RefPyramid(p["cam"], nl, cb, cd)
        case = {"seed": seed, "w": w, "h": h, "levels": nl, "cam": [float(x) for x in p["cam"]], "pyramid": [], "track_level": [], }
        out[f"s{seed}_key_bgr"], out[f"s{seed}_cur_bgr"] = kb, cb
        out[f"s{seed}_key_depth16"], out[f"s{seed}_cur_depth16"] = k16, c16
        for l in range(nl):
            case["pyramid"].append({what: digest(rk.get(what, l)) for what in ("gray", "depth", "edges", "edges_orig", "hist", "edges3d", "dt", "opt")}
                                   | {"cur_edges3d": digest(rc.get("edges3d", l)), "n_key": int(len(rk.get("edges3d", l))),
                                      "n_cur": int(len(rc.get("edges3d", l)))})
        R, T = np.eye(3, dtype=np.float32), np.zeros(3, np.float32)
        for lvl in (2, 1, 0):
            r = RF.opt_track_level(rk, rc, ocfg, lvl, R, T)
            R, T = r["R"], r["T"]
            e = RF.opt_eval(rk, rc, ocfg, lvl, R, T)
            out[f"s{seed}_l{lvl}_R"], out[f"s{seed}_l{lvl}_T"] = r["R"], r["T"]
            out[f"s{seed}_l{lvl}_A"], out[f"s{seed}_l{lvl}_b"] = e["A"], e["b"]
            case["track_level"].append({"lvl": lvl, "n_evals": r["n_evals"], "error": float(np.float32(r["error"])), "good": r["good"], "bad": r["bad"],
                                        "sum_w": float(np.float32(r["sum_w"])), "sum_unw": float(np.float32(r["sum_unw"])),
                                        "trace_good_bad": [[g, b] for g, b, _ in r["trace"]], "trace_error_6dec": [er for _, _, er in r["trace"]],
                                        "eval_good": e["good"], "eval_bad": e["bad"], "eval_sum_w": float(np.float32(e["sum_w"]))})
        trk = RF.RefTracker(rk, ocfg)
        bad = synth.se3_exp([0.4, 0.3, -0.2, 0.2, -0.15, 0.1])
        case["tracker"] = []
        for name, M in (("identity", np.eye(4)), ("bad_init", bad)):
            R0, T0 = np.asarray(M[:3, :3], np.float32), np.asarray(M[:3, 3], np.float32)
            r = trk.track_frames(rk, rc, R0, T0)
            out[f"s{seed}_trk_{name}_R0"], out[f"s{seed}_trk_{name}_T0"] = R0, T0
            out[f"s{seed}_trk_{name}_R"], out[f"s{seed}_trk_{name}_T"] = r["R"], r["T"]
            case["tracker"].append({"init": name, "status": r["status"], "error": float(np.float32(r["error"])), "n_tries": len(r["trace"]),
                                    "cost_init": float(np.float32(trk.eval_cost(rk, rc, R0, T0, 2)))})
        # quality vote: the key frame and the current frame as "past" clouds under small pose offsets
        rng = np.random.default_rng(seed)
        poses = [synth.se3_exp(rng.normal(0, 0.01, 6)).astype(np.float32) for _ in range(3)]
        est = synth.se3_exp(rng.normal(0, 0.01, 6)).astype(np.float32)
        votes = []
        for k, (pyr, P) in enumerate(zip((rk, rc, rk), poses)):
            trk.add_old(pyr, P, float(k))
            votes.append(trk.assess(rc, est))
        out[f"s{seed}_vote_poses"], out[f"s{seed}_vote_est"] = np.stack(poses), est
        case["votes"] = votes
        meta["cases"].append(case)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
