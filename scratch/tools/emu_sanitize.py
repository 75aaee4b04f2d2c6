#!/usr/bin/env python
"""Data-race / memory check of the kernels WITHOUT a GPU: the emulation layer (tests/_cuda_emu.py) runs one OS thread per
CUDA thread and synchronises them with pthread barriers and atomics, so ThreadSanitizer sees the kernels' shared-memory and
global-memory accesses with their real synchronisation -- a racecheck on the CPU -- and AddressSanitizer sees every
out-of-bounds access to the (heap) buffers the emulation allocates.

  python scratch/tools/emu_sanitize.py thread     # or: address
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TESTS = ["tests/test_kernel_on_host.py::test_k_track_on_host_matches_oracle_after_same_iterations",
         "tests/test_kernel_on_host.py::test_k_track_on_host_multi_cta_cluster_exchange",
         "tests/test_canny_on_host.py::test_canny_kernels_on_host_match_opencv",
         "tests/test_canny_on_host.py::test_canny_kernels_on_host_wide_and_tall_images",
         "tests/test_pyramid_on_host.py::test_pyramid_kernels_on_host_bit_exact",
         "tests/test_pyramid_on_host.py::test_pyramid_kernels_on_host_fill_in_and_group_compaction",
         "tests/test_pyramid_on_host.py::test_quality_vote_kernels_on_host"]


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "thread"
    rt = subprocess.run(["/usr/bin/g++", f"-print-file-name=lib{'tsan' if kind == 'thread' else 'asan'}.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ)
    env["REVO_EMU_CXXFLAGS"] = f"-fsanitize={kind} -g -fno-omit-frame-pointer"
    env["LD_PRELOAD"] = rt
    env["TSAN_OPTIONS"] = "halt_on_error=0 report_signal_unsafe=0 history_size=4 second_deadlock_stack=0"
    env["ASAN_OPTIONS"] = "detect_leaks=0 halt_on_error=0"
    tests = sys.argv[2:] or TESTS
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-s", "-p", "no:cacheprovider", *tests], cwd=ROOT, env=env, capture_output=True, text=True)
    out = r.stdout + r.stderr
    # one report = from a WARNING / ERROR line to the next "====" line; keep those whose stacks touch the emulated kernels
    # (cv2's own thread pool is not instrumented and shows up as false positives through the interceptors)
    reports, cur = [], None
    for line in out.splitlines():
        if "WARNING: ThreadSanitizer" in line or "ERROR: AddressSanitizer" in line:
            cur = [line]
        elif cur is not None:
            cur.append(line)
            if line.startswith("=================="):
                reports.append("\n".join(cur))
                cur = None
    ours = [r_ for r_ in reports if "_emu.so" in r_ or "_emu.cpp" in r_]
    seen = {}
    for r_ in ours:                                   # group by the source lines of the two accesses
        key = tuple(sorted(set(__import__("re").findall(r"(\w+_emu\.cpp:\d+)", r_)))[:4])
        seen.setdefault(key, r_)
    for key, r_ in seen.items():
        print("---- report at", key)
        print("\n".join(r_.splitlines()[:14]))
    tail = [l for l in out.splitlines() if "passed" in l or "failed" in l or "error" in l.lower()][-3:]
    print("\n".join(tail))
    print(f"[emu_sanitize] {kind}: pytest rc {r.returncode}, reports total {len(reports)}, in the emulated kernels {len(ours)} ({len(seen)} distinct sites)")
    return 1 if ours else 0


if __name__ == "__main__":
    sys.exit(main())
