#!/usr/bin/env python
"""Host-side check of the per-point arithmetic of scratch/experiments/track_lean.cu (no GPU needed).

The function TEXT is taken from the sources (track_common.cuh: unpack_grad, ProjB, LevelConst, project_b, finish_point_b;
track_lean.cu: project_l, finish_point_l, PackedAcc, finish_point_p), wrapped with host shims (fma.rn.f32x2 -> two fmaf,
rcp.approx -> 1/x) and compiled with g++ -mfma -ffp-contract=fast, so that `a * b + c` fuses on the host as it does on
the device.  Checked on random points / records:
  * finish_point_l == finish_point_b (library) with the "bad" count derived as visited - good,
  * finish_point_p (packed accumulators, unpacked to record order) == finish_point_l to float rounding (the host compiler
    is free to fuse the Jacobian expressions differently in the two functions, so bit-equality is not a criterion; a slot
    mapped to the wrong sum would be off by orders of magnitude),
  * project_l == project_b up to the rounding of the re-associated rigid transform.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def grab(text, start_pat, what):
    """Text of a function / struct starting at the line matching start_pat up to its closing brace at column 0."""
    m = re.search(start_pat, text, re.M)
    assert m, what
    i = m.start()
    j = text.index("\n}", i)
    end = text.index("\n", j + 1)
    return text[i:end + 1]


def main():
    common = open(os.path.join(ROOT, "revo_b200", "csrc", "track_common.cuh")).read()
    lean = open(os.path.join(ROOT, "scratch", "experiments", "track_lean.cu")).read()
    parts = [
        grab(common, r"^__device__ __forceinline__ void unpack_grad", "unpack_grad"),
        grab(common, r"^struct ProjB \{", "ProjB"),
        grab(common, r"^struct LevelConst \{", "LevelConst"),
        grab(common, r"^__device__ __forceinline__ ProjB project_b", "project_b"),
        grab(common, r"^__device__ __forceinline__ void finish_point_b", "finish_point_b"),
        grab(lean, r"^__device__ __forceinline__ ProjB project_l", "project_l"),
        grab(lean, r"^__device__ __forceinline__ void finish_point_l", "finish_point_l"),
        grab(lean, r"^struct PackedAcc \{", "PackedAcc"),
        grab(lean, r"^__device__ __forceinline__ void finish_point_p", "finish_point_p"),
    ]
    shim = r'''
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#define __device__
#define __forceinline__ inline
#define __restrict__
struct uint4 { uint32_t x, y, z, w; };
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float rcp_approx(float x) { return 1.0f / x; }
static inline float2 ffma2(float2 a, float2 b, float2 c) { return float2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
static inline float2 fmul2(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
constexpr int kRecA = 0, kRecB = 21, kRecSW = 27, kRecSU = 28, kRecGood = 29, kRecBad = 30;
'''
    main_c = r'''
static uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
int main()
{
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    LevelConst L;
    L.fx = 517.3f; L.fy = 516.5f; L.cx = 318.6f; L.cy = 255.3f; L.umax = 638.f; L.vmax = 478.f; L.w = 640; L.opt = nullptr;
    const float kqfx = L.fx * (1.0f / 32764.0f), kqfy = L.fy * (1.0f / 32764.0f);
    float accB[32] = {0}, accL[32] = {0}, accP[32];
    PackedAcc S;
    S.clear();
    int visited = 0, proj_bad = 0;
    double max_da = 0;
    for (int it = 0; it < 200000; ++it) {
        // pose near identity, point in front of the camera (some project out of bounds, some lie behind it)
        float R[9] = {1, 0.01f * U(rng), -0.02f * U(rng), -0.01f * U(rng), 1, 0.015f * U(rng), 0.02f * U(rng), -0.015f * U(rng), 1};
        float t[3] = {0.05f * (U(rng) - 0.5f), 0.05f * (U(rng) - 0.5f), 0.05f * (U(rng) - 0.5f)};
        const float z = (it % 97 == 0) ? -1.f : 0.5f + 4.f * U(rng);
        const float x = (U(rng) - 0.5f) * 1.6f * z, y = (U(rng) - 0.5f) * 1.2f * z;
        const float ed = (it % 5 == 0) ? 3.f : 30.f, ed_eff = (it % 11 == 0) ? INFINITY : ed;
        const bool use_filter = !(it % 11 == 0);
        const float huber = 0.3f;
        uint4 r0, r1;
        const float base = 6.f * U(rng);
        r0.x = f2u(base + U(rng)); r0.y = f2u(base + U(rng)); r1.x = f2u(base + U(rng)); r1.y = f2u(base + U(rng));
        auto g = [&]() { const int a = (int)(65528.f * U(rng)) - 32764, b = (int)(65528.f * U(rng)) - 32764; return ((uint32_t)a & 0xffffu) | ((uint32_t)b << 16); };
        r0.z = g(); r0.w = g(); r1.z = g(); r1.w = g();
        struct P4 { float x, y, z, w; } p{x, y, z, 1.f};
        const ProjB Pb = project_b(true, p, L, R, t);
        const ProjB Pl = project_l(x, y, z, L, R, t);
        if (Pb.valid != Pl.valid) { ++proj_bad; continue; }      // a pixel-border knife edge of the re-associated transform
        max_da = std::fmax(max_da, std::fabs((double)Pb.a - Pl.a));
        ++visited;
        finish_point_b(Pl, r0, r1, L, ed, use_filter, huber, accB);
        finish_point_l(Pl, r0, r1, kqfx, kqfy, ed_eff, huber, accL);
        finish_point_p(Pl, r0, r1, kqfx, kqfy, ed_eff, huber, S);
    }
    accL[kRecBad] = (float)visited - accL[kRecGood];
    S.unpack(accP, (float)visited);
    int bad_lb = 0, bad_pl = 0;
    for (int i = 0; i < 31; ++i) {
        if (f2u(accL[i]) != f2u(accB[i])) { ++bad_lb; std::printf("l vs b: slot %d %.9g %.9g\n", i, accL[i], accB[i]); }
        if (std::fabs((double)accP[i] - accL[i]) > 1e-5 * (std::fabs((double)accL[i]) + 1.0)) { ++bad_pl; std::printf("p vs l: slot %d %.9g %.9g\n", i, accP[i], accL[i]); }
    }
    std::printf("visited %d good %.0f bad %.0f | project_l vs project_b: %d validity flips, max |da| %.3g | mismatching slots: l/b %d, p/l %d\n",
                visited, accL[kRecGood], accL[kRecBad], proj_bad, max_da, bad_lb, bad_pl);
    return (bad_lb || bad_pl || proj_bad > 20 || max_da > 1e-5 || accL[kRecGood] < 1000) ? 1 : 0;
}
'''
    src = shim + "\n".join(parts).replace("const float4 p,", "const P4T p,") + main_c
    # project_b takes a float4 point: give it a host struct of that shape
    src = src.replace("#define __restrict__", "#define __restrict__\nstruct P4T { float x, y, z, w; };")
    src = src.replace("struct P4 { float x, y, z, w; } p{x, y, z, 1.f};", "P4T p{x, y, z, 1.f};")
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "scratch")) as d:
        cpp, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(cpp, "w").write(src)
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-mfma", "-ffp-contract=fast", cpp, "-o", exe], check=True)
        r = subprocess.run([exe], capture_output=True, text=True)
        print(r.stdout.strip())
        sys.exit(r.returncode)


if __name__ == "__main__":
    main()
