// C++ consumer of revo_b200/host/revo_host.hpp: the reference's call sequence
//   ImgPyramidRGBD(settings, camPyr, rgb, depth, ts) x2 ; kf->makeKeyframe() ; TrackerNew::trackFrames(R, T, error, kf, cur)
// (system/system.cpp:151-188) through the C ABI.
//   test_host --selftest             : no GPU needed; settings/camera rules, error path without a device
//   test_host <fixture.bin> <out.txt>: tracks the pair stored in the fixture and writes status, error, R, T, evals
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../revo_b200/host/revo_host.hpp"
#include "../../revo_b200/host/revo_system.hpp"

// ---- fakes for the main-loop mirror (revo_system.hpp): a tracker that always reports 1 cm along x relative to the keyframe in
// use and whose vote always asks for a new keyframe (the scenario of tests/test_system_host.py::test_forced_keyframe_switch)
struct FakePyr {
    explicit FakePyr(double ts) : ts_(ts) { const revo::Mat4f I = revo::Mat4f::Identity(); std::memcpy(T_, I.data(), sizeof(T_)); }
    void makeKeyframe() { kf = true; }
    void setTwf(const float T[16]) { std::memcpy(T_, T, sizeof(T_)); }
    const float *getTransKFtoWorld() const { return T_; }
    double returnTimestamp() const { return ts_; }
    int frameId = 0;
    bool kf = false;
    double ts_;
    float T_[16];
};
struct FakeTracker {
    int cleared = 0, votes = 0, bad_ref = 0;
    int trackFrames(revo::Mat3f &R, revo::Vec3f &T, float &error, const std::shared_ptr<FakePyr> &ref, const std::shared_ptr<FakePyr> &) {
        if (!ref->kf) ++bad_ref;
        R = revo::Mat3f::Identity(); T = revo::Vec3f{{0.01f, 0.f, 0.f}}; error = 0.1f;
        return revo::STATE_OK;
    }
    int assessTrackingQuality(const float *, const std::shared_ptr<FakePyr> &) { ++votes; return revo::STATE_NEW_KF; }
    void addOldPclAndPose(const std::shared_ptr<FakePyr> &, const float *, double) {}
    void clearUpPastLists() { ++cleared; }
};

// compile check of the instantiation over the CUDA classes (never called here: it needs a device)
__attribute__((unused)) static revo::Mat4f instantiate_real_loop(const std::shared_ptr<TrackerNew> &t, const std::shared_ptr<ImgPyramidRGBD> &p)
{
    REVOLoop loop(t);
    return loop.processFrame(p);
}

static int selftest_main_loop(const char *self)
{
    auto trk = std::make_shared<FakeTracker>();
    revo::REVOLoopT<FakePyr, FakeTracker> sys(trk);
    std::vector<std::shared_ptr<FakePyr>> frames;
    for (int i = 0; i < 6; ++i) {
        frames.push_back(std::make_shared<FakePyr>(0.033 * i));
        sys.processFrame(frames.back());
    }
    // frame 0 is a keyframe; frame 1: vote NEW_KF but justAdded -> no switch; frame 2: switch to frame 1; 3: no; 4: switch to 3
    const bool want_kf[6] = {true, true, false, true, false, false};
    for (int i = 0; i < 6; ++i)
        if (frames[i]->kf != want_kf[i]) return 30 + i;
    if (sys.retracked != std::vector<int>{2, 4} || trk->cleared != 2 || sys.nKeyFrames != 3 || trk->bad_ref) return 40;
    const float want_x[6] = {0.f, 0.01f, 0.02f, 0.02f, 0.03f, 0.03f};
    const std::vector<revo::Mat4f> traj = sys.trajectory();
    if (traj.size() != 6) return 41;
    for (int i = 0; i < 6; ++i)
        if (std::fabs(traj[i](0, 3) - want_x[i]) > 1e-6f) return 50 + i;
    // motion model: the guess of the next frame is T_kf_N * T_NM1_N (system.cpp:268)
    const revo::Mat4f T_init = sys.mPoseGraph.back().T_kf_N() * sys.T_NM1_N;
    if (std::fabs(T_init(0, 3) - sys.T[0]) > 1e-6f || std::fabs(sys.R(0, 0) - 1.f) > 1e-6f) return 60;
    // TUM wire formats: trajectory line and association list
    revo::Mat4f W = revo::Mat4f::Identity();
    W(0, 0) = 0.f; W(0, 1) = -1.f; W(1, 0) = 1.f; W(1, 1) = 0.f; W(0, 3) = 1.5f; W(1, 3) = -2.f; W(2, 3) = 0.25f;   // 90 deg about z
    if (revo::poseToTUMString(W, 1305031102.175304) != "1305031102.175304 1.500000000 -2.000000000 0.250000000 0.000000000 0.000000000 0.707106769 0.707106769")
        return 62;
    {
        const std::string path_s = std::string(self) + ".assoc_selftest.txt";   // next to the binary, removed below
        const char *path = path_s.c_str();
        FILE *f = std::fopen(path, "w");
        if (!f) return 63;
        std::fputs("# comment\n\n1.5 rgb/a.png 1.25 depth/a.png\n2.5 rgb/b.png 2.25 depth/b.png\n", f);
        std::fclose(f);
        const std::vector<revo::Association> a = revo::readAssociations(path);
        if (a.size() != 2 || a[1].rgbFile != "rgb/b.png" || a[0].depthTs != 1.25 || a[1].depthFile != "depth/b.png") return 64;
        if (revo::readAssociations(path, 1).size() != 1) return 65;
        std::remove(path);
    }
    // long sequences: > 1000 keyframe switches chain float32 world poses; the motion-model guess handed to the tracker
    // must stay inside the tracker's orthogonality gate (|R R^T - I|_F < 1e-5) and the trajectory line must keep printing
    {
        struct TurnTracker : FakeTracker {
            float worst = 0.f;
            int trackFrames(revo::Mat3f &R, revo::Vec3f &T, float &error, const std::shared_ptr<FakePyr> &, const std::shared_ptr<FakePyr> &) {
                float e = 0.f;
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        float s = 0.f;
                        for (int k = 0; k < 3; ++k) s += R(i, k) * R(j, k);
                        e += (s - (i == j)) * (s - (i == j));
                    }
                worst = std::max(worst, std::sqrt(e));
                // 1.7 degrees about a skew axis and 1 cm per frame, rounded to float like a tracker result
                const float q[4] = {0.008f, 0.011f, -0.006f, 0.99989f};
                float r9[9];
                revo_quat_to_R9(q, r9);
                std::memcpy(R.m, r9, sizeof(r9));
                T = revo::Vec3f{{0.01f, -0.004f, 0.002f}}; error = 0.1f;
                return revo::STATE_OK;
            }
        };
        auto turn = std::make_shared<TurnTracker>();
        revo::REVOLoopT<FakePyr, TurnTracker> longrun(turn);
        for (int i = 0; i < 2600; ++i) longrun.processFrame(std::make_shared<FakePyr>(0.033 * i));
        if (longrun.nKeyFrames < 1200) return 66;
        if (!(turn->worst < 2e-6f)) { std::printf("orthogonality of the initial guess drifted to %g\n", (double)turn->worst); return 67; }
        if (revo::poseToTUMString(longrun.trajectory().back(), 1.0).size() < 40) return 68;
    }
    // 4x4 inverse
    revo::Mat4f A = revo::Mat4f::Identity();
    A(0, 0) = 0.f; A(0, 1) = -1.f; A(1, 0) = 1.f; A(1, 1) = 0.f; A(0, 3) = 1.f; A(1, 3) = 2.f; A(2, 3) = 3.f;
    const revo::Mat4f P = A * A.inverseRigid();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (std::fabs(P(i, j) - (i == j ? 1.f : 0.f)) > 1e-6f) return 61;
    return 0;
}

static int selftest()
{
    ImgPyramidSettings s;
    s.PYR_MIN_LVL = 3; s.fx = 517.3f; s.fy = 516.5f; s.cx = 318.6f; s.cy = 255.3f;
    if (s.nLevels() != 4) return 1;
    CameraPyr cp(s);
    if (cp.size() != 5 || cp.at(3).width != 80 || cp.at(3).height != 60) return 2;
    if (std::fabs(cp.at(2).fx - 517.3f / 4) > 1e-4f) return 3;
    OptimizerSettings o;
    if (o.USE_EDGE_FILTER || o.maxItsPerLvl[0] != 100 || o.edgeDistanceLvl[0] != 30.f || o.huber_edge != 0.3f) return 4;
    TrackerSettings t;
    if (!t.optimizerSettings.USE_EDGE_FILTER || !t.CHECK_INIT_VALUES) return 5;
    revo_opt_config c = t.optimizerSettings.c_config();
    if (c.use_edge_filter != 1 || c.convergence_eps[2] != 0.999f) return 6;
    // pose helpers of the C ABI (host arithmetic): quaternion <-> column-major rotation, Sophus-style error instead of abort()
    const float q[4] = {0.1f, -0.2f, 0.3f, 0.9f};
    float R[9], q2[4];
    if (revo_quat_to_R9(q, R) != REVO_OK || revo_R9_to_quat(R, q2) != REVO_OK) return 7;
    const float n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i)
        if (std::fabs(q2[i] - q[i] / n) > 1e-6f) return 8;
    R[0] *= 1.01f;
    if (revo_R9_to_quat(R, q2) != REVO_ERR_NOT_ORTHOGONAL) return 9;
    revo::SE3f id = revo::SE3f::Identity();
    if (id.q[3] != 1.f || id.t[0] != 0.f) return 10;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 2 && !std::strcmp(argv[1], "--selftest")) {
        int rc = selftest();
        if (!rc) rc = selftest_main_loop(argv[0]);
        if (rc) { std::printf("selftest failed: %d\n", rc); return rc; }
        try {
            revo::Context ctx(0);
            std::printf("selftest ok (device present)\n");
        } catch (const revo::Error &e) {
            if (e.code != REVO_ERR_NO_DEVICE) { std::printf("unexpected error %d: %s\n", e.code, e.what()); return 20; }
            std::printf("selftest ok (no device: %s)\n", e.what());
        }
        return 0;
    }
    if (argc < 3) { std::printf("usage: test_host --selftest | <fixture.bin> <out.txt>\n"); return 2; }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int hdr[4];   // w, h, n_levels, reserved
    float cam[4];
    if (std::fread(hdr, sizeof(int), 4, f) != 4 || std::fread(cam, sizeof(float), 4, f) != 4) return 4;
    const int w = hdr[0], h = hdr[1];
    std::vector<uint8_t> bgr[2];
    std::vector<float> depth[2];
    for (int i = 0; i < 2; ++i) {
        bgr[i].resize((size_t)w * h * 3);
        depth[i].resize((size_t)w * h);
        if (std::fread(bgr[i].data(), 1, bgr[i].size(), f) != bgr[i].size()) return 5;
        if (std::fread(depth[i].data(), sizeof(float), depth[i].size(), f) != depth[i].size()) return 6;
    }
    std::fclose(f);
    try {
        auto ctx = std::make_shared<revo::Context>(0);
        ImgPyramidSettings s;
        s.PYR_MIN_LVL = hdr[2] - 1; s.width = w; s.height = h; s.fx = cam[0]; s.fy = cam[1]; s.cx = cam[2]; s.cy = cam[3];
        auto camPyr = std::make_shared<CameraPyr>(s);
        auto kf = std::make_shared<ImgPyramidRGBD>(ctx, s, camPyr, bgr[0].data(), 0, 3, depth[0].data(), 0, 0.0);
        auto cur = std::make_shared<ImgPyramidRGBD>(ctx, s, camPyr, bgr[1].data(), 0, 3, depth[1].data(), 0, 1.0 / 30);
        bool threw = false;
        try { cur->returnOptimizationStructure(0); } catch (const revo::Error &e) { threw = e.code == REVO_ERR_NOT_KEYFRAME; }
        kf->makeKeyframe();
        TrackerNew tracker(ctx, TrackerSettings(), s);
        revo::Mat3f R = revo::Mat3f::Identity();
        revo::Vec3f T = revo::Vec3f::Zero();
        float error = 0.f;
        TrackerNew::TrackerStatus st = tracker.trackFrames(R, T, error, kf, cur);
        FILE *o = std::fopen(argv[2], "w");
        std::fprintf(o, "%d %d %.9g %d %d\n", (int)st, threw ? 1 : 0, error, cur->return3DEdgesCount(0), (int)kf->returnEdges(0).data.size());
        for (int i = 0; i < 9; ++i) std::fprintf(o, "%.9g ", R.m[i]);
        std::fprintf(o, "\n%.9g %.9g %.9g\n", T[0], T[1], T[2]);
        for (int l = 0; l < 6; ++l) std::fprintf(o, "%d ", tracker.lastResult.n_evals[l]);
        std::fprintf(o, "\n");
        std::fclose(o);
    } catch (const revo::Error &e) {
        std::printf("error %d: %s\n", e.code, e.what());
        return 10;
    }
    return 0;
}
