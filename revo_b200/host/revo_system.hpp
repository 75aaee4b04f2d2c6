// revo_system.hpp -- C++ mirror of the per-frame body of REVO::start (system/system.cpp:128-283) and of REVO::Pose
// (system/system.h:89-150) over the classes of revo_host.hpp: motion-model initialisation, the tracking-quality vote,
// "take the previous frame as keyframe and track again", pose-graph bookkeeping.  Header only.  It is the caller of the
// hot path, not part of it: plain host logic, templated on the pyramid / tracker types so that the CPU test
// (tests/cpp/test_host.cpp --selftest) can run it over fakes; `revo::REVOLoop` is the instantiation over the CUDA classes.
// The same logic in Python: revo_b200/system.py.  IO, viewer, logging and pose output of the reference loop are out of scope.
//
// Required of PyrT:      void makeKeyframe(); void setTwf(const float T[16]); const float *getTransKFtoWorld() const;
//                        double returnTimestamp() const; int frameId;
// Required of TrackerT:  TrackerStatus-like int trackFrames(revo::Mat3f &R, revo::Vec3f &T, float &error, ref, cur);
//                        int assessTrackingQuality(const float *pose16, cur); void addOldPclAndPose(cur, const float *pose16, double ts);
//                        void clearUpPastLists();
#pragma once
#include <cstdio>
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "revo_host.hpp"

namespace revo {

// Column-major 4x4, exactly Eigen::Matrix4f's storage.
struct Mat4f {
    float m[16];
    static Mat4f Identity() { Mat4f r{}; r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.f; return r; }
    static Mat4f fromPtr(const float *p) { Mat4f r; std::memcpy(r.m, p, sizeof(r.m)); return r; }
    // transformFromRT (utils/...: [R T; 0 1])
    static Mat4f fromRT(const Mat3f &R, const Vec3f &T) {
        Mat4f r = Identity();
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) r(i, j) = R(i, j);
        for (int i = 0; i < 3; ++i) r(i, 3) = T[i];
        return r;
    }
    float &operator()(int i, int j) { return m[j * 4 + i]; }
    float operator()(int i, int j) const { return m[j * 4 + i]; }
    const float *data() const { return m; }
    Mat3f rotation() const { Mat3f R; for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) R(i, j) = (*this)(i, j); return R; }
    Vec3f translation() const { return Vec3f{{(*this)(0, 3), (*this)(1, 3), (*this)(2, 3)}}; }
    friend Mat4f operator*(const Mat4f &a, const Mat4f &b) {
        Mat4f c{};
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) {
                float s = 0.f;
                for (int k = 0; k < 4; ++k) s += a(i, k) * b(k, j);
                c(i, j) = s;
            }
        return c;
    }
    // General 4x4 inverse like the reference's Eigen::Matrix4f::inverse() (system.h:136-139), computed in double and rounded once
    // (the same as revo_b200/system.py).  NOT the transpose shortcut of a rigid transform: world poses are float32 products
    // chained over every keyframe and drift off orthogonality by E; with the true inverse E cancels in T_NM1_N = T_N_W * T_W_N,
    // with the transpose it would double and eventually trip the tracker's orthogonality gate on long sequences.
    Mat4f inverse() const {
        double a[4][8];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) { a[i][j] = (*this)(i, j); a[i][4 + j] = i == j ? 1.0 : 0.0; }
        for (int k = 0; k < 4; ++k) {
            int piv = k;
            for (int i = k + 1; i < 4; ++i)
                if (std::fabs(a[i][k]) > std::fabs(a[piv][k])) piv = i;
            if (piv != k)
                for (int j = 0; j < 8; ++j) std::swap(a[piv][j], a[k][j]);
            const double d = 1.0 / a[k][k];
            for (int j = 0; j < 8; ++j) a[k][j] *= d;
            for (int i = 0; i < 4; ++i) {
                if (i == k) continue;
                const double f = a[i][k];
                if (f != 0.0)
                    for (int j = 0; j < 8; ++j) a[i][j] -= f * a[k][j];
            }
        }
        Mat4f r;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r(i, j) = (float)a[i][4 + j];
        return r;
    }
    Mat4f inverseRigid() const { return inverse(); }   // old name
};

// Eigen::Quaternionf(R) (what REVO::writePose prints, system.cpp:75-79): the trace / largest-diagonal rule, no orthogonality
// requirement and no normalisation -- it never fails, unlike revo_R9_to_quat's Sophus-style gate.
inline void quaternionFromMatrix(const Mat3f &R, float q_xyzw[4]) {
    float t = R(0, 0) + R(1, 1) + R(2, 2);
    if (t > 0.f) {
        t = std::sqrt(t + 1.0f);
        q_xyzw[3] = 0.5f * t;
        t = 0.5f / t;
        q_xyzw[0] = (R(2, 1) - R(1, 2)) * t;
        q_xyzw[1] = (R(0, 2) - R(2, 0)) * t;
        q_xyzw[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
        int i = 0;
        if (R(1, 1) > R(0, 0)) i = 1;
        if (R(2, 2) > R(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0f);
        q_xyzw[i] = 0.5f * t;
        t = 0.5f / t;
        q_xyzw[3] = (R(k, j) - R(j, k)) * t;
        q_xyzw[j] = (R(j, i) + R(i, j)) * t;
        q_xyzw[k] = (R(k, i) + R(i, k)) * t;
    }
}

enum { STATE_OK = 0, STATE_LOST = 1, STATE_NEW_KF = 2, STATE_UNKNOWN = 3 };   // TrackerNew::TrackerStatus, tracker.h:60-65

// REVO::Pose (system/system.h:89-150): pose of a frame relative to its parent keyframe.
template <class PyrT>
class PoseT {
public:
    PoseT(const Mat4f &T_kf_curr, double timestamp, const std::shared_ptr<PyrT> &kfFrame)
        : T_kf_curr_(T_kf_curr), timestamp_(timestamp), kfFrame_(kfFrame) {}
    Mat4f getCurrToWorld() const { return Mat4f::fromPtr(kfFrame_->getTransKFtoWorld()) * T_kf_curr_; }   // T_W_KF * T_KF_CURR (:131-134)
    Mat4f T_W_N() const { return getCurrToWorld(); }
    Mat4f T_N_W() const { return getCurrToWorld().inverse(); }
    const Mat4f &T_kf_N() const { return T_kf_curr_; }
    // only called when the "previous frame" becomes keyframe (system.h:140-146)
    void setKfFrame(const std::shared_ptr<PyrT> &kfFrame) { kfFrame_ = kfFrame; T_kf_curr_ = Mat4f::Identity(); }
    double returnTimestamp() const { return timestamp_; }
    const std::shared_ptr<PyrT> &kfFrame() const { return kfFrame_; }

private:
    Mat4f T_kf_curr_;
    double timestamp_;
    std::shared_ptr<PyrT> kfFrame_;
};

// The tracking part of REVO::start for one stream: feed pyramids in order with processFrame().
template <class PyrT, class TrackerT>
class REVOLoopT {
public:
    explicit REVOLoopT(const std::shared_ptr<TrackerT> &tracker) : mTracker(tracker) {}

    // One iteration of the while loop (system.cpp:147-275).  Returns the frame's pose in the world.
    Mat4f processFrame(const std::shared_ptr<PyrT> &currPyr) {
        TrackerT &trk = *mTracker;
        currPyr->frameId = noFrames;
        if (noFrames == 0) {                                    // first frame -> keyframe (system.cpp:151-175)
            kfPyr = prevPyr = currPyr;
            currPyr->makeKeyframe();
            const Mat4f I = Mat4f::Identity();
            currPyr->setTwf(I.data());
            mPoseGraph.emplace_back(I, currPyr->returnTimestamp(), currPyr);
            ++nKeyFrames; ++noFrames;
            justAddedNewKeyframe = true;
            trk.addOldPclAndPose(currPyr, I.data(), currPyr->returnTimestamp());
            return I;
        }
        ++noFrames;
        Mat3f r = R; Vec3f t = T;
        (void)trk.trackFrames(r, t, error, kfPyr, currPyr);                                            // :188
        Mat4f T_KF_N = Mat4f::fromRT(r, t);
        Mat4f currPoseInWorld = Mat4f::fromPtr(kfPyr->getTransKFtoWorld()) * T_KF_N;                  // :192
        int status = (int)trk.assessTrackingQuality(currPoseInWorld.data(), currPyr);                 // :199
        if (status == STATE_NEW_KF && !justAddedNewKeyframe) {
            // tracking gets inaccurate: take the previous frame as keyframe and optimise again (:203-239)
            kfPyr = prevPyr;
            const Mat4f T_w_prev = mPoseGraph.back().getCurrToWorld();
            kfPyr->setTwf(T_w_prev.data());
            kfPyr->makeKeyframe();
            mPoseGraph.back().setKfFrame(kfPyr);
            ++nKeyFrames;
            trk.clearUpPastLists();
            r = T_NM1_N.rotation(); t = T_NM1_N.translation();
            (void)trk.trackFrames(r, t, error, kfPyr, currPyr);                                        // :225
            T_KF_N = Mat4f::fromRT(r, t);
            currPoseInWorld = Mat4f::fromPtr(kfPyr->getTransKFtoWorld()) * T_KF_N;
            status = (int)trk.assessTrackingQuality(currPoseInWorld.data(), currPyr);
            justAddedNewKeyframe = true;
            retracked.push_back(currPyr->frameId);
        } else {
            justAddedNewKeyframe = false;
        }
        trackerStatus = status;
        // add the frame to the pose graph, remember its edge cloud for the vote (:253-254)
        mPoseGraph.emplace_back(T_KF_N, currPyr->returnTimestamp(), kfPyr);
        trk.addOldPclAndPose(currPyr, currPoseInWorld.data(), currPyr->returnTimestamp());
        // relative motion N-1 -> N and the constant-velocity guess for the next frame (:262-271)
        const size_t n = mPoseGraph.size();
        T_NM1_N = mPoseGraph[n - 2].T_N_W() * mPoseGraph[n - 1].T_W_N();
        const Mat4f T_init = mPoseGraph[n - 1].T_kf_N() * T_NM1_N;
        R = T_init.rotation(); T = T_init.translation();
        prevPyr = currPyr;
        return mPoseGraph.back().getCurrToWorld();
    }

    std::vector<Mat4f> trajectory() const {
        std::vector<Mat4f> out;
        for (const auto &p : mPoseGraph) out.push_back(p.getCurrToWorld());
        return out;
    }

    std::shared_ptr<TrackerT> mTracker;
    std::vector<PoseT<PyrT>> mPoseGraph;
    std::shared_ptr<PyrT> kfPyr, prevPyr;
    int noFrames = 0, nKeyFrames = 0, trackerStatus = STATE_UNKNOWN;
    bool justAddedNewKeyframe = false;
    Mat3f R = Mat3f::Identity();          // initial guess of the next frame relative to the keyframe
    Vec3f T = Vec3f::Zero();
    Mat4f T_NM1_N = Mat4f::Identity();
    float error = 0.f;
    std::vector<int> retracked;           // frame ids at which the previous frame was promoted and tracking repeated
};

// ---- dataset wire formats ------------------------------------------------------------------------
// One line of a TUM association file, "rgb_ts rgb_file depth_ts depth_file" (io/iowrapperRGBD.cpp:301-333).
struct Association {
    double rgbTs, depthTs;
    std::string rgbFile, depthFile;
};
// '#' comments and empty lines are ignored; the first skipFirstN data lines are skipped (SKIP_FIRST_N_FRAMES).
inline std::vector<Association> readAssociations(const std::string &path, int skipFirstN = 0) {
    std::vector<Association> out;
    std::ifstream f(path);
    if (!f) throw Error(REVO_ERR_INVALID_ARG, "cannot open " + path);
    std::string line;
    int n = 0;
    while (std::getline(f, line)) {
        const size_t b = line.find_first_not_of(" \t\r");
        if (b == std::string::npos || line[b] == '#') continue;
        if (++n <= skipFirstN) continue;
        std::istringstream is(line);
        Association a;
        if (!(is >> a.rgbTs >> a.rgbFile >> a.depthTs >> a.depthFile)) throw Error(REVO_ERR_INVALID_ARG, "malformed association line: " + line);
        out.push_back(a);
    }
    return out;
}
// "timestamp tx ty tz qx qy qz qw" as REVO::writePose formats it (system/system.cpp:75-79: std::fixed, 6 decimals for the
// timestamp, setprecision(9) afterwards; quaternion = Eigen::Quaternionf(R)).
inline std::string poseToTUMString(const Mat4f &T_w_c, double timestamp) {
    float q[4];
    quaternionFromMatrix(T_w_c.rotation(), q);
    char buf[256];
    std::snprintf(buf, sizeof(buf), "%.6f %.9f %.9f %.9f %.9f %.9f %.9f %.9f", timestamp, (double)T_w_c(0, 3), (double)T_w_c(1, 3),
                  (double)T_w_c(2, 3), (double)q[0], (double)q[1], (double)q[2], (double)q[3]);
    return buf;
}

}  // namespace revo

// the instantiation over the CUDA classes of revo_host.hpp
using REVOLoop = revo::REVOLoopT<ImgPyramidRGBD, TrackerNew>;
