// Drop-in for CML::Optimization::DSOBundleAdjustment (reference src/cml/optimization/dso/DSOBundleAdjustment.h:26-101, "BA.h") on top of
// libcmlba's C ABI (include/cmlba.h).  It binds CML's Frame / MapPoint graph to the handle: the graph -> handle edge is addNewFrame /
// addPoints / the cameras passed to run(), the handle -> graph edge is the scatter at the end of run() (BA:934-940, 966-982).
// A CML maintainer drops this file next to the reference class, changes the type of Hybrid::mPhotometricBA (slam/modslam/Hybrid.h) and
// links cmlba.  Compiled and executed here by oracle/Makefile (`make adapter`) + oracle/adapter_check.cpp, which run the same synthetic Map
// graph through the reference class and through this one (tests/test_gpu_adapter.py).
//
// Same public member names, argument meaning and error behaviour as the reference class: run() returns false where the reference does;
// anything libcmlba reports as an error (state, CUDA) throws, like the reference's assertThrow.  One caller at a time.
#ifndef CML_DSOBUNDLEADJUSTMENT_B200_H
#define CML_DSOBUNDLEADJUSTMENT_B200_H

#include <cml/config.h>
#include <cml/base/AbstractFunction.h>
#include <cml/map/Map.h>
#include <cml/capture/CaptureImage.h>

#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "cmlba.h"

namespace CML::Optimization {

class DSOBundleAdjustmentB200 : public AbstractFunction {
public:
    explicit DSOBundleAdjustmentB200(Ptr<AbstractFunction, NonNullable> parent, int device = 0) : AbstractFunction(parent), mDevice(device) {}

    ~DSOBundleAdjustmentB200() override { if (mH) cmlba_destroy(mH); }

    std::string getName() final { return "DSO Bundle Adjustment (B200)"; }

    // BA.h:28 createResidual(frame, point): libcmlba creates the residual of a point towards every window frame itself (cmlba_add_points /
    // cmlba_add_frame, BA:399-403, 455-461); kept for source compatibility.
    void createResidual(PFrame, PPoint) {}

    void addPoints(const PointSet &points) {                                                  // BA:382-415
        std::vector<int64_t> id, host; std::vector<float> xy; std::vector<double> idepth;
        for (auto p : points) {
            if (mPoints.count((int64_t) p->getId())) continue;                               // BA:386-388
            id.push_back((int64_t) p->getId()); host.push_back((int64_t) p->getReferenceFrame()->getId());
            DistortedVector2d c = p->getReferenceCorner().point(0);
            xy.push_back((float) c.x()); xy.push_back((float) c.y());
            idepth.push_back(p->getReferenceInverseDepth());
            mPoints.emplace((int64_t) p->getId(), p);
            p->setGroup(ACTIVEPOINT, true);
        }
        if (!id.empty()) check(cmlba_add_points(handle(), (int) id.size(), id.data(), host.data(), xy.data(), idepth.data()));
    }

    void addNewFrame(PFrame frame, int immatureGroup) {                                       // BA:417-462
        if (!mFrames.empty()) {                                                              // flagFramesForMarginalization, BA:428, 603-708
            std::vector<double> cams(12 * mFrames.size()); std::vector<int32_t> imm(mFrames.size());
            for (size_t i = 0; i < mFrames.size(); i++) { pack(mFrames[i]->getCamera(), &cams[12 * i]); imm[i] = (int32_t) mFrames[i]->getReferenceGroupMapPoints(immatureGroup).size(); }
            int n = 0;
            check(cmlba_flag_frames_for_marginalization(handle(), cams.data(), imm.data(), nullptr, &n));
        }
        if (!mCalibSet) {                                                                    // BA:419-425
            Vector4 k = frame->getCalibration().getPinhole(0).getParameters();                 // fx fy cx cy
            check(cmlba_set_calib(handle(), k[0], k[1], k[2], k[3], (int) frame->getWidth(0), (int) frame->getHeight(0)));
            mCalibSet = true;
        }
        double w2c[12]; pack(frame->getCamera(), w2c);
        const auto &g = frame->getCaptureFrame().getDerivativeImage(0);                       // AoS (I, dx, dy) floats, image/Array2D.h
        Vector2 ab = frame->getExposure().getParameters();
        static_assert(sizeof(g.data()[0]) == 3 * sizeof(float), "GradientImage texel is three floats");
        check(cmlba_add_frame(handle(), (int64_t) frame->getId(), w2c, ab[0], ab[1], frame->getCaptureFrame().getExposureTime(),
                              reinterpret_cast<const float *>(g.data()), frame->isGroup(getMap().INITFRAME) ? 1 : 0));
        mFrames.push_back(frame);
        frame->setGroup(ACTIVEKEYFRAME, true);
    }

    void marginalizeFrame(PFrame frame) {                                                     // BA:464-601: one flagged frame leaves the window
        for (auto it = mFrames.begin(); it != mFrames.end(); ++it) if (*it == frame) { check(cmlba_remove_frame(handle(), (int64_t) frame->getId())); frameLeft(frame); mFrames.erase(it); return; }
    }

    List<PFrame> marginalizeFrames() {                                                        // BA:710-742
        int n = cmlba_num_frames(handle()); std::vector<int64_t> ids(std::max(n, 1));
        check(cmlba_marginalize_frames(handle(), ids.data(), &n));
        List<PFrame> gone;
        for (int i = 0; i < n; i++)
            for (auto it = mFrames.begin(); it != mFrames.end(); ++it)
                if ((int64_t) (*it)->getId() == ids[i]) { frameLeft(*it); gone.push_back(*it); mFrames.erase(it); break; }
        return gone;
    }

    void tryMarginalize() {                                                                   // BA:2240-2363
        int nd = 0, nm = 0;
        check(cmlba_try_marginalize(handle(), &nd, &nm));
        collectOutliers();
        feedStatistics();
    }

    bool run(bool updatePointsOnly = false) {                                                 // BA:744-910
        if (mFrames.empty()) return true;
        std::vector<double> cams(12 * mFrames.size());
        for (size_t i = 0; i < mFrames.size(); i++) pack(mFrames[i]->getCamera(), &cams[12 * i]);        // updateCamera, BA.h:54-60, BA:753-755
        cmlba_run_result res;
        const int rc = cmlba_run(handle(), cams.data(), mNumIterations.i(), updatePointsOnly ? 1 : 0, &res);
        if (rc == CMLBA_ERR_NUMERIC) return false;                                           // the reference's `return false`
        check(rc);
        mLastResult = res;
        // scatter into the graph (BA:934-940, 966-982; DSOPoint.h:107-118)
        const int n = cmlba_num_frames(handle()), np = cmlba_num_points(handle());
        std::vector<int64_t> fid(std::max(n, 1)), pid(std::max(np, 1)); std::vector<double> w2c(12 * std::max(n, 1)), ab(2 * std::max(n, 1)), idp(std::max(np, 1)), unc(std::max(np, 1));
        std::vector<int32_t> good(std::max(np, 1));
        check(cmlba_get_frames(handle(), fid.data(), w2c.data(), ab.data(), nullptr, nullptr, nullptr));
        check(cmlba_get_points(handle(), pid.data(), idp.data(), unc.data(), nullptr, nullptr, nullptr, good.data()));
        {
            LockGuard lg(mLastOptimizedCameraMutex);
            for (int i = 0; i < n; i++) {
                Camera cam = unpack(&w2c[12 * i]);
                mFrames[i]->setCamera(cam);
                mFrames[i]->setExposureParameters(Exposure(mFrames[i]->getExposure().getExposureFromCamera(), ab[2 * i], ab[2 * i + 1]));
                mLastOptimizedCamera[mFrames[i]] = cam;
            }
        }
        mGoodForTracking = PointSet();
        for (int i = 0; i < np; i++) {
            auto p = mPoints.at(pid[i]);
            p->setReferenceInverseDepth(idp[i]); p->setUncertainty(unc[i]);
            if (good[i]) mGoodForTracking.insert(p);
        }
        mOutliers = PointSet();
        collectOutliers();
        feedStatistics();
        return true;
    }

    void marginalizePointsF() {                                                               // BA:2466-2513
        int n = cmlba_num_points(handle()); std::vector<int64_t> ids(std::max(n, 1));
        check(cmlba_marginalize_points(handle(), ids.data(), &n));
        for (int i = 0; i < n; i++) { auto p = mPoints.at(ids[i]); p->setMarginalized(true); p->setGroup(ACTIVEPOINT, false); mGoodForTracking.erase(p); mPoints.erase(ids[i]); }
    }

    void computeNullspaces() {}                                                               // BA:2365-2417: recomputed inside every cmlba_run

    const PointSet &getOutliers() { return mOutliers; }                                       // BA.h:50

    void updateCamera(PFrame) {}                                                              // BA.h:54-60: run() reads every frame's camera itself

    Camera getLastOptimizedCamera(PFrame frame) {                                             // BA.h:62-69
        LockGuard lg(mLastOptimizedCameraMutex);
        if (mLastOptimizedCamera.count(frame) == 0) return frame->getCamera();
        return mLastOptimizedCamera[frame];
    }

    FrameHashMap<Camera> getLastOptimizedCameras() {                                          // BA.h:71-74
        LockGuard lg(mLastOptimizedCameraMutex);
        return mLastOptimizedCamera;
    }

    PointSet getGoodPointsForTracking() { return mGoodForTracking; }                          // BA.h:76-85 (lastResidual(0) is IN)

    void setNumIterations(int it) { mNumIterations.set(it); }                                 // BA.h:87
    void setNumFrames(int n) { maxFrames.set(n); if (mH) throw std::runtime_error("setNumFrames after the first frame: the window size is fixed when the handle is created"); }
    void setMixedBundleAdjustment(bool b) { if (b) throw std::runtime_error("mixed (indirect + direct) bundle adjustment is not supported by libcmlba"); }
    const PointSet &indirectOptimizedPoints() { return mNoPoints; }                           // BA.h:99

    // DSOContext surface the callers of the class use (DSOContext.h:94-172)
    void removePoint(PPoint p, bool = false) { if (mPoints.erase((int64_t) p->getId())) { check(cmlba_remove_point(handle(), (int64_t) p->getId())); p->setGroup(ACTIVEPOINT, false); mGoodForTracking.erase(p); } }
    const List<PFrame> &getFrames() const { return mFrames; }
    bool have(PPoint p) const { return mPoints.count((int64_t) p->getId()) != 0; }
    size_t numPoints() const { return mPoints.size(); }
    const cmlba_run_result &lastResult() const { return mLastResult; }
    cmlba_handle *handle() {                                                                  // created on first use so that the Parameters can still be set
        if (!mH) {
            cmlba_config cfg; check0(cmlba_default_config(&cfg));
            cfg.iterations = mNumIterations.i(); cfg.max_frames = maxFrames.i();
            cfg.optimize_light_a = mOptimizeA.b(); cfg.optimize_light_b = mOptimizeB.b();
            cfg.force_accept = mForceAccept.b(); cfg.fix_lambda = mFixLambda.b(); cfg.fixed_lambda = mFixedLambda.f();
            cfg.disable_marginalization = mDisableMarginalization.b();
            if (cmlba_create(&cfg, mDevice, &mH) != CMLBA_OK) throw std::runtime_error(std::string("cmlba_create: ") + cmlba_last_error(nullptr));
        }
        return mH;
    }

    // the reference's Parameter names (BA.h:235-288)
    Parameter mNumIterations = createParameter("iterations", 4);
    Parameter maxFrames = createParameter("maxFrames", 6);
    Parameter mOptimizeA = createParameter("optimizeLightA", true);
    Parameter mOptimizeB = createParameter("optimizeLightB", true);
    Parameter mForceAccept = createParameter("forceAccept", true);
    Parameter mFixLambda = createParameter("fixLambda", true);
    Parameter mFixedLambda = createParameter("fixedLambda", 1e-5f);
    Parameter mDisableMarginalization = createParameter("disableMarginalization", true);

    const int ACTIVEKEYFRAME = getMap().createFrameGroup("DSO Active Key Frame");
    const int ACTIVEPOINT = getMap().createMapPointGroup("DSO Active Point");

private:
    static void pack(const Camera &c, double *o) {
        const Matrix33 &R = c.getRotationMatrix(); const Vector3 &t = c.getTranslation();
        for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) o[r * 3 + k] = R(r, k);
        for (int k = 0; k < 3; k++) o[9 + k] = t[k];
    }
    static Camera unpack(const double *o) {
        Matrix33 R; for (int r = 0; r < 3; r++) for (int k = 0; k < 3; k++) R(r, k) = o[r * 3 + k];
        return Camera(Vector3(o[9], o[10], o[11]), R);
    }
    void check(int rc) { if (rc != CMLBA_OK) throw std::runtime_error(std::string("libcmlba: ") + cmlba_last_error(mH)); }
    static void check0(int rc) { if (rc != CMLBA_OK) throw std::runtime_error("libcmlba: cmlba_default_config failed"); }
    void frameLeft(PFrame f) {                                                                // the points a leaving frame hosts leave with it (CTX:131-172)
        f->setGroup(ACTIVEKEYFRAME, false);
        for (auto it = mPoints.begin(); it != mPoints.end();) if (it->second->getReferenceFrame() == f) { it->second->setGroup(ACTIVEPOINT, false); mGoodForTracking.erase(it->second); it = mPoints.erase(it); } else ++it;
    }
    void collectOutliers() {                                                                  // points that lost all residuals / were dropped (BA:1636-1640, 2342-2346)
        int no = 0; check(cmlba_get_outliers(handle(), nullptr, &no));
        std::vector<int64_t> out(std::max(no, 1)); int cap = no;
        check(cmlba_get_outliers(handle(), out.data(), &cap));
        for (int i = 0; i < no; i++) { auto it = mPoints.find(out[i]); if (it == mPoints.end()) continue; mOutliers.insert(it->second); it->second->setGroup(ACTIVEPOINT, false); mGoodForTracking.erase(it->second); mPoints.erase(it); }
    }
    void feedStatistics() {                                                                   // the 19 series of BA.h:215-233 under their own names
        double v[CMLBA_NUM_STATISTICS];
        if (cmlba_get_statistics(handle(), v) != CMLBA_OK) return;
        if (mStatistics.empty()) for (int i = 0; i < CMLBA_NUM_STATISTICS; i++) mStatistics.push_back(createStatistic(cmlba_statistic_name(i)));
        for (int i = 0; i < CMLBA_NUM_STATISTICS; i++) mStatistics[i]->addValue(v[i]);
    }

    cmlba_handle *mH = nullptr;
    int mDevice = 0;
    bool mCalibSet = false;
    List<PFrame> mFrames;
    std::unordered_map<int64_t, PPoint> mPoints;
    PointSet mOutliers, mGoodForTracking, mNoPoints;
    Mutex mLastOptimizedCameraMutex;
    FrameHashMap<Camera> mLastOptimizedCamera;
    List<PStatistic> mStatistics;
    cmlba_run_result mLastResult = {};
};

}  // namespace CML::Optimization

#endif
