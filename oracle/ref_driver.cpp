// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference implementation (lizabelos/libCML,
// src/cml/optimization/dso/DSOBundleAdjustment.cpp, compiled in place from /root/reference by
// oracle/Makefile) on a synthetic window read from a CMLW file, and dumps golden vectors.
// Nothing here is linked into, imported by, or executed from the product (libcml_b200/, include/);
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs run it.
//
//   cmlba_ref --window in.cmlw --mode stages|run|maintain|track|bench --out out.cmlw [--repeat K]
//
// mode stages : replays DSOBundleAdjustment::run (BA:744-910) call by call through the class's own
//               (protected) methods and dumps every intermediate quantity SURVEY.md section 8(d) lists.
// mode run    : calls the public run() only and dumps the final state (checks that "stages" == run()).
// mode bench  : times linearizeAll / addToHessianTop+stitchDoubleTop / addToHessianSC+stitchDoubleSC
//               and whole run() with std::chrono, prints one JSON line (single thread = reference behaviour).
#include <cml/config.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>
#include <unistd.h>
#include <sstream>
#include <iostream>
#include <fstream>
#include <map>
#include <set>
#include <list>
#include <memory>
#include <mutex>
#include <thread>
#include <atomic>
#include <functional>
#include <algorithm>
#include <random>
#include <optional>
#include <variant>
#include <any>
#include <queue>
#include <deque>
#include <condition_variable>
#include <future>
#include <regex>
#include <iomanip>
#include <Eigen/Dense>
#include <sophus/se3.hpp>

#define private public
#define protected public
#include <cml/base/AbstractFunction.h>
#include <cml/map/Map.h>
#include <cml/capture/CaptureImage.h>
#include <cml/optimization/dso/DSOBundleAdjustment.h>
#include <cml/optimization/dso/DSOTracker.h>
#include <cml/optimization/dso/DSOTracer.h>
#include <cml/features/corner/PixelSelector.h>
#include <cml/features/corner/FAST.h>
#undef private
#undef protected

#include "cmlw_io.h"

// ---- the two link stubs named in SURVEY.md section 8(c) step 3 -------------------------------
namespace CML {
Atomic<size_t> __array2DCounter = 0;
namespace Evaluation {
double align(const List<Camera> &, const List<Optional<Camera>> &, List<Camera> &) { return 0; }
}  // namespace Evaluation
}  // namespace CML

using namespace CML;
using namespace CML::Optimization;

struct Root : public AbstractFunction {
    Root() : AbstractFunction(nullptr) {}
    Map map;
    Map &getMap() override { return map; }
    std::string getName() override { return "root"; }
};

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RefWindow {
    int W, H, N, P;
    Root *root;
    InternalCalibration *calib;
    CaptureImageGenerator *gen;
    DSOBundleAdjustment *ba;
    std::vector<PFrame> frames;
    std::vector<PPoint> points;
    std::unordered_map<MapPoint *, int> pointIndex;
    std::unordered_map<Frame *, int> frameIndex;
    int iterations;
    bool updatePointsOnly;
};

static Camera cameraFromRt(const double *p) {
    Matrix33 R;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R(r, c) = p[r * 3 + c];
    Vector3 t(p[9], p[10], p[11]);
    return Camera(t, R);
}

static RefWindow *buildWindow(const cmlw::File &in) {
    RefWindow *w = new RefWindow;
    const int32_t *size = in.get("size").as<int32_t>();
    w->W = size[0]; w->H = size[1];
    const double *K = in.get("calib").as<double>();
    w->N = (int) in.get("frame_evalpt").dims[0];
    w->P = (int) in.get("pt_host").dims[0];
    w->iterations = in.has("iterations") ? in.get("iterations").as<int32_t>()[0] : 4;
    w->updatePointsOnly = in.has("update_points_only") ? in.get("update_points_only").as<int32_t>()[0] != 0 : false;

    w->root = new Root;
    w->calib = new InternalCalibration(PinholeUndistorter(Vector2(K[0], K[1]), Vector2(K[2], K[3])), Vector2(w->W, w->H));
    w->gen = new CaptureImageGenerator(w->W, w->H, w->N + 2, w->N + 2);
    w->ba = new DSOBundleAdjustment(w->root);
    // maxFrames: N+2 keeps flagFramesForMarginalization quiet (BA:649); the maintenance goldens set it lower
    w->ba->setNumFrames(in.has("max_frames") ? in.get("max_frames").as<int32_t>()[0] : w->N + 2);
    w->ba->setNumIterations(w->iterations);
    if (in.has("optimize_a")) w->ba->mOptimizeA.set(in.get("optimize_a").as<int32_t>()[0] != 0);
    if (in.has("optimize_b")) w->ba->mOptimizeB.set(in.get("optimize_b").as<int32_t>()[0] != 0);
    if (in.has("force_accept")) w->ba->mForceAccept.set(in.get("force_accept").as<int32_t>()[0] != 0);
    if (in.has("disable_marginalization")) w->ba->mDisableMarginalization.set(in.get("disable_marginalization").as<int32_t>()[0] != 0);
    if (in.has("fixed_lambda")) w->ba->mFixedLambda.set((float) in.get("fixed_lambda").as<double>()[0]);

    Map &map = w->root->getMap();
    int immature = map.createMapPointGroup("immature");

    const double *evalpt = in.get("frame_evalpt").as<double>();
    const double *cam = in.get("frame_cam").as<double>();
    const double *aff = in.get("frame_affine").as<double>();
    const double *expo = in.get("frame_exposure").as<double>();
    const float *gray = in.get("gray").as<float>();
    const uint8_t *isInit = in.has("frame_init") ? in.get("frame_init").as<uint8_t>() : nullptr;

    for (int i = 0; i < w->N; i++) {
        FloatImage img(w->W, w->H);
        memcpy(img.data(), gray + (size_t) i * w->W * w->H, sizeof(float) * w->W * w->H);
        auto cap = w->gen->create().setImage(img).setTime(i).setCalibration(w->calib).setExposure(expo[i]).generate();
        PFrame f = map.createFrame(cap);
        f->setCamera(cameraFromRt(evalpt + 12 * i));
        f->setExposureParameters(Exposure(expo[i], aff[2 * i], aff[2 * i + 1]));
        if (isInit && isInit[i]) f->setGroup(map.INITFRAME, true);
        map.addFrame(f);
        w->frames.push_back(f);
        w->frameIndex[f.p()] = i;
    }
    for (int i = 0; i < w->N; i++) w->ba->addNewFrame(w->frames[i], immature);
    // the "current estimate" at run() time (run() starts with updateCamera -> setStateFromCamera, BA:753-755)
    for (int i = 0; i < w->N; i++) w->frames[i]->setCamera(cameraFromRt(cam + 12 * i));

    const int32_t *host = in.get("pt_host").as<int32_t>();
    const float *xy = in.get("pt_xy").as<float>();
    const double *idepth = in.get("pt_idepth").as<double>();
    std::vector<std::vector<int>> perHost(w->N);
    for (int p = 0; p < w->P; p++) perHost[host[p]].push_back(p);
    w->points.resize(w->P, PPoint());
    PointSet set;
    for (int h = 0; h < w->N; h++) {
        if (perHost[h].empty()) continue;
        List<Corner> corners;
        for (int p : perHost[h]) corners.emplace_back(Corner(DistortedVector2d(xy[2 * p], xy[2 * p + 1])));
        int gid = w->frames[h]->addFeaturePoints(corners);
        for (size_t k = 0; k < perHost[h].size(); k++) {
            int p = perHost[h][k];
            PPoint mp = map.createMapPoint(w->frames[h], FeatureIndex(gid, (short) k), DIRECTTYPE);
            mp->setReferenceInverseDepth(idepth[p]);
            w->points[p] = mp;
            w->pointIndex[mp.p()] = p;
            set.insert(mp);
        }
    }
    w->ba->addPoints(set);
    return w;
}

// ---------------------------------------------------------------------------------------------
template <typename M> static void putMat(cmlw::File &out, const std::string &name, const M &m) {
    std::vector<double> v((size_t) m.rows() * m.cols());
    for (int r = 0; r < m.rows(); r++) for (int c = 0; c < m.cols(); c++) v[(size_t) r * m.cols() + c] = (double) m(r, c);
    out.put<double>(name, v, {(uint64_t) m.rows(), (uint64_t) m.cols()});
}

static void dumpFrames(RefWindow *w, cmlw::File &out, const std::string &pre) {
    int N = w->N;
    std::vector<double> state(N * 10), zero(N * 10), scaled(N * 10), evalpt(N * 12), pre_w2c(N * 12), prior(N * 8), delta(N * 8), dprior(N * 8), th(N), step(N * 10);
    std::vector<double> nsp(N * 36), nss(N * 6), nsa(N * 8), camRt(N * 12), aff(N * 2);
    for (int i = 0; i < N; i++) {
        auto d = w->ba->get(w->frames[i]);
        for (int k = 0; k < 10; k++) { state[i * 10 + k] = d->state[k]; zero[i * 10 + k] = d->state_zero[k]; scaled[i * 10 + k] = d->state_scaled[k]; step[i * 10 + k] = d->step[k]; }
        Matrix33 R0 = d->worldToCam_evalPT.rotationMatrix(); Vector3 t0 = d->worldToCam_evalPT.translation();
        Matrix33 R1 = d->PRE_worldToCam.rotationMatrix(); Vector3 t1 = d->PRE_worldToCam.translation();
        Matrix33 Rc = w->frames[i]->getCamera().getRotationMatrix(); Vector3 tc = w->frames[i]->getCamera().getTranslation();
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { evalpt[i * 12 + r * 3 + c] = R0(r, c); pre_w2c[i * 12 + r * 3 + c] = R1(r, c); camRt[i * 12 + r * 3 + c] = Rc(r, c); }
        for (int r = 0; r < 3; r++) { evalpt[i * 12 + 9 + r] = t0[r]; pre_w2c[i * 12 + 9 + r] = t1[r]; camRt[i * 12 + 9 + r] = tc[r]; }
        for (int k = 0; k < 8; k++) { prior[i * 8 + k] = d->prior[k]; delta[i * 8 + k] = d->delta[k]; dprior[i * 8 + k] = d->delta_prior[k]; }
        th[i] = d->frameEnergyTH;
        for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) nsp[i * 36 + r * 6 + c] = d->nullspaces_pose(r, c);
        for (int r = 0; r < 6; r++) nss[i * 6 + r] = d->nullspaces_scale[r];
        for (int r = 0; r < 4; r++) for (int c = 0; c < 2; c++) nsa[i * 8 + r * 2 + c] = d->nullspaces_affine(r, c);
        aff[i * 2 + 0] = w->frames[i]->getExposure().getParameters()(0);
        aff[i * 2 + 1] = w->frames[i]->getExposure().getParameters()(1);
    }
    out.put<double>(pre + "frame_state", state, {(uint64_t) N, 10});
    out.put<double>(pre + "frame_state_zero", zero, {(uint64_t) N, 10});
    out.put<double>(pre + "frame_state_scaled", scaled, {(uint64_t) N, 10});
    out.put<double>(pre + "frame_step", step, {(uint64_t) N, 10});
    out.put<double>(pre + "frame_evalpt", evalpt, {(uint64_t) N, 12});
    out.put<double>(pre + "frame_pre_w2c", pre_w2c, {(uint64_t) N, 12});
    out.put<double>(pre + "frame_cam", camRt, {(uint64_t) N, 12});
    out.put<double>(pre + "frame_affine", aff, {(uint64_t) N, 2});
    out.put<double>(pre + "frame_prior", prior, {(uint64_t) N, 8});
    out.put<double>(pre + "frame_delta", delta, {(uint64_t) N, 8});
    out.put<double>(pre + "frame_delta_prior", dprior, {(uint64_t) N, 8});
    out.put<double>(pre + "frame_energy_th", th, {(uint64_t) N});
    out.put<double>(pre + "frame_ns_pose", nsp, {(uint64_t) N, 6, 6});
    out.put<double>(pre + "frame_ns_scale", nss, {(uint64_t) N, 6});
    out.put<double>(pre + "frame_ns_affine", nsa, {(uint64_t) N, 4, 2});
}

static void dumpAdjoints(RefWindow *w, cmlw::File &out, const std::string &pre) {
    int N = w->N;
    std::vector<double> ah(N * N * 64), at(N * N * 64), dd(N * N * 8);
    for (int i = 0; i < N * N; i++) {
        for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) { ah[i * 64 + r * 8 + c] = w->ba->mAdHost[i](r, c); at[i * 64 + r * 8 + c] = w->ba->mAdTarget[i](r, c); }
        for (int c = 0; c < 8; c++) dd[i * 8 + c] = w->ba->mAdHTdeltaF[i](0, c);
    }
    out.put<double>(pre + "ad_host", ah, {(uint64_t) N * N, 8, 8});
    out.put<double>(pre + "ad_target", at, {(uint64_t) N * N, 8, 8});
    out.put<double>(pre + "ad_ht_delta", dd, {(uint64_t) N * N, 8});
}

// per-residual dump in mActiveResiduals order; `full` adds the raw Jacobian rJ (what linearize just wrote)
static void dumpResiduals(RefWindow *w, cmlw::File &out, const std::string &pre, bool full, bool useEfs) {
    auto &res = w->ba->mActiveResiduals;
    size_t R = res.size();
    std::vector<int32_t> pt(R), tgt(R), host(R), state(R), nstate(R), good(R), lin(R);
    std::vector<double> e(R), ne(R), neo(R), cpt(R * 3);
    std::vector<float> jpjd(R * 8), resF, Jpdxi, Jpdc, Jpdd, JIdx, JabF, JIdx2, JabJIdx, Jab2;
    if (full) { resF.resize(R * 8); Jpdxi.resize(R * 12); Jpdc.resize(R * 8); Jpdd.resize(R * 2); JIdx.resize(R * 16); JabF.resize(R * 16); JIdx2.resize(R * 4); JabJIdx.resize(R * 4); Jab2.resize(R * 4); }
    for (size_t i = 0; i < R; i++) {
        DSOResidual *r = res[i];
        pt[i] = w->pointIndex.at(r->elements.mapPoint.p());
        tgt[i] = w->frameIndex.at(r->elements.frame.p());
        host[i] = w->frameIndex.at(r->elements.mapPoint->getReferenceFrame().p());
        state[i] = (int) r->getState(); nstate[i] = (int) r->getNewState();
        good[i] = r->isActiveAndIsGoodNEW; lin[i] = r->isLinearized;
        e[i] = r->state_energy; ne[i] = r->state_NewEnergy; neo[i] = r->state_NewEnergyWithOutlier;
        for (int k = 0; k < 3; k++) cpt[i * 3 + k] = r->getCenterProjectedTo()[k];
        for (int k = 0; k < 8; k++) jpjd[i * 8 + k] = r->JpJdF[k];
        if (full) {
            const DSORawResidualJacobian &J = useEfs ? r->efsJ : r->rJ;
            for (int k = 0; k < 8; k++) resF[i * 8 + k] = J.resF[k];
            for (int a = 0; a < 2; a++) {
                for (int k = 0; k < 6; k++) Jpdxi[i * 12 + a * 6 + k] = J.Jpdxi[a][k];
                for (int k = 0; k < 4; k++) Jpdc[i * 8 + a * 4 + k] = J.Jpdc[a][k];
                Jpdd[i * 2 + a] = J.Jpdd[a];
                for (int k = 0; k < 8; k++) { JIdx[i * 16 + a * 8 + k] = J.JIdx[a][k]; JabF[i * 16 + a * 8 + k] = J.JabF[a][k]; }
                for (int b = 0; b < 2; b++) { JIdx2[i * 4 + a * 2 + b] = J.JIdx2(a, b); JabJIdx[i * 4 + a * 2 + b] = J.JabJIdx(a, b); Jab2[i * 4 + a * 2 + b] = J.Jab2(a, b); }
            }
        }
    }
    out.put1<int32_t>(pre + "res_point", pt); out.put1<int32_t>(pre + "res_target", tgt); out.put1<int32_t>(pre + "res_host", host);
    out.put1<int32_t>(pre + "res_state", state); out.put1<int32_t>(pre + "res_new_state", nstate);
    out.put1<int32_t>(pre + "res_good", good); out.put1<int32_t>(pre + "res_linearized", lin);
    out.put1<double>(pre + "res_energy", e); out.put1<double>(pre + "res_new_energy", ne); out.put1<double>(pre + "res_new_energy_wo", neo);
    out.put<double>(pre + "res_center", cpt, {(uint64_t) R, 3});
    out.put<float>(pre + "res_JpJdF", jpjd, {(uint64_t) R, 8});
    if (full) {
        out.put<float>(pre + "rJ_resF", resF, {(uint64_t) R, 8});
        out.put<float>(pre + "rJ_Jpdxi", Jpdxi, {(uint64_t) R, 2, 6});
        out.put<float>(pre + "rJ_Jpdc", Jpdc, {(uint64_t) R, 2, 4});
        out.put<float>(pre + "rJ_Jpdd", Jpdd, {(uint64_t) R, 2});
        out.put<float>(pre + "rJ_JIdx", JIdx, {(uint64_t) R, 2, 8});
        out.put<float>(pre + "rJ_JabF", JabF, {(uint64_t) R, 2, 8});
        out.put<float>(pre + "rJ_JIdx2", JIdx2, {(uint64_t) R, 2, 2});
        out.put<float>(pre + "rJ_JabJIdx", JabJIdx, {(uint64_t) R, 2, 2});
        out.put<float>(pre + "rJ_Jab2", Jab2, {(uint64_t) R, 2, 2});
    }
}

static void dumpPoints(RefWindow *w, cmlw::File &out, const std::string &pre) {
    int P = w->P;
    std::vector<double> idepth(P), step(P), unc(P), colors(P * 8), weights(P * 8);
    std::vector<float> hdd(P), bd(P), hcd(P * 4), hdi(P), bdsum(P), deltaF(P), priorF(P), idz(P), idh(P), mrb(P);
    std::vector<int32_t> alive(P), ngood(P), nres(P), last0(P), last1(P), outlier(P);
    for (int p = 0; p < P; p++) {
        PPoint mp = w->points[p];
        alive[p] = w->ba->have(mp) && w->ba->getPoints().count(mp) > 0;
        idepth[p] = mp->getReferenceInverseDepth();
        unc[p] = mp->getUncertainty();
        outlier[p] = w->ba->mOutliers.count(mp) > 0;
        if (!w->ba->have(mp)) continue;
        auto d = w->ba->get(mp);
        step[p] = d->step;
        hdd[p] = d->Hdd_accAF; bd[p] = d->bd_accAF; for (int k = 0; k < 4; k++) hcd[p * 4 + k] = d->Hcd_accAF[k];
        hdi[p] = d->HdiF; bdsum[p] = d->bdSumF; deltaF[p] = d->deltaF; priorF[p] = d->priorF; idz[p] = d->idepth_zero;
        idh[p] = d->getInverseDepthHessian(); mrb[p] = d->getMaxRelBaseline();
        ngood[p] = d->numGoodResiduals; nres[p] = (int) d->getResiduals().size();
        last0[p] = d->getLastResidual(0).first ? (int) d->getLastResidual(0).second : -1;
        last1[p] = d->getLastResidual(1).first ? (int) d->getLastResidual(1).second : -1;
        for (int k = 0; k < 8; k++) { colors[p * 8 + k] = d->colors[k]; weights[p * 8 + k] = d->weights.size() == 8 ? d->weights[k] : 0; }
    }
    out.put1<double>(pre + "pt_idepth", idepth); out.put1<double>(pre + "pt_step", step); out.put1<double>(pre + "pt_uncertainty", unc);
    out.put1<float>(pre + "pt_Hdd", hdd); out.put1<float>(pre + "pt_bd", bd); out.put<float>(pre + "pt_Hcd", hcd, {(uint64_t) P, 4});
    out.put1<float>(pre + "pt_HdiF", hdi); out.put1<float>(pre + "pt_bdSumF", bdsum); out.put1<float>(pre + "pt_deltaF", deltaF);
    out.put1<float>(pre + "pt_priorF", priorF); out.put1<float>(pre + "pt_idepth_zero", idz); out.put1<float>(pre + "pt_idepth_hessian", idh);
    out.put1<float>(pre + "pt_max_rel_baseline", mrb);
    out.put1<int32_t>(pre + "pt_alive", alive); out.put1<int32_t>(pre + "pt_num_good", ngood); out.put1<int32_t>(pre + "pt_num_res", nres);
    out.put1<int32_t>(pre + "pt_last0", last0); out.put1<int32_t>(pre + "pt_last1", last1); out.put1<int32_t>(pre + "pt_outlier", outlier);
    out.put<double>(pre + "pt_colors", colors, {(uint64_t) P, 8}); out.put<double>(pre + "pt_weights", weights, {(uint64_t) P, 8});
}

static void dumpSystem(RefWindow *w, cmlw::File &out, const std::string &pre) {
    int N = w->N;
    auto *ba = w->ba;
    std::vector<float> acc((size_t) N * N * 169); std::vector<int32_t> accNum(N * N);
    for (int i = 0; i < N * N; i++) {
        accNum[i] = (int) ba->mAccumulatorActive[i].num;
        for (int r = 0; r < 13; r++) for (int c = 0; c < 13; c++) acc[(size_t) i * 169 + r * 13 + c] = ba->mAccumulatorActive[i].H(r, c);
    }
    out.put<float>(pre + "acc_active", acc, {(uint64_t) N * N, 13, 13});
    out.put1<int32_t>(pre + "acc_active_num", accNum);
    std::vector<float> accE((size_t) N * N * 32), accEB((size_t) N * N * 8), accD((size_t) N * N * N * 64);
    for (int i = 0; i < N * N; i++) {
        for (int r = 0; r < 8; r++) { for (int c = 0; c < 4; c++) accE[(size_t) i * 32 + r * 4 + c] = ba->mAccE[i].A1m(r, c); accEB[(size_t) i * 8 + r] = ba->mAccEB[i].A1m[r]; }
    }
    for (int i = 0; i < N * N * N; i++) for (int r = 0; r < 8; r++) for (int c = 0; c < 8; c++) accD[(size_t) i * 64 + r * 8 + c] = ba->mAccD[i].A1m(r, c);
    out.put<float>(pre + "acc_E", accE, {(uint64_t) N * N, 8, 4});
    out.put<float>(pre + "acc_EB", accEB, {(uint64_t) N * N, 8});
    out.put<float>(pre + "acc_D", accD, {(uint64_t) N * N * N, 8, 8});
    putMat(out, pre + "HA_top", ba->HA_top); putMat(out, pre + "HL_top", ba->HL_top); putMat(out, pre + "H_sc", ba->H_sc);
    putMat(out, pre + "bA_top", ba->bA_top); putMat(out, pre + "bL_top", ba->bL_top); putMat(out, pre + "b_sc", ba->b_sc); putMat(out, pre + "bM_top", ba->bM_top);
    // x = -(step) for every frame (BA:1439); calibration step is identically 0 when !optimizeCalibration
    std::vector<double> x(8 * N + 4, 0.0);
    for (int k = 0; k < 4; k++) x[k] = -ba->mCalibStep[k];
    for (int i = 0; i < N; i++) { auto d = ba->get(w->frames[i]); for (int k = 0; k < 8; k++) x[4 + 8 * i + k] = -d->step[k]; }
    out.put1<double>(pre + "x", x);
}

// rebuild mActiveResiduals exactly as run() does (BA:764-779)
static void collectActive(RefWindow *w) {
    auto *ba = w->ba;
    ba->mActiveResiduals.clear();
    for (auto r : ba->getResiduals()) {
        if (!r->isLinearized) { ba->mActiveResiduals.push_back(r); r->resetOOB(); }
        else if (ba->mAddLinearizedPoints.b()) { ba->mActiveResiduals.push_back(r); r->resetOOB(); r->isLinearized = false; }
    }
}

static void dumpFinal(RefWindow *w, cmlw::File &out, const std::string &pre, bool ok) {
    out.scalar<int32_t>(pre + "ok", ok ? 1 : 0);
    dumpFrames(w, out, pre);
    dumpPoints(w, out, pre);
    // surviving residuals (point,target) with their committed state
    std::vector<int32_t> pt, tgt, state;
    std::vector<double> e;
    for (auto r : w->ba->getResiduals()) {
        pt.push_back(w->pointIndex.at(r->elements.mapPoint.p())); tgt.push_back(w->frameIndex.at(r->elements.frame.p()));
        state.push_back((int) r->getState()); e.push_back(r->state_energy);
    }
    out.put1<int32_t>(pre + "alive_res_point", pt); out.put1<int32_t>(pre + "alive_res_target", tgt);
    out.put1<int32_t>(pre + "alive_res_state", state); out.put1<double>(pre + "alive_res_energy", e);
    std::vector<int32_t> goodTrack;
    for (auto p : w->ba->getGoodPointsForTracking()) goodTrack.push_back(w->pointIndex.at(p.p()));
    std::sort(goodTrack.begin(), goodTrack.end());
    out.put1<int32_t>(pre + "good_points_for_tracking", goodTrack);
}

static int runStages(RefWindow *w, cmlw::File &out) {
    auto *ba = w->ba;
    // level-0 derivative image and gray of frame 0 as the reference built them (capture/CaptureImage.cpp:209-262)
    {
        std::vector<float> g((size_t) w->N * w->W * w->H * 3);
        for (int i = 0; i < w->N; i++) {
            const GradientImage &im = w->frames[i]->getCaptureFrame().getDerivativeImage(0);
            for (int y = 0; y < w->H; y++) for (int x = 0; x < w->W; x++) {
                Vector3f v = im.get(x, y);
                size_t o = (((size_t) i * w->H + y) * w->W + x) * 3;
                g[o] = v[0]; g[o + 1] = v[1]; g[o + 2] = v[2];
            }
        }
        out.put<float>("grad_images", g, {(uint64_t) w->N, (uint64_t) w->H, (uint64_t) w->W, 3});
    }
    // --- run() prologue (BA:753-790)
    for (auto f : ba->getFrames()) ba->updateCamera(f);
    ba->mOutliers = PointSet();
    collectActive(w);
    ba->computeAdjoints();
    ba->computeDelta();
    dumpFrames(w, out, "pre_");
    dumpAdjoints(w, out, "pre_");
    dumpPoints(w, out, "pre_");

    Vector3 lastEnergy = ba->linearizeAll(false);
    double lastEnergyL = ba->calcLEnergy(), lastEnergyM = ba->calcMEnergy();
    out.scalar<double>("lin0_energy", lastEnergy[0]);
    dumpResiduals(w, out, "lin0_", true, false);
    { std::vector<double> th(w->N); for (int i = 0; i < w->N; i++) th[i] = ba->get(w->frames[i])->frameEnergyTH; out.put1<double>("lin0_frame_energy_th", th); }
    ba->applyActiveRes(true);
    dumpResiduals(w, out, "app0_", false, false);

    int numIterations = ba->mNumIterations.i();
    double lambda = ba->mFixedLambda.f();
    int itDone = 0;
    std::vector<int32_t> accepted;
    for (int it = 0; it < numIterations; it++) {
        std::string s = std::to_string(it);
        ba->backupState();
        if (!ba->solveSystem(it, lambda)) { out.scalar<int32_t>("failed_at", it); return 1; }
        dumpSystem(w, out, "sol" + s + "_");
        dumpPoints(w, out, "sol" + s + "_");
        { // nullspace matrix as orthogonalize() would stack it (BA:1204-1220), for K3 parity
            int n = 8 * w->N + 4; std::vector<double> ns((size_t) 7 * n);
            for (int k = 0; k < 6; k++) for (int r = 0; r < n; r++) ns[(size_t) k * n + r] = ba->mLastNullspaces_pose[k][r];
            for (int r = 0; r < n; r++) ns[(size_t) 6 * n + r] = ba->mLastNullspaces_scale[0][r];
            out.put<double>("sol" + s + "_nullspaces", ns, {7, (uint64_t) n});
        }
        bool canbreak = ba->doStepFromBackup(w->updatePointsOnly);
        out.scalar<int32_t>("step" + s + "_canbreak", canbreak ? 1 : 0);
        dumpFrames(w, out, "step" + s + "_");
        { std::vector<double> id(w->P); for (int p = 0; p < w->P; p++) id[p] = w->points[p]->getReferenceInverseDepth(); out.put1<double>("step" + s + "_pt_idepth", id); }

        Vector3 newEnergy = ba->linearizeAll(false);
        double newEnergyL = ba->calcLEnergy(), newEnergyM = ba->calcMEnergy();
        out.scalar<double>("lin" + std::to_string(it + 1) + "_energy", newEnergy[0]);
        dumpResiduals(w, out, "lin" + std::to_string(it + 1) + "_", it + 1 == numIterations || it == 0, false);
        { std::vector<double> th(w->N); for (int i = 0; i < w->N; i++) th[i] = ba->get(w->frames[i])->frameEnergyTH; out.put1<double>("lin" + std::to_string(it + 1) + "_frame_energy_th", th); }
        double newTotal = newEnergy[0] + newEnergy[1] + newEnergyL + newEnergyM;
        double lastTotal = lastEnergy[0] + lastEnergy[1] + lastEnergyL + lastEnergyM;
        if (!std::isfinite(newTotal)) { out.scalar<int32_t>("failed_at", it); return 1; }
        if (newTotal < lastTotal || ba->mForceAccept.b()) {
            ba->applyActiveRes(true);
            lastEnergy = newEnergy; lastEnergyL = newEnergyL; lastEnergyM = newEnergyM;
            lambda *= 0.25; accepted.push_back(1);
        } else {
            ba->loadSateBackup();
            lastEnergy = ba->linearizeAll(false); lastEnergyL = ba->calcLEnergy(); lastEnergyM = ba->calcMEnergy();
            lambda *= 1e2; accepted.push_back(0);
        }
        itDone = it + 1;
        if (canbreak && it >= 1) break;
    }
    out.scalar<int32_t>("iterations_done", itDone);
    out.put1<int32_t>("accepted", accepted);
    // --- run() epilogue (BA:885-896)
    auto back = ba->get(ba->getFrames().back());
    Vector<10> nz = Vector<10>::Zero();
    nz.segment<2>(6) = back->get_state().segment<2>(6);
    back->setEvalPT(back->PRE_worldToCam, nz, ba->mScaleTranslation.f(), ba->mScaleRotation.f(), ba->mScaleLightA.f(), ba->mScaleLightB.f());
    ba->mDeltaValid = false; ba->mAdjointsValid = false;
    ba->computeAdjoints(); ba->computeDelta();
    // per-residual view of the final linearization BEFORE residuals get deleted inside linearizeAll(true):
    // not reachable from outside, so keep (point,target) of the active list and report survivors afterwards.
    std::vector<int32_t> apt, atg;
    for (auto r : ba->mActiveResiduals) { apt.push_back(w->pointIndex.at(r->elements.mapPoint.p())); atg.push_back(w->frameIndex.at(r->elements.frame.p())); }
    out.put1<int32_t>("fin_active_point", apt); out.put1<int32_t>("fin_active_target", atg);
    lastEnergy = ba->linearizeAll(true);
    ba->mActiveResiduals.clear();  // dangling after linearizeAll(true) (SURVEY 8c gotcha)
    out.scalar<double>("fin_energy", lastEnergy[0]);
    bool ok = std::isfinite((double) lastEnergy[0]);
    dumpFinal(w, out, "fin_", ok);
    return 0;
}


// ---------------------------------------------------------------------------------------------
// mode maintain: the window-maintenance flow of Hybrid::directMap (slam/modslam/direct/Mapping.cpp:61-100):
// run() x runs, tryMarginalize, removePoint(outliers), computeNullspaces, marginalizePointsF, marginalizeFrames,
// then one more run() on the reduced window.  Dumps every decision (which frames / points go, counters).
static void dumpMaintState(RefWindow *w, cmlw::File &out, const std::string &pre) {
    auto *ba = w->ba;
    int N = w->N, P = w->P;
    std::vector<int32_t> inWin(N, 0), flagged(N, 0), nMarg(N, 0), nOut(N, 0), nRes(N, 0), keyid(N, 0);
    for (auto f : ba->getFrames()) {
        int i = w->frameIndex.at(f.p()); auto d = ba->get(f);
        inWin[i] = 1; flagged[i] = d->flaggedForMarginalization ? 1 : 0; nMarg[i] = (int) d->getNumMarginalized(); nOut[i] = (int) d->getNumResidualsOut();
        nRes[i] = (int) d->getResiduals().size(); keyid[i] = (int) d->keyid;
    }
    out.put1<int32_t>(pre + "frame_in_window", inWin); out.put1<int32_t>(pre + "frame_flagged", flagged); out.put1<int32_t>(pre + "frame_num_marginalized", nMarg);
    out.put1<int32_t>(pre + "frame_num_residuals_out", nOut); out.put1<int32_t>(pre + "frame_num_residuals", nRes); out.put1<int32_t>(pre + "frame_keyid", keyid);
    std::vector<int32_t> alive(P, 0), outl(P, 0), toMarg(P, 0), marg(P, 0), ngood(P, 0), last0(P, -2), last1(P, -2);
    std::vector<float> idh(P, 0.f);
    for (int p = 0; p < P; p++) {
        PPoint mp = w->points[p];
        alive[p] = ba->getPoints().count(mp) > 0;
        outl[p] = ba->mOutliers.count(mp) > 0;
        toMarg[p] = mp->isGroup(ba->DSOTOMARGINALIZE) ? 1 : 0;
        marg[p] = mp->isGroup(ba->DSOMARGINALIZED) ? 1 : 0;
        if (alive[p]) { auto d = ba->get(mp); ngood[p] = d->numGoodResiduals; idh[p] = d->getInverseDepthHessian(); last0[p] = (int) d->getLastResidual(0).second; last1[p] = (int) d->getLastResidual(1).second; }
    }
    out.put1<int32_t>(pre + "pt_alive", alive); out.put1<int32_t>(pre + "pt_outlier", outl); out.put1<int32_t>(pre + "pt_to_marginalize", toMarg);
    out.put1<int32_t>(pre + "pt_marginalized", marg); out.put1<int32_t>(pre + "pt_num_good", ngood); out.put1<float>(pre + "pt_idepth_hessian", idh);
    out.put1<int32_t>(pre + "pt_last0_state", last0); out.put1<int32_t>(pre + "pt_last1_state", last1);
    std::vector<int32_t> rp, rt, rs;
    for (auto r : ba->getResiduals()) { rp.push_back(w->pointIndex.at(r->elements.mapPoint.p())); rt.push_back(w->frameIndex.at(r->elements.frame.p())); rs.push_back((int) r->getState()); }
    out.put1<int32_t>(pre + "res_point", rp); out.put1<int32_t>(pre + "res_target", rt); out.put1<int32_t>(pre + "res_state", rs);
}

static int runMaintain(RefWindow *w, const cmlw::File &in, cmlw::File &out) {
    auto *ba = w->ba;
    const int runs = in.has("runs") ? in.get("runs").as<int32_t>()[0] : 1;
    bool ok = true;
    for (int k = 0; k < runs; k++) { ok = ba->run(w->updatePointsOnly) && ok; ba->mActiveResiduals.clear(); }
    out.scalar<int32_t>("runs_ok", ok ? 1 : 0);
    dumpMaintState(w, out, "m0_");                 // after the runs (flags were set inside addNewFrame)
    dumpFrames(w, out, "m0_"); dumpPoints(w, out, "m0_");
    ba->tryMarginalize();
    dumpMaintState(w, out, "m1_");                 // after tryMarginalize
    { std::vector<PPoint> o(ba->getOutliers().begin(), ba->getOutliers().end()); for (auto pnt : o) ba->removePoint(pnt); }
    ba->computeNullspaces();
    ba->marginalizePointsF();
    dumpMaintState(w, out, "m2_");                 // after marginalizePointsF
    putMat(out, "m2_HM", ba->mMarginalizedHessian); putMat(out, "m2_bM", ba->mMarginalizedB);
    auto removed = ba->marginalizeFrames();
    std::vector<int32_t> rem; for (auto f : removed) rem.push_back(w->frameIndex.at(f.p()));
    out.put1<int32_t>("m3_removed_frames", rem);
    putMat(out, "m3_HM", ba->mMarginalizedHessian); putMat(out, "m3_bM", ba->mMarginalizedB);
    dumpMaintState(w, out, "m3_");                 // after marginalizeFrames
    // one more run() on the reduced window: the reduced window's frames keep their original indices in the dump
    ok = ba->run(w->updatePointsOnly); ba->mActiveResiduals.clear();
    out.scalar<int32_t>("m4_ok", ok ? 1 : 0);
    dumpMaintState(w, out, "m4_");
    {   // final states of the frames that are still in the window + inverse depths
        std::vector<double> pre_w2c(w->N * 12, 0.0), aff(w->N * 2, 0.0), idepth(w->P, 0.0);
        for (auto f : ba->getFrames()) {
            int i = w->frameIndex.at(f.p()); auto d = ba->get(f);
            Matrix33 R1 = d->PRE_worldToCam.rotationMatrix(); Vector3 t1 = d->PRE_worldToCam.translation();
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) pre_w2c[i * 12 + r * 3 + c] = R1(r, c);
            for (int r = 0; r < 3; r++) pre_w2c[i * 12 + 9 + r] = t1[r];
            aff[i * 2] = f->getExposure().getParameters()[0]; aff[i * 2 + 1] = f->getExposure().getParameters()[1];
        }
        for (int p = 0; p < w->P; p++) idepth[p] = w->points[p]->getReferenceInverseDepth();
        out.put<double>("m4_frame_pre_w2c", pre_w2c, {(uint64_t) w->N, 12}); out.put<double>("m4_frame_affine", aff, {(uint64_t) w->N, 2});
        out.put1<double>("m4_pt_idepth", idepth);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// mode track: the coarse-to-fine direct image alignment of DSOTracker (NEXT #1): makeCoarseDepthL0 on the reference
// keyframe (optimization/dso/DSOTracker.cpp:494-725) and optimize() of a new frame against it (:15-246).
// Window keys: track_ref, track_new (frame indices), track_init_cam [12] (initial world->cam guess of the new frame),
// pt_uncertainty [P], track_new_affine [2] (initial a,b of the new frame).  With --repeat K also times optimize().
static int runTrack(RefWindow *w, const cmlw::File &in, cmlw::File &out, int repeat) {
    const int refIdx = in.get("track_ref").as<int32_t>()[0], newIdx = in.get("track_new").as<int32_t>()[0];
    PFrame ref = w->frames[refIdx], nf = w->frames[newIdx];
    DSOTracker tracker(w->root);
    const double *unc = in.get("pt_uncertainty").as<double>();
    PointSet pts;
    for (int p = 0; p < w->P; p++) {
        if (w->frameIndex.at(w->points[p]->getReferenceFrame().p()) == newIdx) continue;      // points hosted in the frame to track do not exist yet
        w->points[p]->setUncertainty(unc[p]);
        pts.insert(w->points[p]);
    }
    tracker.makeCoarseDepthL0(ref, pts);
    DSOTrackerPrivate *cd = tracker.get(ref);
    const int L = nf->getCaptureFrame().getPyramidLevels();
    std::vector<int32_t> wh(2 * L), npc(L);
    std::vector<double> Ks(4 * L);
    for (int l = 0; l < L; l++) {
        wh[2 * l] = nf->getWidth(l); wh[2 * l + 1] = nf->getHeight(l);
        Matrix33 K = nf->getK(l);
        Ks[4 * l] = K(0, 0); Ks[4 * l + 1] = K(1, 1); Ks[4 * l + 2] = K(0, 2); Ks[4 * l + 3] = K(1, 2);
        DSOTrackerPrivateLevel &lv = cd->level[l];
        npc[l] = lv.n();
        std::vector<float> pc((size_t) 4 * lv.n());
        for (int i = 0; i < lv.n(); i++) { pc[4 * i] = lv.PCu(i); pc[4 * i + 1] = lv.PCv(i); pc[4 * i + 2] = lv.PCidepth(i); pc[4 * i + 3] = lv.PCcolor(i); }
        out.put<float>("trk_pc" + std::to_string(l), pc, {(uint64_t) lv.n(), 4});
        // the new frame's derivative image and the reference's gray image of the level (checked against the numpy restatement, then dropped)
        const GradientImage &g = nf->getCaptureFrame().getDerivativeImage(l);
        std::vector<float> gi((size_t) wh[2 * l] * wh[2 * l + 1] * 3);
        for (int y = 0; y < wh[2 * l + 1]; y++) for (int x = 0; x < wh[2 * l]; x++) for (int c = 0; c < 3; c++) gi[((size_t) y * wh[2 * l] + x) * 3 + c] = g.get(x, y)[c];
        out.put<float>("trk_grad" + std::to_string(l), gi, {(uint64_t) wh[2 * l + 1], (uint64_t) wh[2 * l], 3});
    }
    out.put1<int32_t>("trk_levels_wh", wh); out.put1<int32_t>("trk_pc_n", npc); out.put<double>("trk_K", Ks, {(uint64_t) L, 4});
    const double *ic = in.get("track_init_cam").as<double>();
    const double *na = in.get("track_new_affine").as<double>();
    DSOTracker::Residual r;
    Camera cam; Exposure ex(nf->getExposure().getExposureFromCamera(), na[0], na[1]);
    double best = 1e30;
    for (int rep = 0; rep < std::max(1, repeat); rep++) {
        cam = cameraFromRt(ic);
        ex.setParametersAndExposure(Exposure(nf->getExposure().getExposureFromCamera(), na[0], na[1]));
        tracker.mLastResidual = DSOTracker::Residual();
        double a = now_s();
        r = tracker.optimize(0, nf, ref, cam, ex);
        best = std::min(best, now_s() - a);
    }
    out.scalar<double>("trk_seconds", best);
    std::vector<double> cm(12);
    Matrix33 R = cam.getRotationMatrix(); Vector3 t = cam.getTranslation();
    for (int rr = 0; rr < 3; rr++) for (int c = 0; c < 3; c++) cm[rr * 3 + c] = R(rr, c);
    for (int k = 0; k < 3; k++) cm[9 + k] = t[k];
    out.put1<double>("trk_cam", cm);
    std::vector<double> ab = {ex.getParameters()[0], ex.getParameters()[1]};
    out.put1<double>("trk_affine", ab);
    std::vector<double> E(r.E.begin(), r.E.end()), rep_(r.levelCutoffRepeat.begin(), r.levelCutoffRepeat.end());
    std::vector<int32_t> nT(r.numTermsInE.begin(), r.numTermsInE.end()), nS(r.numSaturated.begin(), r.numSaturated.end()), nR(r.numRobust.begin(), r.numRobust.end());
    out.put1<double>("trk_E", E); out.put1<int32_t>("trk_numTermsInE", nT); out.put1<int32_t>("trk_numSaturated", nS); out.put1<int32_t>("trk_numRobust", nR);
    out.put1<double>("trk_levelCutoffRepeat", rep_);
    std::vector<double> fl = {r.flowVector[0], r.flowVector[1], r.flowVector[2]}, ra = {r.relAff[0], r.relAff[1]}, cov(6);
    for (int k = 0; k < 6; k++) cov[k] = r.covariance[k];
    out.put1<double>("trk_flow", fl); out.put1<double>("trk_relAff", ra); out.put1<double>("trk_covariance", cov);
    out.scalar<int32_t>("trk_isCorrect", r.isCorrect ? 1 : 0); out.scalar<int32_t>("trk_tooManySaturated", r.tooManySaturated ? 1 : 0);
    printf("{\"track_seconds\": %.6f, \"levels\": %d, \"points_l0\": %d}\n", best, L, npc[0]);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// mode trace: DSOTracer (SURVEY 8f NEXT #2).  Immature points `im_host` / `im_xy` are created by the reference's own
// makeNewTracesFrom; every later frame f traces all points hosted before f (trace(), DSOTracer.cpp:585-832), the state of
// every point is dumped after each pass; finally optimizeImmaturePoint (DSOTracer.cpp:280-411) on the points with a finite
// depth interval.  Frames = the window's frames at `frame_cam`, exposures `frame_exposure`/`frame_affine`.
static int runTrace(const cmlw::File &in, cmlw::File &out, int repeat) {
    const int32_t *size = in.get("size").as<int32_t>();
    const int W = size[0], H = size[1];
    const double *K = in.get("calib").as<double>();
    const int N = (int) in.get("frame_cam").dims[0];
    const int P = (int) in.get("im_host").dims[0];
    Root *root = new Root;
    auto *calib = new InternalCalibration(PinholeUndistorter(Vector2(K[0], K[1]), Vector2(K[2], K[3])), Vector2(W, H));
    auto *gen = new CaptureImageGenerator(W, H, N + 2, N + 2);
    Map &map = root->getMap();
    DSOTracer tracer(root);
    if (in.has("min_idepth_h_act")) tracer.mMinIDepthHAct.set((float) in.get("min_idepth_h_act").as<double>()[0]);
    const int fg = map.createFrameGroup("trace window");
    const double *cam = in.get("frame_cam").as<double>();
    const double *aff = in.get("frame_affine").as<double>();
    const double *expo = in.get("frame_exposure").as<double>();
    const float *gray = in.get("gray").as<float>();
    std::vector<PFrame> frames;
    for (int i = 0; i < N; i++) {
        FloatImage img(W, H);
        memcpy(img.data(), gray + (size_t) i * W * H, sizeof(float) * W * H);
        auto cap = gen->create().setImage(img).setTime(i).setCalibration(calib).setExposure(expo[i]).generate();
        PFrame f = map.createFrame(cap);
        f->setCamera(cameraFromRt(cam + 12 * i));
        f->setExposureParameters(Exposure(expo[i], aff[2 * i], aff[2 * i + 1]));
        map.addFrame(f);
        f->setGroup(fg, true);
        frames.push_back(f);
    }
    const int32_t *host = in.get("im_host").as<int32_t>();
    const float *xy = in.get("im_xy").as<float>();
    std::vector<PPoint> pts(P, PPoint());
    for (int h = 0; h < N; h++) {
        std::vector<int> mine;
        List<Corner> corners;
        for (int p = 0; p < P; p++) if (host[p] == h) { mine.push_back(p); corners.emplace_back(Corner(DistortedVector2d(xy[2 * p], xy[2 * p + 1]))); }
        if (mine.empty()) continue;
        const int gid = frames[h]->addFeaturePoints(corners);
        tracer.makeNewTracesFrom(frames[h], gid);
        for (auto mp : map.getGroupMapPoints(tracer.IMMATUREPOINT)) {
            if (mp->getReferenceFrame() != frames[h]) continue;
            auto pd = tracer.getPrivateData(mp);
            if (pd->referenceIndex.group != gid) continue;
            pts[mine[pd->referenceIndex.index]] = mp;
        }
    }
    // point initialisation (makeNewTracesFrom): gradH, weights, energyTH
    {
        std::vector<double> gh((size_t) P * 4), wt((size_t) P * 8), eth(P);
        std::vector<int32_t> ok(P);
        for (int p = 0; p < P; p++) {
            ok[p] = pts[p].isNotNull() ? 1 : 0;
            if (!ok[p]) continue;
            auto pd = tracer.getPrivateData(pts[p]);
            gh[4 * p] = pd->gradH(0, 0); gh[4 * p + 1] = pd->gradH(0, 1); gh[4 * p + 2] = pd->gradH(1, 0); gh[4 * p + 3] = pd->gradH(1, 1);
            for (int k = 0; k < 8; k++) wt[8 * p + k] = pd->weights[k];
            eth[p] = pd->energyTH;
        }
        out.put1<int32_t>("trc_created", ok); out.put<double>("trc_gradH", gh, {(uint64_t) P, 4}); out.put<double>("trc_weights", wt, {(uint64_t) P, 8});
        out.put1<double>("trc_energyTH", eth);
    }
    double tTrace = 0; long nTraced = 0;
    for (int f = 1; f < N; f++) {
        double a = now_s();
        for (int p = 0; p < P; p++) if (pts[p].isNotNull() && host[p] < f) { tracer.trace(frames[f], pts[p]); nTraced++; }
        tTrace += now_s() - a;
        std::vector<int32_t> st(P, -1);
        std::vector<double> v((size_t) P * 6, 0.0);
        for (int p = 0; p < P; p++) {
            if (!pts[p].isNotNull()) continue;
            auto pd = tracer.getPrivateData(pts[p]);
            st[p] = (int32_t) pd->lastTraceStatus;
            v[6 * p] = pd->iDepthMin; v[6 * p + 1] = pd->iDepthMax; v[6 * p + 2] = pd->lastTraceUV[0]; v[6 * p + 3] = pd->lastTraceUV[1];
            v[6 * p + 4] = pd->lastTracePixelInterval; v[6 * p + 5] = pd->quality;
        }
        out.put1<int32_t>("trc_status_f" + std::to_string(f), st);
        out.put<double>("trc_state_f" + std::to_string(f), v, {(uint64_t) P, 6});
    }
    if (in.has("act_host")) {
        // activatePoints (DSOTracer.cpp:62-278): `act_*` = the active points of the window (their projections feed the distance map); the iteration order of
        // the immature set (a hash set) is dumped so that the same order can be replayed
        const int pg = map.createMapPointGroup("active");
        const int32_t *ah = in.get("act_host").as<int32_t>();
        const float *axy = in.get("act_xy").as<float>();
        const double *aid = in.get("act_idepth").as<double>();
        const int A = (int) in.get("act_host").dims[0];
        if (in.has("desired_density")) tracer.mSettingsDesiredPointDensity.set(in.get("desired_density").as<int32_t>()[0]);
        for (int h = 0; h < N; h++) {
            List<Corner> corners; std::vector<int> mine;
            for (int p = 0; p < A; p++) if (ah[p] == h) { mine.push_back(p); corners.emplace_back(Corner(DistortedVector2d(axy[2 * p], axy[2 * p + 1]))); }
            if (mine.empty()) continue;
            const int gid = frames[h]->addFeaturePoints(corners);
            for (size_t k = 0; k < mine.size(); k++) {
                PPoint mp = map.createMapPoint(frames[h], FeatureIndex(gid, (short) k), DIRECTTYPE);
                mp->setReferenceInverseDepth(aid[mine[k]]);
                mp->setGroup(pg, true);
            }
        }
        std::unordered_map<MapPoint *, int> index;
        for (int p = 0; p < P; p++) if (pts[p].isNotNull()) index[pts[p].p()] = p;
        std::vector<int32_t> order, status_before(P, -1);
        for (auto mp : map.getGroupMapPoints(tracer.IMMATUREPOINT)) order.push_back(index.at(mp.p()));
        for (int p = 0; p < P; p++) if (pts[p].isNotNull()) status_before[p] = (int32_t) tracer.getPrivateData(pts[p])->lastTraceStatus;
        const double a0 = now_s();
        PointSet mapped = tracer.activatePoints(fg, pg);
        const double tAct = now_s() - a0;
        std::vector<int32_t> isMapped(P, 0), stillImmature(P, 0);
        std::vector<double> idOut(P, 0.0);
        for (auto mp : mapped) { const int p = index.at(mp.p()); isMapped[p] = 1; idOut[p] = mp->getReferenceInverseDepth(); }
        for (auto mp : map.getGroupMapPoints(tracer.IMMATUREPOINT)) stillImmature[index.at(mp.p())] = 1;
        out.put1<int32_t>("act_order", order); out.put1<int32_t>("act_mapped", isMapped); out.put1<int32_t>("act_still_immature", stillImmature);
        out.put1<double>("act_idepth_out", idOut);
        out.scalar<double>("act_min_distance", tracer.mCurrentMinimumDistance); out.scalar<int32_t>("act_urgent", tracer.mUrgentlyNeedNewPoints ? 1 : 0);
        out.scalar<double>("act_seconds", tAct);
        printf("{\"traces\": %ld, \"trace_seconds\": %.6f, \"activate_seconds\": %.6f, \"mapped\": %zu}\n", nTraced, tTrace, tAct, mapped.size());
        return 0;
    }
    // activation: optimizeImmaturePoint on every point with a finite interval
    const int minObs = in.has("min_obs") ? in.get("min_obs").as<int32_t>()[0] : 1;
    std::vector<int32_t> rc(P, -2);
    std::vector<double> idp(P, 0.0);
    double a = now_s(); long nOpt = 0;
    for (int p = 0; p < P; p++) {
        if (!pts[p].isNotNull()) continue;
        auto pd = tracer.getPrivateData(pts[p]);
        if (!std::isfinite(pd->iDepthMax) || !std::isfinite(pd->iDepthMin)) continue;
        rc[p] = tracer.optimizeImmaturePoint(pts[p], minObs, fg); nOpt++;
        if (rc[p] == 1) idp[p] = pts[p]->getReferenceInverseDepth();
    }
    const double tOpt = now_s() - a;
    out.put1<int32_t>("trc_opt_rc", rc); out.put1<double>("trc_opt_idepth", idp);
    out.scalar<double>("trc_trace_seconds", tTrace); out.scalar<double>("trc_opt_seconds", tOpt);
    printf("{\"traces\": %ld, \"trace_seconds\": %.6f, \"optimized\": %ld, \"opt_seconds\": %.6f}\n", nTraced, tTrace, nOpt, tOpt);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// mode prepare: image preparation (SURVEY 8f NEXT #3): CaptureImageGenerator::generate with a response LUT, an inverse vignette and a
// radial-tangential pre-undistorter (CaptureImage.cpp:108-262): dumps the undistortion map and, per level, gray / derivative / weighted
// gradient-norm images.
static int runPrepare(const cmlw::File &in, cmlw::File &out, int repeat) {
    const int32_t *sin = in.get("size_in").as<int32_t>(), *sout = in.get("size_out").as<int32_t>();
    const double *kin = in.get("calib_in").as<double>(), *kout = in.get("calib_out").as<double>(), *rt = in.get("radtan").as<double>();
    const int Wi = sin[0], Hi = sin[1], Wo = sout[0], Ho = sout[1];
    auto *calib = new InternalCalibration(PinholeUndistorter(Vector2(kin[0], kin[1]), Vector2(kin[2], kin[3])), Vector2(Wi, Hi), new RadtanUndistorter(rt[0], rt[1], rt[2], rt[3]),
                                          PinholeUndistorter(Vector2(kout[0], kout[1]), Vector2(kout[2], kout[3])), Vector2(Wo, Ho));
    auto *gen = new CaptureImageGenerator(Wo, Ho, 4, 4);
    GrayLookupTable *lut = nullptr;
    if (in.has("lut")) { Vectorf<256> v; const float *l = in.get("lut").as<float>(); for (int i = 0; i < 256; i++) v[i] = l[i]; lut = new GrayLookupTable(v); }
    FloatImage raw(Wi, Hi);
    memcpy(raw.data(), in.get("raw").as<float>(), sizeof(float) * Wi * Hi);
    Array2D<float> vig(Wi, Hi);
    const bool hasVig = in.has("inv_vignette");
    if (hasVig) memcpy(vig.data(), in.get("inv_vignette").as<float>(), sizeof(float) * Wi * Hi);
    double best = 1e30;
    Ptr<CaptureImage, NonNullable> cap = gen->create().setImage(raw).setTime(0).setCalibration(calib).setExposure(1).generate();
    for (int rep = 0; rep < std::max(1, repeat); rep++) {
        auto mk = gen->create();
        mk.setImage(raw).setTime(rep + 1).setCalibration(calib).setExposure(1);
        if (lut) mk.setLut(lut);
        if (hasVig) mk.setInverseVignette(vig);
        const double a = now_s();
        cap = mk.generate();
        best = std::min(best, now_s() - a);
    }
    {
        std::vector<float> m((size_t) Wo * Ho * 2);
        for (int y = 0; y < Ho; y++) for (int x = 0; x < Wo; x++) { const Vector2f v = calib->mUndistortMap(x, y); m[((size_t) y * Wo + x) * 2] = v[0]; m[((size_t) y * Wo + x) * 2 + 1] = v[1]; }
        out.put<float>("prep_map", m, {(uint64_t) Ho, (uint64_t) Wo, 2});
    }
    const int L = cap->getPyramidLevels();
    std::vector<int32_t> wh(2 * L);
    for (int l = 0; l < L; l++) {
        const int w = cap->getWidth(l), h = cap->getHeight(l);
        wh[2 * l] = w; wh[2 * l + 1] = h;
        std::vector<float> g((size_t) w * h), d((size_t) w * h * 3), n((size_t) w * h);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) {
            g[(size_t) y * w + x] = cap->getGrayImage(l).get(x, y);
            const Vector3f v = cap->getDerivativeImage(l).get(x, y);
            for (int c = 0; c < 3; c++) d[((size_t) y * w + x) * 3 + c] = v[c];
            n[(size_t) y * w + x] = cap->getWeightedGradientNorm(l).get(x, y);
        }
        out.put<float>("prep_gray" + std::to_string(l), g, {(uint64_t) h, (uint64_t) w});
        out.put<float>("prep_grad" + std::to_string(l), d, {(uint64_t) h, (uint64_t) w, 3});
        out.put<float>("prep_wgn" + std::to_string(l), n, {(uint64_t) h, (uint64_t) w});
    }
    out.put1<int32_t>("prep_levels_wh", wh);
    out.scalar<double>("prep_seconds", best);
    printf("{\"prepare_seconds\": %.6f, \"levels\": %d}\n", best, L);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// mode fast: Features::FAST::compute (features/corner/FAST.cpp:3-13: fast9_detect, fast9_score, nonmax_suppression) on the 8-bit image `gray_u8`
// for every threshold in `thresholds`.
static int runFast(const cmlw::File &in, cmlw::File &out, int repeat) {
    const int H = (int) in.get("gray_u8").dims[0], W = (int) in.get("gray_u8").dims[1];
    GrayImage img(W, H);
    memcpy(img.data(), in.get("gray_u8").as<uint8_t>(), (size_t) W * H);
    const int32_t *th = in.get("thresholds").as<int32_t>();
    const int nt = (int) in.get("thresholds").dims[0];
    double best = 1e30;
    for (int k = 0; k < nt; k++) {
        List<Corner> corners;
        for (int rep = 0; rep < std::max(1, repeat); rep++) {
            Features::FAST fast;
            corners.clear();
            const double a = now_s();
            fast.compute(img, corners, th[k]);
            best = std::min(best, now_s() - a);
        }
        std::vector<int32_t> xy(corners.size() * 2), sc(corners.size());
        for (size_t i = 0; i < corners.size(); i++) { xy[2 * i] = (int32_t) corners[i].point(0).x(); xy[2 * i + 1] = (int32_t) corners[i].point(0).y(); sc[i] = (int32_t) corners[i].response(); }
        out.put<int32_t>("fast_xy" + std::to_string(k), xy, {(uint64_t) corners.size(), 2});
        out.put1<int32_t>("fast_score" + std::to_string(k), sc);
    }
    out.scalar<double>("fast_seconds", best);
    printf("{\"fast_seconds\": %.6f}\n", best);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// mode select: DSO pixel selector (SURVEY 8f NEXT #4, PixelSelector part): PixelSelector::compute on the prepared frame `gray`
// (features/corner/PixelSelector.cpp:367-384 -> makeMaps :121-213 -> makeHists :41-118, select :217-365), for every density in `densities`
// on the SAME selector instance (currentPotential carries over, like successive keyframes).
static int runSelect(const cmlw::File &in, cmlw::File &out, int repeat) {
    const int32_t *size = in.get("size").as<int32_t>();
    const int W = size[0], H = size[1];
    const double *K = in.get("calib").as<double>();
    Root *root = new Root;
    auto *calib = new InternalCalibration(PinholeUndistorter(Vector2(K[0], K[1]), Vector2(K[2], K[3])), Vector2(W, H));
    auto *gen = new CaptureImageGenerator(W, H, 4, 4);
    FloatImage img(W, H);
    memcpy(img.data(), in.get("gray").as<float>(), sizeof(float) * W * H);
    auto cap = gen->create().setImage(img).setTime(0).setCalibration(calib).setExposure(1).generate();
    Features::PixelSelector sel(root, W, H);
    const double *dens = in.get("densities").as<double>();
    const int nd = (int) in.get("densities").dims[0];
    double best = 1e30;
    for (int d = 0; d < nd; d++) {
        List<Corner> corners; List<float> types;
        const int potBefore = sel.currentPotential;
        const double a = now_s();
        sel.compute(*cap.p(), corners, types, (float) dens[d]);
        best = std::min(best, now_s() - a);
        std::vector<float> xy(corners.size() * 2), ty(types.begin(), types.end());
        for (size_t i = 0; i < corners.size(); i++) { xy[2 * i] = corners[i].point(0).x(); xy[2 * i + 1] = corners[i].point(0).y(); }
        out.put<float>("sel_xy" + std::to_string(d), xy, {(uint64_t) corners.size(), 2});
        out.put1<float>("sel_type" + std::to_string(d), ty);
        out.scalar<int32_t>("sel_pot_before" + std::to_string(d), potBefore);
        out.scalar<int32_t>("sel_pot_after" + std::to_string(d), sel.currentPotential);
        if (d == 0) {
            const int w32 = W / 32, h32 = H / 32;
            std::vector<float> ths(sel.ths, sel.ths + w32 * h32), sm(sel.thsSmoothed, sel.thsSmoothed + w32 * h32);
            out.put<float>("sel_ths", ths, {(uint64_t) h32, (uint64_t) w32}); out.put<float>("sel_ths_smoothed", sm, {(uint64_t) h32, (uint64_t) w32});
        }
    }
    out.scalar<double>("sel_seconds", best);
    printf("{\"select_seconds\": %.6f}\n", best);
    return 0;
}

int main(int argc, char **argv) {
    std::string window, mode = "stages", outPath;
    int repeat = 3;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--window" && i + 1 < argc) window = argv[++i];
        else if (a == "--mode" && i + 1 < argc) mode = argv[++i];
        else if (a == "--out" && i + 1 < argc) outPath = argv[++i];
        else if (a == "--repeat" && i + 1 < argc) repeat = atoi(argv[++i]);
    }
    if (window.empty()) { fprintf(stderr, "usage: cmlba_ref --window in.cmlw --mode stages|run|bench [--out out.cmlw] [--repeat K]\n"); return 2; }
    initCML();
    cmlw::File in;
    if (!in.load(window)) { fprintf(stderr, "cannot read %s\n", window.c_str()); return 2; }

    int rc = 0;
    if (mode == "fast") {
        cmlw::File out;
        rc = runFast(in, out, repeat);
        if (!outPath.empty()) out.save(outPath);
        fflush(stdout);
        _exit(rc);
    }
    if (mode == "select") {
        cmlw::File out;
        rc = runSelect(in, out, repeat);
        if (!outPath.empty()) out.save(outPath);
        fflush(stdout);
        _exit(rc);
    }
    if (mode == "prepare") {
        cmlw::File out;
        rc = runPrepare(in, out, repeat);
        if (!outPath.empty()) out.save(outPath);
        fflush(stdout);
        _exit(rc);
    }
    if (mode == "trace") {
        cmlw::File out;
        rc = runTrace(in, out, repeat);
        if (!outPath.empty()) out.save(outPath);
        fflush(stdout);
        _exit(rc);
    }
    if (mode == "stages") {
        RefWindow *w = buildWindow(in);
        cmlw::File out;
        rc = runStages(w, out);
        if (!outPath.empty()) out.save(outPath);
    } else if (mode == "run") {
        RefWindow *w = buildWindow(in);
        cmlw::File out;
        double t0 = now_s();
        bool ok = w->ba->run(w->updatePointsOnly);
        double t1 = now_s();
        w->ba->mActiveResiduals.clear();
        out.scalar<double>("run_seconds", t1 - t0);
        // one P-energy sample before the loop (BA:798) + one per ACCEPTED iteration (BA:847)
        out.scalar<int32_t>("accepted_count", (int) w->ba->mStatisticEnergyP->mWaitingValues.size() - 1);
        dumpFinal(w, out, "fin_", ok);
        if (!outPath.empty()) out.save(outPath);
    } else if (mode == "maintain") {
        RefWindow *w = buildWindow(in);
        cmlw::File out;
        rc = runMaintain(w, in, out);
        if (!outPath.empty()) out.save(outPath);
    } else if (mode == "track") {
        RefWindow *w = buildWindow(in);
        cmlw::File out;
        rc = runTrack(w, in, out, repeat);
        if (!outPath.empty()) out.save(outPath);
    } else if (mode == "bench") {
        // min over `repeat` freshly built windows, 1 thread (the reference BA is single-threaded)
        double tLin = 1e30, tTop = 1e30, tSC = 1e30, tRun = 1e30; size_t R = 0; int itDone = 0;
        for (int rep = 0; rep < repeat; rep++) {
            RefWindow *w = buildWindow(in);
            auto *ba = w->ba;
            for (auto f : ba->getFrames()) ba->updateCamera(f);
            collectActive(w);
            R = ba->mActiveResiduals.size();
            ba->computeAdjoints(); ba->computeDelta();
            double a = now_s(); ba->linearizeAll(false); double b = now_s();
            tLin = std::min(tLin, b - a);
            ba->applyActiveRes(true);
            ba->setZero(); ba->computeNullspaces();
            auto pts = ba->getPointsAsList();
            a = now_s();
            for (size_t i = 0; i < pts.size(); i++) ba->addToHessianTop(pts[i].first, pts[i].second, DSORES_ACTIVE);
            ba->stitchDoubleTop(ba->mAccumulatorActive, ba->HA_top, ba->bA_top, false);
            b = now_s(); tTop = std::min(tTop, b - a);
            a = now_s();
            for (size_t i = 0; i < pts.size(); i++) ba->addToHessianSC(pts[i].first, pts[i].second, true);
            ba->stitchDoubleSC(ba->H_sc, ba->b_sc);
            b = now_s(); tSC = std::min(tSC, b - a);
            // whole run() on a second fresh window
            RefWindow *w2 = buildWindow(in);
            a = now_s(); w2->ba->run(w2->updatePointsOnly); b = now_s();
            tRun = std::min(tRun, b - a);
            // GN iterations actually executed (run() leaves early on convergence, BA:879): one P-energy sample before the loop (BA:798) + one per accepted iteration (BA:847)
            itDone = (int) w2->ba->mStatisticEnergyP->mWaitingValues.size() - 1;
        }
        printf("{\"residuals\": %zu, \"t_linearize\": %.6f, \"t_top\": %.6f, \"t_sc\": %.6f, \"t_run\": %.6f, \"iterations\": %d, \"threads\": 1, \"repeat\": %d}\n",
               R, tLin, tTop, tSC, tRun, itDone, repeat);
    } else {
        fprintf(stderr, "unknown mode %s\n", mode.c_str());
        rc = 2;
    }
    fflush(stdout);
    _exit(rc);  // destructor order trips ~PrivateData's abort
}
