// TEST INFRASTRUCTURE (oracle/): runs ONE synthetic Map graph through the unmodified reference class DSOBundleAdjustment and through the
// adapter class DSOBundleAdjustmentB200 (adapter/DSOBundleAdjustmentB200.h, libcmlba behind it) and compares what both leave in the graph:
// run() -> [tryMarginalize -> removePoint(outliers) -> marginalizePointsF -> marginalizeFrames -> run()] (the window-maintenance flow of
// Hybrid::directMap, slam/modslam/direct/Mapping.cpp:61-100).  Prints one JSON line; exit code 0 = within tolerance.
//     oracle/_ref/cmlba_adapter_check --window w.cmlw [--maintain 1]
// Built by `make -C oracle adapter` from the reference sources where they lie (nothing is copied); needs a GPU at run time.
#define main cmlba_ref_driver_main
#include "ref_driver.cpp"
#undef main
#include "../adapter/DSOBundleAdjustmentB200.h"
#include <cstdarg>

struct B200Window {
    Root *root; InternalCalibration *calib; CaptureImageGenerator *gen; DSOBundleAdjustmentB200 *ba;
    std::vector<PFrame> frames; std::vector<PPoint> points;
};

// the same graph construction as buildWindow() of ref_driver.cpp, on a second Map, bound to the adapter class
static B200Window *buildB200(const cmlw::File &in) {
    B200Window *w = new B200Window;
    const int32_t *size = in.get("size").as<int32_t>();
    const int W = size[0], H = size[1], N = (int) in.get("frame_evalpt").dims[0], P = (int) in.get("pt_host").dims[0];
    const double *K = in.get("calib").as<double>();
    w->root = new Root;
    w->calib = new InternalCalibration(PinholeUndistorter(Vector2(K[0], K[1]), Vector2(K[2], K[3])), Vector2(W, H));
    w->gen = new CaptureImageGenerator(W, H, N + 2, N + 2);
    w->ba = new DSOBundleAdjustmentB200(w->root);
    w->ba->maxFrames.set(in.has("max_frames") ? in.get("max_frames").as<int32_t>()[0] : N + 2);
    w->ba->setNumIterations(in.has("iterations") ? in.get("iterations").as<int32_t>()[0] : 4);
    if (in.has("optimize_a")) w->ba->mOptimizeA.set(in.get("optimize_a").as<int32_t>()[0] != 0);
    if (in.has("optimize_b")) w->ba->mOptimizeB.set(in.get("optimize_b").as<int32_t>()[0] != 0);
    if (in.has("force_accept")) w->ba->mForceAccept.set(in.get("force_accept").as<int32_t>()[0] != 0);
    if (in.has("disable_marginalization")) w->ba->mDisableMarginalization.set(in.get("disable_marginalization").as<int32_t>()[0] != 0);
    if (in.has("fixed_lambda")) w->ba->mFixedLambda.set((float) in.get("fixed_lambda").as<double>()[0]);
    Map &map = w->root->getMap();
    const int immature = map.createMapPointGroup("immature");
    const double *evalpt = in.get("frame_evalpt").as<double>(), *cam = in.get("frame_cam").as<double>(), *aff = in.get("frame_affine").as<double>(), *expo = in.get("frame_exposure").as<double>();
    const float *gray = in.get("gray").as<float>();
    const uint8_t *isInit = in.has("frame_init") ? in.get("frame_init").as<uint8_t>() : nullptr;
    for (int i = 0; i < N; i++) {
        FloatImage img(W, H);
        memcpy(img.data(), gray + (size_t) i * W * H, sizeof(float) * W * H);
        auto cap = w->gen->create().setImage(img).setTime(i).setCalibration(w->calib).setExposure(expo[i]).generate();
        PFrame f = map.createFrame(cap);
        f->setCamera(cameraFromRt(evalpt + 12 * i));
        f->setExposureParameters(Exposure(expo[i], aff[2 * i], aff[2 * i + 1]));
        if (isInit && isInit[i]) f->setGroup(map.INITFRAME, true);
        map.addFrame(f);
        w->frames.push_back(f);
    }
    for (int i = 0; i < N; i++) w->ba->addNewFrame(w->frames[i], immature);
    for (int i = 0; i < N; i++) w->frames[i]->setCamera(cameraFromRt(cam + 12 * i));
    const int32_t *host = in.get("pt_host").as<int32_t>();
    const float *xy = in.get("pt_xy").as<float>();
    const double *idepth = in.get("pt_idepth").as<double>();
    std::vector<std::vector<int>> perHost(N);
    for (int p = 0; p < P; p++) perHost[host[p]].push_back(p);
    w->points.resize(P, PPoint());
    PointSet set;
    for (int h = 0; h < N; h++) {
        if (perHost[h].empty()) continue;
        List<Corner> corners;
        for (int p : perHost[h]) corners.emplace_back(Corner(DistortedVector2d(xy[2 * p], xy[2 * p + 1])));
        const int gid = w->frames[h]->addFeaturePoints(corners);
        for (size_t k = 0; k < perHost[h].size(); k++) {
            const int p = perHost[h][k];
            PPoint mp = map.createMapPoint(w->frames[h], FeatureIndex(gid, (short) k), DIRECTTYPE);
            mp->setReferenceInverseDepth(idepth[p]);
            w->points[p] = mp;
            set.insert(mp);
        }
    }
    w->ba->addPoints(set);
    return w;
}

struct Cmp { double rot = 0, trans = 0, aff = 0, idepth = 0; int alive_ref = 0, alive_b200 = 0, alive_diff = 0, good_ref = 0, good_b200 = 0, good_diff = 0, outliers_ref = 0, outliers_b200 = 0; };

// what both classes left in their graphs: frame cameras / exposure, inverse depths of the points both still hold, set sizes
template <typename RefHave, typename B2Have>
static Cmp compare(RefWindow *r, B200Window *b, const std::vector<PFrame> &rf, const std::vector<PFrame> &bf, RefHave refHave, B2Have b2Have) {
    Cmp c;
    double tn = 1e-30;
    for (size_t i = 0; i < rf.size(); i++) tn = std::max(tn, rf[i]->getCamera().getTranslation().norm());
    for (size_t i = 0; i < rf.size(); i++) {
        const Camera a = rf[i]->getCamera(), d = bf[i]->getCamera();
        c.rot = std::max(c.rot, (a.getRotationMatrix() - d.getRotationMatrix()).cwiseAbs().maxCoeff());
        c.trans = std::max(c.trans, (a.getTranslation() - d.getTranslation()).norm() / tn);
        c.aff = std::max(c.aff, (rf[i]->getExposure().getParameters() - bf[i]->getExposure().getParameters()).cwiseAbs().maxCoeff());
    }
    double idmax = 1e-30;
    for (size_t p = 0; p < r->points.size(); p++) if (refHave(r->points[p])) idmax = std::max(idmax, std::abs(r->points[p]->getReferenceInverseDepth()));
    for (size_t p = 0; p < r->points.size(); p++) {
        const bool hr = refHave(r->points[p]), hb = b2Have(b->points[p]);
        c.alive_ref += hr; c.alive_b200 += hb; c.alive_diff += hr != hb;
        if (hr && hb) c.idepth = std::max(c.idepth, std::abs(r->points[p]->getReferenceInverseDepth() - b->points[p]->getReferenceInverseDepth()) / idmax);
    }
    return c;
}

static std::string g_json;
static void jprintf(const char *fmt, ...) { char buf[2048]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap); g_json += buf; }
static void printCmp(const char *name, const Cmp &c, bool last) {
    jprintf("\"%s\": {\"rot_abs\": %.3g, \"trans_rel\": %.3g, \"affine_abs\": %.3g, \"idepth_rel\": %.3g, \"alive_ref\": %d, \"alive_b200\": %d, \"alive_diff\": %d, "
           "\"good_for_tracking_ref\": %d, \"good_for_tracking_b200\": %d, \"good_for_tracking_diff\": %d, \"outliers_ref\": %d, \"outliers_b200\": %d}%s",
           name, c.rot, c.trans, c.aff, c.idepth, c.alive_ref, c.alive_b200, c.alive_diff, c.good_ref, c.good_b200, c.good_diff, c.outliers_ref, c.outliers_b200, last ? "" : ", ");
}

int main(int argc, char **argv) {
    std::string window; int maintain = 1;
    for (int i = 1; i + 1 < argc; i += 2) { std::string k = argv[i]; if (k == "--window") window = argv[i + 1]; else if (k == "--maintain") maintain = atoi(argv[i + 1]); }
    if (window.empty()) { fprintf(stderr, "usage: cmlba_adapter_check --window w.cmlw [--maintain 0|1]\n"); return 2; }
    initCML();
    cmlw::File in; if (!in.load(window)) { fprintf(stderr, "cannot read %s\n", window.c_str()); return 2; }
    fprintf(stderr, "[adapter_check] building reference window\n");
    RefWindow *r = buildWindow(in);
    fprintf(stderr, "[adapter_check] building adapter window\n");
    B200Window *b = nullptr;
    try { b = buildB200(in); } catch (const std::exception &e) { printf("{\"ok\": false, \"error\": \"%s\"}\n", e.what()); return 3; }
    auto goodCmp = [&](Cmp &c) {
        PointSet gr = r->ba->getGoodPointsForTracking(), gb = b->ba->getGoodPointsForTracking();
        c.good_ref = (int) gr.size(); c.good_b200 = (int) gb.size();
        for (size_t p = 0; p < r->points.size(); p++) c.good_diff += (gr.count(r->points[p]) != 0) != (gb.count(b->points[p]) != 0);
        c.outliers_ref = (int) r->ba->getOutliers().size(); c.outliers_b200 = (int) b->ba->getOutliers().size();
    };
    // "in the window" = member of the context's point set and not waiting in getOutliers() for the caller's removePoint (the reference keeps a
    // point that lost all residuals listed until then, BA:1636-1640; libcmlba drops it right away and only reports it)
    auto refHave = [&](PPoint p) { return r->ba->getPoints().count(p) != 0 && r->ba->getOutliers().count(p) == 0; };
    auto b2Have = [&](PPoint p) { return b->ba->have(p); };
    bool ok = true;
    jprintf("{");
    // ---- run()
    const bool okr = r->ba->run(r->updatePointsOnly); r->ba->mActiveResiduals.clear();
    const bool okb = b->ba->run(r->updatePointsOnly);
    Cmp c1 = compare(r, b, r->frames, b->frames, refHave, b2Have); goodCmp(c1);
    jprintf("\"run_ok_ref\": %s, \"run_ok_b200\": %s, ", okr ? "true" : "false", okb ? "true" : "false");
    printCmp("after_run", c1, false);
    auto within = [](const Cmp &c) { return c.rot < 1e-4 && c.trans < 3e-4 && c.aff < 1e-4 && c.idepth < 1e-3 && c.alive_diff == 0 && c.good_diff <= std::max(1, c.good_ref / 1000) && c.outliers_ref == c.outliers_b200; };
    ok = ok && okr == okb && within(c1);
    if (maintain) {
        // ---- the maintenance flow of Hybrid::directMap on both
        r->ba->tryMarginalize(); b->ba->tryMarginalize();
        { std::vector<PPoint> o(r->ba->getOutliers().begin(), r->ba->getOutliers().end()); for (auto p : o) r->ba->removePoint(p); }
        { std::vector<PPoint> o(b->ba->getOutliers().begin(), b->ba->getOutliers().end()); for (auto p : o) b->ba->removePoint(p); }
        r->ba->computeNullspaces(); b->ba->computeNullspaces();
        r->ba->marginalizePointsF(); b->ba->marginalizePointsF();
        auto gone_r = r->ba->marginalizeFrames(); auto gone_b = b->ba->marginalizeFrames();
        std::vector<int> ir, ib;
        for (auto f : gone_r) for (size_t i = 0; i < r->frames.size(); i++) if (r->frames[i] == f) ir.push_back((int) i);
        for (auto f : gone_b) for (size_t i = 0; i < b->frames.size(); i++) if (b->frames[i] == f) ib.push_back((int) i);
        std::sort(ir.begin(), ir.end()); std::sort(ib.begin(), ib.end());
        jprintf("\"frames_marginalized_ref\": %d, \"frames_marginalized_b200\": %d, \"same_frames_marginalized\": %s, ", (int) ir.size(), (int) ib.size(), ir == ib ? "true" : "false");
        ok = ok && ir == ib;
        const bool okr2 = r->ba->run(r->updatePointsOnly); r->ba->mActiveResiduals.clear();
        const bool okb2 = b->ba->run(r->updatePointsOnly);
        std::vector<PFrame> rf, bf;
        for (size_t i = 0; i < r->frames.size(); i++) if (!std::binary_search(ir.begin(), ir.end(), (int) i)) { rf.push_back(r->frames[i]); bf.push_back(b->frames[i]); }
        Cmp c2 = compare(r, b, rf, bf, refHave, b2Have); goodCmp(c2);
        printCmp("after_maintenance_and_second_run", c2, false);
        ok = ok && okr2 == okb2 && within(c2);
    }
    // the Statistic series exist under the reference's names on the adapter
    int nstat = 0, named = 0;
    auto sr = r->ba->getStatistics(); auto sb = b->ba->getStatistics();
    for (auto s : sb) { nstat++; for (auto t : sr) if (t->getName() == s->getName()) { named++; break; } }
    jprintf("\"adapter_statistics\": %d, \"adapter_statistics_with_a_reference_name\": %d, ", nstat, named);
    ok = ok && nstat == 19 && named == 19;
    jprintf("\"ok\": %s}", ok ? "true" : "false");
    fflush(stdout); printf("\n%s\n", g_json.c_str()); fflush(stdout);
    return ok ? 0 : 1;
}
