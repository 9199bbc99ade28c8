// adapter_check - runs the REFERENCE'S OWN per-frame call sequence (ROFTFilter::filtering_step, ROFTFilter.cpp:255-367,
// transcribed below statement by statement, without the OpenGL outlier test) over the fine-grained adapter classes of
// roft_b200/host/roft_adapters.h, and the fused batched loop (ROFT::ROFTFilter -> roftb_filter_step) over a second set of
// sources reading the same Fast-YCB-format directory, and compares the two beliefs every frame.
//
//   adapter_check --sequence <dir> [--frames N] [--stride S] [--desired-fps F] [--no-resync]
// exit code 0 iff every frame agrees within 1e-6 relative - two orders of magnitude inside the 1e-4 parity tolerance.  The two
// paths run the same kernels at different granularity (one operator per reference method vs one fused step), so the twists
// differ in the last bits (1e-14); the pose UKF's eigen-decomposition of a covariance with clustered eigenvalues can
// amplify that to 1e-8.
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <regex>
#include <sstream>

#include "roft_adapters.h"

using namespace ROFT;
using bfl::Gaussian;

static double json_number(const std::string& text, const std::string& key, double dflt) {
    std::smatch m;
    if (std::regex_search(text, m, std::regex("\"" + key + "\"\\s*:\\s*([-+0-9.eE]+)"))) return std::stod(m[1]);
    return dflt;
}

struct Sources {
    std::shared_ptr<CameraMeasurement> camera;
    std::shared_ptr<Segmentation> segmentation;
    std::shared_ptr<ImageOpticalFlowSource> flow;
    std::shared_ptr<DatasetTransformDelayed> pose;
};

static Sources make_sources(const std::string& seq, const CameraParameters& cam, double fps, double desired_fps) {
    Sources s;
    s.camera = std::make_shared<CameraMeasurement>(seq, cam, 0, 0);
    s.segmentation = std::make_shared<DatasetImageSegmentationDelayed>(float(fps), float(desired_fps), true, seq, "pgm", cam.width, cam.height,
                                                                       "gt", "003_cracker_box", 0, 0);
    s.pose = std::make_shared<DatasetTransformDelayed>(float(fps), float(desired_fps), true, seq + "/gt/poses.txt", 0, 0, 7);
    s.flow = std::make_shared<DatasetImageOpticalFlow>(seq, "nvof", cam.width, cam.height, 0, 0);
    return s;
}

// ROFTFilter with the members and the filtering_step of the reference, over the adapter classes
template <class T>
struct FineGrainedFilter {
    std::shared_ptr<CameraMeasurement> camera_;
    std::shared_ptr<ImageSegmentationMeasurement> segmentation_;
    std::shared_ptr<RobotsIO::Utils::SpatialVelocityBuffer> velocity_;
    std::unique_ptr<bfl::GaussianPrediction> p_prediction_, v_prediction_;
    std::unique_ptr<bfl::GaussianCorrection> p_correction_, v_correction_;
    Gaussian p_pred_belief_{9, 1, true}, p_corr_belief_{9, 1, true}, buffered_belief_{9, 1, true};
    Gaussian v_pred_belief_{6}, v_corr_belief_{6};
    double sample_time_, last_camera_stamp_ = -1;
    bool pose_resync_;
    // ROFTFilter.h: outlier rejection members
    bool outlier_rejection_ = false, outlier_rejection_features_initialized_ = false;
    double outlier_rejection_gain_ = 1.0;
    int divider_ = 4;
    DepthImage buffered_depth_;
    MaskImage buffered_segmentation_;
    std::shared_ptr<B200Context> ctx_;
    std::vector<int> selected_;  // diagnostics: choices of pick_best_alternative

    FineGrainedFilter(const Sources& s, const CameraParameters& cam, std::shared_ptr<B200Context> ctx, const double* x0, double sample_time,
                      bool pose_resync, double stride, double max_depth, bool weighting)
        : sample_time_(sample_time), pose_resync_(pose_resync), ctx_(ctx) {
        camera_ = s.camera;
        divider_ = cam.width == 640 ? 2 : 4;  // ROFTFilter.cpp:191-193
        // ROFTFilter.cpp:118-128: the segmentation source is wrapped in the flow-aided one
        auto of_aided = std::make_shared<ImageSegmentationOFAidedSource<T>>(s.segmentation, s.flow, cam, false, ctx);
        segmentation_ = std::make_shared<ImageSegmentationMeasurement>(of_aided);
        velocity_ = std::make_shared<RobotsIO::Utils::SpatialVelocityBuffer>();
        // :64-106 initial beliefs / covariances (config_fast_ycb.cfg)
        for (int i = 0; i < 13; ++i) p_corr_belief_.mean()(i, 0) = x0[i];
        for (int i = 0; i < 12; ++i) p_corr_belief_.covariance()(i, i) = 1e-3;
        for (int i = 0; i < 6; ++i) v_corr_belief_.covariance()(i, i) = 1e-3;
        buffered_belief_ = p_corr_belief_;  // :231
        Eigen::MatrixXd psd = Eigen::MatrixXd::Identity(3, 3), sw = Eigen::MatrixXd::Identity(3, 3);
        Eigen::MatrixXd qv = Eigen::MatrixXd::Identity(3, 3), qw = Eigen::MatrixXd::Identity(3, 3);
        for (int i = 0; i < 3; ++i) { qv(i, i) = 0.1; qw(i, i) = 0.1; }
        // :163-171 predictions
        p_prediction_ = std::make_unique<UKFPrediction>(std::make_unique<CartesianQuaternionModel>(psd, sw, sample_time), ctx);
        v_prediction_ = std::make_unique<bfl::KFPrediction>(std::make_unique<SpatialVelocityModel>(qv, qw));
        // :174-182 corrections
        auto pose_meas = std::make_unique<CartesianQuaternionMeasurement>(s.pose, velocity_, false, true, true);
        p_correction_ = std::make_unique<UKFCorrection>(std::move(pose_meas), 1.0, 2.0, 0.0, ctx);
        Eigen::MatrixXd r = Eigen::MatrixXd::Identity(2, 2);
        auto flow_meas = std::make_unique<ImageOpticalFlowMeasurement<T>>(s.flow, camera_, segmentation_, std::size_t(stride), max_depth, r, false, ctx);
        v_correction_ = std::make_unique<SKFCorrection>(std::move(flow_meas), 2, weighting);
    }

    // ROFTFilter.cpp:255-367
    bool filtering_step() {
        bool data_in;
        if (!(data_in = camera_->freeze())) return false;                                             // :261-266
        double elapsed_time = sample_time_;                                                           // :273
        double camera_stamp;
        std::tie(std::ignore, camera_stamp) = camera_->camera_time_stamp_rgb();                       // :275
        if (last_camera_stamp_ != -1) elapsed_time = camera_stamp - last_camera_stamp_;               // :276-277
        last_camera_stamp_ = camera_stamp;
        p_prediction_->getStateModel().setSamplingTime(elapsed_time);                                 // :279
        using FT = ImageOpticalFlowMeasurementBase::FreezeType;
        v_correction_->getMeasurementModel().freeze(std::make_pair(FT::OnlyStepSource, elapsed_time));            // :283
        data_in &= segmentation_->freeze();                                                           // :286
        data_in &= v_correction_->getMeasurementModel().freeze(std::make_pair(FT::ExceptStepSource, elapsed_time));  // :289
        if (data_in) {                                                                                // :291-302
            Gaussian v_corr_belief_copy = v_corr_belief_;
            v_prediction_->predict(v_corr_belief_, v_pred_belief_);
            v_correction_->correct(v_pred_belief_, v_corr_belief_);
            if (!v_correction_->getMeasurementModel().setProperty("check_observability")) v_corr_belief_ = v_corr_belief_copy;
        }
        double tw[6];
        for (int i = 0; i < 6; ++i) tw[i] = v_corr_belief_.mean()(i, 0);
        velocity_->set_twist(tw, tw + 3);                                                             // :305
        if (outlier_rejection_ && pose_resync_ && !outlier_rejection_features_initialized_) {          // :313-321
            if (!buffer_outlier_rejection_features()) throw std::runtime_error("cannot initialize the outlier rejection features");
            outlier_rejection_features_initialized_ = true;
        }
        p_prediction_->predict(p_corr_belief_, p_pred_belief_);                                       // :325
        using MM = CartesianQuaternionMeasurement::MeasurementMode;
        if (p_correction_->getMeasurementModel().freeze(MM::Standard)) {                              // :327
            if (p_correction_->getMeasurementModel().getMeasurementDescription().total_size() == 13) {  // :329
                if (pose_resync_) {                                                                   // :331-354
                    Gaussian buffered_belief_copy = buffered_belief_;
                    buffered_belief_ = p_corr_belief_;
                    p_corr_belief_ = buffered_belief_copy;
                    while (p_correction_->getMeasurementModel().freeze(MM::PopBufferedMeasurement)) {
                        p_prediction_->predict(p_corr_belief_, p_pred_belief_);
                        if (outlier_rejection_ && p_correction_->getMeasurementModel().getMeasurementDescription().total_size() == 13)
                            p_corr_belief_ = correct_outlier_rejection(p_pred_belief_, true);           // :346-347
                        else
                            p_correction_->correct(p_pred_belief_, p_corr_belief_);
                    }
                    if (outlier_rejection_) buffer_outlier_rejection_features();                       // :352-353
                } else {
                    if (outlier_rejection_)
                        p_corr_belief_ = correct_outlier_rejection(p_pred_belief_, false);              // :357-358
                    else
                        p_correction_->correct(p_pred_belief_, p_corr_belief_);                       // :360
                }
            } else {
                p_correction_->correct(p_pred_belief_, p_corr_belief_);                               // :364
            }
        } else {
            p_corr_belief_ = p_pred_belief_;                                                          // :367
        }
        return true;
    }

    // ROFTFilter.cpp:624-646
    bool buffer_outlier_rejection_features() {
        const auto d = camera_->measure();
        if (!d.first) return false;
        buffered_depth_ = *d.second;
        const auto m = segmentation_->measure();
        if (!m.first) return false;
        buffered_segmentation_ = bfl::any::any_cast<std::pair<bool, MaskImage>>(m.second).second;
        return true;
    }

    // ROFTFilter.cpp:467-621: SICAD render of both alternatives + masked depth L1 + choice, through the C ABI
    std::pair<bool, Gaussian> pick_best_alternative(const std::vector<Gaussian>& alternatives, bool use_buffered_features) {
        DepthImage depth;
        MaskImage segmentation;
        if (use_buffered_features) {
            depth = buffered_depth_;
            segmentation = buffered_segmentation_;
        } else {
            const auto d = camera_->measure();
            if (!d.first) return {false, Gaussian()};
            depth = *d.second;
            const auto m = segmentation_->measure();
            if (!m.first) return {false, Gaussian()};
            segmentation = bfl::any::any_cast<std::pair<bool, MaskImage>>(m.second).second;
        }
        double alts[26];
        for (int k = 0; k < 2; ++k)
            for (int i = 0; i < 13; ++i) alts[13 * k + i] = alternatives[k].mean()(i, 0);
        std::int32_t selected = 0;
        double lik[2];
        if (roftb_pick_best_alternative(ctx_->get(), 1, segmentation.data.data(), depth.data.data(), alts, divider_, outlier_rejection_gain_,
                                        &selected, lik) != 0)
            return {false, Gaussian()};
        selected_.push_back(selected);
        return {true, alternatives[selected]};
    }

    // ROFTFilter.cpp:649-676
    Gaussian correct_outlier_rejection(const Gaussian& prediction, bool use_buffered_features) {
        Gaussian corr_belief_v = prediction, corr_belief_p_v = prediction;
        p_correction_->correct(prediction, corr_belief_p_v);
        p_correction_->getMeasurementModel().freeze(CartesianQuaternionMeasurement::MeasurementMode::RepeatOnlyVelocity);
        p_correction_->correct(prediction, corr_belief_v);
        const auto best = pick_best_alternative({corr_belief_p_v, corr_belief_v}, use_buffered_features);
        return best.first ? best.second : corr_belief_p_v;
    }
};

template <class T>
static int run(const std::string& seq, const CameraParameters& cam, int flow_type, std::size_t grid, float scale, int frames, double stride,
               double fps, double desired_fps, bool resync, const std::string& mesh_path) {
    const double cov_flow[2] = {1.0, 1.0};
    const double p_model[6] = {1, 1, 1, 1, 1, 1};
    const double p_meas[12] = {0.1, 0.1, 0.1, 1e-4, 1e-4, 1e-4, 1e-3, 1e-3, 1e-3, 1e-4, 1e-4, 1e-4};
    const int delay = int(fps / desired_fps);
    auto ctx = std::make_shared<B200Context>(cam, flow_type, grid, scale, stride, 2.0, true, cov_flow, p_model, p_meas, 1.0, 2.0, 0.0, delay);
    // initial condition = first pose of the pose file
    DatasetTransformDelayed init(float(fps), float(fps), false, seq + "/gt/poses.txt", 0, 0, 7);
    std::vector<double> x0(13, 0.0);
    x0[9] = 1.0;
    if (init.freeze(false)) std::copy(init.transform(), init.transform() + 7, x0.begin() + 6);

    FineGrainedFilter<T> fine(make_sources(seq, cam, fps, desired_fps), cam, ctx, x0.data(), 0.033333333333, resync, stride, 2.0, true);
    const bool outrej = !mesh_path.empty();
    if (outrej) {  // the fine-grained loop renders through the operators of the same library
        std::vector<float> mv;
        std::vector<std::int32_t> mf;
        read_obj_mesh(mesh_path, mv, mf);
        if (roftb_set_mesh(ctx->get(), mv.data(), int(mv.size() / 3), mf.data(), int(mf.size() / 3)) != 0) throw std::runtime_error(roftb_last_error(ctx->get()));
        fine.outlier_rejection_ = true;
    }

    Sources s2 = make_sources(seq, cam, fps, desired_fps);
    TrackSources ts;
    ts.camera = s2.camera; ts.segmentation = s2.segmentation; ts.flow = s2.flow; ts.pose = s2.pose;
    ts.initial_condition_p = x0;
    ts.initial_condition_v.assign(6, 0.0);
    std::vector<TrackSources> tracks;
    tracks.push_back(std::move(ts));
    const std::vector<double> p_cov0(12, 1e-3), v_cov0(6, 1e-3), v_q(6, 0.1), v_r{1.0, 1.0};
    ROFTFilter fused(std::move(tracks), p_cov0, std::vector<double>(p_model, p_model + 6), std::vector<double>(p_meas, p_meas + 12), v_cov0, v_q, v_r,
                     1.0, 2.0, 0.0, 0.033333333333, true, resync, true, true, true, 2.0, stride, false, ".", "", 0, outrej, true, mesh_path);
    fused.initialization_step();

    double worst = 0.0;
    int k = 0;
    for (; frames < 0 || k < frames; ++k) {
        const bool a = fine.filtering_step(), b = fused.filtering_step();
        if (a != b) { std::cerr << "frame " << k << ": one loop ended before the other" << std::endl; return 1; }
        if (!a) break;
        auto rel = [](const double* x, const double* y, int n) {
            double d = 0, m = 0;
            for (int i = 0; i < n; ++i) { d += (x[i] - y[i]) * (x[i] - y[i]); m += y[i] * y[i]; }
            return std::sqrt(d) / std::max(std::sqrt(m), 1e-12);
        };
        double pf[13], vf[6];
        for (int i = 0; i < 13; ++i) pf[i] = fine.p_corr_belief_.mean()(i, 0);
        for (int i = 0; i < 6; ++i) vf[i] = fine.v_corr_belief_.mean()(i, 0);
        const double ev = rel(vf, fused.velocity_mean().data(), 6);
        const double ep = rel(pf, fused.pose_mean().data(), 13);
        const bool v_small = std::sqrt(vf[0] * vf[0] + vf[1] * vf[1] + vf[2] * vf[2] + vf[3] * vf[3] + vf[4] * vf[4] + vf[5] * vf[5]) < 1e-12;
        worst = std::max(worst, std::max(v_small ? 0.0 : ev, ep));
        if ((!v_small && ev > 1e-6) || ep > 1e-6) {
            std::cerr << "frame " << k << ": velocity rel " << ev << ", pose rel " << ep << std::endl;
            return 1;
        }
    }
    std::cout << "adapter_check: " << k << " frames, reference call sequence over the adapters == fused batched step, max rel diff " << worst
              << std::endl;
    if (outrej) {
        std::cout << "outlier rejection choices:";
        for (int c : fine.selected_) std::cout << ' ' << c;
        std::cout << std::endl;
    }
    return 0;
}

int main(int argc, char** argv) {
    std::string seq;
    int frames = -1;
    double stride = 35.0, fps = 30.0, desired_fps = 5.0;
    bool resync = true;
    std::string mesh_path;  // non-empty: outlier rejection on, in both loops
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
        if (a == "--sequence") seq = next();
        else if (a == "--frames") frames = std::atoi(next().c_str());
        else if (a == "--stride") stride = std::atof(next().c_str());
        else if (a == "--desired-fps") desired_fps = std::atof(next().c_str());
        else if (a == "--no-resync") resync = false;
        else if (a == "--mesh") mesh_path = next();
        else { std::cerr << "unknown option " << a << std::endl; return 2; }
    }
    if (seq.empty()) { std::cerr << "usage: adapter_check --sequence <dir> [--frames N] [--stride S] [--desired-fps F] [--no-resync]" << std::endl; return 2; }
    try {
        std::ifstream kf(seq + "/cam_K.json");
        std::stringstream ks;
        ks << kf.rdbuf();
        CameraParameters cam;
        cam.width = std::size_t(json_number(ks.str(), "width", 1280));
        cam.height = std::size_t(json_number(ks.str(), "height", 720));
        cam.fx = json_number(ks.str(), "fx", 1229.4285612615463);
        cam.fy = json_number(ks.str(), "fy", 1229.4285612615463);
        cam.cx = json_number(ks.str(), "cx", 640.0);
        cam.cy = json_number(ks.str(), "cy", 360.0);
        // ROFTFilter.cpp:122-149: the flow element type decides the template argument
        DatasetImageOpticalFlow probe(seq, "nvof", cam.width, cam.height, 0, 0);
        if (probe.get_matrix_type() == ROFTB_FLOW_S16)
            return run<cv::Vec2s>(seq, cam, ROFTB_FLOW_S16, probe.get_grid_size(), probe.get_scaling_factor(), frames, stride, fps, desired_fps, resync, mesh_path);
        return run<cv::Vec2f>(seq, cam, ROFTB_FLOW_F32, probe.get_grid_size(), probe.get_scaling_factor(), frames, stride, fps, desired_fps, resync, mesh_path);
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
}
