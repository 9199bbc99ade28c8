// roft_b200_tracker - minimal counterpart of the ROFT-tracker executable (src/roft/src/main.cpp:39-427) for the hot
// path: wires dataset sources into the batched ROFTFilter and runs initialization_step + filtering_step until the
// depth stream ends.  Two front-ends:
//   roft_b200_tracker --from <file.cfg> [--group::key value ...]     the reference's own: a libconfig file with every
//        leaf overridable on the command line (src/roft/src/main.cpp:41-147, ConfigParser.cpp:8-169) - one track;
//   roft_b200_tracker --sequence <dir> [--sequence <dir> ...] [...]   several tracks batched through one context, every
//        track sharing the camera / filter parameters (defaults = config/config_fast_ycb.cfg).
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <regex>
#include <sstream>

#include "config_parser.h"
#include "roft_host.h"

using namespace ROFT;

static double json_number(const std::string& text, const std::string& key, double dflt) {
    std::smatch m;
    if (std::regex_search(text, m, std::regex("\"" + key + "\"\\s*:\\s*([-+0-9.eE]+)"))) return std::stod(m[1]);
    return dflt;
}

// main.cpp:41-424 over the batched filter (one track): same setting names, same wiring
static int main_from_config(int argc, char** argv) {
    ConfigParser conf(argc, argv);
    double sample_time; conf("sample_time", sample_time);
    CameraParameters cam;
    int iv;
    conf("camera_dataset.width", iv); cam.width = std::size_t(iv);
    conf("camera_dataset.height", iv); cam.height = std::size_t(iv);
    conf("camera_dataset.fx", cam.fx); conf("camera_dataset.fy", cam.fy);
    conf("camera_dataset.cx", cam.cx); conf("camera_dataset.cy", cam.cy);
    std::string camera_path; conf("camera_dataset.path", camera_path);
    int camera_heading_zeros, camera_index_offset;
    conf("camera_dataset.heading_zeros", camera_heading_zeros); conf("camera_dataset.index_offset", camera_index_offset);
    std::vector<double> p_v_0, p_w_0, p_x_0, p_aa_0, p_cov_v_0, p_cov_w_0, p_cov_x_0, p_cov_q_0, v_v_0, v_w_0, v_cov_v_0, v_cov_w_0;
    conf("initial_condition.pose.v", p_v_0); conf("initial_condition.pose.w", p_w_0); conf("initial_condition.pose.x", p_x_0);
    conf("initial_condition.pose.axis_angle", p_aa_0);
    conf("initial_condition.pose.cov_v", p_cov_v_0); conf("initial_condition.pose.cov_w", p_cov_w_0);
    conf("initial_condition.pose.cov_x", p_cov_x_0); conf("initial_condition.pose.cov_q", p_cov_q_0);
    conf("initial_condition.velocity.v", v_v_0); conf("initial_condition.velocity.w", v_w_0);
    conf("initial_condition.velocity.cov_v", v_cov_v_0); conf("initial_condition.velocity.cov_w", v_cov_w_0);
    std::vector<double> psd_lin_acc, sigma_ang_vel, kin_q_v, kin_q_w;
    conf("kinematic_model.pose.sigma_linear", psd_lin_acc); conf("kinematic_model.pose.sigma_angular", sigma_ang_vel);
    conf("kinematic_model.velocity.sigma_linear", kin_q_v); conf("kinematic_model.velocity.sigma_angular", kin_q_w);
    bool enable_log; conf("log.enable", enable_log);
    std::string log_path; conf("log.path", log_path);
    if (enable_log && log_path == "") {
        std::cout << "Invalid log path. Disabling log." << std::endl;
        enable_log = false;
    }
    std::vector<double> m_cov_v, m_cov_w, m_cov_x, m_cov_q, v_meas_cov_flow;
    conf("measurement_model.pose.cov_v", m_cov_v); conf("measurement_model.pose.cov_w", m_cov_w);
    conf("measurement_model.pose.cov_x", m_cov_x); conf("measurement_model.pose.cov_q", m_cov_q);
    conf("measurement_model.velocity.cov_flow", v_meas_cov_flow);
    double depth_maximum, subsampling_radius;
    conf("measurement_model.velocity.depth_maximum", depth_maximum);
    conf("measurement_model.velocity.subsampling_radius", subsampling_radius);
    bool flow_weighting, use_pose, use_pose_resync, use_velocity;
    conf("measurement_model.velocity.weight_flow", flow_weighting);
    conf("measurement_model.use_pose", use_pose); conf("measurement_model.use_pose_resync", use_pose_resync);
    conf("measurement_model.use_velocity", use_velocity);
    std::string model_name; conf("model.name", model_name);
    std::string of_path, of_set; conf("optical_flow_dataset.path", of_path); conf("optical_flow_dataset.set", of_set);
    int of_heading_zeros, of_index_offset;
    conf("optical_flow_dataset.heading_zeros", of_heading_zeros); conf("optical_flow_dataset.index_offset", of_index_offset);
    bool outlier_rejection_enable; conf("outlier_rejection.enable", outlier_rejection_enable);
    double outlier_rejection_gain = 0.0;
    if (conf.exists("outlier_rejection.gain")) conf("outlier_rejection.gain", outlier_rejection_gain);
    // ModelParameters (main.cpp:224-235): only an external Wavefront OBJ is understood here (no Assimp, no internal DB)
    std::string model_mesh_path;
    if (outlier_rejection_enable) {
        if (conf.exists("model.external_path")) conf("model.external_path", model_mesh_path);
        if (model_mesh_path.empty() || !std::ifstream(model_mesh_path).is_open()) {
            std::cout << "outlier_rejection.enable: no readable mesh at model.external_path (Wavefront OBJ); running without the "
                         "render-and-compare pose test (as --outlier_rejection::enable false)." << std::endl;
            outlier_rejection_enable = false;
        }
    }
    std::string pose_path; conf("pose_dataset.path", pose_path);
    int pose_skip_rows, pose_skip_cols; conf("pose_dataset.skip_rows", pose_skip_rows); conf("pose_dataset.skip_cols", pose_skip_cols);
    bool pose_fps_reduction, pose_delay; conf("pose_dataset.fps_reduction", pose_fps_reduction); conf("pose_dataset.delay", pose_delay);
    double pose_original_fps, pose_desired_fps;
    conf("pose_dataset.original_fps", pose_original_fps); conf("pose_dataset.desired_fps", pose_desired_fps);
    std::string seg_path, seg_format, seg_set;
    conf("segmentation_dataset.path", seg_path); conf("segmentation_dataset.format", seg_format); conf("segmentation_dataset.set", seg_set);
    int seg_heading_zeros, seg_index_offset;
    conf("segmentation_dataset.heading_zeros", seg_heading_zeros); conf("segmentation_dataset.index_offset", seg_index_offset);
    double seg_original_fps, seg_desired_fps;
    conf("segmentation_dataset.original_fps", seg_original_fps); conf("segmentation_dataset.desired_fps", seg_desired_fps);
    bool seg_fps_reduction, seg_delay, flow_aided;
    conf("segmentation_dataset.fps_reduction", seg_fps_reduction); conf("segmentation_dataset.delay", seg_delay);
    conf("segmentation_dataset.flow_aided", flow_aided);
    double ut_alpha, ut_beta, ut_kappa;
    conf("unscented_transform.alpha", ut_alpha); conf("unscented_transform.beta", ut_beta); conf("unscented_transform.kappa", ut_kappa);

    TrackSources s;
    s.camera = std::make_shared<CameraMeasurement>(camera_path, cam, std::size_t(camera_heading_zeros), std::size_t(camera_index_offset));
    if (pose_delay || pose_fps_reduction) {  // main.cpp:347-356
        if (!pose_fps_reduction) pose_desired_fps = pose_original_fps;
        s.pose = std::make_shared<DatasetTransformDelayed>(float(pose_original_fps), float(pose_desired_fps), pose_delay, pose_path,
                                                           std::size_t(pose_skip_rows), std::size_t(pose_skip_cols), 7);
    } else {
        s.pose = std::make_shared<DatasetTransformDelayed>(float(pose_original_fps), float(pose_original_fps), false, pose_path,
                                                           std::size_t(pose_skip_rows), std::size_t(pose_skip_cols), 7);
    }
    if (seg_delay || seg_fps_reduction) {  // main.cpp:359-381
        if (!seg_fps_reduction) seg_desired_fps = seg_original_fps;
        s.segmentation = std::make_shared<DatasetImageSegmentationDelayed>(float(seg_original_fps), float(seg_desired_fps), seg_delay, seg_path,
                                                                           seg_format, cam.width, cam.height, seg_set, model_name,
                                                                           std::size_t(seg_heading_zeros), std::size_t(seg_index_offset));
    } else {
        s.segmentation = std::make_shared<DatasetImageSegmentation>(seg_path, seg_format, cam.width, cam.height, seg_set, model_name,
                                                                    std::size_t(seg_heading_zeros), std::size_t(seg_index_offset));
    }
    s.flow = std::make_shared<DatasetImageOpticalFlow>(of_path, of_set, cam.width, cam.height, std::size_t(of_heading_zeros), std::size_t(of_index_offset));
    // initial condition: (v, w, x, axis-angle -> quaternion) (main.cpp:283-296)
    s.initial_condition_p.assign(13, 0.0);
    for (int i = 0; i < 3; ++i) { s.initial_condition_p[i] = p_v_0[i]; s.initial_condition_p[3 + i] = p_w_0[i]; s.initial_condition_p[6 + i] = p_x_0[i]; }
    {
        const double n = std::sqrt(p_aa_0[0] * p_aa_0[0] + p_aa_0[1] * p_aa_0[1] + p_aa_0[2] * p_aa_0[2]);
        const double sn = n > 0 ? std::sin(p_aa_0[3] / 2) / n : 0.0;
        s.initial_condition_p[9] = std::cos(p_aa_0[3] / 2);
        for (int i = 0; i < 3; ++i) s.initial_condition_p[10 + i] = sn * p_aa_0[i];
    }
    s.initial_condition_v = {v_v_0[0], v_v_0[1], v_v_0[2], v_w_0[0], v_w_0[1], v_w_0[2]};
    auto cat = [](std::initializer_list<const std::vector<double>*> parts) {
        std::vector<double> out;
        for (const auto* p : parts) out.insert(out.end(), p->begin(), p->end());
        return out;
    };
    std::vector<TrackSources> tracks;
    tracks.push_back(std::move(s));
    ROFTFilter filter(std::move(tracks), cat({&p_cov_v_0, &p_cov_w_0, &p_cov_x_0, &p_cov_q_0}), cat({&sigma_ang_vel, &psd_lin_acc}),
                      cat({&m_cov_v, &m_cov_w, &m_cov_x, &m_cov_q}), cat({&v_cov_v_0, &v_cov_w_0}), cat({&kin_q_v, &kin_q_w}), v_meas_cov_flow,
                      ut_alpha, ut_beta, ut_kappa, sample_time, use_pose, use_pose_resync, use_velocity, flow_weighting, flow_aided,
                      depth_maximum, subsampling_radius, enable_log, log_path, "", 0, outlier_rejection_enable,
                      bool(outlier_rejection_gain), model_mesh_path);  // (the gain goes through `const bool`, ROFTFilter.cpp:54)
    filter.initialization_step();
    int k = 0;
    while (filter.filtering_step()) ++k;
    std::cout << "tracked " << k << " frames x " << filter.n_tracks() << " tracks" << std::endl;
    return 0;
}

int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i)
        if (std::string(argv[i]) == "--from") {
            try {
                return main_from_config(argc, argv);
            } catch (const std::exception& e) {
                std::cerr << e.what() << std::endl;
                return 1;
            }
        }
    std::vector<std::string> sequences;
    std::string object = "003_cracker_box", log_path = ".", flow_set = "nvof", mask_set = "gt", pose_set = "gt", mesh_path, mask_format = "pgm";
    bool outlier_rejection = false;
    int frames = -1, device = 0;
    double stride = 35.0, fps = 30.0, desired_fps = 5.0;
    bool delay = true, weight = true, resync = true, flow_aided = true;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
        if (a == "--sequence") sequences.push_back(next());
        else if (a == "--object") object = next();
        else if (a == "--log") log_path = next();
        else if (a == "--mesh") mesh_path = next();
        else if (a == "--mask-format") mask_format = next();
        else if (a == "--outlier-rejection") outlier_rejection = true;
        else if (a == "--flow-set") flow_set = next();
        else if (a == "--mask-set") mask_set = next();
        else if (a == "--pose-set") pose_set = next();
        else if (a == "--frames") frames = std::atoi(next().c_str());
        else if (a == "--device") device = std::atoi(next().c_str());
        else if (a == "--stride") stride = std::atof(next().c_str());
        else if (a == "--desired-fps") desired_fps = std::atof(next().c_str());
        else if (a == "--no-delay") delay = false;
        else if (a == "--no-weight") weight = false;
        else if (a == "--no-resync") resync = false;
        else if (a == "--no-flow-aid") flow_aided = false;
        else { std::cerr << "unknown option " << a << std::endl; return 2; }
    }
    if (sequences.empty()) {
        std::cerr << "usage: roft_b200_tracker --sequence <dir> [--sequence <dir> ...] [--object name] [--log dir] [--frames N] "
                     "[--stride S] [--flow-set s] [--mask-set s] [--pose-set s] [--no-delay] [--no-weight] [--no-resync] [--outlier-rejection --mesh file.obj] [--mask-format pgm|png]" << std::endl;
        return 2;
    }
    try {
        std::vector<TrackSources> tracks;
        for (const std::string& seq : sequences) {
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
            TrackSources s;
            s.camera = std::make_shared<CameraMeasurement>(seq, cam, 0, 0);
            if (delay) {
                s.segmentation = std::make_shared<DatasetImageSegmentationDelayed>(float(fps), float(desired_fps), true, seq, mask_format, cam.width,
                                                                                   cam.height, mask_set, object, 0, 0);
                s.pose = std::make_shared<DatasetTransformDelayed>(float(fps), float(desired_fps), true, seq + "/" + pose_set + "/poses.txt", 0, 0, 7);
            } else {
                s.segmentation = std::make_shared<DatasetImageSegmentation>(seq, mask_format, cam.width, cam.height, mask_set, object, 0, 0);
                s.pose = std::make_shared<DatasetTransformDelayed>(float(fps), float(fps), false, seq + "/" + pose_set + "/poses.txt", 0, 0, 7);
            }
            s.flow = std::make_shared<DatasetImageOpticalFlow>(seq, flow_set, cam.width, cam.height, 0, 0);
            // initial condition = first pose of the pose file (what test/test.sh:120-123 injects)
            DatasetTransformDelayed init(float(fps), float(fps), false, seq + "/" + pose_set + "/poses.txt", 0, 0, 7);
            s.initial_condition_p.assign(13, 0.0);
            s.initial_condition_p[9] = 1.0;
            if (init.freeze(false)) std::copy(init.transform(), init.transform() + 7, s.initial_condition_p.begin() + 6);
            s.initial_condition_v.assign(6, 0.0);
            tracks.push_back(std::move(s));
        }
        // config/config_fast_ycb.cfg
        const std::vector<double> p_cov0(12, 1e-3), v_cov0(6, 1e-3), v_q(6, 0.1), v_r{1.0, 1.0};
        const std::vector<double> p_model{1.0, 1.0, 1.0, 1.0, 1.0, 1.0};
        const std::vector<double> p_meas{0.1, 0.1, 0.1, 1e-4, 1e-4, 1e-4, 1e-3, 1e-3, 1e-3, 1e-4, 1e-4, 1e-4};
        ROFTFilter filter(std::move(tracks), p_cov0, p_model, p_meas, v_cov0, v_q, v_r, 1.0, 2.0, 0.0, 0.033333333333, true, resync, true,
                          weight, flow_aided, 2.0, stride, true, log_path, "", device, outlier_rejection, true, mesh_path);
        filter.initialization_step();
        int k = 0;
        while ((frames < 0 || k < frames) && filter.filtering_step()) ++k;
        std::cout << "tracked " << k << " frames x " << filter.n_tracks() << " tracks" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
