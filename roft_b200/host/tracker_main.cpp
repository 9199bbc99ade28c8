// roft_b200_tracker - minimal counterpart of the ROFT-tracker executable (src/roft/src/main.cpp:39-427) for the hot
// path: wires dataset sources into the batched ROFTFilter and runs initialization_step + filtering_step until the
// depth stream ends.  One --sequence per track; every track shares the camera / filter parameters (defaults =
// config/config_fast_ycb.cfg).  The libconfig / tclap front-end of the reference is out of scope.
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <regex>
#include <sstream>

#include "roft_host.h"

using namespace ROFT;

static double json_number(const std::string& text, const std::string& key, double dflt) {
    std::smatch m;
    if (std::regex_search(text, m, std::regex("\"" + key + "\"\\s*:\\s*([-+0-9.eE]+)"))) return std::stod(m[1]);
    return dflt;
}

int main(int argc, char** argv) {
    std::vector<std::string> sequences;
    std::string object = "003_cracker_box", log_path = ".", flow_set = "nvof", mask_set = "gt", pose_set = "gt";
    int frames = -1, device = 0;
    double stride = 35.0, fps = 30.0, desired_fps = 5.0;
    bool delay = true, weight = true, resync = true, flow_aided = true;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : std::string(); };
        if (a == "--sequence") sequences.push_back(next());
        else if (a == "--object") object = next();
        else if (a == "--log") log_path = next();
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
                     "[--stride S] [--flow-set s] [--mask-set s] [--pose-set s] [--no-delay] [--no-weight] [--no-resync]" << std::endl;
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
                s.segmentation = std::make_shared<DatasetImageSegmentationDelayed>(float(fps), float(desired_fps), true, seq, "pgm", cam.width,
                                                                                   cam.height, mask_set, object, 0, 0);
                s.pose = std::make_shared<DatasetTransformDelayed>(float(fps), float(desired_fps), true, seq + "/" + pose_set + "/poses.txt", 0, 0, 7);
            } else {
                s.segmentation = std::make_shared<DatasetImageSegmentation>(seq, "pgm", cam.width, cam.height, mask_set, object, 0, 0);
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
                          weight, flow_aided, 2.0, stride, true, log_path, "", device);
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
