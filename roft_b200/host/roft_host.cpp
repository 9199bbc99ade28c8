// roft_host.cpp - see roft_host.h.  Host IO (dataset formats of SURVEY.md 5.1), the delay schedules, and the
// batched ROFTFilter that forwards every frame to libroft_b200.so through the C ABI.
#include "roft_host.h"

#include <zlib.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>

namespace ROFT {

namespace {
std::string compose_file_name(int index, std::size_t number_of_digits) {  // DatasetImageOpticalFlow / Segmentation ::compose_file_name
    std::ostringstream ss;
    ss << std::setw(int(number_of_digits)) << std::setfill('0') << index;
    return ss.str();
}
std::string with_slash(std::string root) {
    if (!root.empty() && root.back() != '/') root += '/';
    return root;
}
}  // namespace

// ---- OpticalFlowUtils ------------------------------------------------------------------------------------
std::pair<bool, FlowFrame> OpticalFlowUtils::read_flow(const std::string& file_name) {
    const std::string log_name = "ROFT::OpticalFlowUtils::read_flow";
    std::FILE* in = std::fopen(file_name.c_str(), "rb");
    if (in == nullptr) {
        std::cout << log_name << " Error: cannot load flow frame " + file_name << std::endl;
        return {false, FlowFrame()};
    }
    FlowFrame f;
    int frame_type = 0;
    std::size_t frame_size[2] = {0, 0};
    if (std::fread(&frame_type, sizeof(frame_type), 1, in) != 1 || std::fread(frame_size, sizeof(frame_size), 1, in) != 1 ||
        (frame_type != ROFTB_FLOW_F32 && frame_type != ROFTB_FLOW_S16)) {
        std::cout << log_name << " Error: cannot load flow frame header for frame " + file_name << std::endl;
        std::fclose(in);
        return {false, FlowFrame()};
    }
    f.type = frame_type;
    f.cols = frame_size[0];
    f.rows = frame_size[1];
    f.data.resize(f.cols * f.rows * f.elem_size());
    const std::size_t n = 2 * f.cols * f.rows;
    if (std::fread(f.data.data(), f.elem_size() / 2, n, in) != n) {
        std::cout << log_name << " Error: cannot load flow data for frame " + file_name << std::endl;
        std::fclose(in);
        return {false, FlowFrame()};
    }
    std::fclose(in);
    return {true, std::move(f)};
}

bool OpticalFlowUtils::save_flow(const FlowFrame& flow, const std::string& output_path) {
    std::FILE* out = std::fopen(output_path.c_str(), "wb");
    if (out == nullptr) return false;
    const int type = flow.type;
    const std::size_t dims[2] = {flow.cols, flow.rows};
    bool ok = std::fwrite(&type, sizeof(type), 1, out) == 1 && std::fwrite(dims, sizeof(dims), 1, out) == 1 &&
              std::fwrite(flow.data.data(), 1, flow.data.size(), out) == flow.data.size();
    std::fclose(out);
    return ok;
}

// ---- DatasetImageOpticalFlow -----------------------------------------------------------------------------
DatasetImageOpticalFlow::DatasetImageOpticalFlow(const std::string& dataset_path, const std::string& set, std::size_t width,
                                                 std::size_t height, std::size_t heading_zeros, std::size_t index_offset)
    : width_(width), height_(height), head_(-1 + int(index_offset)), index_offset_(index_offset), heading_zeros_(heading_zeros) {
    dataset_path_ = with_slash(dataset_path) + "optical_flow/" + set + "/";
    // find the parameters of this dataset from the first readable frame (:38-50); frame 0 has no flow
    bool valid = false;
    FlowFrame tmp;
    for (int counter = 0; !valid && counter < 64; ++counter)
        std::tie(valid, tmp) = OpticalFlowUtils::read_flow(dataset_path_ + compose_file_name(counter, heading_zeros_) + ".float");
    if (!valid) throw std::runtime_error("DatasetImageOpticalFlow::ctor. Error: cannot find any flow frame in " + dataset_path_);
    grid_size_ = width_ / tmp.cols;
    matrix_type_ = tmp.type;
    scaling_factor_ = 1;
    if (matrix_type_ == ROFTB_FLOW_S16) scaling_factor_ = float(1 << 5);
}
bool DatasetImageOpticalFlow::reset() {
    head_ = -1 + int(index_offset_);
    return true;
}
bool DatasetImageOpticalFlow::step_frame() {
    head_++;
    const auto t0 = std::chrono::steady_clock::now();
    std::tie(output_valid_, output_) = OpticalFlowUtils::read_flow(dataset_path_ + compose_file_name(head_, heading_zeros_) + ".float");
    data_loading_time_ = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    return true;
}
std::tuple<bool, const FlowFrame*> DatasetImageOpticalFlow::flow(const bool&) { return {output_valid_, &output_}; }

// ---- segmentation sources --------------------------------------------------------------------------------
DatasetImageSegmentation::DatasetImageSegmentation(const std::string& dataset_path, const std::string& format, std::size_t width,
                                                   std::size_t height, const std::string& segmentation_set,
                                                   const std::string& object_name, std::size_t heading_zeros,
                                                   std::size_t index_offset)
    : head_(-1 + int(index_offset)), format_(format), object_name_(object_name), width_(width), height_(height),
      heading_zeros_(heading_zeros), index_offset_(index_offset) {
    dataset_path_ = with_slash(dataset_path) + "masks/" + segmentation_set + "/";
}
bool DatasetImageSegmentation::reset() {
    head_ = -1 + int(index_offset_);
    return true;
}
bool DatasetImageSegmentation::step_frame() {
    head_++;
    return true;
}
std::pair<bool, MaskImage> DatasetImageSegmentation::segmentation(const bool&) { return read_file(std::size_t(head_)); }

// Greyscale PNG -> 8-bit plane, what cv::imread(IMREAD_UNCHANGED) + convertTo(CV_8UC1) give for the masks of the Fast-YCB /
// HO-3D layouts (DatasetImageSegmentation.cpp:130-132): colour type 0, bit depth 8 or 16 (16-bit samples saturate to 255
// like cv::saturate_cast), non-interlaced.  zlib does the inflate; the five scanline filters are undone here.
bool read_png_gray8(const std::string& file_name, MaskImage& out) {
    std::ifstream in(file_name, std::ios::binary);
    if (!in) return false;
    std::vector<unsigned char> file((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) return false;
    auto be32 = [&](std::size_t o) { return (std::uint32_t(file[o]) << 24) | (std::uint32_t(file[o + 1]) << 16) | (std::uint32_t(file[o + 2]) << 8) | file[o + 3]; };
    std::size_t w = 0, h = 0;
    int depth = 0, colour = -1, interlace = 0;
    std::vector<unsigned char> idat;
    for (std::size_t o = 8; o + 12 <= file.size();) {
        const std::size_t len = be32(o);
        if (o + 12 + len > file.size()) return false;
        const std::string type(reinterpret_cast<const char*>(&file[o + 4]), 4);
        if (type == "IHDR" && len >= 13) {
            w = be32(o + 8); h = be32(o + 12);
            depth = file[o + 16]; colour = file[o + 17]; interlace = file[o + 20];
        } else if (type == "IDAT") {
            idat.insert(idat.end(), file.begin() + long(o + 8), file.begin() + long(o + 8 + len));
        } else if (type == "IEND") {
            break;
        }
        o += 12 + len;
    }
    if (!w || !h || colour != 0 || (depth != 8 && depth != 16) || interlace != 0) return false;
    const std::size_t bpp = std::size_t(depth / 8), stride = w * bpp;
    std::vector<unsigned char> raw((stride + 1) * h);
    uLongf raw_len = uLongf(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), uLong(idat.size())) != Z_OK || raw_len != raw.size()) return false;
    std::vector<unsigned char> prev(stride, 0), cur(stride);
    out.cols = w; out.rows = h;
    out.data.resize(w * h);
    for (std::size_t y = 0; y < h; ++y) {
        const unsigned char* line = &raw[y * (stride + 1)];
        const int filter = line[0];
        for (std::size_t x = 0; x < stride; ++x) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = prev[x], cc = x >= bpp ? prev[x - bpp] : 0;
            int pred = 0;
            switch (filter) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) / 2; break;
                case 4: {
                    const int p = a + b - cc, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - cc);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : cc);
                    break;
                }
                default: return false;
            }
            cur[x] = static_cast<unsigned char>(line[1 + x] + pred);
        }
        for (std::size_t x = 0; x < w; ++x) {
            const unsigned v = bpp == 1 ? cur[x] : ((unsigned(cur[2 * x]) << 8) | cur[2 * x + 1]);
            out.data[y * w + x] = static_cast<std::uint8_t>(v > 255u ? 255u : v);
        }
        prev.swap(cur);
    }
    return true;
}

std::pair<bool, MaskImage> DatasetImageSegmentation::read_file(std::size_t index) {
    const std::string file_name = dataset_path_ + object_name_ + "_" + compose_file_name(int(index), heading_zeros_) + "." + format_;
    auto fail = [&]() {
        std::cout << "DatasetImageSegmentation::segmentation. Error: cannot load segmentation data for frame " + file_name << std::endl;
        return std::pair<bool, MaskImage>{false, MaskImage()};
    };
    MaskImage m;
    if (format_ == "png") {
        if (!read_png_gray8(file_name, m) || m.cols != width_ || m.rows != height_) return fail();
        return {true, std::move(m)};
    }
    std::ifstream in(file_name, std::ios::binary);
    if (!in) return fail();
    // binary PGM (P5), maxval 255 (the format the synthetic sequences of tests / bench are written in)
    std::string magic;
    std::size_t w = 0, h = 0;
    int maxval = 0;
    in >> magic >> w >> h >> maxval;
    in.get();
    if (magic != "P5" || maxval != 255 || w != width_ || h != height_) return {false, MaskImage()};
    m.cols = w;
    m.rows = h;
    m.data.resize(w * h);
    in.read(reinterpret_cast<char*>(m.data.data()), std::streamsize(m.data.size()));
    if (!in) return {false, MaskImage()};
    return {true, std::move(m)};
}

DatasetImageSegmentationDelayed::DatasetImageSegmentationDelayed(float fps, float simulated_fps, bool simulate_inference_time,
                                                                 const std::string& dataset_path, const std::string& format,
                                                                 std::size_t width, std::size_t height,
                                                                 const std::string& segmentation_set, const std::string& object_name,
                                                                 std::size_t heading_zeros, std::size_t index_offset)
    : DatasetImageSegmentation(dataset_path, format, width, height, segmentation_set, object_name, heading_zeros, index_offset),
      fps_(fps), simulated_fps_(simulated_fps), simulate_inference_time_(simulate_inference_time), head_0_(head_ + 1),
      delay_(static_cast<int>(fps / simulated_fps)) {}

int DatasetImageSegmentationDelayed::delivered_index(int head, int delay, int head_0, bool simulate_inference_time) {
    // DatasetImageSegmentationDelayed.cpp:42-52
    int index = head;
    if (simulate_inference_time) index -= delay;
    if (((index - head_0) % delay) != 0) return -1;
    if (index < 0) index = head_0;
    return index;
}

std::pair<bool, MaskImage> DatasetImageSegmentationDelayed::segmentation(const bool&) {
    const int index = delivered_index(head_, delay_, head_0_, simulate_inference_time_);
    if (index < 0) return {false, MaskImage()};
    const auto t0 = std::chrono::steady_clock::now();
    auto output = read_file(std::size_t(index));
    data_loading_time_ = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    return output;
}

// ---- camera / pose ----------------------------------------------------------------------------------------
CameraMeasurement::CameraMeasurement(const std::string& path, const CameraParameters& parameters, std::size_t heading_zeros,
                                     std::size_t index_offset)
    : path_(with_slash(path)), parameters_(parameters), heading_zeros_(heading_zeros), index_offset_(index_offset),
      head_(-1 + int(index_offset)) {
    // data.txt: "stamp_rgb stamp_depth x y z ax ay az angle" per frame
    std::ifstream in(path_ + "data.txt");
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ss(line);
        double stamp = 0;
        if (ss >> stamp) stamps_.push_back(stamp);
    }
}
bool CameraMeasurement::reset() {
    head_ = -1 + int(index_offset_);
    valid_ = false;
    return true;
}
bool CameraMeasurement::freeze() {
    head_++;
    valid_ = false;
    // depth/<i>.float: u64 width, u64 height, H*W float32 metres row-major (tools/dataset/conversion/ho3d_utils.py:74-79)
    std::FILE* in = std::fopen((path_ + "depth/" + compose_file_name(head_, heading_zeros_) + ".float").c_str(), "rb");
    if (in == nullptr) return false;
    std::size_t dims[2] = {0, 0};
    bool ok = std::fread(dims, sizeof(dims), 1, in) == 1 && dims[0] == parameters_.width && dims[1] == parameters_.height;
    if (ok) {
        depth_.cols = dims[0];
        depth_.rows = dims[1];
        depth_.data.resize(dims[0] * dims[1]);
        ok = std::fread(depth_.data.data(), sizeof(float), depth_.data.size(), in) == depth_.data.size();
    }
    std::fclose(in);
    stamp_valid_ = std::size_t(head_) < stamps_.size();
    if (stamp_valid_) stamp_ = stamps_[std::size_t(head_)];
    valid_ = ok;
    return ok;
}
std::pair<bool, const DepthImage*> CameraMeasurement::measure() const { return {valid_, &depth_}; }

DatasetTransformDelayed::DatasetTransformDelayed(float fps, float simulated_fps, bool simulate_delay, const std::string& file_path,
                                                 std::size_t skip_rows, std::size_t skip_cols, std::size_t expected_cols)
    : delay_(static_cast<int>(fps / simulated_fps)), simulate_(simulate_delay) {
    std::ifstream in(file_path);
    if (!in) throw std::runtime_error("DatasetTransformDelayed::ctor. Error: cannot open " + file_path);
    std::string line;
    std::size_t row = 0;
    while (std::getline(in, line)) {
        if (row++ < skip_rows) continue;
        std::istringstream ss(line);
        std::vector<double> v;
        double x;
        while (ss >> x) v.push_back(x);
        if (v.size() < skip_cols + expected_cols) continue;
        rows_.emplace_back(v.begin() + long(skip_cols), v.begin() + long(skip_cols + expected_cols));
    }
}
bool DatasetTransformDelayed::reset() {
    head_ = -1;
    return true;
}
bool DatasetTransformDelayed::freeze(bool) {
    head_++;
    int index = head_;
    if (delay_ > 0) {
        index = DatasetImageSegmentationDelayed::delivered_index(head_, delay_, head_0_, simulate_);
        if (index < 0) return false;
    }
    if (std::size_t(index) >= rows_.size()) return false;
    const std::vector<double>& r = rows_[std::size_t(index)];
    bool all_zero = true;
    for (double v : r) all_zero &= (v == 0.0);
    if (all_zero) return false;  // invalid pose (tools/dataset/dope_pose_finder/pose_finder.py:20)
    // axis-angle -> quaternion (Eigen AngleAxisd -> Quaterniond)
    const double n = std::sqrt(r[3] * r[3] + r[4] * r[4] + r[5] * r[5]);
    const double s = n > 0 ? std::sin(r[6] / 2) / n : 0.0;
    pose_[0] = r[0]; pose_[1] = r[1]; pose_[2] = r[2];
    pose_[3] = std::cos(r[6] / 2); pose_[4] = s * r[3]; pose_[5] = s * r[4]; pose_[6] = s * r[5];
    return true;
}

// ---- mesh input of the depth renderer -----------------------------------------------------------------------
void read_obj_mesh(const std::string& path, std::vector<float>& vertices, std::vector<std::int32_t>& faces) {
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("read_obj_mesh. Error: cannot open " + path);
    vertices.clear();
    faces.clear();
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tag;
        if (!(ls >> tag)) continue;
        if (tag == "v") {
            float x, y, z;
            if (!(ls >> x >> y >> z)) throw std::runtime_error("read_obj_mesh. Error: malformed vertex in " + path);
            vertices.push_back(x); vertices.push_back(y); vertices.push_back(z);
        } else if (tag == "f") {
            std::vector<std::int32_t> idx;
            std::string tok;
            while (ls >> tok) {
                const long v = std::strtol(tok.c_str(), nullptr, 10);  // "v", "v/vt", "v//vn", "v/vt/vn"
                const long nv = long(vertices.size() / 3);
                const long k = v > 0 ? v - 1 : nv + v;
                if (v == 0 || k < 0 || k >= nv) throw std::runtime_error("read_obj_mesh. Error: face index out of range in " + path);
                idx.push_back(std::int32_t(k));
            }
            for (std::size_t i = 2; i < idx.size(); ++i) { faces.push_back(idx[0]); faces.push_back(idx[i - 1]); faces.push_back(idx[i]); }
        }
    }
    if (vertices.empty() || faces.empty()) throw std::runtime_error("read_obj_mesh. Error: no triangles in " + path);
}

// ---- ROFTFilter ---------------------------------------------------------------------------------------------
ROFTFilter::ROFTFilter(std::vector<TrackSources> tracks, const std::vector<double>& initial_covariance_p,
                       const std::vector<double>& model_covariance_p, const std::vector<double>& measurement_covariance_p,
                       const std::vector<double>& initial_covariance_v, const std::vector<double>& model_covariance_v,
                       const std::vector<double>& measurement_covariance_v, double ut_alpha, double ut_beta, double ut_kappa,
                       double sample_time, bool pose_meas, bool pose_resync, bool velocity_meas, bool flow_weighting,
                       bool flow_aided_segmentation, double maximum_depth, double subsampling_radius, bool enable_log,
                       const std::string& log_path, const std::string& log_prefix, int device, bool pose_outlier_rejection,
                       bool pose_outlier_rejection_gain, const std::string& model_mesh_path)
    : tracks_(std::move(tracks)), sample_time_(sample_time), enable_log_(enable_log) {
    const std::string log_name = "ROFTFilter";
    if (tracks_.empty()) throw std::runtime_error(log_name + "::ctor. Error: no tracks.");
    if (initial_covariance_p.size() != 12 || model_covariance_p.size() != 6 || measurement_covariance_p.size() != 12 ||
        initial_covariance_v.size() != 6 || model_covariance_v.size() != 6 || measurement_covariance_v.size() != 2)
        throw std::runtime_error(log_name + "::ctor. Error: wrong size of a covariance vector.");
    bool valid = false;
    CameraParameters cam;
    std::tie(valid, cam) = tracks_[0].camera->camera_parameters();
    if (!valid) throw std::runtime_error(log_name + "::ctor. Error: cannot get camera parameters.");
    roftb_config_default(&cfg_);
    cfg_.n_tracks = int(tracks_.size());
    cfg_.width = int(cam.width); cfg_.height = int(cam.height);
    cfg_.fx = cam.fx; cfg_.fy = cam.fy; cfg_.cx = cam.cx; cfg_.cy = cam.cy;
    cfg_.sample_time = sample_time;
    // flow format from the source, like ROFTFilter.cpp:122-149 picks cv::Vec2f / cv::Vec2s
    cfg_.flow_format = tracks_[0].flow->get_matrix_type();
    cfg_.flow_grid = int(tracks_[0].flow->get_grid_size());
    cfg_.flow_scale = tracks_[0].flow->get_scaling_factor();
    cfg_.cov_flow[0] = measurement_covariance_v[0]; cfg_.cov_flow[1] = measurement_covariance_v[1];
    cfg_.depth_maximum = maximum_depth;
    cfg_.subsampling_radius = int(std::size_t(subsampling_radius));  // double -> size_t ctor argument (hpp:137)
    cfg_.weight_flow = flow_weighting;
    for (int i = 0; i < 6; ++i) { cfg_.v_sigma[i] = model_covariance_v[i]; cfg_.v_cov0[i] = initial_covariance_v[i]; }
    for (int i = 0; i < 12; ++i) cfg_.p_cov0[i] = initial_covariance_p[i];
    for (int i = 0; i < 3; ++i) {
        cfg_.p_sigma_angular[i] = model_covariance_p[i];     // ROFTFilter.cpp:89-90: head = angular, tail = linear psd
        cfg_.p_sigma_linear[i] = model_covariance_p[3 + i];
        cfg_.cov_v[i] = measurement_covariance_p[i]; cfg_.cov_w[i] = measurement_covariance_p[3 + i];
        cfg_.cov_x[i] = measurement_covariance_p[6 + i]; cfg_.cov_q[i] = measurement_covariance_p[9 + i];
    }
    cfg_.ut_alpha = ut_alpha; cfg_.ut_beta = ut_beta; cfg_.ut_kappa = ut_kappa;
    cfg_.use_pose = pose_meas; cfg_.use_pose_resync = pose_resync; cfg_.use_velocity = velocity_meas;
    cfg_.flow_aided = flow_aided_segmentation;
    cfg_.segm_delay = tracks_[0].segmentation->get_frames_between_iterations();
    cfg_.pose_delay = tracks_[0].pose ? tracks_[0].pose->get_frames_between_iterations() : 0;
    cfg_.device = device;
    std::vector<float> mesh_vertices;
    std::vector<std::int32_t> mesh_faces;
    if (pose_outlier_rejection) {
        read_obj_mesh(model_mesh_path, mesh_vertices, mesh_faces);
        cfg_.outlier_rejection = 1;
        cfg_.outlier_rejection_gain = pose_outlier_rejection_gain ? 1.0 : 0.0;  // the reference's `const bool` (ROFTFilter.cpp:54)
        if (cfg_.outlier_rejection_gain == 0.0) cfg_.outlier_rejection_gain = 1.0;  // (a zero gain divides by zero there; the choice ignores it)
    }
    if (roftb_create(&cfg_, &ctx_) != 0) throw std::runtime_error(log_name + "::ctor. Error: " + roftb_last_error(nullptr));
    if (pose_outlier_rejection &&
        roftb_set_mesh(ctx_, mesh_vertices.data(), int(mesh_vertices.size() / 3), mesh_faces.data(), int(mesh_faces.size() / 3)) != 0)
        throw std::runtime_error(log_name + "::ctor. Error: " + roftb_last_error(ctx_));
    const std::size_t T = tracks_.size(), HW = cam.width * cam.height;
    flow_bytes_per_track_ = (cam.width / cfg_.flow_grid) * (cam.height / cfg_.flow_grid) * (cfg_.flow_format == ROFTB_FLOW_S16 ? 4 : 8);
    depth_.resize(T * HW);
    mask_.resize(T * HW);
    flow_.resize(T * flow_bytes_per_track_);
    flow_valid_.resize(T); mask_valid_.resize(T); pose_valid_.resize(T);
    pose_.resize(T * 7); dt_.resize(T);
    last_camera_stamp_.assign(T, -1.0);
    p_mean_.assign(T * 13, 0.0); v_mean_.assign(T * 6, 0.0);
    if (enable_log_) {
        for (std::size_t t = 0; t < T; ++t) {
            const std::string prefix = log_prefix + (T > 1 ? "track" + std::to_string(t) + "_" : "");
            auto names = log_file_names(log_path, prefix);
            log_pose_.emplace_back(names[0] + ".txt");
            log_velocity_.emplace_back(names[1] + ".txt");
            log_time_.emplace_back(names[2] + ".txt");
        }
    }
}

ROFTFilter::~ROFTFilter() { roftb_destroy(ctx_); }

std::vector<std::string> ROFTFilter::log_file_names(const std::string& prefix_path, const std::string& prefix_name) {
    return {prefix_path + "/" + prefix_name + "pose_estimate", prefix_path + "/" + prefix_name + "velocity_estimate",
            prefix_path + "/" + prefix_name + "execution_times"};
}

bool ROFTFilter::initialization_step() {
    const std::size_t T = tracks_.size();
    std::vector<double> p0(T * 13, 0.0), v0(T * 6, 0.0);
    for (std::size_t t = 0; t < T; ++t) {
        p0[t * 13 + 9] = 1.0;
        if (tracks_[t].initial_condition_p.size() == 13) std::memcpy(&p0[t * 13], tracks_[t].initial_condition_p.data(), 13 * sizeof(double));
        if (tracks_[t].initial_condition_v.size() == 6) std::memcpy(&v0[t * 6], tracks_[t].initial_condition_v.data(), 6 * sizeof(double));
        tracks_[t].segmentation->reset();
    }
    return roftb_filter_init(ctx_, p0.data(), v0.data()) == 0;
}

bool ROFTFilter::filtering_step() {
    const std::size_t T = tracks_.size();
    const std::size_t HW = std::size_t(cfg_.width) * cfg_.height;
    const auto time0 = std::chrono::steady_clock::now();
    bool any_flow = false, any_mask = false, any_pose = false;
    for (std::size_t t = 0; t < T; ++t) {
        TrackSources& s = tracks_[t];
        // camera_->freeze(RGBD): no depth -> teardown (ROFTFilter.cpp:261-266)
        if (!s.camera->freeze()) {
            std::cout << "ROFTFilter::filteringStep. Error: cannot continue without a continuous depth stream" << std::endl;
            return false;
        }
        const DepthImage* depth = s.camera->measure().second;
        std::memcpy(&depth_[t * HW], depth->data.data(), HW * sizeof(float));
        // elapsed time between camera stamps (:272-279)
        double elapsed = sample_time_;
        bool sv; double stamp;
        std::tie(sv, stamp) = s.camera->camera_time_stamp_rgb();
        if (sv) {
            if (last_camera_stamp_[t] != -1) elapsed = stamp - last_camera_stamp_[t];
            last_camera_stamp_[t] = stamp;
        }
        dt_[t] = elapsed;
        // flow source stepping (:283), segmentation source (:286), pose source
        if (s.flow->is_stepping_required()) s.flow->step_frame();
        bool fv; const FlowFrame* ff;
        std::tie(fv, ff) = s.flow->flow(false);
        flow_valid_[t] = fv && ff && ff->data.size() == flow_bytes_per_track_;
        if (flow_valid_[t]) std::memcpy(&flow_[t * flow_bytes_per_track_], ff->data.data(), flow_bytes_per_track_);
        any_flow |= flow_valid_[t] != 0;
        if (s.segmentation->is_stepping_required()) s.segmentation->step_frame();
        auto seg = s.segmentation->segmentation(false);
        mask_valid_[t] = seg.first && seg.second.data.size() == HW;
        if (mask_valid_[t]) std::memcpy(&mask_[t * HW], seg.second.data.data(), HW);
        any_mask |= mask_valid_[t] != 0;
        pose_valid_[t] = s.pose ? s.pose->freeze(false) : false;
        if (pose_valid_[t]) std::memcpy(&pose_[t * 7], s.pose->transform(), 7 * sizeof(double));
        any_pose |= pose_valid_[t] != 0;
    }
    const double load_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - time0).count();
    const auto t_start = std::chrono::steady_clock::now();  // start_time_count() excludes loading (:270)
    roftb_frame f;
    std::memset(&f, 0, sizeof(f));
    f.memory = ROFTB_MEM_HOST;
    f.depth = depth_.data(); f.depth_track_stride = (int64_t)HW;
    f.flow = any_flow ? flow_.data() : nullptr;
    f.flow_track_stride = (int64_t)(flow_bytes_per_track_ / (cfg_.flow_format == ROFTB_FLOW_S16 ? 2 : 4));
    f.mask = any_mask ? mask_.data() : nullptr; f.mask_track_stride = (int64_t)HW;
    f.flow_valid = flow_valid_.data(); f.mask_valid = mask_valid_.data();
    f.pose = any_pose ? pose_.data() : nullptr; f.pose_valid = pose_valid_.data();
    f.dt = dt_.data();
    if (roftb_filter_step(ctx_, &f) != 0) throw std::runtime_error(std::string("ROFTFilter::filteringStep. Error: ") + roftb_last_error(ctx_));
    if (roftb_get_state(ctx_, p_mean_.data(), nullptr, v_mean_.data(), nullptr) != 0)
        throw std::runtime_error(std::string("ROFTFilter::filteringStep. Error: ") + roftb_last_error(ctx_));
    const double exec_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count();
    if (enable_log_) {
        for (std::size_t t = 0; t < T; ++t) {
            // pose logged as (v, w, x, axis, angle) (ROFTFilter.cpp:386-392); Eigen's AngleAxisd(Quaterniond)
            const double* m = &p_mean_[t * 13];
            double n = std::sqrt(m[10] * m[10] + m[11] * m[11] + m[12] * m[12]);
            double axis[3] = {1, 0, 0}, angle = 0;
            if (n != 0) {
                angle = 2 * std::atan2(n, std::fabs(m[9]));
                if (m[9] < 0) n = -n;
                axis[0] = m[10] / n; axis[1] = m[11] / n; axis[2] = m[12] / n;
            }
            log_pose_[t] << std::setprecision(17);
            for (int i = 0; i < 9; ++i) log_pose_[t] << m[i] << " ";
            log_pose_[t] << axis[0] << " " << axis[1] << " " << axis[2] << " " << angle << "\n";
            log_velocity_[t] << std::setprecision(17);
            for (int i = 0; i < 6; ++i) log_velocity_[t] << v_mean_[t * 6 + i] << (i < 5 ? " " : "\n");
            log_time_[t] << exec_ms << " " << load_ms << "\n";
        }
    }
    return true;
}

}  // namespace ROFT

// ---- small C hooks for the Python tests (no GPU needed) ---------------------------------------------------------
extern "C" {
int rofth_flow_roundtrip(const char* in_path, const char* out_path, int* type, unsigned long long* cols, unsigned long long* rows) {
    auto r = ROFT::OpticalFlowUtils::read_flow(in_path);
    if (!r.first) return -1;
    *type = r.second.type; *cols = r.second.cols; *rows = r.second.rows;
    return ROFT::OpticalFlowUtils::save_flow(r.second, out_path) ? 0 : -2;
}
int rofth_delivered_index(int head, int delay, int head_0, int simulate) {
    return ROFT::DatasetImageSegmentationDelayed::delivered_index(head, delay, head_0, simulate != 0);
}
int rofth_is_flow_valid(float fx, float fy) { return ROFT::OpticalFlowUtils::is_flow_valid(fx, fy) ? 1 : 0; }
// PNG mask -> out[capacity] (row-major 8-bit); returns 0 and the size, -1 on a reader error, -2 if it does not fit
int rofth_read_png(const char* path, unsigned char* out, unsigned long long capacity, unsigned long long* cols, unsigned long long* rows) {
    ROFT::MaskImage m;
    if (!ROFT::read_png_gray8(path, m)) return -1;
    *cols = m.cols; *rows = m.rows;
    if (m.data.size() > capacity) return -2;
    std::memcpy(out, m.data.data(), m.data.size());
    return 0;
}
}
