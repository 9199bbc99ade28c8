// roft_host.h - host-side C++ mirror of the reference's operator / plugin interface for the hot path,
// implemented over the C ABI of libroft_b200.so (include/roft_b200.h).
//
// The reference's host code is C++ over Eigen / OpenCV / BayesFilters / RobotsIO, none of which exist in
// this image, so this layer keeps the reference's CLASS NAMES, METHOD NAMES, ARGUMENT MEANING and ERROR
// BEHAVIOUR (bool / pair<bool, ...> for soft failure, std::runtime_error from constructors) with plain
// std:: containers where the reference uses cv::Mat / Eigen types.  With the real dependencies available
// the same bodies compile against them by swapping the three POD structs below for cv::Mat /
// Eigen::MatrixXf.  Batching: ROFTFilter drives N independent tracks (one set of
// sources each) through ONE roftb_ctx, i.e. one CUDA launch sequence per frame for all tracks.
//
// Reference interfaces mirrored (paths under hsp-iit/roft):
//   ROFT::OpticalFlowUtils::{is_flow_valid, read_flow, save_flow}  src/roft-lib/include/ROFT/OpticalFlowUtilities.h:19-31
//   ROFT::ImageOpticalFlowSource                                   src/roft-lib/include/ROFT/ImageOpticalFlowSource.h:20-48
//   ROFT::DatasetImageOpticalFlow                                  src/roft-lib/src/DatasetImageOpticalFlow.cpp:24-100
//   RobotsIO::Utils::Segmentation (interface as overridden at)     src/roft-lib/include/ROFT/DatasetImageSegmentationDelayed.h:21-31
//   ROFT::DatasetImageSegmentation(Delayed)                        src/roft-lib/src/DatasetImageSegmentation(Delayed).cpp
//   RobotsIO DatasetCamera / DatasetTransform(Delayed) (UPSTREAM-RECALL, formats SURVEY.md 5.1)
//   ROFT::ROFTFilter                                               src/roft-lib/src/ROFTFilter.cpp:32-452
#pragma once

#include <cstddef>
#include <cstdint>
#include <fstream>
#include <memory>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "roft_b200.h"

namespace ROFT {

// ---- stand-ins for cv::Mat / Eigen::MatrixXf (row-major, owning) ---------------------------------------
struct FlowFrame {            // cv::Mat of type CV_32FC2 (13) or CV_16SC2 (11)
    int type = 0;
    std::size_t cols = 0, rows = 0;
    std::vector<std::uint8_t> data;
    bool empty() const { return data.empty(); }
    std::size_t elem_size() const { return type == ROFTB_FLOW_S16 ? 4 : 8; }
};
struct MaskImage {            // cv::Mat CV_8UC1
    std::size_t cols = 0, rows = 0;
    std::vector<std::uint8_t> data;
    bool empty() const { return data.empty(); }
};
struct DepthImage {           // Eigen::MatrixXf indexed (v, u); stored row-major like the .float files
    std::size_t cols = 0, rows = 0;
    std::vector<float> data;
    bool empty() const { return data.empty(); }
};

namespace OpticalFlowUtils {
inline bool is_flow_valid(const float& f_x, const float& f_y) {  // OpticalFlowUtilities.h:19-22
    return !(f_x != f_x) && !(f_y != f_y) && (f_x < 0 ? -f_x : f_x) < 1e9 && (f_y < 0 ? -f_y : f_y) < 1e9;
}
std::pair<bool, FlowFrame> read_flow(const std::string& file_name);   // OpticalFlowUtilities.cpp:26-74
bool save_flow(const FlowFrame& flow, const std::string& output_path);  // OpticalFlowUtilities.cpp:77-136
}  // namespace OpticalFlowUtils

// ---- sources ---------------------------------------------------------------------------------------------
class ImageOpticalFlowSource {  // ImageOpticalFlowSource.h:20-48
public:
    virtual ~ImageOpticalFlowSource() = default;
    virtual bool reset() { return true; }
    virtual bool step_frame() { return true; }
    virtual bool is_stepping_required() const = 0;
    virtual double get_data_loading_time() const { return 0.0; }
    virtual std::tuple<bool, const FlowFrame*> flow(const bool& blocking) = 0;
    virtual std::size_t get_grid_size() const = 0;
    virtual float get_scaling_factor() const = 0;
    virtual int get_matrix_type() const = 0;
};

class DatasetImageOpticalFlow : public ImageOpticalFlowSource {  // DatasetImageOpticalFlow.cpp:24-100
public:
    DatasetImageOpticalFlow(const std::string& dataset_path, const std::string& set, std::size_t width, std::size_t height,
                            std::size_t heading_zeros, std::size_t index_offset);
    bool reset() override;
    bool step_frame() override;
    bool is_stepping_required() const override { return true; }
    double get_data_loading_time() const override { return data_loading_time_; }
    std::tuple<bool, const FlowFrame*> flow(const bool& blocking) override;
    std::size_t get_grid_size() const override { return grid_size_; }
    float get_scaling_factor() const override { return scaling_factor_; }
    int get_matrix_type() const override { return matrix_type_; }

private:
    std::string dataset_path_;
    std::size_t width_, height_;
    int head_;
    std::size_t index_offset_, heading_zeros_;
    std::size_t grid_size_ = 1;
    float scaling_factor_ = 1;
    int matrix_type_ = 0;
    bool output_valid_ = false;
    FlowFrame output_;
    double data_loading_time_ = 0.0;
};

class Segmentation {  // RobotsIO::Utils::Segmentation as used by ROFT (UPSTREAM-RECALL)
public:
    virtual ~Segmentation() = default;
    virtual bool reset() { return true; }
    virtual bool step_frame() { return true; }
    virtual bool is_stepping_required() const = 0;
    virtual void reset_data_loading_time() {}
    virtual double get_data_loading_time() const { return 0.0; }
    virtual int get_frames_between_iterations() const { return -1; }
    virtual std::pair<bool, MaskImage> segmentation(const bool& blocking) = 0;
};

// Reads masks/<set>/<object>_<index>.pgm (binary P5; PNG decoding is host IO outside the scope)
class DatasetImageSegmentation : public Segmentation {  // DatasetImageSegmentation.cpp:25-147
public:
    DatasetImageSegmentation(const std::string& dataset_path, const std::string& format, std::size_t width, std::size_t height,
                             const std::string& segmentation_set, const std::string& object_name, std::size_t heading_zeros,
                             std::size_t index_offset);
    bool reset() override;
    bool step_frame() override;
    bool is_stepping_required() const override { return true; }
    std::pair<bool, MaskImage> segmentation(const bool& blocking) override;
    double get_data_loading_time() const override { return data_loading_time_; }
    void reset_data_loading_time() override { data_loading_time_ = 0.0; }

protected:
    std::pair<bool, MaskImage> read_file(std::size_t index);  // :128-147
    int head_;
    double data_loading_time_ = 0.0;

private:
    std::string dataset_path_, format_, object_name_;
    std::size_t width_, height_, heading_zeros_, index_offset_;
};

class DatasetImageSegmentationDelayed : public DatasetImageSegmentation {  // DatasetImageSegmentationDelayed.cpp:18-81
public:
    DatasetImageSegmentationDelayed(float fps, float simulated_fps, bool simulate_inference_time, const std::string& dataset_path,
                                    const std::string& format, std::size_t width, std::size_t height,
                                    const std::string& segmentation_set, const std::string& object_name,
                                    std::size_t heading_zeros, std::size_t index_offset);
    std::pair<bool, MaskImage> segmentation(const bool& blocking) override;
    int get_frames_between_iterations() const override { return int(fps_ / simulated_fps_); }
    // the schedule on its own (frame index delivered at `head`, or -1), used by tests
    static int delivered_index(int head, int delay, int head_0, bool simulate_inference_time);

private:
    float fps_, simulated_fps_;
    bool simulate_inference_time_;
    int head_0_, delay_;
};

struct CameraParameters { std::size_t width = 0, height = 0; double fx = 0, fy = 0, cx = 0, cy = 0; };

// RobotsIO::Camera::DatasetCamera + ROFT::CameraMeasurement in one (CameraMeasurement.cpp:28-94): depth/<i>.float,
// data.txt stamps
class CameraMeasurement {
public:
    CameraMeasurement(const std::string& path, const CameraParameters& parameters, std::size_t heading_zeros, std::size_t index_offset);
    bool freeze();                                             // step_frame + read depth (CameraMeasurementType::RGBD)
    std::pair<bool, const DepthImage*> measure() const;        // depth of the frozen frame
    std::pair<bool, CameraParameters> camera_parameters() const { return {true, parameters_}; }
    std::pair<bool, double> camera_time_stamp_rgb() const { return {stamp_valid_, stamp_}; }
    bool reset();

private:
    std::string path_;
    CameraParameters parameters_;
    std::size_t heading_zeros_, index_offset_;
    int head_;
    std::vector<double> stamps_;
    DepthImage depth_;
    bool valid_ = false, stamp_valid_ = false;
    double stamp_ = 0.0;
};

// RobotsIO::Utils::DatasetTransform(Delayed): poses.txt rows "x y z ax ay az angle", all-zero row = invalid
class DatasetTransformDelayed {
public:
    DatasetTransformDelayed(float fps, float simulated_fps, bool simulate_delay, const std::string& file_path, std::size_t skip_rows,
                            std::size_t skip_cols, std::size_t expected_cols);
    bool freeze(bool blocking);                     // steps; true iff a valid pose is delivered at this frame
    const double* transform() const { return pose_; }  // (x, q wxyz)
    int get_frames_between_iterations() const { return delay_; }
    bool reset();

private:
    std::vector<std::vector<double>> rows_;
    int head_ = -1, head_0_ = 0, delay_ = 0;
    bool simulate_ = true;
    double pose_[7] = {0, 0, 0, 1, 0, 0, 0};
};

// ---- the filter -------------------------------------------------------------------------------------------
// 8-bit plane of a greyscale PNG (colour type 0, 8 or 16 bit, non-interlaced): cv::imread + convertTo(CV_8UC1) for masks
bool read_png_gray8(const std::string& file_name, struct MaskImage& out);

// Wavefront OBJ triangles (v / f records, polygons fan-triangulated, 1-based or negative indices): the mesh input of the
// depth renderer.  Stands in for MeshResource + Assimp (SICAD's model loading); throws std::runtime_error.
void read_obj_mesh(const std::string& path, std::vector<float>& vertices, std::vector<std::int32_t>& faces);

struct TrackSources {  // what main.cpp:328-388 wires into one ROFTFilter
    std::shared_ptr<CameraMeasurement> camera;
    std::shared_ptr<Segmentation> segmentation;
    std::shared_ptr<ImageOpticalFlowSource> flow;
    std::shared_ptr<DatasetTransformDelayed> pose;
    std::vector<double> initial_condition_p;  // 13: v, w, x, q(wxyz)
    std::vector<double> initial_condition_v;  // 6
};

class ROFTFilter {  // ROFTFilter.h:42-73, ROFTFilter.cpp:32-452; argument order and meaning as the reference's ctor
public:
    ROFTFilter(std::vector<TrackSources> tracks, const std::vector<double>& initial_covariance_p,
               const std::vector<double>& model_covariance_p, const std::vector<double>& measurement_covariance_p,
               const std::vector<double>& initial_covariance_v, const std::vector<double>& model_covariance_v,
               const std::vector<double>& measurement_covariance_v, double ut_alpha, double ut_beta, double ut_kappa,
               double sample_time, bool pose_meas, bool pose_resync, bool velocity_meas, bool flow_weighting,
               bool flow_aided_segmentation, double maximum_depth, double subsampling_radius, bool enable_log,
               const std::string& log_path, const std::string& log_prefix, int device = 0,
               // render-and-compare pose outlier rejection (ROFTFilter.cpp:52-54, 184-199): the reference takes them after
               // pose_resync; they trail here so that existing callers compile.  model_mesh_path: Wavefront OBJ of the object
               // (what ModelParameters / MeshResource resolve to); the gain is a `const bool` in the reference
               bool pose_outlier_rejection = false, bool pose_outlier_rejection_gain = false, const std::string& model_mesh_path = "");
    ~ROFTFilter();
    bool initialization_step();            // ROFTFilter.cpp:216-237
    bool filtering_step();                 // ROFTFilter.cpp:255-452; false = teardown (no depth)
    bool run_condition() { return true; }
    std::vector<std::string> log_file_names(const std::string& prefix_path, const std::string& prefix_name);  // :247-252
    // beliefs after the last step, [n_tracks][13] / [n_tracks][6]
    const std::vector<double>& pose_mean() const { return p_mean_; }
    const std::vector<double>& velocity_mean() const { return v_mean_; }
    std::size_t n_tracks() const { return tracks_.size(); }

private:
    std::vector<TrackSources> tracks_;
    roftb_ctx* ctx_ = nullptr;
    roftb_config cfg_;
    double sample_time_;
    std::vector<double> last_camera_stamp_;
    std::vector<double> p_mean_, v_mean_;
    bool enable_log_;
    std::vector<std::ofstream> log_pose_, log_velocity_, log_time_;
    // staging buffers for one frame of all tracks
    std::vector<float> depth_;
    std::vector<std::uint8_t> flow_, mask_, flow_valid_, mask_valid_, pose_valid_;
    std::vector<double> pose_, dt_;
    std::size_t flow_bytes_per_track_ = 0;
};

}  // namespace ROFT
