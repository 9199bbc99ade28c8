// roft_adapters.h - the reference's L2 plug-in classes, same names and virtual overrides, each forwarding
// to one C-ABI operator of libroft_b200.so.  They are what lets ROFTFilter::filtering_step (ROFTFilter.cpp:255-367) run
// UNCHANGED over the B200 kernels: tests/cpp/adapter_check.cpp transcribes that call sequence over these classes and
// compares every frame with the fused batched loop (roftb_filter_step).
//
//   ROFT::ImageSegmentationOFAidedSource<T>   include/ROFT/ImageSegmentationOFAidedSource.hpp:37-53,128-281  -> roftb_mask_sync
//   ROFT::ImageSegmentationOFAidedSourceStamped<T>  include/ROFT/ImageSegmentationOFAidedSourceStamped.hpp:153-318   -> roftb_mask_sync
//   ROFT::OpticalFlowQueueHandler             src/OpticalFlowQueueHandler.cpp:18-57 (time-stamped flow queue, host)
//   ROFT::ImageSegmentationMeasurement        src/ImageSegmentationMeasurement.cpp:30-81 (threshold :61-65)
//   ROFT::ImageOpticalFlowMeasurement<T>      include/ROFT/ImageOpticalFlowMeasurement.hpp:47-71,168-375     -> roftb_flow_measurement_export
//   ROFT::SKFCorrection                       include/ROFT/SKFCorrection.h:26-33, src/SKFCorrection.cpp:37-153 -> roftb_flow_velocity
//   ROFT::SpatialVelocityModel                src/SpatialVelocityModel.cpp:15-27 (with bfl::KFPrediction)
//   ROFT::CartesianQuaternionModel            include/ROFT/CartesianQuaternionModel.h:26-42 (parameters + sampling time)
//   ROFT::UKFPrediction                       replaces bfl::UKFPrediction(CartesianQuaternionModel), ROFTFilter.cpp:163-166 -> roftb_ukf_predict
//   ROFT::CartesianQuaternionMeasurement      include/ROFT/CartesianQuaternionMeasurement.h:33-49, .cpp:92-348 (mode machine, host)
//   ROFT::UKFCorrection                       include/ROFT/UKFCorrection.h:27-38, src/UKFCorrection.cpp:54-133 -> roftb_ukf_correct
//   RobotsIO::Utils::SpatialVelocityBuffer    (UPSTREAM-RECALL) set_twist / freeze / linear_velocity_origin / angular_velocity
//
// cv::Mat / Eigen::MatrixXf image payloads are the POD stand-ins of roft_host.h (MaskImage, FlowFrame, DepthImage); with
// OpenCV / Eigen present the bodies are the same with `.data` read from the cv::Mat / MatrixXf.  T = cv::Vec2f or
// cv::Vec2s only selects the flow element type, as in the reference (ROFTFilter.cpp:122-149).
#pragma once

#include <deque>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "bfl_shim.h"
#include "roft_host.h"

namespace cv {
struct Vec2f { float v[2]; };
struct Vec2s { short v[2]; };
}  // namespace cv

namespace RobotsIO {
namespace Utils {
class SpatialVelocityBuffer {  // UPSTREAM-RECALL: holds the last twist handed over by the velocity filter
public:
    void set_twist(const double* linear, const double* angular) {
        for (int i = 0; i < 3; ++i) { v_[i] = linear[i]; w_[i] = angular[i]; }
        fresh_ = true;
    }
    bool freeze(bool) { const bool f = fresh_; fresh_ = false; return f; }
    const double* linear_velocity_origin() const { return v_; }
    const double* angular_velocity() const { return w_; }

private:
    double v_[3] = {0, 0, 0}, w_[3] = {0, 0, 0};
    bool fresh_ = false;
};
}  // namespace Utils
}  // namespace RobotsIO

namespace ROFT {

// one single-track context of libroft_b200.so shared by the adapters of a filter
class B200Context {
public:
    B200Context(const CameraParameters& camera, int flow_type, std::size_t flow_grid, float flow_scale, double subsampling_radius,
                double maximum_depth, bool flow_weighting, const double* cov_flow, const double* sigma_pose_model,
                const double* sigma_pose_measurement, double ut_alpha, double ut_beta, double ut_kappa, int segm_delay, int device = 0);
    ~B200Context();
    roftb_ctx* get() const { return ctx_; }
    const roftb_config& config() const { return cfg_; }

private:
    roftb_ctx* ctx_ = nullptr;
    roftb_config cfg_;
};

// ---- segmentation ----------------------------------------------------------------------------------------
template <class T>
class ImageSegmentationOFAidedSource : public Segmentation {
public:
    ImageSegmentationOFAidedSource(std::shared_ptr<Segmentation> segmentation_source, std::shared_ptr<ImageOpticalFlowSource> flow_source,
                                   const CameraParameters& camera_parameters, const bool& wait_source_initialization,
                                   std::shared_ptr<B200Context> ctx);
    bool step_frame() override;                                   // hpp:128-231
    bool is_stepping_required() const override { return true; }
    bool reset() override;
    void reset_data_loading_time() override { segmentation_->reset_data_loading_time(); }
    double get_data_loading_time() const override { return segmentation_->get_data_loading_time(); }
    int get_frames_between_iterations() const override { return segmentation_->get_frames_between_iterations(); }
    std::pair<bool, MaskImage> segmentation(const bool& blocking = false) override;  // hpp:284-295

private:
    bool warp(const std::vector<const FlowFrame*>& flows, bool zero_origin);  // map() + cv::remap, hpp:235-281
    std::shared_ptr<Segmentation> segmentation_;
    std::shared_ptr<ImageOpticalFlowSource> flow_;
    std::shared_ptr<B200Context> ctx_;
    std::vector<FlowFrame> flow_buffer_;
    MaskImage mask_;
    bool segmentation_available_ = false, is_first_frame_ = true;
    int segm_frames_between_iterations_;
};

// time-stamped queue of the most recent flow frames (OpticalFlowQueueHandler.cpp:18-57)
class OpticalFlowQueueHandler {
public:
    explicit OpticalFlowQueueHandler(const std::size_t& window_size) : window_size_(window_size) {}
    void add_flow(const FlowFrame& frame, const double& time_stamp);                       // :18-26
    std::vector<const FlowFrame*> get_buffer_region(const double& initial_time_stamp);     // :29-51: the frames AFTER the stamped one
    void clear() { buffer_.clear(); }
    std::size_t size() const { return buffer_.size(); }

private:
    struct Entry { FlowFrame frame; double timestamp; };
    std::size_t window_size_;
    std::deque<Entry> buffer_;
};

// a segmentation source that stamps its masks (RobotsIO::Utils::Segmentation::get_time_stamp, UPSTREAM-RECALL)
class StampedSegmentation : public Segmentation {
public:
    virtual double get_time_stamp() = 0;
};

// asynchronous (time-stamp matched) variant: a new mask is warped through the queued flows that FOLLOW the frame it was
// computed on; host side of ImageSegmentationOFAidedSourceStamped.hpp:153-268, same kernels as the un-stamped source
template <class T>
class ImageSegmentationOFAidedSourceStamped : public Segmentation {
public:
    ImageSegmentationOFAidedSourceStamped(std::shared_ptr<StampedSegmentation> segmentation_source, std::shared_ptr<ImageOpticalFlowSource> flow_source,
                                          const CameraParameters& camera_parameters, const bool& wait_source_initialization,
                                          std::shared_ptr<B200Context> ctx);
    void set_rgb_image_time_stamp(const double& timestamp) { rgb_image_time_stamp_ = timestamp; }  // set_rgb_image(image, stamp), hpp:145-150
    bool step_frame() override;                                   // hpp:153-268
    bool is_stepping_required() const override { return true; }
    bool reset() override;
    int get_frames_between_iterations() const override { return segmentation_->get_frames_between_iterations(); }
    std::pair<bool, MaskImage> segmentation(const bool& blocking = false) override { (void)blocking; return std::make_pair(segmentation_available_, mask_); }

private:
    void warp(const std::vector<const FlowFrame*>& flows, bool zero_origin);  // map() hpp:272-318 + cv::remap
    std::shared_ptr<StampedSegmentation> segmentation_;
    std::shared_ptr<ImageOpticalFlowSource> flow_;
    std::shared_ptr<B200Context> ctx_;
    OpticalFlowQueueHandler flow_handler_{30};  // flow_queue_max_size_, hpp:99
    MaskImage mask_;
    double rgb_image_time_stamp_ = 0.0;
    bool segmentation_available_ = false, is_first_frame_ = true;
    int segm_frames_between_iterations_;
};

class ImageSegmentationMeasurement {  // (bfl::MeasurementModel in the reference; only freeze / measure are implemented there)
public:
    explicit ImageSegmentationMeasurement(std::shared_ptr<Segmentation> segmentation_source) : segmentation_source_(std::move(segmentation_source)) {}
    bool freeze(const bfl::Data& data = bfl::Data());            // .cpp:30-75
    std::pair<bool, bfl::Data> measure(const bfl::Data& data = bfl::Data()) const;  // pair<bool new, MaskImage>
    void reset();
    double get_data_loading_time() const { return segmentation_source_->get_data_loading_time(); }

private:
    std::shared_ptr<Segmentation> segmentation_source_;
    MaskImage segmentation_;
    bool segmentation_available_ = false, new_segmentation_ = false;
};

// ---- velocity ---------------------------------------------------------------------------------------------
class ImageOpticalFlowMeasurementBase {
public:
    enum class FreezeType { OnlyStepSource, ExceptStepSource, Complete };
};

template <class T>
class ImageOpticalFlowMeasurement : public bfl::LinearMeasurementModel, public ImageOpticalFlowMeasurementBase {
public:
    ImageOpticalFlowMeasurement(std::shared_ptr<ImageOpticalFlowSource> flow_source, std::shared_ptr<CameraMeasurement> camera_measurement,
                                std::shared_ptr<ImageSegmentationMeasurement> segmentation, const std::size_t& segmentation_radius,
                                const double& maximum_depth, const Eigen::Ref<const Eigen::MatrixXd>& covariance,
                                const bool use_full_covariance_matrix, std::shared_ptr<B200Context> ctx);
    bool freeze(const bfl::Data& data = bfl::Data()) override;                                       // hpp:168-294
    std::pair<bool, bfl::Data> measure(const bfl::Data& data = bfl::Data()) const override;           // hpp:297-301
    std::pair<bool, bfl::Data> predictedMeasure(const Eigen::Ref<const Eigen::MatrixXd>& cur_states) const override;  // :305-311
    std::pair<bool, bfl::Data> innovation(const bfl::Data& predicted_measurements, const bfl::Data& measurements) const override;
    Eigen::MatrixXd getMeasurementMatrix() const override;
    std::pair<bool, Eigen::MatrixXd> getNoiseCovarianceMatrix() const override;
    bfl::VectorDescription getInputDescription() const override;
    bfl::VectorDescription getMeasurementDescription() const override;
    bool setProperty(const std::string& property) override;                                           // :360-375
    // what the fused correction needs instead of the materialised (z, H): the frame the measurement is taken on
    const MaskImage& measurement_mask() const { return used_segmentation_; }
    const DepthImage& measurement_depth() const { return used_depth_; }
    const FlowFrame* measurement_flow() const { return used_flow_; }
    double sample_time() const { return sample_time_; }
    void set_valid_count(int n) { n_valid_ = n; }
    std::shared_ptr<B200Context> context() const { return ctx_; }

private:
    void materialise() const;  // z, H through roftb_flow_measurement_export (only when a bfl-style caller asks for them)
    std::shared_ptr<ImageOpticalFlowSource> flow_;
    std::shared_ptr<CameraMeasurement> camera_;
    std::shared_ptr<ImageSegmentationMeasurement> segmentation_;
    std::shared_ptr<B200Context> ctx_;
    Eigen::MatrixXd covariance_;
    MaskImage previous_segmentation_, used_segmentation_;
    DepthImage previous_depth_, used_depth_;
    const FlowFrame* used_flow_ = nullptr;
    mutable Eigen::MatrixXd measurement_matrix_, measurement_;
    mutable bool materialised_ = false;
    int n_valid_ = 0;
    double sample_time_ = 0.0;
    bool flow_available_ = false, is_first_frame_ = true;
    FreezeType freeze_type_ = FreezeType::Complete;
};

class SpatialVelocityModel : public bfl::LinearStateModel {  // SpatialVelocityModel.cpp:15-27: F = I6, Q = diag(sigma_v, sigma_w)
public:
    SpatialVelocityModel(const Eigen::Ref<const Eigen::MatrixXd>& sigma_v, const Eigen::Ref<const Eigen::MatrixXd>& sigma_w);
    Eigen::MatrixXd getStateTransitionMatrix() override { return F_; }
    Eigen::MatrixXd getNoiseCovarianceMatrix() override { return Q_; }
    bfl::VectorDescription getInputDescription() override { return bfl::VectorDescription(6, 0, 6); }
    bfl::VectorDescription getStateDescription() override { return bfl::VectorDescription(6, 0); }

private:
    Eigen::MatrixXd F_, Q_;
};

class SKFCorrection : public bfl::GaussianCorrection {
public:
    SKFCorrection(std::unique_ptr<bfl::LinearMeasurementModel> measurement_model, const std::size_t measurement_sub_size,
                  const bool use_laplacian_reweighting = false);
    bfl::MeasurementModel& getMeasurementModel() override { return *measurement_model_; }

protected:
    void correctStep(const bfl::GaussianMixture& pred_state, bfl::GaussianMixture& corr_state) override;  // SKFCorrection.cpp:37-153

private:
    std::unique_ptr<bfl::LinearMeasurementModel> measurement_model_;
    std::size_t measurement_sub_size_;
    bool use_laplacian_reweighting_;
};

// ---- pose -------------------------------------------------------------------------------------------------
class CartesianQuaternionModel : public bfl::StateModel {  // parameters and sampling time (CartesianQuaternionModel.cpp:127-170)
public:
    CartesianQuaternionModel(const Eigen::Ref<const Eigen::MatrixXd>& psd_linear_acceleration,
                             const Eigen::Ref<const Eigen::MatrixXd>& sigma_angular_velocity, const double sample_time);
    bool setSamplingTime(const double& sample_time) override { sample_time_ = sample_time; return true; }
    double sampling_time() const { return sample_time_; }
    Eigen::MatrixXd getNoiseCovarianceMatrix() override;          // Q(T), .cpp:127-141
    bfl::VectorDescription getInputDescription() override { return bfl::VectorDescription(9, 1, 9, bfl::VectorDescription::CircularType::Quaternion); }
    bfl::VectorDescription getStateDescription() override { return bfl::VectorDescription(9, 1, 0, bfl::VectorDescription::CircularType::Quaternion); }

private:
    Eigen::MatrixXd psd_, sigma_w_;
    double sample_time_;
};

class UKFPrediction : public bfl::GaussianPrediction {
public:
    UKFPrediction(std::unique_ptr<CartesianQuaternionModel> state_model, std::shared_ptr<B200Context> ctx);
    bfl::StateModel& getStateModel() override { return *state_model_; }

protected:
    void predictStep(const bfl::GaussianMixture& prev_state, bfl::GaussianMixture& pred_state) override;

private:
    std::unique_ptr<CartesianQuaternionModel> state_model_;
    std::shared_ptr<B200Context> ctx_;
};

class CartesianQuaternionMeasurement : public bfl::MeasurementModel {
public:
    enum class MeasurementMode { Standard, RepeatOnlyVelocity, PopBufferedMeasurement };
    CartesianQuaternionMeasurement(std::shared_ptr<DatasetTransformDelayed> pose_measurement,
                                   std::shared_ptr<RobotsIO::Utils::SpatialVelocityBuffer> velocity_measurement, const bool use_screw_velocity,
                                   const bool use_pose_measurement, const bool use_velocity_measurement);
    bool freeze(const bfl::Data& data = bfl::Data()) override;                               // .cpp:92-348
    std::pair<bool, bfl::Data> measure(const bfl::Data& data = bfl::Data()) const override;   // the measurement vector (6, 7 or 13)
    std::pair<bool, bfl::Data> predictedMeasure(const Eigen::Ref<const Eigen::MatrixXd>&) const override;  // fused into roftb_ukf_correct
    std::pair<bool, bfl::Data> innovation(const bfl::Data&, const bfl::Data&) const override;
    bfl::VectorDescription getInputDescription() const override { return input_description_; }
    bfl::VectorDescription getMeasurementDescription() const override { return measurement_description_; }
    int measurement_type() const { return int(measurement_type_); }  // ROFTB_MEAS_*

private:
    enum class MeasurementType { None = ROFTB_MEAS_NONE, Velocity = ROFTB_MEAS_VELOCITY, Pose = ROFTB_MEAS_POSE, PoseVelocity = ROFTB_MEAS_POSE_VELOCITY };
    void set_type(MeasurementType t);
    std::shared_ptr<DatasetTransformDelayed> pose_measurement_;
    std::shared_ptr<RobotsIO::Utils::SpatialVelocityBuffer> velocity_measurement_;
    Eigen::MatrixXd measurement_;
    double last_linear_velocity_[3] = {0, 0, 0}, last_angular_velocity_[3] = {0, 0, 0};
    double last_pose_[7] = {0, 0, 0, 1, 0, 0, 0};
    std::deque<std::vector<double>> buffer_velocities_;
    MeasurementType measurement_type_ = MeasurementType::None;
    bfl::VectorDescription input_description_, measurement_description_;
    bool is_first_velocity_in_ = false, is_pose_ = false;
    bool use_pose_measurement_, use_velocity_measurement_;
    int pose_frames_between_iterations_;
};

class UKFCorrection : public bfl::GaussianCorrection {
public:
    UKFCorrection(std::unique_ptr<bfl::MeasurementModel> meas_model, const double alpha, const double beta, const double kappa,
                  std::shared_ptr<B200Context> ctx);
    bfl::MeasurementModel& getMeasurementModel() override { return *measurement_model_; }

protected:
    void correctStep(const bfl::GaussianMixture& pred_state, bfl::GaussianMixture& corr_state) override;  // UKFCorrection.cpp:54-133

private:
    std::unique_ptr<bfl::MeasurementModel> measurement_model_;
    std::shared_ptr<B200Context> ctx_;
};

}  // namespace ROFT
