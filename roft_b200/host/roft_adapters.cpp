// roft_adapters.cpp - see roft_adapters.h.  Each adapter keeps the host-side state machine of the reference class it
// names (transcribed from the cited lines) and hands the per-pixel / sigma-point work to one operator of the C ABI.
#include "roft_adapters.h"

#include <cstring>
#include <iostream>

namespace ROFT {

namespace {
void check(int rc, roftb_ctx* ctx, const char* what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + roftb_last_error(ctx));
}
}  // namespace

// ---- context --------------------------------------------------------------------------------------------
B200Context::B200Context(const CameraParameters& camera, int flow_type, std::size_t flow_grid, float flow_scale, double subsampling_radius,
                         double maximum_depth, bool flow_weighting, const double* cov_flow, const double* sigma_pose_model,
                         const double* sigma_pose_measurement, double ut_alpha, double ut_beta, double ut_kappa, int segm_delay, int device) {
    roftb_config_default(&cfg_);
    cfg_.n_tracks = 1;
    cfg_.width = int(camera.width); cfg_.height = int(camera.height);
    cfg_.fx = camera.fx; cfg_.fy = camera.fy; cfg_.cx = camera.cx; cfg_.cy = camera.cy;
    cfg_.flow_format = flow_type; cfg_.flow_grid = int(flow_grid); cfg_.flow_scale = flow_scale;
    cfg_.subsampling_radius = int(subsampling_radius);
    cfg_.depth_maximum = maximum_depth;
    cfg_.weight_flow = flow_weighting ? 1 : 0;
    cfg_.cov_flow[0] = cov_flow[0]; cfg_.cov_flow[1] = cov_flow[1];
    for (int i = 0; i < 3; ++i) {
        // model: (sigma_angular, psd_linear) as ROFTFilter.cpp:89-90 unpacks; measurement: (v, w, x, q) as :97-104
        cfg_.p_sigma_angular[i] = sigma_pose_model[i]; cfg_.p_sigma_linear[i] = sigma_pose_model[3 + i];
        cfg_.cov_v[i] = sigma_pose_measurement[i]; cfg_.cov_w[i] = sigma_pose_measurement[3 + i];
        cfg_.cov_x[i] = sigma_pose_measurement[6 + i]; cfg_.cov_q[i] = sigma_pose_measurement[9 + i];
    }
    cfg_.ut_alpha = ut_alpha; cfg_.ut_beta = ut_beta; cfg_.ut_kappa = ut_kappa;
    cfg_.segm_delay = segm_delay;
    cfg_.device = device;
    if (roftb_create(&cfg_, &ctx_)) throw std::runtime_error(std::string("B200Context: ") + roftb_last_error(nullptr));
}
B200Context::~B200Context() { roftb_destroy(ctx_); }

// ---- ImageSegmentationOFAidedSource ------------------------------------------------------------------------
template <class T>
ImageSegmentationOFAidedSource<T>::ImageSegmentationOFAidedSource(std::shared_ptr<Segmentation> segmentation_source,
                                                                  std::shared_ptr<ImageOpticalFlowSource> flow_source,
                                                                  const CameraParameters&, const bool&, std::shared_ptr<B200Context> ctx)
    : segmentation_(std::move(segmentation_source)), flow_(std::move(flow_source)), ctx_(std::move(ctx)) {
    segm_frames_between_iterations_ = segmentation_->get_frames_between_iterations();  // hpp:107
}

template <class T>
bool ImageSegmentationOFAidedSource<T>::reset() {
    segmentation_available_ = false;
    is_first_frame_ = true;
    flow_buffer_.clear();
    return segmentation_->reset();
}

template <class T>
bool ImageSegmentationOFAidedSource<T>::warp(const std::vector<const FlowFrame*>& flows, bool zero_origin) {
    // map(): only the last `segm_frames_between_iterations_` flows are chained (hpp:239-245)
    std::size_t start = 0;
    if (segm_frames_between_iterations_ > 0 && flows.size() > std::size_t(segm_frames_between_iterations_))
        start = flows.size() - std::size_t(segm_frames_between_iterations_);
    const std::size_t n = flows.size() - start;
    std::vector<std::uint8_t> packed;
    for (std::size_t j = start; j < flows.size(); ++j) packed.insert(packed.end(), flows[j]->data.begin(), flows[j]->data.end());
    MaskImage out = mask_;
    check(roftb_mask_sync(ctx_->get(), 1, mask_.data.data(), n ? packed.data() : nullptr, int(n), zero_origin ? 1 : 0, out.data.data(), nullptr),
          ctx_->get(), "ImageSegmentationOFAidedSource::map");
    mask_ = std::move(out);
    return true;
}

template <class T>
bool ImageSegmentationOFAidedSource<T>::step_frame() {
    if (segmentation_->is_stepping_required()) segmentation_->step_frame();
    bool valid_segmentation = false;
    MaskImage mask;
    std::tie(valid_segmentation, mask) = segmentation_->segmentation(false);
    if (!segmentation_available_ && valid_segmentation) {  // hpp:169-178: initialisation, not treated as a new mask
        segmentation_available_ = true;
        mask_ = mask;
        valid_segmentation = false;
    }
    if (valid_segmentation) {  // hpp:180-197: an uninformative mask is skipped
        bool any = false;
        for (std::uint8_t b : mask.data) any |= b != 0;
        if (!any) {
            valid_segmentation = false;
            if (segm_frames_between_iterations_ <= 0) flow_buffer_.clear();
        }
    }
    bool valid_flow = false;
    const FlowFrame* flow = nullptr;
    std::tie(valid_flow, flow) = flow_->flow(false);
    valid_flow &= !is_first_frame_;
    if (valid_flow) flow_buffer_.push_back(*flow);  // flow.clone() (hpp:208)
    if (valid_segmentation) {  // hpp:211-219
        mask_ = mask;
        std::vector<const FlowFrame*> fl;
        for (const FlowFrame& f : flow_buffer_) fl.push_back(&f);
        warp(fl, false);
        flow_buffer_.clear();
    } else if (valid_flow) {  // hpp:221-226
        warp({flow}, true);
    }
    is_first_frame_ = false;
    return true;
}

template <class T>
std::pair<bool, MaskImage> ImageSegmentationOFAidedSource<T>::segmentation(const bool&) {
    return std::make_pair(segmentation_available_, mask_);
}

template class ImageSegmentationOFAidedSource<cv::Vec2f>;
template class ImageSegmentationOFAidedSource<cv::Vec2s>;

// ---- OpticalFlowQueueHandler / ImageSegmentationOFAidedSourceStamped -----------------------------------------
void OpticalFlowQueueHandler::add_flow(const FlowFrame& frame, const double& time_stamp) {
    buffer_.push_back(Entry{frame, time_stamp});
    if (buffer_.size() > window_size_) buffer_.pop_front();  // enforce the maximum size
}

std::vector<const FlowFrame*> OpticalFlowQueueHandler::get_buffer_region(const double& initial_time_stamp) {
    std::vector<const FlowFrame*> output_region;
    std::size_t index;
    bool found = false;
    for (index = 0; index < buffer_.size(); index++)
        if (std::fabs(buffer_[index].timestamp - initial_time_stamp) < 1e-3) {
            found = true;
            break;
        }
    if (!found) return output_region;
    index++;  // the optical flow always refers to the previous RGB image, thus the next frame is needed
    for (; index < buffer_.size(); index++) output_region.push_back(&buffer_[index].frame);
    return output_region;
}

template <class T>
ImageSegmentationOFAidedSourceStamped<T>::ImageSegmentationOFAidedSourceStamped(std::shared_ptr<StampedSegmentation> segmentation_source,
                                                                                std::shared_ptr<ImageOpticalFlowSource> flow_source,
                                                                                const CameraParameters&, const bool&, std::shared_ptr<B200Context> ctx)
    : segmentation_(std::move(segmentation_source)), flow_(std::move(flow_source)), ctx_(std::move(ctx)) {
    segm_frames_between_iterations_ = segmentation_->get_frames_between_iterations();
}

template <class T>
bool ImageSegmentationOFAidedSourceStamped<T>::reset() {
    segmentation_available_ = false;
    is_first_frame_ = true;
    flow_handler_.clear();
    return segmentation_->reset();
}

template <class T>
void ImageSegmentationOFAidedSourceStamped<T>::warp(const std::vector<const FlowFrame*>& flows, bool zero_origin) {
    std::size_t start = 0;  // map(): only the last `segm_frames_between_iterations_` flows are chained (hpp:276-282)
    if (segm_frames_between_iterations_ > 0 && flows.size() > std::size_t(segm_frames_between_iterations_))
        start = flows.size() - std::size_t(segm_frames_between_iterations_);
    const std::size_t n = flows.size() - start;
    if (n > ROFTB_MAX_CHAIN) throw std::runtime_error("ImageSegmentationOFAidedSourceStamped::map: flow chain longer than ROFTB_MAX_CHAIN");
    std::vector<std::uint8_t> packed;
    for (std::size_t j = start; j < flows.size(); ++j) packed.insert(packed.end(), flows[j]->data.begin(), flows[j]->data.end());
    MaskImage out = mask_;
    check(roftb_mask_sync(ctx_->get(), 1, mask_.data.data(), n ? packed.data() : nullptr, int(n), zero_origin ? 1 : 0, out.data.data(), nullptr),
          ctx_->get(), "ImageSegmentationOFAidedSourceStamped::map");
    mask_ = std::move(out);
}

template <class T>
bool ImageSegmentationOFAidedSourceStamped<T>::step_frame() {
    if (segmentation_->is_stepping_required()) segmentation_->step_frame();
    bool valid_segmentation = false;
    MaskImage mask;
    std::tie(valid_segmentation, mask) = segmentation_->segmentation(false);
    const double mask_time_stamp = segmentation_->get_time_stamp();
    if (!segmentation_available_ && valid_segmentation) {  // hpp:212-221: initialisation, not treated as a new mask
        segmentation_available_ = true;
        mask_ = mask;
        valid_segmentation = false;
    }
    if (valid_segmentation) {  // hpp:223-230: an uninformative mask is skipped
        bool any = false;
        for (std::uint8_t b : mask.data) any |= b != 0;
        if (!any) valid_segmentation = false;
    }
    bool valid_flow = false;
    const FlowFrame* flow = nullptr;
    std::tie(valid_flow, flow) = flow_->flow(false);
    valid_flow &= !is_first_frame_;
    if (valid_flow) flow_handler_.add_flow(*flow, rgb_image_time_stamp_);  // hpp:237-241
    if (valid_segmentation) {  // hpp:243-258
        mask_ = mask;
        const std::vector<const FlowFrame*> buffer = flow_handler_.get_buffer_region(mask_time_stamp);
        if (!buffer.empty()) {
            warp(buffer, false);
        } else if (flow) {
            warp({flow}, true);  // no queued flow follows the mask's frame: propagate with the current flow (mask_(0,0) = 0 first)
        }
    } else if (valid_flow) {  // hpp:260-265
        warp({flow}, true);
    }
    is_first_frame_ = false;
    return true;
}

template class ImageSegmentationOFAidedSourceStamped<cv::Vec2f>;
template class ImageSegmentationOFAidedSourceStamped<cv::Vec2s>;

// ---- ImageSegmentationMeasurement --------------------------------------------------------------------------
bool ImageSegmentationMeasurement::freeze(const bfl::Data&) {
    if (segmentation_source_->is_stepping_required()) segmentation_source_->step_frame();
    MaskImage segmentation;
    new_segmentation_ = false;
    std::tie(new_segmentation_, segmentation) = segmentation_source_->segmentation(false);
    if (new_segmentation_) {
        segmentation_available_ = true;
        segmentation_ = std::move(segmentation);
        for (std::uint8_t& b : segmentation_.data) b = b > 1 ? 255 : 0;  // cv::threshold(.., 1, 255, THRESH_BINARY), .cpp:61-65
    }
    return segmentation_available_;
}
std::pair<bool, bfl::Data> ImageSegmentationMeasurement::measure(const bfl::Data&) const {
    return std::make_pair(segmentation_available_, bfl::Data(std::make_pair(new_segmentation_, segmentation_)));
}
void ImageSegmentationMeasurement::reset() {
    segmentation_available_ = false;
    new_segmentation_ = false;
    segmentation_source_->reset();
}

// ---- ImageOpticalFlowMeasurement ---------------------------------------------------------------------------
template <class T>
ImageOpticalFlowMeasurement<T>::ImageOpticalFlowMeasurement(std::shared_ptr<ImageOpticalFlowSource> flow_source,
                                                            std::shared_ptr<CameraMeasurement> camera_measurement,
                                                            std::shared_ptr<ImageSegmentationMeasurement> segmentation, const std::size_t&,
                                                            const double&, const Eigen::Ref<const Eigen::MatrixXd>& covariance, const bool,
                                                            std::shared_ptr<B200Context> ctx)
    : flow_(std::move(flow_source)), camera_(std::move(camera_measurement)), segmentation_(std::move(segmentation)), ctx_(std::move(ctx)),
      covariance_(covariance) {}

template <class T>
bool ImageOpticalFlowMeasurement<T>::freeze(const bfl::Data& data) {
    std::tie(freeze_type_, sample_time_) = bfl::any::any_cast<std::pair<FreezeType, double>>(data);
    if (freeze_type_ != FreezeType::ExceptStepSource)
        if (flow_->is_stepping_required()) flow_->step_frame();
    if (freeze_type_ == FreezeType::OnlyStepSource) return true;
    bfl::Data segmentation_data;
    bool valid_segmentation = false;
    std::tie(valid_segmentation, segmentation_data) = segmentation_->measure();
    if (!valid_segmentation) return false;
    const MaskImage segmentation = bfl::any::any_cast<std::pair<bool, MaskImage>>(segmentation_data).second;
    bool valid_data = false;
    const DepthImage* depth = nullptr;
    std::tie(valid_data, depth) = camera_->measure();
    if (!valid_data) return false;
    flow_available_ = false;
    const FlowFrame* flow = nullptr;
    std::tie(flow_available_, flow) = flow_->flow(false);
    materialised_ = false;
    if (!flow_available_ || is_first_frame_) {  // hpp:217-229
        previous_depth_ = *depth;
        previous_segmentation_ = segmentation;
        is_first_frame_ = false;
        return false;
    }
    // the measurement is taken on the PREVIOUS mask and depth with the current flow (hpp:231-283) ...
    used_segmentation_ = std::move(previous_segmentation_);
    used_depth_ = std::move(previous_depth_);
    used_flow_ = flow;
    // ... which are then replaced by the current ones (hpp:286-287)
    previous_depth_ = *depth;
    previous_segmentation_ = segmentation;
    return flow_available_;
}

template <class T>
void ImageOpticalFlowMeasurement<T>::materialise() const {
    if (materialised_) return;
    int capacity = 0;
    for (std::uint8_t b : used_segmentation_.data) capacity += b != 0;
    std::vector<double> z(std::size_t(capacity) * 2 + 2), H(std::size_t(capacity) * 12 + 12);
    int n = 0;
    check(roftb_flow_measurement_export(ctx_->get(), used_segmentation_.data.data(), used_depth_.data.data(), used_flow_->data.data(), sample_time_,
                                        capacity, z.data(), H.data(), &n),
          ctx_->get(), "ImageOpticalFlowMeasurement::measure");
    measurement_.resize(2 * n, 1);
    measurement_matrix_.resize(2 * n, 6);
    for (int i = 0; i < 2 * n; ++i) {
        measurement_(i, 0) = z[i];
        for (int j = 0; j < 6; ++j) measurement_matrix_(i, j) = H[std::size_t(i) * 6 + j];
    }
    materialised_ = true;
}

template <class T>
std::pair<bool, bfl::Data> ImageOpticalFlowMeasurement<T>::measure(const bfl::Data&) const {
    if (flow_available_) materialise();
    return std::make_pair(flow_available_, bfl::Data(measurement_));
}
template <class T>
std::pair<bool, bfl::Data> ImageOpticalFlowMeasurement<T>::predictedMeasure(const Eigen::Ref<const Eigen::MatrixXd>& cur_states) const {
    if (!flow_available_) return std::make_pair(false, bfl::Data());
    materialise();
    return std::make_pair(true, bfl::Data(Eigen::MatrixXd(measurement_matrix_ * cur_states)));
}
template <class T>
std::pair<bool, bfl::Data> ImageOpticalFlowMeasurement<T>::innovation(const bfl::Data& predicted_measurements, const bfl::Data& measurements) const {
    return std::make_pair(true, bfl::Data(Eigen::MatrixXd(bfl::any::any_cast<Eigen::MatrixXd>(measurements) -
                                                          bfl::any::any_cast<Eigen::MatrixXd>(predicted_measurements))));
}
template <class T>
Eigen::MatrixXd ImageOpticalFlowMeasurement<T>::getMeasurementMatrix() const {
    materialise();
    return measurement_matrix_;
}
template <class T>
std::pair<bool, Eigen::MatrixXd> ImageOpticalFlowMeasurement<T>::getNoiseCovarianceMatrix() const { return std::make_pair(true, covariance_); }
template <class T>
bfl::VectorDescription ImageOpticalFlowMeasurement<T>::getInputDescription() const { return bfl::VectorDescription(6, 0, std::size_t(2 * n_valid_)); }
template <class T>
bfl::VectorDescription ImageOpticalFlowMeasurement<T>::getMeasurementDescription() const { return bfl::VectorDescription(std::size_t(2 * n_valid_), 0); }
template <class T>
bool ImageOpticalFlowMeasurement<T>::setProperty(const std::string& property) {
    if (property == "check_observability") return n_valid_ >= 3;  // hpp:363-366
    if (property == "reset") {
        flow_available_ = false;
        is_first_frame_ = true;
        return true;
    }
    return false;
}
template class ImageOpticalFlowMeasurement<cv::Vec2f>;
template class ImageOpticalFlowMeasurement<cv::Vec2s>;

// ---- velocity model / correction -----------------------------------------------------------------------------
SpatialVelocityModel::SpatialVelocityModel(const Eigen::Ref<const Eigen::MatrixXd>& sigma_v, const Eigen::Ref<const Eigen::MatrixXd>& sigma_w)
    : F_(Eigen::MatrixXd::Identity(6, 6)), Q_(Eigen::MatrixXd::Zero(6, 6)) {
    for (int i = 0; i < 3; ++i) { Q_(i, i) = sigma_v(i, i); Q_(3 + i, 3 + i) = sigma_w(i, i); }
}

SKFCorrection::SKFCorrection(std::unique_ptr<bfl::LinearMeasurementModel> measurement_model, const std::size_t measurement_sub_size,
                             const bool use_laplacian_reweighting)
    : measurement_model_(std::move(measurement_model)), measurement_sub_size_(measurement_sub_size),
      use_laplacian_reweighting_(use_laplacian_reweighting) {}

namespace {
template <class T>
bool skf_fused(bfl::LinearMeasurementModel* mm, const bfl::GaussianMixture& pred, bfl::GaussianMixture& corr) {
    auto* m = dynamic_cast<ImageOpticalFlowMeasurement<T>*>(mm);
    if (!m) return false;
    // sum_j l_j H_j^T R^-1 H_j and sum_j l_j H_j^T R^-1 z_j on the GPU (the per-pixel loop of SKFCorrection.cpp:129-149 in
    // information form, with the Laplacian weights of :91-116 taken against the predicted mean)
    double xp[6], lambda[36], eta[6], dt = m->sample_time();
    for (int i = 0; i < 6; ++i) xp[i] = pred.mean()(i, 0);
    int count = 0;
    roftb_ctx* ctx = m->context()->get();
    check(roftb_flow_velocity(ctx, 1, m->measurement_mask().data.data(), m->measurement_depth().data.data(), m->measurement_flow()->data.data(), xp,
                              &dt, lambda, eta, &count),
          ctx, "SKFCorrection::correctStep");
    m->set_valid_count(count);
    corr = pred;
    if (count == 0) return true;  // "measurement is empty" (SKFCorrection.cpp:60-68)
    Eigen::MatrixXd Pinv = pred.covariance().inverse(), L(6, 6), e(6, 1);
    for (int i = 0; i < 6; ++i) {
        e(i, 0) = eta[i];
        for (int j = 0; j < 6; ++j) L(i, j) = lambda[i * 6 + j];
    }
    const Eigen::MatrixXd P = (Pinv + L).inverse();
    corr.covariance() = P;
    corr.mean() = P * (Pinv * pred.mean() + e);
    return true;
}
}  // namespace

void SKFCorrection::correctStep(const bfl::GaussianMixture& pred_state, bfl::GaussianMixture& corr_state) {
    if (skf_fused<cv::Vec2f>(measurement_model_.get(), pred_state, corr_state)) return;
    if (skf_fused<cv::Vec2s>(measurement_model_.get(), pred_state, corr_state)) return;
    throw std::runtime_error("SKFCorrection::correctStep: the measurement model is not a ROFT::ImageOpticalFlowMeasurement");
}

// ---- pose model / prediction ----------------------------------------------------------------------------------
CartesianQuaternionModel::CartesianQuaternionModel(const Eigen::Ref<const Eigen::MatrixXd>& psd_linear_acceleration,
                                                   const Eigen::Ref<const Eigen::MatrixXd>& sigma_angular_velocity, const double sample_time)
    : psd_(psd_linear_acceleration), sigma_w_(sigma_angular_velocity), sample_time_(sample_time) {}

Eigen::MatrixXd CartesianQuaternionModel::getNoiseCovarianceMatrix() {
    const double T = sample_time_;
    Eigen::MatrixXd Q = Eigen::MatrixXd::Zero(9, 9);
    for (int i = 0; i < 3; ++i) {
        Q(i, i) = psd_(i, i) * T;
        Q(3 + i, 3 + i) = sigma_w_(i, i);
        Q(6 + i, 6 + i) = psd_(i, i) * T * T * T / 3.0;
        Q(i, 6 + i) = Q(6 + i, i) = psd_(i, i) * T * T / 2.0;
    }
    return Q;
}

UKFPrediction::UKFPrediction(std::unique_ptr<CartesianQuaternionModel> state_model, std::shared_ptr<B200Context> ctx)
    : state_model_(std::move(state_model)), ctx_(std::move(ctx)) {}

namespace {
void belief_to_arrays(const bfl::GaussianMixture& g, double* mean, double* cov) {
    for (int i = 0; i < 13; ++i) mean[i] = g.mean()(i, 0);
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) cov[i * 12 + j] = g.covariance()(i, j);
}
void arrays_to_belief(const double* mean, const double* cov, bfl::GaussianMixture& g) {
    for (int i = 0; i < 13; ++i) g.mean()(i, 0) = mean[i];
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) g.covariance()(i, j) = cov[i * 12 + j];
}
}  // namespace

void UKFPrediction::predictStep(const bfl::GaussianMixture& prev_state, bfl::GaussianMixture& pred_state) {
    double mean[13], cov[144], dt = state_model_->sampling_time();
    belief_to_arrays(prev_state, mean, cov);
    check(roftb_ukf_predict(ctx_->get(), 1, mean, cov, &dt), ctx_->get(), "UKFPrediction::predictStep");
    pred_state = prev_state;
    arrays_to_belief(mean, cov, pred_state);
}

// ---- pose measurement / correction -----------------------------------------------------------------------------
CartesianQuaternionMeasurement::CartesianQuaternionMeasurement(std::shared_ptr<DatasetTransformDelayed> pose_measurement,
                                                               std::shared_ptr<RobotsIO::Utils::SpatialVelocityBuffer> velocity_measurement,
                                                               const bool, const bool use_pose_measurement, const bool use_velocity_measurement)
    : pose_measurement_(std::move(pose_measurement)), velocity_measurement_(std::move(velocity_measurement)),
      use_pose_measurement_(use_pose_measurement), use_velocity_measurement_(use_velocity_measurement) {
    pose_frames_between_iterations_ = pose_measurement_ ? pose_measurement_->get_frames_between_iterations() : -1;  // .cpp:74-75
}

void CartesianQuaternionMeasurement::set_type(MeasurementType t) {
    using VD = bfl::VectorDescription;
    measurement_type_ = t;
    switch (t) {
        case MeasurementType::PoseVelocity:
            input_description_ = VD(9, 1, 12, VD::CircularType::Quaternion);
            measurement_description_ = VD(9, 1, 0, VD::CircularType::Quaternion);
            measurement_.resize(13, 1);
            for (int i = 0; i < 3; ++i) { measurement_(i, 0) = last_linear_velocity_[i]; measurement_(3 + i, 0) = last_angular_velocity_[i]; }
            for (int i = 0; i < 7; ++i) measurement_(6 + i, 0) = last_pose_[i];
            break;
        case MeasurementType::Velocity:
            input_description_ = VD(9, 1, 6, VD::CircularType::Quaternion);
            measurement_description_ = VD(6);
            measurement_.resize(6, 1);
            for (int i = 0; i < 3; ++i) { measurement_(i, 0) = last_linear_velocity_[i]; measurement_(3 + i, 0) = last_angular_velocity_[i]; }
            break;
        case MeasurementType::Pose:
            input_description_ = VD(9, 1, 6, VD::CircularType::Quaternion);
            measurement_description_ = VD(3, 1, 0, VD::CircularType::Quaternion);
            measurement_.resize(7, 1);
            for (int i = 0; i < 7; ++i) measurement_(i, 0) = last_pose_[i];
            break;
        default:
            input_description_ = VD(0, 0, 0);
            measurement_description_ = VD(0, 0, 0);
    }
}

bool CartesianQuaternionMeasurement::freeze(const bfl::Data& data) {
    const MeasurementMode mode = bfl::any::any_cast<MeasurementMode>(data);
    if (mode == MeasurementMode::PopBufferedMeasurement) {  // .cpp:97-152
        if (pose_frames_between_iterations_ > 0)
            while (int(buffer_velocities_.size()) > pose_frames_between_iterations_ + 1) buffer_velocities_.pop_front();
        if (buffer_velocities_.empty()) {
            std::vector<double> v(6);
            for (int i = 0; i < 6; ++i) v[i] = measurement_(i, 0);
            buffer_velocities_.push_back(v);
            return false;
        }
        const std::vector<double> buffered = buffer_velocities_.front();
        buffer_velocities_.pop_front();
        for (int i = 0; i < 3; ++i) { last_linear_velocity_[i] = buffered[i]; last_angular_velocity_[i] = buffered[3 + i]; }
        if (is_pose_) {
            set_type(MeasurementType::PoseVelocity);
            is_pose_ = false;  // once consumed, the pose is not valid anymore
        } else {
            set_type(MeasurementType::Velocity);
        }
        return true;
    }
    if (mode == MeasurementMode::RepeatOnlyVelocity) {  // .cpp:154-174
        if (is_first_velocity_in_) set_type(MeasurementType::Velocity);
        return true;
    }
    // Standard (.cpp:176-347)
    if (use_velocity_measurement_ && velocity_measurement_->freeze(true)) {
        is_first_velocity_in_ = true;
        for (int i = 0; i < 3; ++i) {
            last_linear_velocity_[i] = velocity_measurement_->linear_velocity_origin()[i];
            last_angular_velocity_[i] = velocity_measurement_->angular_velocity()[i];
        }
    }
    is_pose_ = false;
    if (use_pose_measurement_) {
        is_pose_ = pose_measurement_->freeze(false);
        if (is_pose_) std::memcpy(last_pose_, pose_measurement_->transform(), sizeof(last_pose_));
    }
    bool valid_freeze = true;
    if (is_first_velocity_in_ && is_pose_) {
        set_type(MeasurementType::PoseVelocity);
        buffer_velocities_.push_back({measurement_(0, 0), measurement_(1, 0), measurement_(2, 0), measurement_(3, 0), measurement_(4, 0), measurement_(5, 0)});
    } else if (is_first_velocity_in_) {
        set_type(MeasurementType::Velocity);
        buffer_velocities_.push_back({measurement_(0, 0), measurement_(1, 0), measurement_(2, 0), measurement_(3, 0), measurement_(4, 0), measurement_(5, 0)});
    } else if (is_pose_) {
        set_type(MeasurementType::Pose);
    } else {
        set_type(MeasurementType::None);
        valid_freeze = false;
    }
    return valid_freeze;
}

std::pair<bool, bfl::Data> CartesianQuaternionMeasurement::measure(const bfl::Data&) const {
    return std::make_pair(measurement_type_ != MeasurementType::None, bfl::Data(measurement_));
}
std::pair<bool, bfl::Data> CartesianQuaternionMeasurement::predictedMeasure(const Eigen::Ref<const Eigen::MatrixXd>&) const {
    throw std::runtime_error("CartesianQuaternionMeasurement::predictedMeasure: evaluated on the device by ROFT::UKFCorrection (roftb_ukf_correct)");
}
std::pair<bool, bfl::Data> CartesianQuaternionMeasurement::innovation(const bfl::Data&, const bfl::Data&) const {
    throw std::runtime_error("CartesianQuaternionMeasurement::innovation: evaluated on the device by ROFT::UKFCorrection (roftb_ukf_correct)");
}

UKFCorrection::UKFCorrection(std::unique_ptr<bfl::MeasurementModel> meas_model, const double, const double, const double,
                             std::shared_ptr<B200Context> ctx)
    : measurement_model_(std::move(meas_model)), ctx_(std::move(ctx)) {}

void UKFCorrection::correctStep(const bfl::GaussianMixture& pred_state, bfl::GaussianMixture& corr_state) {
    auto* m = dynamic_cast<CartesianQuaternionMeasurement*>(measurement_model_.get());
    if (!m) throw std::runtime_error("UKFCorrection::correctStep: the measurement model is not a ROFT::CartesianQuaternionMeasurement");
    bool valid = false;
    bfl::Data d;
    std::tie(valid, d) = m->measure();
    corr_state = pred_state;
    if (!valid) return;  // UKFCorrection.cpp:60-66
    const Eigen::MatrixXd z = bfl::any::any_cast<Eigen::MatrixXd>(d);
    double meas[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0};
    int32_t type = m->measurement_type();
    if (type == ROFTB_MEAS_POSE)
        for (int i = 0; i < 7; ++i) meas[6 + i] = z(i, 0);
    else
        for (int i = 0; i < z.rows(); ++i) meas[i] = z(i, 0);
    double mean[13], cov[144];
    belief_to_arrays(pred_state, mean, cov);
    check(roftb_ukf_correct(ctx_->get(), 1, mean, cov, meas, &type), ctx_->get(), "UKFCorrection::correctStep");
    arrays_to_belief(mean, cov, corr_state);
}

}  // namespace ROFT

// ---- test hook: the stamped source driven from arrays (tests/test_host.py) ----------------------------------------------
namespace {
using namespace ROFT;
struct ArrayFlowSource : ImageOpticalFlowSource {
    std::vector<FlowFrame> frames;
    std::vector<uint8_t> valid;
    int head = -1, type = ROFTB_FLOW_F32;
    std::size_t grid = 1;
    float scale = 1.f;
    bool step_frame() override { ++head; return true; }
    bool is_stepping_required() const override { return true; }
    std::tuple<bool, const FlowFrame*> flow(const bool&) override {
        const bool ok = head >= 0 && head < int(frames.size()) && valid[head];
        return std::make_tuple(ok, ok ? &frames[head] : nullptr);
    }
    std::size_t get_grid_size() const override { return grid; }
    float get_scaling_factor() const override { return scale; }
    int get_matrix_type() const override { return type; }
};
struct ArrayStampedSegmentation : StampedSegmentation {
    std::vector<MaskImage> masks;
    std::vector<uint8_t> valid;
    std::vector<double> stamps;
    int head = -1, between = -1;
    bool step_frame() override { ++head; return true; }
    bool is_stepping_required() const override { return true; }
    int get_frames_between_iterations() const override { return between; }
    std::pair<bool, MaskImage> segmentation(const bool&) override {
        const bool ok = head >= 0 && head < int(masks.size()) && valid[head];
        return std::make_pair(ok, ok ? masks[head] : MaskImage());
    }
    double get_time_stamp() override { return head >= 0 && head < int(stamps.size()) ? stamps[head] : -1.0; }
};
}  // namespace

extern "C" int rofth_stamped_sync_run(int width, int height, int flow_type, int flow_grid, float flow_scale, int n_frames,
                                      const uint8_t* masks, const uint8_t* mask_valid, const double* mask_stamp, const uint8_t* flows,
                                      const uint8_t* flow_valid, const double* rgb_stamp, int frames_between, uint8_t* out_masks,
                                      uint8_t* out_available) {
    try {
        CameraParameters cam;
        cam.width = std::size_t(width); cam.height = std::size_t(height);
        cam.fx = cam.fy = 600.0; cam.cx = width / 2.0; cam.cy = height / 2.0;
        const double cov_flow[2] = {1, 1}, pm[6] = {1, 1, 1, 1, 1, 1}, pz[12] = {0.1, 0.1, 0.1, 1e-4, 1e-4, 1e-4, 1e-3, 1e-3, 1e-3, 1e-4, 1e-4, 1e-4};
        auto ctx = std::make_shared<B200Context>(cam, flow_type, std::size_t(flow_grid), flow_scale, 1.0, 2.0, true, cov_flow, pm, pz, 1.0, 2.0, 0.0,
                                                 frames_between > 0 && frames_between <= ROFTB_MAX_DELAY ? frames_between : 0);
        const std::size_t HW = std::size_t(width) * height;
        const std::size_t fbytes = std::size_t(width / flow_grid) * (height / flow_grid) * (flow_type == ROFTB_FLOW_S16 ? 4 : 8);
        auto fs = std::make_shared<ArrayFlowSource>();
        fs->type = flow_type; fs->grid = std::size_t(flow_grid); fs->scale = flow_scale;
        auto ss = std::make_shared<ArrayStampedSegmentation>();
        ss->between = frames_between;
        for (int k = 0; k < n_frames; ++k) {
            FlowFrame f;
            f.type = flow_type; f.cols = std::size_t(width / flow_grid); f.rows = std::size_t(height / flow_grid);
            f.data.assign(flows + k * fbytes, flows + (k + 1) * fbytes);
            fs->frames.push_back(std::move(f));
            fs->valid.push_back(flow_valid[k]);
            MaskImage m;
            m.cols = std::size_t(width); m.rows = std::size_t(height);
            m.data.assign(masks + k * HW, masks + (k + 1) * HW);
            ss->masks.push_back(std::move(m));
            ss->valid.push_back(mask_valid[k]);
            ss->stamps.push_back(mask_stamp[k]);
        }
        ImageSegmentationOFAidedSourceStamped<cv::Vec2f> f32(ss, fs, cam, false, ctx);
        ImageSegmentationOFAidedSourceStamped<cv::Vec2s> s16(ss, fs, cam, false, ctx);
        for (int k = 0; k < n_frames; ++k) {
            fs->step_frame();  // the filter steps the flow source before the segmentation (ROFTFilter.cpp:283-286)
            std::pair<bool, MaskImage> r;
            if (flow_type == ROFTB_FLOW_S16) {
                s16.set_rgb_image_time_stamp(rgb_stamp[k]);
                s16.step_frame();
                r = s16.segmentation(false);
            } else {
                f32.set_rgb_image_time_stamp(rgb_stamp[k]);
                f32.step_frame();
                r = f32.segmentation(false);
            }
            out_available[k] = r.first ? 1 : 0;
            if (r.first) std::memcpy(out_masks + k * HW, r.second.data.data(), HW);
            else std::memset(out_masks + k * HW, 0, HW);
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "rofth_stamped_sync_run: " << e.what() << std::endl;
        return -1;
    }
}
