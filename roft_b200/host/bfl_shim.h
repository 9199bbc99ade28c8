// bfl_shim.h - the smallest stand-in for the third-party interfaces ROFT's plug-ins are written against, so that the
// adapter classes of roft_adapters.h carry the reference's exact class names and virtual signatures and can be compiled
// and tested in an image that has neither Eigen nor BayesFilters.  With the real libraries installed this
// header is simply not included: the adapters only use the members declared here.
//
//   Eigen::MatrixXd / VectorXd / MatrixXf / Ref        (column-major storage like Eigen's default)
//   bfl::Data, bfl::any::any_cast                      BayesFilters/Data.h, any.h
//   bfl::VectorDescription                             BayesFilters/VectorDescription.h
//   bfl::Gaussian, bfl::GaussianMixture                BayesFilters/Gaussian(Mixture).h   (one component)
//   bfl::MeasurementModel, LinearMeasurementModel      as overridden at ImageOpticalFlowMeasurement.hpp:47-71,
//                                                      CartesianQuaternionMeasurement.h:33-49
//   bfl::StateModel, LinearStateModel                  as overridden at CartesianQuaternionModel.h:26-42, SpatialVelocityModel.h
//   bfl::GaussianPrediction / GaussianCorrection       as overridden at SKFCorrection.h:26-33, UKFCorrection.h:27-38
//   bfl::KFPrediction                                  x <- F x, P <- F P F^T + Q  (UPSTREAM-RECALL)
// Everything here is UPSTREAM-RECALL (robotology/bayes-filters-lib is not vendored by the reference).
#pragma once

#include <any>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace Eigen {

template <class S>
class DenseXX {  // dynamic dense matrix, COLUMN-major
public:
    DenseXX() = default;
    DenseXX(int r, int c) : r_(r), c_(c), d_(std::size_t(r) * c, S(0)) {}
    int rows() const { return r_; }
    int cols() const { return c_; }
    int size() const { return r_ * c_; }
    S* data() { return d_.data(); }
    const S* data() const { return d_.data(); }
    S& operator()(int i, int j) { return d_[std::size_t(j) * r_ + i]; }
    const S& operator()(int i, int j) const { return d_[std::size_t(j) * r_ + i]; }
    S& operator()(int i) { return d_[i]; }
    const S& operator()(int i) const { return d_[i]; }
    void resize(int r, int c) { r_ = r; c_ = c; d_.assign(std::size_t(r) * c, S(0)); }
    void setZero() { d_.assign(d_.size(), S(0)); }
    static DenseXX Zero(int r, int c) { return DenseXX(r, c); }
    static DenseXX Identity(int r, int c) {
        DenseXX m(r, c);
        for (int i = 0; i < (r < c ? r : c); ++i) m(i, i) = S(1);
        return m;
    }
    DenseXX transpose() const {
        DenseXX t(c_, r_);
        for (int i = 0; i < r_; ++i)
            for (int j = 0; j < c_; ++j) t(j, i) = (*this)(i, j);
        return t;
    }
    DenseXX operator*(const DenseXX& o) const {
        DenseXX m(r_, o.c_);
        for (int i = 0; i < r_; ++i)
            for (int j = 0; j < o.c_; ++j) {
                S s = 0;
                for (int k = 0; k < c_; ++k) s += (*this)(i, k) * o(k, j);
                m(i, j) = s;
            }
        return m;
    }
    DenseXX operator+(const DenseXX& o) const { DenseXX m = *this; for (int i = 0; i < size(); ++i) m.d_[i] += o.d_[i]; return m; }
    DenseXX operator-(const DenseXX& o) const { DenseXX m = *this; for (int i = 0; i < size(); ++i) m.d_[i] -= o.d_[i]; return m; }
    DenseXX inverse() const {  // Gauss-Jordan with partial pivoting
        const int n = r_;
        DenseXX a = *this, inv = Identity(n, n);
        for (int col = 0; col < n; ++col) {
            int piv = col;
            for (int i = col + 1; i < n; ++i)
                if (std::fabs(a(i, col)) > std::fabs(a(piv, col))) piv = i;
            for (int j = 0; j < n; ++j) { std::swap(a(col, j), a(piv, j)); std::swap(inv(col, j), inv(piv, j)); }
            const S p = S(1) / a(col, col);
            for (int j = 0; j < n; ++j) { a(col, j) *= p; inv(col, j) *= p; }
            for (int i = 0; i < n; ++i)
                if (i != col) {
                    const S f = a(i, col);
                    for (int j = 0; j < n; ++j) { a(i, j) -= f * a(col, j); inv(i, j) -= f * inv(col, j); }
                }
        }
        return inv;
    }

private:
    int r_ = 0, c_ = 0;
    std::vector<S> d_;
};

using MatrixXd = DenseXX<double>;
using MatrixXf = DenseXX<float>;
class VectorXd : public MatrixXd {
public:
    VectorXd() = default;
    explicit VectorXd(int n) : MatrixXd(n, 1) {}
    VectorXd(const MatrixXd& m) : MatrixXd(m) {}
};
// Eigen::Ref<const MatrixXd> parameters are passed as `const Ref<const MatrixXd>&`: a plain const reference here
template <class T>
using Ref = std::remove_const_t<T>;

}  // namespace Eigen

namespace bfl {

using Data = std::any;
namespace any {
using std::any_cast;
}

class VectorDescription {
public:
    enum class CircularType { Euler, Quaternion };
    VectorDescription(std::size_t linear = 0, std::size_t circular = 0, std::size_t noise = 0, CircularType type = CircularType::Euler)
        : linear_(linear), circular_(circular), noise_(noise), type_(type) {}
    std::size_t linear_size() const { return linear_; }
    std::size_t circular_size() const { return circular_; }
    std::size_t noise_size() const { return noise_; }
    std::size_t linear_components() const { return type_ == CircularType::Quaternion ? 4 : 1; }
    std::size_t total_size() const { return linear_ + circular_ * linear_components() + noise_; }
    std::size_t dof_size() const { return linear_ + circular_ * (type_ == CircularType::Quaternion ? 3 : 1) + noise_; }

private:
    std::size_t linear_, circular_, noise_;
    CircularType type_;
};

class GaussianMixture {  // one component: mean (dim_linear + 4 dim_circular) x 1, covariance over the dof
public:
    GaussianMixture() = default;
    GaussianMixture(std::size_t dim_linear, std::size_t dim_circular = 0, bool use_quaternion = false)
        : dim_linear(dim_linear), dim_circular(dim_circular), use_quaternion(use_quaternion),
          mean_(int(dim_linear + dim_circular * (use_quaternion ? 4 : 1)), 1),
          cov_(int(dim_linear + dim_circular * (use_quaternion ? 3 : 1)), int(dim_linear + dim_circular * (use_quaternion ? 3 : 1))) {}
    Eigen::MatrixXd& mean() { return mean_; }
    const Eigen::MatrixXd& mean() const { return mean_; }
    Eigen::MatrixXd& covariance() { return cov_; }
    const Eigen::MatrixXd& covariance() const { return cov_; }
    std::size_t components = 1, dim_linear = 0, dim_circular = 0;
    bool use_quaternion = false;

private:
    Eigen::MatrixXd mean_, cov_;
};
class Gaussian : public GaussianMixture {
public:
    using GaussianMixture::GaussianMixture;
};

class MeasurementModel {
public:
    virtual ~MeasurementModel() = default;
    virtual bool freeze(const Data& data = Data()) = 0;
    virtual std::pair<bool, Data> measure(const Data& data = Data()) const = 0;
    virtual std::pair<bool, Data> predictedMeasure(const Eigen::Ref<const Eigen::MatrixXd>& cur_states) const = 0;
    virtual std::pair<bool, Data> innovation(const Data& predicted_measurements, const Data& measurements) const = 0;
    virtual std::pair<bool, Eigen::MatrixXd> getNoiseCovarianceMatrix() const { return {false, Eigen::MatrixXd()}; }
    virtual VectorDescription getInputDescription() const = 0;
    virtual VectorDescription getMeasurementDescription() const = 0;
    virtual bool setProperty(const std::string&) { return false; }
};
class LinearMeasurementModel : public MeasurementModel {
public:
    virtual Eigen::MatrixXd getMeasurementMatrix() const = 0;
};

class StateModel {
public:
    virtual ~StateModel() = default;
    virtual bool setSamplingTime(const double&) { return false; }
    virtual bool setProperty(const std::string&) { return false; }
    virtual Eigen::MatrixXd getNoiseCovarianceMatrix() = 0;
    virtual VectorDescription getInputDescription() = 0;
    virtual VectorDescription getStateDescription() = 0;
};
class LinearStateModel : public StateModel {
public:
    virtual Eigen::MatrixXd getStateTransitionMatrix() = 0;
};

class GaussianPrediction {
public:
    virtual ~GaussianPrediction() = default;
    void predict(const GaussianMixture& prev_state, GaussianMixture& pred_state) { predictStep(prev_state, pred_state); }
    virtual StateModel& getStateModel() = 0;

protected:
    virtual void predictStep(const GaussianMixture& prev_state, GaussianMixture& pred_state) = 0;
};
class GaussianCorrection {
public:
    virtual ~GaussianCorrection() = default;
    void correct(const GaussianMixture& pred_state, GaussianMixture& corr_state) { correctStep(pred_state, corr_state); }
    virtual MeasurementModel& getMeasurementModel() = 0;

protected:
    virtual void correctStep(const GaussianMixture& pred_state, GaussianMixture& corr_state) = 0;
};

class KFPrediction : public GaussianPrediction {  // x <- F x, P <- F P F^T + Q
public:
    explicit KFPrediction(std::unique_ptr<LinearStateModel> state_model) : state_model_(std::move(state_model)) {}
    StateModel& getStateModel() override { return *state_model_; }

protected:
    void predictStep(const GaussianMixture& prev_state, GaussianMixture& pred_state) override {
        const Eigen::MatrixXd F = state_model_->getStateTransitionMatrix();
        pred_state = prev_state;
        pred_state.mean() = F * prev_state.mean();
        pred_state.covariance() = F * prev_state.covariance() * F.transpose() + state_model_->getNoiseCovarianceMatrix();
    }

private:
    std::unique_ptr<LinearStateModel> state_model_;
};

}  // namespace bfl
