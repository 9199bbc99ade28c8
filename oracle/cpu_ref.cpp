// cpu_ref.cpp - dependency-free C++17 CPU restatement of ROFT's per-frame hot path.
//
// TEST INFRASTRUCTURE / REPORTED CPU BASELINE - NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// PARITY STATUS: parity unpinned (the reference has no golden vectors and cannot be built here: Eigen3,
// OpenCV C++, BayesFilters, RobotsIO, ... are absent).  This file is cross-checked against
// oracle/roft_oracle.py, whose mask path is pinned by the genuine OpenCV primitives (cv2).
//
// Single-threaded per track, FP64, with the SEQUENTIAL per-pixel Kalman update exactly as the reference
// (SKFCorrection.cpp:129-149) - not the information form.  It omits Eigen's dynamic allocations, bfl::any
// deep copies and the OpenGL render, so its frame time is a LOWER bound on the real reference's.
//
// Reference files restated (paths under hsp-iit/roft):
//   src/roft-lib/include/ROFT/ImageOpticalFlowMeasurement.hpp:168-294   cref_flow_measurement
//   src/roft-lib/include/ROFT/OpticalFlowUtilities.h:19-22               flow_valid
//   src/roft-lib/src/SKFCorrection.cpp:37-153                            cref_skf_correct
//   src/roft-lib/src/SpatialVelocityModel.cpp:15-27 (+ bfl::KFPrediction) Filter::step
//   src/roft-lib/include/ROFT/ImageSegmentationOFAidedSource.hpp:128-281 cref_mask_warp, Filter::seg_step
//   src/roft-lib/src/ImageSegmentationMeasurement.cpp:56-68              threshold
//   src/roft-lib/src/CartesianQuaternionModel.cpp:86-141                 ukf_predict
//   src/roft-lib/src/CartesianQuaternionMeasurement.cpp:92-487           PoseMeas, ukf_correct
//   src/roft-lib/src/UKFCorrection.cpp:54-133                            ukf_correct
//   src/roft-lib/src/ROFTFilter.cpp:216-367                              Filter::step
// bfl pieces (UPSTREAM-RECALL): UT weights, sigma points (U sqrt(S) by Jacobi),
// unscented transform, mean_quaternion, diff_quaternion, sum_quaternion_rotation_vector.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <thread>
#include <vector>

namespace {

struct Cfg {
    int W, H;
    double fx, fy, cx, cy, sample_time;
    int flow_s16, grid;
    float scale;
    double cov_flow[2], depth_max;
    int stride, weight_flow;
    double v_sigma[6], v_cov0[6];
    double psd_lin[3], sigma_ang[3], p_cov0[12];
    double cov_v[3], cov_w[3], cov_x[3], cov_q[3];
    double alpha, beta, kappa;
    int use_pose, use_pose_resync, use_velocity, flow_aided, segm_delay, pose_delay;
};

inline bool flow_valid(float fx, float fy) {
    return !std::isnan(fx) && !std::isnan(fy) && std::fabs(fx) < 1e9 && std::fabs(fy) < 1e9;
}

inline void flow_at(const Cfg& c, const void* flow, long r, long col, float& dx, float& dy) {
    const long idx = (r * (c.W / c.grid) + col) * 2;
    if (c.flow_s16) {
        const int16_t* f = static_cast<const int16_t*>(flow);
        dx = float(f[idx]) / c.scale;
        dy = float(f[idx + 1]) / c.scale;
    } else {
        const float* f = static_cast<const float*>(flow);
        dx = float(f[idx]) / c.scale;
        dy = float(f[idx + 1]) / c.scale;
    }
}

// x86-64 `int(float)` (cvttss2si): NaN / out of range -> INT_MIN.  Written out so the result does not depend on UB.
inline int cvt_int(float t) {
    if (!(std::fabs(t) < 2147483648.0f)) return std::numeric_limits<int>::min();
    return static_cast<int>(t);
}

// ImageOpticalFlowMeasurement<T>::freeze, hpp:231-283. z: 2n, Hm: 2n x 6 row-major. Returns n.
int flow_measurement(const Cfg& c, const uint8_t* mask, const float* depth, const void* flow, double dt,
                     std::vector<double>& z, std::vector<double>& Hm) {
    std::vector<int> coords;  // cv::findNonZero: row-major
    coords.reserve(1 << 16);
    const int n_px = c.W * c.H;
    for (int i = 0; i < n_px; ++i)
        if (mask[i]) coords.push_back(i);
    z.clear();
    Hm.clear();
    const float radius = float(c.stride);
    for (std::size_t i = 0; i < coords.size(); i += radius) {
        const int u = coords[i] % c.W, v = coords[i] / c.W;
        const float d = depth[coords[i]];
        float dx, dy;
        flow_at(c, flow, v / c.grid, u / c.grid, dx, dy);
        if (!(flow_valid(dx, dy) && d > 0 && d < c.depth_max)) continue;
        z.push_back(dx);
        z.push_back(dy);
        const double uu = u - c.cx, vv = v - c.cy;
        const double row[12] = {c.fx / d, 0.0, -uu / d, -uu * vv / c.fy, c.fx + uu * uu / c.fx, -vv * c.fx / c.fy,
                                0.0, c.fy / d, -vv / d, -(c.fy + vv * vv / c.fy), vv * uu / c.fx, uu * c.fy / c.fx};
        for (int k = 0; k < 12; ++k) Hm.push_back(row[k] * dt);
    }
    return int(z.size() / 2);
}

// SKFCorrection::correctStep, SKFCorrection.cpp:37-153 (x, P = predicted belief in, corrected out)
void skf_correct(const Cfg& c, double* x, double* P, const std::vector<double>& z, const std::vector<double>& Hm) {
    const int n = int(z.size() / 2);
    if (n == 0) return;
    std::vector<double> lik;
    if (c.weight_flow) {
        // innovations w.r.t. the predicted mean, interleaved [dx_0, dy_0, dx_1, dy_1, ...] (:74-84)
        std::vector<double> nu(2 * std::size_t(n)), norms(n), sorted(n);
        for (int j = 0; j < n; ++j)
            for (int r = 0; r < 2; ++r) {
                double p = 0;
                for (int k = 0; k < 6; ++k) p += Hm[(2 * j + r) * 6 + k] * x[k];
                nu[2 * j + r] = z[2 * j + r] - p;
            }
        // Reference quirk Q3 (:93-94): Map<MatrixXd>(nu.data(), n, 2) is COLUMN-major, so the row norms that feed the
        // median and b are sqrt(nu[i]^2 + nu[n + i]^2); the per-pixel norm is only used for the likelihoods (:111).
        for (int i = 0; i < n; ++i) sorted[i] = std::sqrt(nu[i] * nu[i] + nu[n + i] * nu[n + i]);
        for (int j = 0; j < n; ++j) norms[j] = std::sqrt(nu[2 * j] * nu[2 * j] + nu[2 * j + 1] * nu[2 * j + 1]);
        std::sort(sorted.begin(), sorted.end());
        double mi = sorted[n / 2];
        if (n % 2 == 0) mi = 0.5 * (sorted[n / 2 - 1] + sorted[n / 2]);
        double b = 0;
        for (int j = 0; j < n; ++j) b += std::fabs(sorted[j] - mi);
        b /= n;
        lik.assign(n, 1.0);
        if (b > 1e-4) {
            double mx = 0;
            for (int j = 0; j < n; ++j) {
                lik[j] = std::max(1.0 / (2 * b) * std::exp(-std::fabs(norms[j] - mi) / b), 1e-6);
                mx = std::max(mx, lik[j]);
            }
            for (int j = 0; j < n; ++j) lik[j] /= mx;
        }
    }
    for (int j = 0; j < n; ++j) {
        const double* Hj = &Hm[2 * j * 6];
        double R0 = c.cov_flow[0], R1 = c.cov_flow[1];
        if (c.weight_flow) { R0 /= lik[j]; R1 /= lik[j]; }
        // PHt = P Hj^T (6x2), Py = Hj P Hj^T + R (2x2)
        double PHt[12];
        for (int i = 0; i < 6; ++i)
            for (int r = 0; r < 2; ++r) {
                double s = 0;
                for (int k = 0; k < 6; ++k) s += P[i * 6 + k] * Hj[r * 6 + k];
                PHt[i * 2 + r] = s;
            }
        double Py[4];
        for (int r = 0; r < 2; ++r)
            for (int q = 0; q < 2; ++q) {
                double s = 0;
                for (int k = 0; k < 6; ++k) s += Hj[r * 6 + k] * PHt[k * 2 + q];
                Py[r * 2 + q] = s;
            }
        Py[0] += R0;
        Py[3] += R1;
        const double det = Py[0] * Py[3] - Py[1] * Py[2];
        const double Pi[4] = {Py[3] / det, -Py[1] / det, -Py[2] / det, Py[0] / det};
        double K[12];
        for (int i = 0; i < 6; ++i)
            for (int q = 0; q < 2; ++q) K[i * 2 + q] = PHt[i * 2] * Pi[q] + PHt[i * 2 + 1] * Pi[2 + q];
        double in[2];
        for (int r = 0; r < 2; ++r) {
            double p = 0;
            for (int k = 0; k < 6; ++k) p += Hj[r * 6 + k] * x[k];
            in[r] = z[2 * j + r] - p;
        }
        for (int i = 0; i < 6; ++i) x[i] += K[i * 2] * in[0] + K[i * 2 + 1] * in[1];
        // P = (I - K Hj) P
        double M[36], Pn[36];
        for (int i = 0; i < 6; ++i)
            for (int k = 0; k < 6; ++k) M[i * 6 + k] = (i == k ? 1.0 : 0.0) - (K[i * 2] * Hj[k] + K[i * 2 + 1] * Hj[6 + k]);
        for (int i = 0; i < 6; ++i)
            for (int k = 0; k < 6; ++k) {
                double s = 0;
                for (int m = 0; m < 6; ++m) s += M[i * 6 + m] * P[m * 6 + k];
                Pn[i * 6 + k] = s;
            }
        std::memcpy(P, Pn, sizeof(Pn));
    }
}

// ImageSegmentationOFAidedSource<T>::map + cv::remap, hpp:211-226,235-281.  mask is warped in place.
void mask_warp(const Cfg& c, std::vector<uint8_t>& mask, const std::vector<const void*>& flows) {
    const int W = c.W, H = c.H;
    std::vector<float> map(std::size_t(W) * H * 2, 0.0f);  // cv::Mat(CV_32FC2, Scalar(0,0)) every call (hpp:237)
    int start = 0;
    if (c.segm_delay > 0) start = std::max(0, int(flows.size()) - c.segm_delay);
    for (int p = 0; p < W * H; ++p) {
        if (!mask[p]) continue;
        const int px = p % W, py = p / W;
        float tx = float(px), ty = float(py);
        bool error = false;
        for (int j = start; j < int(flows.size()); ++j) {
            if (cvt_int(tx) < 0 || cvt_int(tx) >= W || cvt_int(ty) < 0 || cvt_int(ty) >= H) { error = true; break; }
            float dx, dy;
            flow_at(c, flows[j], cvt_int(ty / float(c.grid)), cvt_int(tx / float(c.grid)), dx, dy);
            tx += dx;
            ty += dy;
        }
        if (error || cvt_int(tx) < 0 || cvt_int(tx) >= W || cvt_int(ty) < 0 || cvt_int(ty) >= H) continue;
        const std::size_t o = (std::size_t(cvt_int(ty)) * W + cvt_int(tx)) * 2;
        map[o] = float(px);
        map[o + 1] = float(py);
    }
    // cv::remap with integer-valued coordinates is an exact gather; remap copies the source when in place
    const std::vector<uint8_t> src = mask;
    for (int p = 0; p < W * H; ++p) mask[p] = src[std::size_t(map[2 * p + 1]) * W + std::size_t(map[2 * p])];
}

// ---- small dense helpers -------------------------------------------------------------------
using Vec = std::vector<double>;
struct Mat {
    int r = 0, c = 0;
    Vec a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a(std::size_t(r_) * c_, 0.0) {}
    double& operator()(int i, int j) { return a[std::size_t(i) * c + j]; }
    double operator()(int i, int j) const { return a[std::size_t(i) * c + j]; }
};

constexpr double kJacobiRelTol = 1e-14;
constexpr int kJacobiMaxSweeps = 24;

// cyclic Jacobi eigen-decomposition (same criterion as oracle/roft_oracle.py jacobi_eigh)
void jacobi(Mat& A, Mat& V) {
    const int n = A.r;
    V = Mat(n, n);
    for (int i = 0; i < n; ++i) V(i, i) = 1.0;
    for (int sweep = 0; sweep < kJacobiMaxSweeps; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A(p, q);
                if (std::fabs(apq) <= kJacobiRelTol * std::sqrt(std::fabs(A(p, p) * A(q, q)))) continue;
                rotated = true;
                const double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
                const double t = std::copysign(1.0, theta) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < n; ++k) {
                    const double akp = A(k, p), akq = A(k, q);
                    A(k, p) = cs * akp - sn * akq;
                    A(k, q) = sn * akp + cs * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A(p, k), aqk = A(q, k);
                    A(p, k) = cs * apk - sn * aqk;
                    A(q, k) = sn * apk + cs * aqk;
                }
                A(p, q) = 0.0;
                A(q, p) = 0.0;
                for (int k = 0; k < n; ++k) {
                    const double vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = cs * vkp - sn * vkq;
                    V(k, q) = sn * vkp + cs * vkq;
                }
            }
        if (!rotated) break;
    }
}

Mat cov_sqrt(const Mat& P) {
    Mat A = P, V;
    for (int i = 0; i < A.r; ++i)
        for (int j = 0; j < A.c; ++j) A(i, j) = 0.5 * (P(i, j) + P(j, i));
    jacobi(A, V);
    Mat out(P.r, P.c);
    for (int i = 0; i < P.r; ++i)
        for (int j = 0; j < P.c; ++j) out(i, j) = V(i, j) * std::sqrt(std::max(A(j, j), 0.0));
    return out;
}

void qmul(const double* a, const double* b, double* o) {
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    const double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
void rotvec_to_quat(const double* r, double* q) {
    const double n = std::sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (n > 0) {
        const double k = std::sin(n / 2) / n;
        q[0] = std::cos(n / 2); q[1] = k * r[0]; q[2] = k * r[1]; q[3] = k * r[2];
    } else {
        q[0] = 1; q[1] = q[2] = q[3] = 0;
    }
}
void quat_diff(const double* a, const double* b, double* r) {  // log(a (x) conj(b)), short way round
    const double bc[4] = {b[0], -b[1], -b[2], -b[3]};
    double p[4];
    qmul(a, bc, p);
    if (p[0] < 0) for (double& v : p) v = -v;
    const double n = std::sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    if (n > 0) {
        const double k = 2.0 * std::acos(std::min(1.0, std::max(-1.0, p[0]))) / n;
        r[0] = k * p[1]; r[1] = k * p[2]; r[2] = k * p[3];
    } else {
        r[0] = r[1] = r[2] = 0;
    }
}

struct UtW { double wm0, wc0, wi, c; };
UtW ut_weights(int n, const Cfg& c) {
    const double lam = c.alpha * c.alpha * (n + c.kappa) - n;
    return {lam / (n + lam), lam / (n + lam) + (1 - c.alpha * c.alpha + c.beta), 1.0 / (2 * (n + lam)), n + lam};
}

// sigma points of the augmented state: rows = points, cols = 13 + k
Mat sigma_points(const double* mean, const Mat& cov12, const Mat& noise, double c) {
    const int k = noise.r, n = 12 + k;
    Mat aug(n, n);
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) aug(i, j) = cov12(i, j);
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j) aug(12 + i, 12 + j) = noise(i, j);
    const Mat A = cov_sqrt(aug);
    const double sc = std::sqrt(c);
    Mat sp(2 * n + 1, 13 + k);
    for (int i = 0; i < 2 * n + 1; ++i) {
        double pert[24] = {0};
        if (i > 0) {
            const int col = (i - 1) % n;
            const double sg = (i - 1) < n ? sc : -sc;
            for (int r = 0; r < n; ++r) pert[r] = sg * A(r, col);
        }
        for (int r = 0; r < 9; ++r) sp(i, r) = mean[r] + pert[r];
        double dq[4], q[4];
        rotvec_to_quat(&pert[9], dq);
        qmul(dq, &mean[9], q);
        for (int r = 0; r < 4; ++r) sp(i, 9 + r) = q[r];
        for (int r = 0; r < k; ++r) sp(i, 13 + r) = pert[12 + r];
    }
    return sp;
}

void mean_quaternion(const Mat& Y, int qoff, const UtW& w, double* out) {
    Mat M(4, 4), V;
    for (int i = 0; i < Y.r; ++i) {
        const double wi = i == 0 ? w.wm0 : w.wi;
        for (int r = 0; r < 4; ++r)
            for (int c = r; c < 4; ++c) M(r, c) += wi * (Y(i, qoff + r) * Y(i, qoff + c));
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < r; ++c) M(r, c) = M(c, r);
    jacobi(M, V);
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (M(i, i) > M(best, best)) best = i;
    double dot = 0, nn = 0;
    for (int i = 0; i < 4; ++i) { dot += V(i, best) * Y(0, qoff + i); nn += V(i, best) * V(i, best); }
    const double k = (dot < 0 ? -1.0 : 1.0) / std::sqrt(nn);
    for (int i = 0; i < 4; ++i) out[i] = k * V(i, best);
}

// bfl::UKFPrediction through CartesianQuaternionModel::motion (.cpp:86-141)
void ukf_predict(const Cfg& c, double* mean, Mat& cov, double T) {
    Mat Q(9, 9);
    for (int i = 0; i < 3; ++i) {
        Q(i, i) = c.psd_lin[i] * T;
        Q(3 + i, 3 + i) = c.sigma_ang[i];
        Q(6 + i, 6 + i) = c.psd_lin[i] * (std::pow(T, 3.0) / 3.0);
        Q(i, 6 + i) = c.psd_lin[i] * (std::pow(T, 2.0) / 2.0);
        Q(6 + i, i) = c.psd_lin[i] * (std::pow(T, 2.0) / 2.0);
    }
    const UtW w = ut_weights(21, c);
    const Mat sp = sigma_points(mean, cov, Q, w.c);
    const int np = sp.r;
    Mat Y(np, 13);
    for (int i = 0; i < np; ++i) {
        for (int r = 0; r < 9; ++r) Y(i, r) = sp(i, r) + sp(i, 13 + r);
        for (int r = 0; r < 3; ++r) Y(i, 6 + r) += sp(i, r) * T;
        const double wx = sp(i, 3), wy = sp(i, 4), wz = sp(i, 5);
        const double nw = std::sqrt(wx * wx + wy * wy + wz * wz) + std::numeric_limits<double>::epsilon();
        const double k = std::sin(nw * T / 2.0) / nw;
        const double dq[4] = {std::cos(nw * T / 2.0), k * wx, k * wy, k * wz};
        const double q[4] = {sp(i, 9), sp(i, 10), sp(i, 11), sp(i, 12)};
        double qo[4];
        qmul(dq, q, qo);
        for (int r = 0; r < 4; ++r) Y(i, 9 + r) = qo[r];
    }
    double m[13] = {0};
    for (int i = 0; i < np; ++i)
        for (int r = 0; r < 9; ++r) m[r] += (i == 0 ? w.wm0 : w.wi) * Y(i, r);
    mean_quaternion(Y, 9, w, &m[9]);
    Mat D(np, 12);
    for (int i = 0; i < np; ++i) {
        for (int r = 0; r < 9; ++r) D(i, r) = Y(i, r) - m[r];
        double d[3];
        const double q[4] = {Y(i, 9), Y(i, 10), Y(i, 11), Y(i, 12)};
        quat_diff(q, &m[9], d);
        for (int r = 0; r < 3; ++r) D(i, 9 + r) = d[r];
    }
    for (int r = 0; r < 12; ++r)
        for (int s = 0; s < 12; ++s) {
            double v = 0;
            for (int i = 0; i < np; ++i) v += (i == 0 ? w.wc0 : w.wi) * D(i, r) * D(i, s);
            cov(r, s) = v;
        }
    std::memcpy(mean, m, sizeof(m));
}

enum { MEAS_NONE = 0, MEAS_VELOCITY = 1, MEAS_POSE = 2, MEAS_POSE_VELOCITY = 3 };

// ROFT::UKFCorrection::correctStep (UKFCorrection.cpp:54-133); meas laid out (v, w, x, q)
void ukf_correct(const Cfg& c, double* mean, Mat& cov, const double* meas, int mtype) {
    if (mtype == MEAS_NONE) return;
    const bool has_v = mtype == MEAS_VELOCITY || mtype == MEAS_POSE_VELOCITY;
    const bool has_p = mtype == MEAS_POSE || mtype == MEAS_POSE_VELOCITY;
    const int k = (has_v ? 6 : 0) + (has_p ? 6 : 0), n = 12 + k;
    const int nlin = (has_v ? 6 : 0) + (has_p ? 3 : 0);
    Mat R(k, k);
    int o = 0;
    if (has_v) { for (int i = 0; i < 3; ++i) { R(i, i) = c.cov_v[i]; R(3 + i, 3 + i) = c.cov_w[i]; } o = 6; }
    if (has_p) for (int i = 0; i < 3; ++i) { R(o + i, o + i) = c.cov_x[i]; R(o + 3 + i, o + 3 + i) = c.cov_q[i]; }
    const UtW w = ut_weights(n, c);
    const Mat sp = sigma_points(mean, cov, R, w.c);
    const int np = sp.r, ny = nlin + (has_p ? 4 : 0);
    Mat Y(np, ny);
    for (int i = 0; i < np; ++i) {
        int oo = 0;
        if (has_v) {
            const double px = -sp(i, 6), py = -sp(i, 7), pz = -sp(i, 8);
            const double wx = sp(i, 3), wy = sp(i, 4), wz = sp(i, 5);
            Y(i, 0) = sp(i, 0) + (wy * pz - wz * py) + sp(i, 13 + 0);
            Y(i, 1) = sp(i, 1) + (wz * px - wx * pz) + sp(i, 13 + 1);
            Y(i, 2) = sp(i, 2) + (wx * py - wy * px) + sp(i, 13 + 2);
            for (int r = 0; r < 3; ++r) Y(i, 3 + r) = sp(i, 3 + r) + sp(i, 13 + 3 + r);
            oo = 6;
        }
        if (has_p) {
            const int no = has_v ? 6 : 0;
            for (int r = 0; r < 3; ++r) Y(i, oo + r) = sp(i, 6 + r) + sp(i, 13 + no + r);
            const double nz[3] = {sp(i, 13 + no + 3), sp(i, 13 + no + 4), sp(i, 13 + no + 5)};
            double dq[4], qo[4];
            const double q[4] = {sp(i, 9), sp(i, 10), sp(i, 11), sp(i, 12)};
            rotvec_to_quat(nz, dq);
            qmul(dq, q, qo);
            for (int r = 0; r < 4; ++r) Y(i, oo + 3 + r) = qo[r];
        }
    }
    double ym[13] = {0};
    for (int i = 0; i < np; ++i)
        for (int r = 0; r < nlin; ++r) ym[r] += (i == 0 ? w.wm0 : w.wi) * Y(i, r);
    if (has_p) mean_quaternion(Y, nlin, w, &ym[nlin]);
    const int m = k;
    Mat DY(np, m), DX(np, 12);
    double innov[12];
    for (int i = 0; i < np; ++i) {
        for (int r = 0; r < nlin; ++r) DY(i, r) = Y(i, r) - ym[r];
        if (has_p) {
            const double q[4] = {Y(i, nlin), Y(i, nlin + 1), Y(i, nlin + 2), Y(i, nlin + 3)};
            double d[3];
            quat_diff(q, &ym[nlin], d);
            for (int r = 0; r < 3; ++r) DY(i, nlin + r) = d[r];
        }
        for (int r = 0; r < 9; ++r) DX(i, r) = sp(i, r) - mean[r];
        const double q[4] = {sp(i, 9), sp(i, 10), sp(i, 11), sp(i, 12)};
        double d[3];
        quat_diff(q, &mean[9], d);
        for (int r = 0; r < 3; ++r) DX(i, 9 + r) = d[r];
    }
    for (int r = 0; r < nlin; ++r) innov[r] = meas[(mtype == MEAS_POSE ? 6 : 0) + r] - ym[r];
    if (has_p) quat_diff(&meas[9], &ym[nlin], &innov[nlin]);
    Mat Py(m, m), Pxy(12, m);
    for (int r = 0; r < m; ++r)
        for (int s = 0; s < m; ++s) {
            double v = 0;
            for (int i = 0; i < np; ++i) v += (i == 0 ? w.wc0 : w.wi) * DY(i, r) * DY(i, s);
            Py(r, s) = v;
        }
    for (int r = 0; r < 12; ++r)
        for (int s = 0; s < m; ++s) {
            double v = 0;
            for (int i = 0; i < np; ++i) v += (i == 0 ? w.wc0 : w.wi) * DX(i, r) * DY(i, s);
            Pxy(r, s) = v;
        }
    // Py^-1 by Gauss-Jordan with partial pivoting
    Mat A = Py, Inv(m, m);
    for (int i = 0; i < m; ++i) Inv(i, i) = 1.0;
    for (int col = 0; col < m; ++col) {
        int piv = col;
        for (int r = col + 1; r < m; ++r)
            if (std::fabs(A(r, col)) > std::fabs(A(piv, col))) piv = r;
        if (piv != col)
            for (int j = 0; j < m; ++j) { std::swap(A(piv, j), A(col, j)); std::swap(Inv(piv, j), Inv(col, j)); }
        const double d = 1.0 / A(col, col);
        for (int j = 0; j < m; ++j) { A(col, j) *= d; Inv(col, j) *= d; }
        for (int r = 0; r < m; ++r) {
            if (r == col) continue;
            const double f = A(r, col);
            for (int j = 0; j < m; ++j) { A(r, j) -= f * A(col, j); Inv(r, j) -= f * Inv(col, j); }
        }
    }
    Mat K(12, m);
    for (int r = 0; r < 12; ++r)
        for (int s = 0; s < m; ++s) {
            double v = 0;
            for (int j = 0; j < m; ++j) v += Pxy(r, j) * Inv(j, s);
            K(r, s) = v;
        }
    double Kn[12];
    for (int r = 0; r < 12; ++r) {
        double v = 0;
        for (int j = 0; j < m; ++j) v += K(r, j) * innov[j];
        Kn[r] = v;
    }
    Mat KPy(12, m);
    for (int r = 0; r < 12; ++r)
        for (int s = 0; s < m; ++s) {
            double v = 0;
            for (int j = 0; j < m; ++j) v += K(r, j) * Py(j, s);
            KPy(r, s) = v;
        }
    for (int r = 0; r < 12; ++r)
        for (int s = 0; s < 12; ++s) {
            double v = 0;
            for (int j = 0; j < m; ++j) v += KPy(r, j) * K(s, j);
            cov(r, s) -= v;
        }
    double dq[4], qn[4];
    rotvec_to_quat(&Kn[9], dq);
    qmul(dq, &mean[9], qn);
    for (int i = 0; i < 9; ++i) mean[i] += Kn[i];
    for (int i = 0; i < 4; ++i) mean[9 + i] = qn[i];
}

// CartesianQuaternionMeasurement::freeze mode machine (.cpp:92-348), use_screw_velocity == false
struct PoseMeas {
    std::deque<std::vector<double>> buffer;
    bool is_pose = false, is_first_velocity_in = false;
    double last_v[6] = {0};
    double last_pose[7] = {0, 0, 0, 1, 0, 0, 0};
    int mtype = MEAS_NONE;
    double meas[13] = {0};
    void set(int t) {
        mtype = t;
        std::memcpy(meas, last_v, sizeof(double) * 6);
        std::memcpy(meas + 6, last_pose, sizeof(double) * 7);
    }
    bool freeze_standard(const Cfg& c, const double* vel, const double* pose) {
        if (c.use_velocity && vel) { is_first_velocity_in = true; std::memcpy(last_v, vel, sizeof(double) * 6); }
        is_pose = false;
        if (c.use_pose && pose) { is_pose = true; std::memcpy(last_pose, pose, sizeof(double) * 7); }
        if (is_first_velocity_in && is_pose) { set(MEAS_POSE_VELOCITY); buffer.emplace_back(meas, meas + 6); }
        else if (is_first_velocity_in) { set(MEAS_VELOCITY); buffer.emplace_back(meas, meas + 6); }
        else if (is_pose) set(MEAS_POSE);
        else { mtype = MEAS_NONE; return false; }
        return true;
    }
    bool freeze_pop(const Cfg& c) {
        if (c.pose_delay > 0)
            while (int(buffer.size()) > c.pose_delay + 1) buffer.pop_front();
        if (buffer.empty()) { buffer.emplace_back(meas, meas + 6); return false; }
        std::memcpy(last_v, buffer.front().data(), sizeof(double) * 6);
        buffer.pop_front();
        if (is_pose) { set(MEAS_POSE_VELOCITY); is_pose = false; }
        else set(MEAS_VELOCITY);
        return true;
    }
};

// ROFTFilter (ROFTFilter.cpp:216-367) for one track, without the GL render-and-compare
struct Filter {
    Cfg c;
    double v_mean[6], v_cov[36], p_mean[13];
    Mat p_cov{12, 12};
    double b_mean[13];
    Mat b_cov{12, 12};
    // ImageSegmentationOFAidedSource
    bool seg_src_available = false, of_first_frame = true;
    std::vector<uint8_t> mask_state;
    std::vector<std::vector<uint8_t>> flow_buffer;  // clones, like hpp:208
    // ImageSegmentationMeasurement / ImageOpticalFlowMeasurement
    bool segmeas_available = false, fm_first_frame = true;
    std::vector<uint8_t> seg, prev_seg;
    std::vector<float> prev_depth;
    PoseMeas pm;
    int last_n_valid = 0;
    std::vector<double> z, Hm;

    explicit Filter(const Cfg& cfg, const double* p0) : c(cfg) {
        std::memset(v_mean, 0, sizeof(v_mean));
        std::memset(v_cov, 0, sizeof(v_cov));
        for (int i = 0; i < 6; ++i) v_cov[i * 7] = c.v_cov0[i];
        std::memset(p_mean, 0, sizeof(p_mean));
        p_mean[9] = 1.0;
        if (p0) std::memcpy(p_mean, p0, sizeof(p_mean));
        for (int i = 0; i < 12; ++i) p_cov(i, i) = c.p_cov0[i];
        std::memcpy(b_mean, p_mean, sizeof(p_mean));
        b_cov = p_cov;
    }
    std::size_t flow_bytes() const { return std::size_t(c.W / c.grid) * (c.H / c.grid) * 2 * (c.flow_s16 ? 2 : 4); }

    void seg_step(const uint8_t* new_mask, const void* flow) {  // hpp:128-231
        const std::size_t HW = std::size_t(c.W) * c.H;
        bool valid_seg = new_mask != nullptr;
        if (!seg_src_available && valid_seg) {
            seg_src_available = true;
            mask_state.assign(new_mask, new_mask + HW);
            valid_seg = false;
        }
        if (valid_seg) {
            bool any = false;
            for (std::size_t i = 0; i < HW && !any; ++i) any = new_mask[i] != 0;
            if (!any) {
                valid_seg = false;
                if (c.segm_delay <= 0) flow_buffer.clear();
            }
        }
        const bool valid_flow = flow != nullptr && !of_first_frame;
        if (valid_flow) {
            const uint8_t* fb = static_cast<const uint8_t*>(flow);
            flow_buffer.emplace_back(fb, fb + flow_bytes());
        }
        if (valid_seg) {
            mask_state.assign(new_mask, new_mask + HW);
            std::vector<const void*> fl;
            for (auto& f : flow_buffer) fl.push_back(f.data());
            mask_warp(c, mask_state, fl);
            flow_buffer.clear();
        } else if (valid_flow && seg_src_available) {
            mask_state[0] = 0;
            mask_warp(c, mask_state, {flow});
        }
        of_first_frame = false;
    }

    void step(const float* depth, const void* flow, const uint8_t* new_mask, const double* pose, double dt) {
        const std::size_t HW = std::size_t(c.W) * c.H;
        // segmentation_->freeze()
        if (c.flow_aided) {
            seg_step(new_mask, flow);
            if (seg_src_available) {
                segmeas_available = true;
                seg.resize(HW);
                for (std::size_t i = 0; i < HW; ++i) seg[i] = mask_state[i] > 1 ? 255 : 0;  // cv::threshold(..,1,255,BINARY)
            }
        } else if (new_mask) {
            segmeas_available = true;
            seg.resize(HW);
            for (std::size_t i = 0; i < HW; ++i) seg[i] = new_mask[i] > 1 ? 255 : 0;
        }
        bool data_in = segmeas_available;
        last_n_valid = 0;
        if (segmeas_available) {
            if (!flow || fm_first_frame) {
                fm_first_frame = false;
                data_in = false;
            } else {
                last_n_valid = flow_measurement(c, prev_seg.data(), prev_depth.data(), flow, dt, z, Hm);
            }
            prev_depth.assign(depth, depth + HW);  // previous_depth_ = depth (deep copy, hpp:286)
            prev_seg = seg;
        }
        if (data_in) {
            double xp[6], Pp[36];
            std::memcpy(xp, v_mean, sizeof(xp));
            std::memcpy(Pp, v_cov, sizeof(Pp));
            for (int i = 0; i < 6; ++i) Pp[i * 7] += c.v_sigma[i];  // KFPrediction, F = I
            skf_correct(c, xp, Pp, z, Hm);
            if (last_n_valid >= 3) {  // check_observability (hpp:363-366, ROFTFilter.cpp:297-301)
                std::memcpy(v_mean, xp, sizeof(xp));
                std::memcpy(v_cov, Pp, sizeof(Pp));
            }
        }
        // pose UKF (ROFTFilter.cpp:325-367)
        double pm_mean[13];
        Mat pm_cov = p_cov;
        std::memcpy(pm_mean, p_mean, sizeof(pm_mean));
        ukf_predict(c, pm_mean, pm_cov, dt);
        if (pm.freeze_standard(c, v_mean, pose)) {
            if (pm.mtype == MEAS_POSE_VELOCITY && c.use_pose_resync) {
                double cm[13];
                Mat cc = b_cov;
                std::memcpy(cm, b_mean, sizeof(cm));
                std::memcpy(b_mean, p_mean, sizeof(b_mean));
                b_cov = p_cov;
                while (pm.freeze_pop(c)) {
                    ukf_predict(c, cm, cc, dt);
                    ukf_correct(c, cm, cc, pm.meas, pm.mtype);
                }
                std::memcpy(p_mean, cm, sizeof(cm));
                p_cov = cc;
            } else {
                ukf_correct(c, pm_mean, pm_cov, pm.meas, pm.mtype);
                std::memcpy(p_mean, pm_mean, sizeof(pm_mean));
                p_cov = pm_cov;
            }
        } else {
            std::memcpy(p_mean, pm_mean, sizeof(pm_mean));
            p_cov = pm_cov;
        }
    }
};

}  // namespace

extern "C" {

struct cref_config {  // mirrors Cfg field by field
    int32_t W, H;
    double fx, fy, cx, cy, sample_time;
    int32_t flow_s16, grid;
    float scale;
    double cov_flow[2], depth_max;
    int32_t stride, weight_flow;
    double v_sigma[6], v_cov0[6];
    double psd_lin[3], sigma_ang[3], p_cov0[12];
    double cov_v[3], cov_w[3], cov_x[3], cov_q[3];
    double alpha, beta, kappa;
    int32_t use_pose, use_pose_resync, use_velocity, flow_aided, segm_delay, pose_delay;
};

static Cfg to_cfg(const cref_config* s) {
    Cfg c;
    c.W = s->W; c.H = s->H; c.fx = s->fx; c.fy = s->fy; c.cx = s->cx; c.cy = s->cy; c.sample_time = s->sample_time;
    c.flow_s16 = s->flow_s16; c.grid = s->grid; c.scale = s->scale;
    c.cov_flow[0] = s->cov_flow[0]; c.cov_flow[1] = s->cov_flow[1]; c.depth_max = s->depth_max;
    c.stride = s->stride; c.weight_flow = s->weight_flow;
    std::memcpy(c.v_sigma, s->v_sigma, sizeof(c.v_sigma)); std::memcpy(c.v_cov0, s->v_cov0, sizeof(c.v_cov0));
    std::memcpy(c.psd_lin, s->psd_lin, sizeof(c.psd_lin)); std::memcpy(c.sigma_ang, s->sigma_ang, sizeof(c.sigma_ang));
    std::memcpy(c.p_cov0, s->p_cov0, sizeof(c.p_cov0));
    std::memcpy(c.cov_v, s->cov_v, 24); std::memcpy(c.cov_w, s->cov_w, 24); std::memcpy(c.cov_x, s->cov_x, 24); std::memcpy(c.cov_q, s->cov_q, 24);
    c.alpha = s->alpha; c.beta = s->beta; c.kappa = s->kappa;
    c.use_pose = s->use_pose; c.use_pose_resync = s->use_pose_resync; c.use_velocity = s->use_velocity;
    c.flow_aided = s->flow_aided; c.segm_delay = s->segm_delay; c.pose_delay = s->pose_delay;
    return c;
}

void* cref_filter_create(const cref_config* cfg, const double* p_mean0) { return new Filter(to_cfg(cfg), p_mean0); }
void cref_filter_destroy(void* f) { delete static_cast<Filter*>(f); }
void cref_filter_step(void* fp, const float* depth, const void* flow, const uint8_t* mask, const double* pose, double dt) {
    static_cast<Filter*>(fp)->step(depth, flow, mask, pose, dt);
}
void cref_filter_state(void* fp, double* p_mean, double* p_cov, double* v_mean, double* v_cov, int32_t* n_valid) {
    Filter* f = static_cast<Filter*>(fp);
    if (p_mean) std::memcpy(p_mean, f->p_mean, sizeof(f->p_mean));
    if (p_cov) std::memcpy(p_cov, f->p_cov.a.data(), sizeof(double) * 144);
    if (v_mean) std::memcpy(v_mean, f->v_mean, sizeof(f->v_mean));
    if (v_cov) std::memcpy(v_cov, f->v_cov, sizeof(f->v_cov));
    if (n_valid) *n_valid = f->last_n_valid;
}
void cref_filter_mask(void* fp, uint8_t* raw, uint8_t* thr) {
    Filter* f = static_cast<Filter*>(fp);
    if (raw && !f->mask_state.empty()) std::memcpy(raw, f->mask_state.data(), f->mask_state.size());
    if (thr && !f->seg.empty()) std::memcpy(thr, f->seg.data(), f->seg.size());
}

// stateless pieces for cross-checks against roft_oracle.py
int32_t cref_flow_measurement(const cref_config* cfg, const uint8_t* mask, const float* depth, const void* flow, double dt,
                              int32_t capacity, double* z, double* Hm) {
    const Cfg c = to_cfg(cfg);
    std::vector<double> zz, hh;
    const int n = flow_measurement(c, mask, depth, flow, dt, zz, hh);
    const int k = std::min(n, capacity);
    if (z) std::memcpy(z, zz.data(), sizeof(double) * 2 * k);
    if (Hm) std::memcpy(Hm, hh.data(), sizeof(double) * 12 * k);
    return n;
}
void cref_skf_correct(const cref_config* cfg, double* x, double* P, const double* z, const double* Hm, int32_t n) {
    const Cfg c = to_cfg(cfg);
    std::vector<double> zz(z, z + 2 * n), hh(Hm, Hm + 12 * n);
    skf_correct(c, x, P, zz, hh);
}
void cref_mask_warp(const cref_config* cfg, uint8_t* mask, const void* const* flows, int32_t n_flows, int32_t zero_origin) {
    const Cfg c = to_cfg(cfg);
    std::vector<uint8_t> m(mask, mask + std::size_t(c.W) * c.H);
    if (zero_origin) m[0] = 0;
    std::vector<const void*> fl(flows, flows + n_flows);
    mask_warp(c, m, fl);
    std::memcpy(mask, m.data(), m.size());
}
void cref_ukf_predict(const cref_config* cfg, double* mean, double* cov, double T) {
    const Cfg c = to_cfg(cfg);
    Mat P(12, 12);
    std::memcpy(P.a.data(), cov, sizeof(double) * 144);
    ukf_predict(c, mean, P, T);
    std::memcpy(cov, P.a.data(), sizeof(double) * 144);
}
void cref_ukf_correct(const cref_config* cfg, double* mean, double* cov, const double* meas13, int32_t mtype) {
    const Cfg c = to_cfg(cfg);
    Mat P(12, 12);
    std::memcpy(P.a.data(), cov, sizeof(double) * 144);
    ukf_correct(c, mean, P, meas13, mtype);
    std::memcpy(cov, P.a.data(), sizeof(double) * 144);
}

// Timed baseline: n_tracks independent filters over n_frames resident frames, n_threads workers (one track
// at a time per worker), stops after ~seconds.  Frame k of track t: depth[k][t], flow[k][t] (k >= 1),
// mask / pose delivered with the delay schedule of DatasetImageSegmentationDelayed.cpp:42-63.
// Timing convention of ROFTFilter.cpp:270,370: compute only, inputs already in memory.
// Returns tracked frames; *elapsed_s = wall time of the slowest worker.
int64_t cref_timed_run(const cref_config* cfg, int32_t n_tracks, int32_t n_frames, int32_t n_threads, double seconds,
                       const float* depth, const void* flow, const uint8_t* mask, const double* pose,
                       const uint8_t* pose_valid, double* elapsed_s, double* out_v_mean) {
    const Cfg c = to_cfg(cfg);
    const std::size_t HW = std::size_t(c.W) * c.H;
    const std::size_t fe = std::size_t(c.W / c.grid) * (c.H / c.grid) * 2 * (c.flow_s16 ? 2 : 4);
    std::atomic<int> next{0};
    std::atomic<int64_t> frames{0};
    std::vector<double> worker_s(n_threads, 0.0);
    const auto t_start = std::chrono::steady_clock::now();
    auto work = [&](int wid) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const int job = next.fetch_add(1);
            const int t = job % n_tracks;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > seconds && job >= n_threads) break;
            double p0[13] = {0};
            std::memcpy(p0 + 6, pose + (std::size_t(0) * n_tracks + t) * 7, sizeof(double) * 7);
            Filter f(c, p0);
            const int D = c.segm_delay;
            for (int k = 0; k < n_frames; ++k) {
                int idx = k - D;
                const bool deliver = D <= 0 ? true : (idx % D == 0);
                if (D <= 0) idx = k;
                if (idx < 0) idx = 0;
                const uint8_t* m = deliver ? mask + (std::size_t(idx) * n_tracks + t) * HW : nullptr;
                const double* ps = (deliver && pose_valid[std::size_t(idx) * n_tracks + t]) ? pose + (std::size_t(idx) * n_tracks + t) * 7 : nullptr;
                const void* fl = k > 0 ? static_cast<const char*>(flow) + (std::size_t(k) * n_tracks + t) * fe : nullptr;
                f.step(depth + (std::size_t(k) * n_tracks + t) * HW, fl, m, ps, c.sample_time);
                frames.fetch_add(1);
            }
            if (out_v_mean) std::memcpy(out_v_mean + std::size_t(t) * 6, f.v_mean, sizeof(double) * 6);
        }
        worker_s[wid] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
    double mx = 0;
    for (double s : worker_s) mx = std::max(mx, s);
    if (elapsed_s) *elapsed_s = mx;
    return frames.load();
}

}  // extern "C"
