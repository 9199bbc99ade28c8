// Batched quaternion UKF over the 13-vector / 12-dof state (v, w, x, q): one warp per track, a small
// per-track op list (predict / correct / swap-with-buffered-belief) per launch so that the pose
// re-synchronisation replay of ROFTFilter.cpp:331-354 never returns to the host.
//
// Replaces
//   bfl::UKFPrediction::predictStep through CartesianQuaternionModel::motion
//        src/roft-lib/src/CartesianQuaternionModel.cpp:86-141           (43 sigma points)
//   ROFT::UKFCorrection::correctStep   src/roft-lib/src/UKFCorrection.cpp:54-133
//   CartesianQuaternionMeasurement::{predictedMeasure, innovation}
//        src/roft-lib/src/CartesianQuaternionMeasurement.cpp:357-487     (37 / 49 sigma points)
// and the bfl utilities they call (sigma_point, unscented_transform, mean_quaternion, diff_quaternion,
// sum_quaternion_rotation_vector - UPSTREAM-RECALL).
//
// All arithmetic is FP64 (the work is a few KB per track; the kernel is latency-bound and is meant to
// hide under the streaming kernels).  The covariance square root is the symmetric eigen-decomposition
// A = U sqrt(S) by round-robin Jacobi with a relative rotation threshold (see DESIGN.md "UKF square root").
#include "roftb_internal.cuh"

namespace roftb {
namespace {

constexpr int kMaxSp = 49;
constexpr double kJacobiRelTol = 1e-14;
constexpr int kJacobiMaxSweeps = 24;

struct UkfSmem {
    double A[12][12];    // Jacobi work matrix
    double V[12][12];    // eigenvectors
    double AP[12][12];   // sqrt(c) * U sqrt(S) of the state covariance
    double AN[12][12];   // same for the noise covariance
    double P[12][12];    // covariance of the belief being processed
    double mean[13];
    double Y[kMaxSp][13];
    double DX[kMaxSp][12];
    double DY[kMaxSp][12];
    double ymean[13];
    double Py[12][24];   // [Py | I] for the Gauss-Jordan inverse
    double Pxy[12][12];
    double K[12][12];
    double KPy[12][12];
    double innov[12];
    double Kn[12];
    double jc[6], js[6];  // rotations of the current Jacobi round
    int jp[6], jq[6];
};

__device__ __forceinline__ void qmul(const double* a, const double* b, double* o) {
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    const double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}

// rotation_vector_to_quaternion: |r| > 0 ? (cos(|r|/2), sin(|r|/2) r/|r|) : (1,0,0,0)
__device__ __forceinline__ void rotvec_to_quat(const double* r, double* q) {
    const double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (n > 0.0) {
        double s, c;
        sincos(0.5 * n, &s, &c);
        const double k = s / n;
        q[0] = c; q[1] = k * r[0]; q[2] = k * r[1]; q[3] = k * r[2];
    } else {
        q[0] = 1.0; q[1] = 0.0; q[2] = 0.0; q[3] = 0.0;
    }
}

// diff_quaternion(a, b) = log(a (x) conj(b)) as a rotation vector, short way round
__device__ __forceinline__ void quat_diff(const double* a, const double* b, double* r) {
    const double bc[4] = {b[0], -b[1], -b[2], -b[3]};
    double p[4];
    qmul(a, bc, p);
    if (p[0] < 0.0) { p[0] = -p[0]; p[1] = -p[1]; p[2] = -p[2]; p[3] = -p[3]; }
    const double n = sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    if (n > 0.0) {
        const double k = 2.0 * acos(fmin(1.0, fmax(-1.0, p[0]))) / n;
        r[0] = k * p[1]; r[1] = k * p[2]; r[2] = k * p[3];
    } else {
        r[0] = 0.0; r[1] = 0.0; r[2] = 0.0;
    }
}

// Jacobi eigen-decomposition of the n x n symmetric matrix in s.A (n <= 12); eigenvectors to s.V. Warp-cooperative,
// round-robin ("chess tournament") ordering: the n/2 index pairs of one round are disjoint, so their rotations
// commute and are applied together - 11 dependent rounds per sweep for n = 12 instead of 66 dependent rotations.
// Same per-pair rotation rule and relative threshold as the cyclic sweep; converged when a whole sweep rotates nothing.
__device__ void jacobi_warp(UkfSmem& s, int n, int lane, bool keep_v = false) {
    if (!keep_v)
        for (int i = lane; i < 144; i += 32) s.V[i / 12][i % 12] = (i / 12 == i % 12) ? 1.0 : 0.0;
    __syncwarp();
    const int ne = n + (n & 1);  // players (a dummy index n sits out when n is odd)
    const int np = ne >> 1;      // pairs per round
    const int nr = ne - 1;       // rounds per sweep
    // work items of this lane in the update phases (n * np <= 72 items: at most three per lane), fixed for the call
    int ck[3], cj[3], rj[3], rk[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int e = lane + 32 * i;
        const bool on = e < n * np;
        ck[i] = on ? e / np : -1;  // column phase: row k, pair j
        cj[i] = on ? e % np : 0;
        rj[i] = on ? e / n : -1;   // row phase: pair j, column k
        rk[i] = on ? e % n : 0;
    }
    constexpr double kTol2 = kJacobiRelTol * kJacobiRelTol;
    for (int sweep = 0; sweep < kJacobiMaxSweeps; ++sweep) {
        bool rotated = false;
        for (int r = 0; r < nr; ++r) {
            // pair of lane j (circle method): (r, ne-1) for j = 0, ((r+j) mod nr, (r-j) mod nr) otherwise
            int p = 0, q = 0;
            double c = 1.0, sn = 0.0;
            bool rot = false;
            if (lane < np) {
                int a = r + lane;
                a = a >= nr ? a - nr : a;
                int b = r - lane;
                b = b < 0 ? b + nr : b;
                if (lane == 0) b = ne - 1;
                p = min(a, b);
                q = max(a, b);
                if (q < n) {
                    const double apq = s.A[p][q], app = s.A[p][p], aqq = s.A[q][q];
                    // |apq| > tol * sqrt(|app aqq|), squared (no square root on the critical path)
                    if (apq * apq > kTol2 * fabs(app * aqq)) {
                        rot = true;
                        // rotation angle phi, |phi| <= pi/4, with tan(2 phi) = be / al for al = aqq - app, be = 2 apq - the same
                        // angle as t = sign(theta) / (|theta| + sqrt(theta^2 + 1)), theta = al / be, c = 1 / sqrt(t^2 + 1),
                        // s = t c - from the double-angle identities: cos(2 phi) = |al| / sqrt(al^2 + be^2),
                        // c = sqrt((1 + cos 2phi) / 2), s = sin(2 phi) / (2 c).  Two reciprocal square roots on the critical
                        // path of a round instead of a square root, a division and a reciprocal square root.
                        const double al = aqq - app, be = 2.0 * apq;
                        const double rq = rsqrt(al * al + be * be);
                        const double h = fma(0.5 * fabs(al), rq, 0.5);
                        const double rc = rsqrt(h);
                        c = h * rc;
                        sn = copysign(0.5 * fabs(be) * rq * rc, al * be);
                    }
                }
            }
            const unsigned any = __ballot_sync(0xffffffffu, rot);
            if (any == 0u) continue;  // warp-uniform
            rotated = true;
            if (lane < np) {
                s.jc[lane] = c;
                s.js[lane] = sn;
                s.jp[lane] = rot ? p : -1;
                s.jq[lane] = q;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 3; ++i) {  // columns p, q of A and V
                const int k = ck[i];
                if (k < 0) continue;
                const int pj = s.jp[cj[i]], qj = s.jq[cj[i]];
                if (pj < 0) continue;
                const double cc = s.jc[cj[i]], ss = s.js[cj[i]];
                const double akp = s.A[k][pj], akq = s.A[k][qj];
                const double vkp = s.V[k][pj], vkq = s.V[k][qj];
                s.A[k][pj] = cc * akp - ss * akq;
                s.A[k][qj] = ss * akp + cc * akq;
                s.V[k][pj] = cc * vkp - ss * vkq;
                s.V[k][qj] = ss * vkp + cc * vkq;
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 3; ++i) {  // rows p, q of A
                const int j = rj[i];
                if (j < 0) continue;
                const int pj = s.jp[j], qj = s.jq[j];
                if (pj < 0) continue;
                const double cc = s.jc[j], ss = s.js[j];
                const double apk = s.A[pj][rk[i]], aqk = s.A[qj][rk[i]];
                // the rotated pair of off-diagonal entries is zero by construction: store it as such (exactly symmetric)
                s.A[pj][rk[i]] = rk[i] == qj ? 0.0 : cc * apk - ss * aqk;
                s.A[qj][rk[i]] = rk[i] == pj ? 0.0 : ss * apk + cc * aqk;
            }
            __syncwarp();
        }
        if (!rotated) break;
    }
}

// out = sc * U sqrt(max(S,0)) for the n x n matrix currently in s.A.
// `warm` (12 x 12 matrices only; may be null): eigenvectors found for this track's previous matrix of the same kind
// (posterior before a prediction / prior before a correction).  Consecutive covariances are close, so B = W^T A W is
// nearly diagonal and the sweeps that a cold start spends on the bulk of the off-diagonal mass are skipped; the
// decomposition reached is the same up to rounding (and up to the basis of a degenerate eigenspace, which the unscented
// transform does not see to second order).  Every 64th use starts cold again so that the accumulated product of
// rotations cannot drift away from orthogonality.  warm[144] = counter.
__device__ void cov_sqrt_warp(UkfSmem& s, int n, double sc, double (*out)[12], int lane, double* warm = nullptr) {
    bool keep = false;
    if (warm && n == 12) {
        const int uses = (int)warm[144];
        keep = uses > 0 && (uses & 63) != 0;
        if (keep) {
            for (int i = lane; i < 144; i += 32) s.V[i / 12][i % 12] = warm[i];
            __syncwarp();
            for (int i = lane; i < 144; i += 32) {  // K = A W
                const int r = i / 12, c = i % 12;
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 12; ++k) v = fma(s.A[r][k], s.V[k][c], v);
                s.K[r][c] = v;
            }
            __syncwarp();
            for (int i = lane; i < 78; i += 32) {  // A <- W^T K, upper triangle mirrored (exactly symmetric)
                int r = 0, c = i;
                while (c >= 12 - r) { c -= 12 - r; ++r; }
                c += r;
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 12; ++k) v = fma(s.V[k][r], s.K[k][c], v);
                s.A[r][c] = v;
                s.A[c][r] = v;
            }
            __syncwarp();
        }
        if (lane == 0) warm[144] = (double)(uses + 1);
    }
    jacobi_warp(s, n, lane, keep);
    if (warm && n == 12)
        for (int i = lane; i < 144; i += 32) warm[i] = s.V[i / 12][i % 12];
    for (int i = lane; i < n * n; i += 32) {
        const int r = i / n, c = i % n;
        out[r][c] = sc * s.V[r][c] * sqrt(fmax(s.A[c][c], 0.0));
    }
    __syncwarp();
}

// dominant eigenvector of sum_i w_i q_i q_i^T over the quaternion columns Y[i][qoff..qoff+3]; sign fixed
// against the first sigma point (DESIGN.md "UKF conventions")
__device__ void mean_quaternion_warp(UkfSmem& s, int npts, int qoff, double wm0, double wi, double* out, int lane) {
    if (lane < 10) {  // upper triangle, mirrored so the matrix is exactly symmetric
        int r = 0, c = lane;
        while (c >= 4 - r) { c -= 4 - r; ++r; }
        c += r;
        double m = 0.0;
        for (int i = 0; i < npts; ++i) m += (i == 0 ? wm0 : wi) * (s.Y[i][qoff + r] * s.Y[i][qoff + c]);
        s.A[r][c] = m;
        s.A[c][r] = m;
    }
    __syncwarp();
    jacobi_warp(s, 4, lane);
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (s.A[i][i] > s.A[best][best]) best = i;
    double q[4] = {s.V[0][best], s.V[1][best], s.V[2][best], s.V[3][best]};
    double dot = 0.0, nn = 0.0;
    for (int i = 0; i < 4; ++i) {
        dot += q[i] * s.Y[0][qoff + i];
        nn += q[i] * q[i];
    }
    const double k = (dot < 0.0 ? -1.0 : 1.0) / sqrt(nn);
    for (int i = 0; i < 4; ++i) out[i] = k * q[i];
    __syncwarp();
}

struct UtW { double wm0, wc0, wi, c; };
__device__ __forceinline__ UtW ut_weights(int n, const UkfParams& p) {
    UtW w;
    const double lam = p.alpha * p.alpha * (n + p.kappa) - n;
    w.wm0 = lam / (n + lam);
    w.wc0 = lam / (n + lam) + (1.0 - p.alpha * p.alpha + p.beta);
    w.wi = 1.0 / (2.0 * (n + lam));
    w.c = n + lam;
    return w;
}

// sigma point i of the augmented state: state part (lin[9], q[4]) and noise part nz[k]
__device__ __forceinline__ void make_sigma_point(const UkfSmem& s, int i, int n, int k, double* lin, double* q, double* nz) {
    double pert[12];
    for (int r = 0; r < 12; ++r) pert[r] = 0.0;
    for (int r = 0; r < k; ++r) nz[r] = 0.0;
    if (i > 0) {
        const int col = (i - 1) % n;
        const double sg = (i - 1) < n ? 1.0 : -1.0;
        if (col < 12) {
            for (int r = 0; r < 12; ++r) pert[r] = sg * s.AP[r][col];
        } else {
            for (int r = 0; r < k; ++r) nz[r] = sg * s.AN[r][col - 12];
        }
    }
    for (int r = 0; r < 9; ++r) lin[r] = s.mean[r] + pert[r];
    double dq[4];
    rotvec_to_quat(&pert[9], dq);
    qmul(dq, &s.mean[9], q);
}

// ---- predict ----------------------------------------------------------------------------------
__device__ __noinline__ void ukf_predict_warp(UkfSmem& s, const UkfParams& p, double T, int lane, double* warm = nullptr) {
    const int k = 9, n = 21, npts = 43;
    const UtW w = ut_weights(n, p);
    const double sc = sqrt(w.c);
    for (int i = lane; i < 144; i += 32) s.A[i / 12][i % 12] = s.P[i / 12][i % 12];
    __syncwarp();
    cov_sqrt_warp(s, 12, sc, s.AP, lane, warm);
    // Q(T), CartesianQuaternionModel.cpp:127-141
    for (int i = lane; i < 144; i += 32) s.A[i / 12][i % 12] = 0.0;
    __syncwarp();
    if (lane < 3) {
        s.A[lane][lane] = p.psd_lin[lane] * T;
        s.A[3 + lane][3 + lane] = p.sigma_ang[lane];
        s.A[6 + lane][6 + lane] = p.psd_lin[lane] * (pow(T, 3.0) / 3.0);
        s.A[lane][6 + lane] = p.psd_lin[lane] * (pow(T, 2.0) / 2.0);
        s.A[6 + lane][lane] = p.psd_lin[lane] * (pow(T, 2.0) / 2.0);
    }
    __syncwarp();
    cov_sqrt_warp(s, k, sc, s.AN, lane);

    for (int i = lane; i < npts; i += 32) {
        double lin[9], q[4], nz[12];
        make_sigma_point(s, i, n, k, lin, q, nz);
        // CartesianQuaternionModel::motion (.cpp:86-124)
        double* y = s.Y[i];
        for (int r = 0; r < 9; ++r) y[r] = lin[r] + nz[r];
        for (int r = 0; r < 3; ++r) y[6 + r] += lin[r] * T;
        const double* wv = &lin[3];
        const double nw = sqrt(wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2]) + 2.220446049250313e-16;
        double sn, cs;
        sincos(nw * T / 2.0, &sn, &cs);
        const double kk = sn / nw;
        const double dq[4] = {cs, kk * wv[0], kk * wv[1], kk * wv[2]};
        qmul(dq, q, &y[9]);
    }
    __syncwarp();
    if (lane < 9) {
        double m = 0.0;
        for (int i = 0; i < npts; ++i) m += (i == 0 ? w.wm0 : w.wi) * s.Y[i][lane];
        s.ymean[lane] = m;
    }
    __syncwarp();
    double qm[4];
    mean_quaternion_warp(s, npts, 9, w.wm0, w.wi, qm, lane);
    if (lane == 0)
        for (int i = 0; i < 4; ++i) s.ymean[9 + i] = qm[i];
    __syncwarp();
    for (int i = lane; i < npts; i += 32) {
        for (int r = 0; r < 9; ++r) s.DY[i][r] = s.Y[i][r] - s.ymean[r];
        quat_diff(&s.Y[i][9], &s.ymean[9], &s.DY[i][9]);
    }
    __syncwarp();
    for (int e = lane; e < 144; e += 32) {
        const int r = e / 12, c = e % 12;
        double v = 0.0;
        for (int i = 0; i < npts; ++i) v += (i == 0 ? w.wc0 : w.wi) * s.DY[i][r] * s.DY[i][c];
        s.P[r][c] = v;
    }
    if (lane < 13) s.mean[lane] = s.ymean[lane];
    __syncwarp();
}

// ---- correct ----------------------------------------------------------------------------------
__device__ __noinline__ void ukf_correct_warp(UkfSmem& s, const UkfParams& p, int mtype, const double* meas, int lane, double* warm = nullptr) {
    if (mtype == ROFTB_MEAS_NONE) return;
    const bool has_v = (mtype == ROFTB_MEAS_VELOCITY || mtype == ROFTB_MEAS_POSE_VELOCITY);
    const bool has_p = (mtype == ROFTB_MEAS_POSE || mtype == ROFTB_MEAS_POSE_VELOCITY);
    const int k = (has_v ? 6 : 0) + (has_p ? 6 : 0);  // noise dof = innovation dof
    const int n = 12 + k, npts = 2 * n + 1;
    const int nlin = (has_v ? 6 : 0) + (has_p ? 3 : 0);
    const UtW w = ut_weights(n, p);
    const double sc = sqrt(w.c);
    for (int i = lane; i < 144; i += 32) s.A[i / 12][i % 12] = s.P[i / 12][i % 12];
    __syncwarp();
    cov_sqrt_warp(s, 12, sc, s.AP, lane, warm);
    // R = blkdiag(R_velocity, R_pose) (CartesianQuaternionMeasurement.cpp:49-61)
    for (int i = lane; i < 144; i += 32) s.A[i / 12][i % 12] = 0.0;
    __syncwarp();
    if (lane < 3) {
        int o = 0;
        if (has_v) {
            s.A[lane][lane] = p.cov_v[lane];
            s.A[3 + lane][3 + lane] = p.cov_w[lane];
            o = 6;
        }
        if (has_p) {
            s.A[o + lane][o + lane] = p.cov_x[lane];
            s.A[o + 3 + lane][o + 3 + lane] = p.cov_q[lane];
        }
    }
    __syncwarp();
    cov_sqrt_warp(s, k, sc, s.AN, lane);

    for (int i = lane; i < npts; i += 32) {
        double lin[9], q[4], nz[12];
        make_sigma_point(s, i, n, k, lin, q, nz);
        double* y = s.Y[i];
        int o = 0;
        if (has_v) {  // .cpp:384-414 with use_screw_velocity == false: v + w x (-p) + noise, w + noise
            const double* v = &lin[0];
            const double* wv = &lin[3];
            const double px = -lin[6], py = -lin[7], pz = -lin[8];
            y[0] = v[0] + (wv[1] * pz - wv[2] * py) + nz[0];
            y[1] = v[1] + (wv[2] * px - wv[0] * pz) + nz[1];
            y[2] = v[2] + (wv[0] * py - wv[1] * px) + nz[2];
            y[3] = wv[0] + nz[3];
            y[4] = wv[1] + nz[4];
            y[5] = wv[2] + nz[5];
            o = 6;
        }
        if (has_p) {  // .cpp:369-379
            const int no = has_v ? 6 : 0;
            for (int r = 0; r < 3; ++r) y[o + r] = lin[6 + r] + nz[no + r];
            double dq[4];
            rotvec_to_quat(&nz[no + 3], dq);
            qmul(dq, q, &y[o + 3]);
        }
        for (int r = 0; r < 9; ++r) s.DX[i][r] = lin[r] - s.mean[r];
        quat_diff(q, &s.mean[9], &s.DX[i][9]);
    }
    __syncwarp();
    if (lane < nlin) {
        double m = 0.0;
        for (int i = 0; i < npts; ++i) m += (i == 0 ? w.wm0 : w.wi) * s.Y[i][lane];
        s.ymean[lane] = m;
    }
    __syncwarp();
    if (has_p) {
        double qm[4];
        mean_quaternion_warp(s, npts, nlin, w.wm0, w.wi, qm, lane);
        if (lane == 0)
            for (int i = 0; i < 4; ++i) s.ymean[nlin + i] = qm[i];
        __syncwarp();
    }
    for (int i = lane; i < npts; i += 32) {
        for (int r = 0; r < nlin; ++r) s.DY[i][r] = s.Y[i][r] - s.ymean[r];
        if (has_p) quat_diff(&s.Y[i][nlin], &s.ymean[nlin], &s.DY[i][nlin]);
    }
    // innovation (.cpp:436-487): linear differences; quaternion through diff_quaternion(measured, predicted)
    if (lane < nlin) {
        // measurement layout (v, w, x, q): velocity-only uses [0..5], pose-only uses [6..12]
        const int src = (mtype == ROFTB_MEAS_POSE) ? 6 + lane : lane;
        s.innov[lane] = meas[src] - s.ymean[lane];
    }
    if (has_p && lane == 0) quat_diff(&meas[9], &s.ymean[nlin], &s.innov[nlin]);
    __syncwarp();
    const int m = k;
    for (int e = lane; e < 12 * 24; e += 32) s.Py[e / 24][e % 24] = 0.0;
    __syncwarp();
    for (int e = lane; e < m * m; e += 32) {
        const int r = e / m, c = e % m;
        double v = 0.0;
        for (int i = 0; i < npts; ++i) v += (i == 0 ? w.wc0 : w.wi) * s.DY[i][r] * s.DY[i][c];
        s.Py[r][c] = v;
        s.Py[r][12 + c] = (r == c) ? 1.0 : 0.0;
    }
    for (int e = lane; e < 12 * m; e += 32) {
        const int r = e / m, c = e % m;
        double v = 0.0;
        for (int i = 0; i < npts; ++i) v += (i == 0 ? w.wc0 : w.wi) * s.DX[i][r] * s.DY[i][c];
        s.Pxy[r][c] = v;
    }
    __syncwarp();
    // keep a copy of Py for the covariance update: KPy is computed from it below, so save into A
    for (int e = lane; e < m * m; e += 32) s.A[e / m][e % m] = s.Py[e / m][e % m];
    __syncwarp();
    // Gauss-Jordan inverse of the SPD matrix Py in [Py | I]
    for (int col = 0; col < m; ++col) {
        const double piv = 1.0 / s.Py[col][col];
        __syncwarp();
        if (lane < 24) s.Py[col][lane] *= piv;
        __syncwarp();
        if (lane < m && lane != col) {
            const double f = s.Py[lane][col];
            for (int c = 0; c < 24; ++c) s.Py[lane][c] -= f * s.Py[col][c];
        }
        __syncwarp();
    }
    // K = Pxy Py^-1 (UKFCorrection.cpp:118)
    for (int e = lane; e < 12 * m; e += 32) {
        const int r = e / m, c = e % m;
        double v = 0.0;
        for (int j = 0; j < m; ++j) v += s.Pxy[r][j] * s.Py[j][12 + c];
        s.K[r][c] = v;
    }
    __syncwarp();
    if (lane < 12) {
        double v = 0.0;
        for (int j = 0; j < m; ++j) v += s.K[lane][j] * s.innov[j];
        s.Kn[lane] = v;
    }
    for (int e = lane; e < 12 * m; e += 32) {
        const int r = e / m, c = e % m;
        double v = 0.0;
        for (int j = 0; j < m; ++j) v += s.K[r][j] * s.A[j][c];
        s.KPy[r][c] = v;
    }
    __syncwarp();
    // P <- P - K Py K^T (UKFCorrection.cpp:132); mean update (:125,:128)
    for (int e = lane; e < 144; e += 32) {
        const int r = e / 12, c = e % 12;
        double v = 0.0;
        for (int j = 0; j < m; ++j) v += s.KPy[r][j] * s.K[c][j];
        s.P[r][c] -= v;
    }
    if (lane == 0) {
        double dq[4], qn[4];
        rotvec_to_quat(&s.Kn[9], dq);
        qmul(dq, &s.mean[9], qn);
        for (int i = 0; i < 9; ++i) s.mean[i] += s.Kn[i];
        for (int i = 0; i < 4; ++i) s.mean[9 + i] = qn[i];
    }
    __syncwarp();
}

// One warp per track, 27 KB of shared memory per track, to be resident BESIDE the two CTAs per SM of
// the velocity kernel so that the latency-bound pose filter runs in the issue slots the streaming kernel leaves idle
// instead of after it.  What decides that is the register file of an SM SUB-PARTITION (16384 registers): two velocity
// CTAs put four warps of R registers x 32 lanes there, and a pose warp of U x 32 fits beside them only if
// 4 * 32 R + 32 U <= 16384.  Left to itself ptxas gives this kernel 164 registers; next to it the second velocity CTA
// of every SM it touches cannot start (measured: 33 instead of 71 clusters in flight while a pose kernel runs, and a
// re-sync replay of ~1 ms halves the occupancy of a whole step).  Capped at 128 (no spills) it pairs with R = 96.
struct UkfWarpSmem { UkfSmem s; double meas[16]; };

template <int REGS, int WARPS>
__global__ void __maxnreg__(REGS) k_ukf_batch(UkfArgs a) {
    extern __shared__ __align__(16) unsigned char ukf_smem_raw[];
    const int t = blockIdx.x * WARPS + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= a.n_tracks) return;
    const int nops = a.n_ops[t];
    if (nops <= 0) return;
    span_stamp(a.span_clock, false);
    UkfWarpSmem& wsm = reinterpret_cast<UkfWarpSmem*>(ukf_smem_raw)[threadIdx.x >> 5];
    UkfSmem& s = wsm.s;
    double* meas = wsm.meas;
    double* gm = a.mean + (long long)t * 13;
    double* gc = a.cov + (long long)t * 144;
    double* wv = a.warm ? a.warm + (long long)t * 290 : nullptr;  // [2][144 eigenvectors + use counter]
    const int o_first = a.resume ? a.resume[t] : 0;
    if (o_first < 0 || o_first >= nops) {  // (second launch of a step: this track had no render-and-compare test)
        if (a.resume && lane == 0) a.resume[t] = -1;
        return;
    }
    if (lane < 13) s.mean[lane] = gm[lane];
    for (int i = lane; i < 144; i += 32) s.P[i / 12][i % 12] = gc[i];
    __syncwarp();
    int o_next = -1;
    for (int o = o_first; o < nops; ++o) {
        const UkfOp* op = a.ops + (long long)t * a.max_ops + o;
        const int kind = op->kind;
        if (kind == kOpPredict) {
            ukf_predict_warp(s, a.p, op->dt, lane, wv);
        } else if (kind == kOpCorrect) {
            if (lane < 13) {
                double v = op->meas[lane];
                if (lane < 6 && op->vel_slot >= 0 && a.vel_hist)
                    v = a.vel_hist[((long long)t * a.hist_ring + op->vel_slot) * 6 + lane];
                meas[lane] = v;
            }
            __syncwarp();
            ukf_correct_warp(s, a.p, op->meas_type, meas, lane, wv ? wv + 145 : nullptr);
        } else if (kind == kOpCorrectBoth && a.cand_mean) {
            if (lane < 13) {
                double v = op->meas[lane];
                if (lane < 6 && op->vel_slot >= 0 && a.vel_hist)
                    v = a.vel_hist[((long long)t * a.hist_ring + op->vel_slot) * 6 + lane];
                meas[lane] = v;
            }
            __syncwarp();
            // park the prediction in candidate slot 1, correct with (velocity, pose) -> candidate 0, then restore the
            // prediction and correct with the velocity only (CartesianQuaternionMeasurement.cpp:154-174) -> candidate 1
            double* cm = a.cand_mean + (long long)t * 26;
            double* cc = a.cand_cov + (long long)t * 288;
            if (lane < 13) cm[13 + lane] = s.mean[lane];
            for (int i = lane; i < 144; i += 32) cc[144 + i] = s.P[i / 12][i % 12];
            __syncwarp();
            ukf_correct_warp(s, a.p, ROFTB_MEAS_POSE_VELOCITY, meas, lane, wv ? wv + 145 : nullptr);
            if (lane < 13) {
                cm[lane] = s.mean[lane];
                s.mean[lane] = cm[13 + lane];
            }
            for (int i = lane; i < 144; i += 32) {
                cc[i] = s.P[i / 12][i % 12];
                s.P[i / 12][i % 12] = cc[144 + i];
            }
            __syncwarp();
            ukf_correct_warp(s, a.p, ROFTB_MEAS_VELOCITY, meas, lane, wv ? wv + 145 : nullptr);
            if (lane < 13) cm[13 + lane] = s.mean[lane];
            for (int i = lane; i < 144; i += 32) cc[144 + i] = s.P[i / 12][i % 12];
            __syncwarp();
            // the belief continues from candidate 0 unless k_or_select overwrites it (ROFTFilter.cpp:670-673)
            if (lane < 13) s.mean[lane] = cm[lane];
            for (int i = lane; i < 144; i += 32) s.P[i / 12][i % 12] = cc[i];
            __syncwarp();
            o_next = o + 1;
            break;
        } else if (kind == kOpSwapBuffered && a.buf_mean) {
            // ROFTFilter.cpp:334-340: buffered_belief_ <-> p_corr_belief_
            double* bm = a.buf_mean + (long long)t * 13;
            double* bc = a.buf_cov + (long long)t * 144;
            if (lane < 13) {
                const double v = bm[lane];
                bm[lane] = s.mean[lane];
                s.mean[lane] = v;
            }
            for (int i = lane; i < 144; i += 32) {
                const double v = bc[i];
                bc[i] = s.P[i / 12][i % 12];
                s.P[i / 12][i % 12] = v;
            }
            __syncwarp();
        }
    }
    if (lane < 13) gm[lane] = s.mean[lane];
    for (int i = lane; i < 144; i += 32) gc[i] = s.P[i / 12][i % 12];
    if (a.resume && lane == 0) a.resume[t] = o_next;
    span_stamp(a.span_clock, true);
}

}  // namespace

// Tracks (warps) per CTA.  While a pose warp is resident in an SM sub-partition the second velocity CTA of that SM cannot
// start (see above), so the question is how many SMs the pose kernel touches: one track per CTA spreads 256 tracks over
// all 148 SMs, four per CTA (one warp per sub-partition, 108 KB of shared memory) over 64.  Measured on B200 at 256
// tracks: 53 instead of 33 velocity clusters in flight while the previous step's pose kernel runs, 1.093 instead of
// 1.117 ms per step.  (Eight per CTA would need 216 KB of shared memory and evict both velocity CTAs.)  ROFTB_UKF_WARPS=1|4.
static int ukf_warps() {
    static const int v = [] { const char* e = getenv("ROFTB_UKF_WARPS"); return (e && atoi(e) < 4) ? 1 : 4; }();
    return v;
}

int launch_ukf(const UkfArgs& a, cudaStream_t s) {
    static const int regs = [] { const char* e = getenv("ROFTB_UKF_REGS"); return e ? atoi(e) : 128; }();
    const int w = ukf_warps();
    const int smem = (int)sizeof(UkfWarpSmem) * w;
    const int grid = (a.n_tracks + w - 1) / w;
    if (w == 4) {
        if (regs >= 168)
            ROFTB_LAUNCH((k_ukf_batch<168, 4>), grid, 128, smem, s, a);
        else if (regs >= 128)
            ROFTB_LAUNCH((k_ukf_batch<128, 4>), grid, 128, smem, s, a);
        else
            ROFTB_LAUNCH((k_ukf_batch<96, 4>), grid, 128, smem, s, a);
    } else {
        if (regs >= 168)
            ROFTB_LAUNCH((k_ukf_batch<168, 1>), grid, 32, smem, s, a);
        else if (regs >= 128)
            ROFTB_LAUNCH((k_ukf_batch<128, 1>), grid, 32, smem, s, a);
        else
            ROFTB_LAUNCH((k_ukf_batch<96, 1>), grid, 32, smem, s, a);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// per device (called from roftb_create after cudaSetDevice): the four-track CTAs need the shared-memory opt-in
int ukf_prepare_device() {
    static_assert(sizeof(UkfWarpSmem) <= 48 * 1024, "one track per CTA must fit the default shared-memory limit");
    const int bytes = (int)sizeof(UkfWarpSmem) * 4;
    cudaError_t e = cudaFuncSetAttribute(k_ukf_batch<168, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ukf_batch<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ukf_batch<96, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e == cudaSuccess ? 0 : -1;
}

}  // namespace roftb
