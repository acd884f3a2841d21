// se3.h -- fp64 SE(3) exp / log / adjoint usable from host and device code.
// Conventions follow Sophus (thirdparty/Sophus/sophus/se3.hpp in the reference): tangent = (upsilon, omega),
// exp(xi) = (R = exp(omega^), t = V(omega) upsilon).  Rotation matrices are row-major double[9].
// Used for DSOFrame::setState (DSOFrame.h:110-124), setStateFromCamera (:143-151), setStateZero (:153-186)
// and computeAdjoints (DSOBundleAdjustment.cpp:1071-1092).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SE3_HD __host__ __device__ __forceinline__
#else
#define SE3_HD inline
#endif

namespace cmlba {

struct Pose {       // world->camera (or any rigid transform): x' = R x + t
    double R[9];
    double t[3];
};

SE3_HD void mat3_mul(const double *A, const double *B, double *C) {
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C[r * 3 + c] = A[r * 3 + 0] * B[0 * 3 + c] + A[r * 3 + 1] * B[1 * 3 + c] + A[r * 3 + 2] * B[2 * 3 + c];
}
SE3_HD void mat3_vec(const double *A, const double *v, double *o) {
    for (int r = 0; r < 3; r++) o[r] = A[r * 3 + 0] * v[0] + A[r * 3 + 1] * v[1] + A[r * 3 + 2] * v[2];
}
SE3_HD Pose pose_mul(const Pose &A, const Pose &B) {  // A o B
    Pose C;
    mat3_mul(A.R, B.R, C.R);
    double Rt[3];
    mat3_vec(A.R, B.t, Rt);
    for (int i = 0; i < 3; i++) C.t[i] = Rt[i] + A.t[i];
    return C;
}
SE3_HD Pose pose_inv(const Pose &A) {
    Pose C;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) C.R[r * 3 + c] = A.R[c * 3 + r];
    double Rt[3];
    mat3_vec(C.R, A.t, Rt);
    for (int i = 0; i < 3; i++) C.t[i] = -Rt[i];
    return C;
}
SE3_HD Pose pose_identity() {
    Pose P;
    for (int i = 0; i < 9; i++) P.R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    P.t[0] = P.t[1] = P.t[2] = 0.0;
    return P;
}

// W = hat(w), W2 = W*W
SE3_HD void hat2(const double *w, double *W, double *W2) {
    W[0] = 0; W[1] = -w[2]; W[2] = w[1];
    W[3] = w[2]; W[4] = 0; W[5] = -w[0];
    W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
    mat3_mul(W, W, W2);
}

SE3_HD Pose se3_exp(const double *xi) {
    const double *u = xi, *w = xi + 3;
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2);
    double W[9], W2[9];
    hat2(w, W, W2);
    double a, b, c;  // R = I + a W + b W2 ; V = I + b W + c W2
    if (th < 1e-10) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; c = (th - sin(th)) / (th2 * th); }
    Pose P;
    double V[9];
    for (int i = 0; i < 9; i++) {
        double I = (i % 4 == 0) ? 1.0 : 0.0;
        P.R[i] = I + a * W[i] + b * W2[i];
        V[i] = I + b * W[i] + c * W2[i];
    }
    mat3_vec(V, u, P.t);
    return P;
}

SE3_HD void so3_log(const double *R, double *w) {
    // through the unit quaternion like Sophus::SO3::log (robust for small angles)
    double tr = R[0] + R[4] + R[8];
    double qw, qx, qy, qz;
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2.0;
        qw = 0.25 * s; qx = (R[7] - R[5]) / s; qy = (R[2] - R[6]) / s; qz = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0;
        qw = (R[7] - R[5]) / s; qx = 0.25 * s; qy = (R[1] + R[3]) / s; qz = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0;
        qw = (R[2] - R[6]) / s; qx = (R[1] + R[3]) / s; qy = 0.25 * s; qz = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0;
        qw = (R[3] - R[1]) / s; qx = (R[2] + R[6]) / s; qy = (R[5] + R[7]) / s; qz = 0.25 * s;
    }
    double n2 = qx * qx + qy * qy + qz * qz;
    double n = sqrt(n2);
    double k;
    if (n < 1e-10) {
        k = 2.0 / qw - (2.0 / 3.0) * n2 / (qw * qw * qw);
    } else if (fabs(qw) < 1e-10) {
        k = (qw >= 0 ? M_PI : -M_PI) / n;
    } else {
        k = 2.0 * atan(n / qw) / n;   // Sophus: atan(n/w), no quadrant wrap for w<0 (angle < pi in BA)
    }
    w[0] = k * qx; w[1] = k * qy; w[2] = k * qz;
}

SE3_HD void se3_log(const Pose &P, double *xi) {
    double *w = xi + 3;
    so3_log(P.R, w);
    double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double th = sqrt(th2);
    double W[9], W2[9];
    hat2(w, W, W2);
    double c;
    if (th < 1e-10) c = 1.0 / 12.0;
    else { double half = 0.5 * th; c = (1.0 - th * cos(half) / (2.0 * sin(half))) / th2; }
    double Vi[9];
    for (int i = 0; i < 9; i++) Vi[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * W[i] + c * W2[i];
    mat3_vec(Vi, P.t, xi);
}

// Adj(T) = [[R, hat(t) R],[0, R]]  (6x6 row-major)
SE3_HD void se3_adj(const Pose &P, double *A) {
    double T[9] = {0, -P.t[2], P.t[1], P.t[2], 0, -P.t[0], -P.t[1], P.t[0], 0};
    double TR[9];
    mat3_mul(T, P.R, TR);
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
            A[r * 6 + c] = P.R[r * 3 + c];
            A[r * 6 + 3 + c] = TR[r * 3 + c];
            A[(r + 3) * 6 + c] = 0.0;
            A[(r + 3) * 6 + 3 + c] = P.R[r * 3 + c];
        }
}

}  // namespace cmlba
