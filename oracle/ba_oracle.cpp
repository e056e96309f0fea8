/*
 * oracle/ba_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference bundle-adjustment path).
 *
 * Restates, in plain FP64 C++, what mage::BundlerLib does through g2o for camera/point/observation problems
 * ("ref" = /root/reference/Dependencies): BundlerLib/Source/BundlerLib.cpp (StepOptimizer, StepBundleAdjustment,
 * setters/getters), g2o/g2o/core/optimization_algorithm_levenberg.cpp (LM control), core/block_solver.hpp
 * (buildSystem / setLambda / Schur solve), core/sparse_optimizer.cpp (active sets, index mapping, update),
 * core/base_binary_edge.hpp + robust_kernel_impl.cpp (quadratic form, Huber), types/sba/types_six_dof_expmap.*
 * (projection error + Jacobians), types/slam3d/se3quat.h (exp, composition), solvers/dense/linear_solver_dense.h
 * + Eigen LDLT (dense solve, positivity test).
 *
 * Parity status: PINNED -- validated against the reference's own code compiled into oracle/_ref/libbundler_ref.so
 * (tests/test_ba_oracle.py compares poses, points, lambda and outlier sets step by step). Tether edges
 * (BundlerLib.cpp:24-90 EdgeScaleConstraint / EdgeRotationConstraint with g2o's numeric BaseMultiEdge Jacobians,
 * core/base_multi_edge.hpp:68-118; g2o EdgeSE3Expmap, types_six_dof_expmap.h:108-127 / .cpp:278-293; setters
 * BundlerLib.cpp:311-350) are restated too. One documented deviation: the reference's closing loop (BundlerLib.cpp:389-427)
 * walks ALL active edges and static_casts vertex(0) of a tether edge to VertexSBAPointXYZ -- it reads a camera vertex as if
 * it were a point, so whether a tether edge's squared error enters the returned mean is undefined behaviour. Here (and in
 * the product) tether edges never enter the mean or the outlier list ("only remove point if it is a camera/point edge").
 *
 * Nothing under mageslam_b200/ may include, link or call this file.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

struct Quat { double x = 0, y = 0, z = 0, w = 1; };
struct Vec3 { double v[3] = {0, 0, 0}; double& operator[](int i) { return v[i]; } double operator[](int i) const { return v[i]; } };

inline Quat qmul(const Quat& a, const Quat& b)
{
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
inline Vec3 qrot(const Quat& q, const Vec3& p)       // Eigen QuaternionBase::_transformVector
{
    double ux = q.y * p[2] - q.z * p[1], uy = q.z * p[0] - q.x * p[2], uz = q.x * p[1] - q.y * p[0];
    ux += ux; uy += uy; uz += uz;
    Vec3 r;
    r[0] = p[0] + q.w * ux + (q.y * uz - q.z * uy);
    r[1] = p[1] + q.w * uy + (q.z * ux - q.x * uz);
    r[2] = p[2] + q.w * uz + (q.x * uy - q.y * ux);
    return r;
}
inline void qtoR(const Quat& q, double R[9])          // row-major, Eigen toRotationMatrix
{
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
template <class T> inline void RtoQ(const T m[9] /*row-major*/, T q[4] /*x y z w*/)   // Eigen quaternionbase_assign_impl
{
    T t = m[0] + m[4] + m[8];
    if (t > T(0)) {
        t = std::sqrt(t + T(1));
        q[3] = T(0.5) * t;
        t = T(0.5) / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + T(1));
        q[i] = T(0.5) * t;
        t = T(0.5) / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}
inline void qnormalizeRotation(Quat& q)                // ref se3quat.h:280-285
{
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

struct Pose { Quat r; Vec3 t; };

// ref se3quat.h:218-257 SE3Quat::exp, then operator* (:99-105): T <- exp(update) * T
inline Pose poseOplus(const Pose& T, const double* u)
{
    double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
    double theta = std::sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
    double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += O[i * 3 + k] * O[k * 3 + j]; O2[i * 3 + j] = s; }
    double a, b, c, d;
    if (theta < 0.00001) { a = 1; b = 0.5; c = 0.5; d = 1. / 6.; }
    else {
        a = std::sin(theta) / theta; b = (1 - std::cos(theta)) / (theta * theta);
        c = b; d = (theta - std::sin(theta)) / (std::pow(theta, 3));
    }
    double R[9], V[9];
    for (int i = 0; i < 9; i++) { double I = (i % 4 == 0) ? 1.0 : 0.0; R[i] = I + a * O[i] + b * O2[i]; V[i] = I + c * O[i] + d * O2[i]; }
    double q4[4];
    RtoQ<double>(R, q4);
    Pose E;
    E.r.x = q4[0]; E.r.y = q4[1]; E.r.z = q4[2]; E.r.w = q4[3];
    for (int i = 0; i < 3; i++) E.t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
    qnormalizeRotation(E.r);                           // SE3Quat(q, t) ctor normalises
    Pose out;
    Vec3 rt = qrot(E.r, T.t);
    for (int i = 0; i < 3; i++) out.t[i] = E.t[i] + rt[i];
    out.r = qmul(E.r, T.r);
    qnormalizeRotation(out.r);
    return out;
}

// ref se3quat.h:120-125 inverse, :99-105 operator*
inline Pose poseInverse(const Pose& T)
{
    Pose r;
    r.r = T.r; r.r.x = -r.r.x; r.r.y = -r.r.y; r.r.z = -r.r.z;
    Vec3 nt; for (int i = 0; i < 3; i++) nt[i] = T.t[i] * -1.;
    r.t = qrot(r.r, nt);
    return r;
}
inline Pose poseMul(const Pose& a, const Pose& b)
{
    Pose r = a;
    Vec3 rt = qrot(a.r, b.t);
    for (int i = 0; i < 3; i++) r.t[i] += rt[i];
    r.r = qmul(a.r, b.r);
    qnormalizeRotation(r.r);
    return r;
}
inline void mat3mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[i * 3 + k] * B[k * 3 + j]; C[i * 3 + j] = s; }
}
// ref se3quat.h:171-210 SE3Quat::log -> (omega, upsilon)
inline void poseLog(const Pose& T, double res[6])
{
    double R[9]; qtoR(T.r, R);
    double d = 0.5 * (R[0] + R[4] + R[8] - 1);
    double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};                 // deltaR (se3_ops.hpp)
    double omega[3], Vinv[9];
    auto skewOf = [](const double* o, double* O) { O[0] = 0; O[1] = -o[2]; O[2] = o[1]; O[3] = o[2]; O[4] = 0; O[5] = -o[0]; O[6] = -o[1]; O[7] = o[0]; O[8] = 0; };
    double O[9], O2[9];
    if (std::abs(d) > 0.99999) {
        for (int i = 0; i < 3; i++) omega[i] = 0.5 * dR[i];
        skewOf(omega, O); mat3mul(O, O, O2);
        for (int i = 0; i < 9; i++) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + (1. / 12.) * O2[i];
    } else {
        double theta = std::acos(d);
        double k = theta / (2 * std::sqrt(1 - d * d));
        for (int i = 0; i < 3; i++) omega[i] = k * dR[i];
        skewOf(omega, O); mat3mul(O, O, O2);
        double c = (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
        for (int i = 0; i < 9; i++) Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * O[i] + c * O2[i];
    }
    for (int i = 0; i < 3; i++) res[i] = omega[i];
    for (int i = 0; i < 3; i++) res[3 + i] = Vinv[i * 3] * T.t[0] + Vinv[i * 3 + 1] * T.t[1] + Vinv[i * 3 + 2] * T.t[2];
}
// ref se3quat.h:217-226 adj(): [R 0; skew(t) R  R], row-major 6x6
inline void poseAdj(const Pose& T, double A[36])
{
    double R[9]; qtoR(T.r, R);
    double Sk[9] = {0, -T.t[2], T.t[1], T.t[2], 0, -T.t[0], -T.t[1], T.t[0], 0}, SR[9];
    mat3mul(Sk, R, SR);
    for (int i = 0; i < 36; i++) A[i] = 0;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { A[r * 6 + c] = R[r * 3 + c]; A[(3 + r) * 6 + 3 + c] = R[r * 3 + c]; A[(3 + r) * 6 + c] = SR[r * 3 + c]; }
}

inline bool inv3(const double* A, double* Ai)           // Eigen 3x3 inverse: cofactors / determinant
{
    double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    double id = 1.0 / det;
    Ai[0] = c00 * id; Ai[1] = (A[2] * A[7] - A[1] * A[8]) * id; Ai[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    Ai[3] = c01 * id; Ai[4] = (A[0] * A[8] - A[2] * A[6]) * id; Ai[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    Ai[6] = c02 * id; Ai[7] = (A[1] * A[6] - A[0] * A[7]) * id; Ai[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    return true;
}

// Eigen::LDLT<MatrixXd> (lower, diagonal pivoting on the not-yet-updated diagonal) + isPositive + solve.
// ref eigen-git-mirror/Eigen/src/Cholesky/LDLT.h:294-400, :558-600; linear_solver_dense.h:104-112
bool ldltSolve(std::vector<double>& A, int n, const double* b, double* x)
{
    if (n == 0) return true;
    std::vector<int> tr(n);
    std::vector<double> temp(n);
    int sign = 0;    // 0 zero, 1 possemidef, -1 negsemidef, 2 indefinite
    auto M = [&](int r, int c) -> double& { return A[(size_t)r * n + c]; };
    if (n == 1) { tr[0] = 0; sign = M(0, 0) > 0 ? 1 : (M(0, 0) < 0 ? -1 : 0); }
    else {
        bool zeroAll = false;
        for (int k = 0; k < n && !zeroAll; ++k) {
            int big = k; double best = std::fabs(M(k, k));
            for (int i = k + 1; i < n; i++) if (std::fabs(M(i, i)) > best) { best = std::fabs(M(i, i)); big = i; }
            tr[k] = big;
            if (k != big) {
                int s = n - big - 1;
                for (int c = 0; c < k; c++) std::swap(M(k, c), M(big, c));
                for (int r = 0; r < s; r++) std::swap(M(big + 1 + r, k), M(big + 1 + r, big));
                std::swap(M(k, k), M(big, big));
                for (int i = k + 1; i < big; ++i) { double tmp = M(i, k); M(i, k) = M(big, i); M(big, i) = tmp; }
            }
            int rs = n - k - 1;
            if (k > 0) {
                for (int c = 0; c < k; c++) temp[c] = M(c, c) * M(k, c);
                double s = 0; for (int c = 0; c < k; c++) s += M(k, c) * temp[c];
                M(k, k) -= s;
                for (int r = 0; r < rs; r++) { double a = 0; for (int c = 0; c < k; c++) a += M(k + 1 + r, c) * temp[c]; M(k + 1 + r, k) -= a; }
            }
            double akk = M(k, k);
            bool valid = std::fabs(akk) > 0;
            if (k == 0 && !valid) { sign = 0; for (int j = 0; j < n; j++) tr[j] = j; zeroAll = true; break; }
            if (rs > 0 && valid) for (int r = 0; r < rs; r++) M(k + 1 + r, k) /= akk;
            if (sign == 1) { if (akk < 0) sign = 2; }
            else if (sign == -1) { if (akk > 0) sign = 2; }
            else if (sign == 0) { if (akk > 0) sign = 1; else if (akk < 0) sign = -1; }
        }
    }
    if (!(sign == 1 || sign == 0)) return false;      // _cholesky.isPositive()
    std::vector<double> y(b, b + n);
    for (int k = 0; k < n; k++) std::swap(y[k], y[tr[k]]);
    for (int i = 0; i < n; i++) { double s = y[i]; for (int c = 0; c < i; c++) s -= M(i, c) * y[c]; y[i] = s; }
    const double tol = 1.0 / std::numeric_limits<double>::max();
    for (int i = 0; i < n; i++) { if (std::fabs(M(i, i)) > tol) y[i] /= M(i, i); else y[i] = 0; }
    for (int i = n - 1; i >= 0; i--) { double s = y[i]; for (int r = i + 1; r < n; r++) s -= M(r, i) * y[r]; y[i] = s; }
    for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[tr[k]]);
    std::copy(y.begin(), y.end(), x);
    return true;
}

struct Camera { Pose T; double f = 0, cx = 0, cy = 0; bool fixed = false; bool set = false; int hidx = -1; };
struct Point { Vec3 X; bool set = false; int hidx = -1; };
struct Obs { double u = 0, v = 0, info = 0; int cam = -1, pt = -1; bool set = false, removed = false; long seq = -1; double err[2] = {0, 0}; };

// tether edges between two cameras: 0 = EdgeScaleConstraint (BundlerLib.cpp:24-55), 1 = EdgeRotationConstraint (:57-90),
// 2 = g2o EdgeSE3Expmap (RelativeTransformConstraints)
struct Tether {
    int type = -1, c1 = -1, c2 = -1, dim = 0;
    double dist = 0, w = 0;       // measurement (type 0), weight (types 0/1: multiplies the error; type 2: information = w I)
    Quat q; Pose C;               // measurement of type 1 / type 2
    bool set = false; long seq = -1;
    double err[6] = {0, 0, 0, 0, 0, 0};
    double J[2][36];              // dim x 6, row-major, per vertex
};

struct BA {
    bool pointsFixed = false;
    std::vector<Camera> cams; std::vector<Point> pts; std::vector<Obs> obs;
    std::vector<Tether> teth[3];
    std::vector<Tether*> activeT;       // active tether edges, insertion order
    long nextSeq = 0;
    // optimizer state (StepOptimizer + OptimizationAlgorithmLevenberg)
    bool dirty = true, useless = false;
    int iteration = 0;
    double lambda = -1, ni = 2, userLambdaInit = 0, huber = 0;
    std::vector<int> active;            // active observation indices, insertion order
    std::vector<int> camOrder, ptOrder; // free cameras / free points in Hessian order
    int sizePoses = 0, sizeLandmarks = 0;
    std::vector<double> x, b;           // full solution / rhs (poses first, then landmarks)
    std::vector<double> Hpp, Hll, W;    // dense (6Kf)^2 row-major; 9 per free point; 18 per active obs (6x3 row-major)
};

inline void computeTetherErrors(BA& s);

// ref types_six_dof_expmap.h:140-147 computeError (+ cam_map, SE3Quat::map)
inline void computeErrors(BA& s)
{
    for (int e : s.active) {
        Obs& o = s.obs[e];
        const Camera& c = s.cams[o.cam];
        Vec3 p = qrot(c.T.r, s.pts[o.pt].X);
        for (int i = 0; i < 3; i++) p[i] += c.T.t[i];
        double pu = p[0] / p[2], pv = p[1] / p[2];
        o.err[0] = o.u - (pu * c.f + c.cx);
        o.err[1] = o.v - (pv * c.f + c.cy);
    }
    computeTetherErrors(s);
}
inline void tetherError(const BA& s, const Tether& t, double* err)
{
    const Pose& T1 = s.cams[t.c1].T; const Pose& T2 = s.cams[t.c2].T;
    if (t.type == 0) {
        double dx = T2.t[0] - T1.t[0], dy = T2.t[1] - T1.t[1], dz = T2.t[2] - T1.t[2];
        err[0] = (t.dist - std::sqrt(dx * dx + dy * dy + dz * dz)) * t.w;
    } else if (t.type == 1) {
        Pose rel = poseMul(poseInverse(T1), T2);
        // Eigen angularDistance: d = this * other.conjugate(); 2 * atan2(d.vec().norm(), |d.w|)
        Quat mc = t.q; mc.x = -mc.x; mc.y = -mc.y; mc.z = -mc.z;
        Quat d = qmul(rel.r, mc);
        err[0] = 2 * std::atan2(std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z), std::abs(d.w)) * t.w;
    } else {
        Pose E = poseMul(poseMul(poseInverse(T2), t.C), T1);
        poseLog(E, err);
    }
}
inline void computeTetherErrors(BA& s)
{
    for (Tether* t : s.activeT) tetherError(s, *t, t->err);
}
inline double tetherChi2(const Tether& t)
{
    double ss = 0;
    for (int i = 0; i < t.dim; i++) ss += t.err[i] * ((t.type == 2 ? t.w : 1.0) * t.err[i]);       // e^T Omega e
    return ss;
}
// ref robust_kernel_impl.cpp:65-78 (Huber on the squared error) and sparse_optimizer.cpp:102-117
inline void huberRho(double e, double delta, double rho[3])
{
    double dsqr = delta * delta;
    if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else { double sq = std::sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e; }
}
inline double robustChi2(const BA& s)
{
    double chi = 0, rho[3];
    for (int e : s.active) {
        const Obs& o = s.obs[e];
        double chi2 = o.info * (o.err[0] * o.err[0] + o.err[1] * o.err[1]);       // e^T (info I) e
        huberRho(chi2, s.huber, rho);
        chi += rho[0];
    }
    for (const Tether* t : s.activeT) chi += tetherChi2(*t);               // no robust kernel on tether edges
    return chi;
}

// ref sparse_optimizer.cpp:208-272 initializeOptimization + :168-192 buildIndexMapping; BundlerLib.cpp:156-166
void initializeOptimization(BA& s)
{
    s.active.clear();
    std::vector<std::pair<long, int>> order;
    for (int e = 0; e < (int)s.obs.size(); e++) {
        const Obs& o = s.obs[e];
        if (!o.set || o.removed) continue;
        if (s.pointsFixed && s.cams[o.cam].fixed) continue;            // allVerticesFixed
        order.push_back({o.seq, e});
    }
    std::sort(order.begin(), order.end());                              // EdgeIDCompare: internalId = insertion order
    for (auto& pr : order) s.active.push_back(pr.second);
    std::vector<char> camActive(s.cams.size(), 0), ptActive(s.pts.size(), 0);
    for (int e : s.active) { camActive[s.obs[e].cam] = 1; ptActive[s.obs[e].pt] = 1; }
    s.activeT.clear();
    for (auto& v : s.teth) for (Tether& t : v) {
        if (!t.set) continue;
        if (s.cams[t.c1].fixed && s.cams[t.c2].fixed) continue;          // allVerticesFixed
        s.activeT.push_back(&t);
        camActive[t.c1] = 1; camActive[t.c2] = 1;
    }
    std::sort(s.activeT.begin(), s.activeT.end(), [](const Tether* a, const Tether* b) { return a->seq < b->seq; });
    for (auto& c : s.cams) c.hidx = -1;
    for (auto& p : s.pts) p.hidx = -1;
    s.camOrder.clear(); s.ptOrder.clear();
    int idx = 0;
    for (int k = 0; k < (int)s.cams.size(); k++) if (camActive[k] && !s.cams[k].fixed) { s.cams[k].hidx = idx++; s.camOrder.push_back(k); }
    // point vertex ids count DOWN from INT_MAX-2 (BundlerLib.cpp:210-218) => ascending id = descending point index
    if (!s.pointsFixed)
        for (int i = (int)s.pts.size() - 1; i >= 0; i--) if (ptActive[i]) { s.pts[i].hidx = idx++; s.ptOrder.push_back(i); }
    s.useless = idx == 0;
    s.iteration = 0;
    s.dirty = false;
}

void buildStructure(BA& s)
{
    s.sizePoses = 6 * (int)s.camOrder.size();
    s.sizeLandmarks = 3 * (int)s.ptOrder.size();
    s.x.assign(s.sizePoses + s.sizeLandmarks, 0.0);
    s.b.assign(s.sizePoses + s.sizeLandmarks, 0.0);
    s.Hpp.assign((size_t)s.sizePoses * s.sizePoses, 0.0);
    s.Hll.assign((size_t)3 * s.sizeLandmarks, 0.0);
    s.W.assign((size_t)18 * s.active.size(), 0.0);
}

void addTethers(BA& s);

// ref block_solver.hpp:463-521 buildSystem; types_six_dof_expmap.cpp:295-331 linearizeOplus;
// base_binary_edge.hpp:62-134 constructQuadraticForm (robust branch)
void buildSystem(BA& s)
{
    std::fill(s.b.begin(), s.b.end(), 0.0);
    std::fill(s.Hpp.begin(), s.Hpp.end(), 0.0);
    std::fill(s.Hll.begin(), s.Hll.end(), 0.0);
    std::fill(s.W.begin(), s.W.end(), 0.0);
    const int np = (int)s.camOrder.size(), n = s.sizePoses;
    for (size_t a = 0; a < s.active.size(); a++) {
        const Obs& o = s.obs[s.active[a]];
        const Camera& c = s.cams[o.cam];
        const Point& P = s.pts[o.pt];
        Vec3 xt = qrot(c.T.r, P.X);
        for (int i = 0; i < 3; i++) xt[i] += c.T.t[i];
        double x = xt[0], y = xt[1], z = xt[2], z2 = z * z, f = c.f;
        double R[9]; qtoR(c.T.r, R);
        double tmp[6] = {f, 0, -x / z * f, 0, f, -y / z * f};
        double Ji[6];   // 2x3 wrt point
        for (int r = 0; r < 2; r++) for (int cc = 0; cc < 3; cc++) {
            double sum = 0; for (int k = 0; k < 3; k++) sum += (-1. / z * tmp[r * 3 + k]) * R[k * 3 + cc];
            Ji[r * 3 + cc] = sum;
        }
        double Jj[12];  // 2x6 wrt pose (omega, upsilon)
        Jj[0] = x * y / z2 * f; Jj[1] = -(1 + (x * x / z2)) * f; Jj[2] = y / z * f; Jj[3] = -1. / z * f; Jj[4] = 0; Jj[5] = x / z2 * f;
        Jj[6] = (1 + y * y / z2) * f; Jj[7] = -x * y / z2 * f; Jj[8] = -x / z * f; Jj[9] = 0; Jj[10] = -1. / z * f; Jj[11] = y / z2 * f;
        double chi2 = o.info * (o.err[0] * o.err[0] + o.err[1] * o.err[1]);
        double rho[3]; huberRho(chi2, s.huber, rho);
        double wOmega = rho[1] * o.info;
        double omr[2] = {-o.info * o.err[0] * rho[1], -o.info * o.err[1] * rho[1]};
        const int hi = P.hidx, hj = c.hidx;
        if (hi >= 0) {
            int li = hi - np;
            double* bl = &s.b[n + 3 * li];
            double* H = &s.Hll[(size_t)9 * li];
            for (int r = 0; r < 3; r++) {
                bl[r] += Ji[r] * omr[0] + Ji[3 + r] * omr[1];
                for (int cc = 0; cc < 3; cc++) H[r * 3 + cc] += wOmega * (Ji[r] * Ji[cc] + Ji[3 + r] * Ji[3 + cc]);
            }
        }
        if (hj >= 0) {
            double* bp = &s.b[6 * hj];
            for (int r = 0; r < 6; r++) {
                bp[r] += Jj[r] * omr[0] + Jj[6 + r] * omr[1];
                for (int cc = 0; cc < 6; cc++) s.Hpp[(size_t)(6 * hj + r) * n + 6 * hj + cc] += wOmega * (Jj[r] * Jj[cc] + Jj[6 + r] * Jj[6 + cc]);
            }
        }
        if (hi >= 0 && hj >= 0) {
            double* W = &s.W[18 * a];       // pose (6) x point (3)
            for (int r = 0; r < 6; r++) for (int cc = 0; cc < 3; cc++) W[r * 3 + cc] += wOmega * (Jj[r] * Ji[cc] + Jj[6 + r] * Ji[3 + cc]);
        }
    }
    addTethers(s);
}

// Tether edges: linearisation (numeric for the BaseMultiEdge types, ref base_multi_edge.hpp:68-118: central differences with
// delta = 1e-9 through oplus on each free vertex; analytic adjoints for EdgeSE3Expmap, ref types_six_dof_expmap.cpp:278-293)
// and quadratic form (ref base_multi_edge.hpp:155-197, base_binary_edge.hpp:75-103): H_ii += A^T O A, b_i += A^T (-O e),
// H_ij += A^T O B.
void linearizeTether(BA& s, Tether& t)
{
    const int cam[2] = {t.c1, t.c2};
    if (t.type == 2) {
        const Pose& Ti = s.cams[t.c1].T; const Pose& Tj = s.cams[t.c2].T;
        Pose invTij = poseInverse(t.C);
        Pose invTj_Tij = poseMul(poseInverse(Tj), t.C);
        Pose invTi_invTij = poseMul(poseInverse(Ti), invTij);
        poseAdj(invTj_Tij, t.J[0]);
        poseAdj(invTi_invTij, t.J[1]);
        for (int i = 0; i < 36; i++) t.J[1][i] = -t.J[1][i];
        return;
    }
    const double delta = 1e-9, scalar = 1 / (2 * delta);
    for (int v = 0; v < 2; v++) {
        Camera& c = s.cams[cam[v]];
        if (c.fixed) continue;
        for (int d = 0; d < 6; d++) {
            double add[6] = {0, 0, 0, 0, 0, 0}, ep[6], em[6];
            const Pose keep = c.T;
            add[d] = delta;  c.T = poseOplus(keep, add); tetherError(s, t, ep); c.T = keep;
            add[d] = -delta; c.T = poseOplus(keep, add); tetherError(s, t, em); c.T = keep;
            for (int r = 0; r < t.dim; r++) t.J[v][r * 6 + d] = scalar * (ep[r] - em[r]);
        }
    }
}
void addTethers(BA& s)
{
    const int n = s.sizePoses;
    for (Tether* tp : s.activeT) {
        Tether& t = *tp;
        linearizeTether(s, t);
        const double om = (t.type == 2) ? t.w : 1.0;                       // information = om * I
        const int h[2] = {s.cams[t.c1].hidx, s.cams[t.c2].hidx};
        const bool freeV[2] = {!s.cams[t.c1].fixed, !s.cams[t.c2].fixed};
        for (int v = 0; v < 2; v++) {
            if (!freeV[v]) continue;
            const double* A = t.J[v];
            for (int r = 0; r < 6; r++) {
                double bb = 0;
                for (int k = 0; k < t.dim; k++) bb += A[k * 6 + r] * (-(om * t.err[k]));
                s.b[6 * h[v] + r] += bb;
                for (int c = 0; c < 6; c++) {
                    double hh = 0;
                    for (int k = 0; k < t.dim; k++) hh += (A[k * 6 + r] * om) * A[k * 6 + c];
                    s.Hpp[(size_t)(6 * h[v] + r) * n + 6 * h[v] + c] += hh;
                }
            }
        }
        if (freeV[0] && freeV[1]) {
            const double* A = t.J[0]; const double* B = t.J[1];
            if (h[0] == h[1]) continue;                                     // (degenerate: both ends the same camera)
            for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
                double hh = 0;
                for (int k = 0; k < t.dim; k++) hh += (A[k * 6 + r] * om) * B[k * 6 + c];
                s.Hpp[(size_t)(6 * h[0] + r) * n + 6 * h[1] + c] += hh;
                s.Hpp[(size_t)(6 * h[1] + c) * n + 6 * h[0] + r] += hh;
            }
        }
    }
}

// ref block_solver.hpp:315-447 (Schur complement, dense solve, back substitution) with lambda on every diagonal (:525-549)
bool solveSystem(BA& s, double lambda)
{
    const int n = s.sizePoses, np = (int)s.camOrder.size(), nl = (int)s.ptOrder.size();
    std::vector<double> S(s.Hpp);
    for (int i = 0; i < n; i++) S[(size_t)i * n + i] += lambda;
    std::vector<double> coeff(n, 0.0), Dinv((size_t)9 * nl), bs(n);
    // per landmark edge lists
    std::vector<std::vector<int>> pe(nl);
    for (size_t a = 0; a < s.active.size(); a++) {
        const Obs& o = s.obs[s.active[a]];
        if (s.pts[o.pt].hidx >= 0 && s.cams[o.cam].hidx >= 0) pe[s.pts[o.pt].hidx - np].push_back((int)a);
    }
    for (int li = 0; li < nl; li++) {
        double D[9];
        for (int i = 0; i < 9; i++) D[i] = s.Hll[(size_t)9 * li + i];
        D[0] += lambda; D[4] += lambda; D[8] += lambda;
        double* Di = &Dinv[(size_t)9 * li];
        inv3(D, Di);
        double db[3];
        for (int r = 0; r < 3; r++) db[r] = Di[r * 3] * s.b[n + 3 * li] + Di[r * 3 + 1] * s.b[n + 3 * li + 1] + Di[r * 3 + 2] * s.b[n + 3 * li + 2];
        for (int a1 : pe[li]) {
            const int i1 = s.cams[s.obs[s.active[a1]].cam].hidx;
            const double* Bi = &s.W[18 * a1];
            double BD[18];
            for (int r = 0; r < 6; r++) for (int c = 0; c < 3; c++) BD[r * 3 + c] = Bi[r * 3] * Di[c] + Bi[r * 3 + 1] * Di[3 + c] + Bi[r * 3 + 2] * Di[6 + c];
            for (int r = 0; r < 6; r++) coeff[6 * i1 + r] += Bi[r * 3] * db[0] + Bi[r * 3 + 1] * db[1] + Bi[r * 3 + 2] * db[2];
            for (int a2 : pe[li]) {
                const int i2 = s.cams[s.obs[s.active[a2]].cam].hidx;
                const double* Bj = &s.W[18 * a2];
                for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++)
                    S[(size_t)(6 * i1 + r) * n + 6 * i2 + c] -= BD[r * 3] * Bj[c * 3] + BD[r * 3 + 1] * Bj[c * 3 + 1] + BD[r * 3 + 2] * Bj[c * 3 + 2];
            }
        }
    }
    for (int i = 0; i < n; i++) bs[i] = s.b[i] - coeff[i];
    if (!ldltSolve(S, n, bs.data(), s.x.data())) return false;        // x keeps its previous content on failure
    for (int li = 0; li < nl; li++) {
        double cl[3] = {s.b[n + 3 * li], s.b[n + 3 * li + 1], s.b[n + 3 * li + 2]};
        for (int a1 : pe[li]) {
            const int i1 = s.cams[s.obs[s.active[a1]].cam].hidx;
            const double* Bi = &s.W[18 * a1];
            for (int c = 0; c < 3; c++) for (int r = 0; r < 6; r++) cl[c] -= Bi[r * 3 + c] * s.x[6 * i1 + r];
        }
        const double* Di = &Dinv[(size_t)9 * li];
        for (int r = 0; r < 3; r++) s.x[n + 3 * li + r] = Di[r * 3] * cl[0] + Di[r * 3 + 1] * cl[1] + Di[r * 3 + 2] * cl[2];
    }
    return true;
}

// ref optimization_algorithm_levenberg.cpp:57-149
bool lmSolve(BA& s)
{
    if (s.iteration == 0) buildStructure(s);
    computeErrors(s);
    double currentChi = robustChi2(s), tempChi = currentChi;
    buildSystem(s);
    const int np = (int)s.camOrder.size();
    if (s.iteration == 0) {
        if (s.userLambdaInit > 0) s.lambda = s.userLambdaInit;
        else {
            double maxDiag = 0;
            for (int i = 0; i < s.sizePoses; i++) maxDiag = std::max(std::fabs(s.Hpp[(size_t)i * s.sizePoses + i]), maxDiag);
            for (size_t li = 0; li < s.ptOrder.size(); li++) for (int j = 0; j < 3; j++) maxDiag = std::max(std::fabs(s.Hll[9 * li + 4 * j]), maxDiag);
            s.lambda = 1e-5 * maxDiag;
        }
        s.ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    std::vector<Pose> camBackup(np); std::vector<Vec3> ptBackup(s.ptOrder.size());
    do {
        for (int k = 0; k < np; k++) camBackup[k] = s.cams[s.camOrder[k]].T;                  // push
        for (size_t i = 0; i < s.ptOrder.size(); i++) ptBackup[i] = s.pts[s.ptOrder[i]].X;
        bool ok2 = solveSystem(s, s.lambda);
        for (int k = 0; k < np; k++) s.cams[s.camOrder[k]].T = poseOplus(s.cams[s.camOrder[k]].T, &s.x[6 * k]);      // update
        for (size_t i = 0; i < s.ptOrder.size(); i++) for (int r = 0; r < 3; r++) s.pts[s.ptOrder[i]].X[r] += s.x[s.sizePoses + 3 * i + r];
        computeErrors(s);
        tempChi = robustChi2(s);
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = currentChi - tempChi;
        double scale = 0;
        for (size_t j = 0; j < s.x.size(); j++) scale += s.x[j] * (s.lambda * s.x[j] + s.b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
            double alpha = 1. - std::pow((2 * rho - 1), 3);
            alpha = std::min(alpha, 2. / 3.);
            double scaleFactor = std::max(1. / 3., alpha);
            s.lambda *= scaleFactor;
            s.ni = 2;
            currentChi = tempChi;
        } else {
            s.lambda *= s.ni;
            s.ni *= 2;
            for (int k = 0; k < np; k++) s.cams[s.camOrder[k]].T = camBackup[k];              // pop
            for (size_t i = 0; i < s.ptOrder.size(); i++) s.pts[s.ptOrder[i]].X = ptBackup[i];
            if (!std::isfinite(s.lambda)) break;
        }
        qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0 || !std::isfinite(s.lambda)) return false;      // Terminate
    return true;
}

// ref BundlerLib.cpp:132-149 StepOptimizer::Step
bool stepOnce(BA& s)
{
    if (s.dirty) initializeOptimization(s);
    if (s.useless) return false;
    bool ok = lmSolve(s);
    s.iteration++;
    return ok;
}

} // namespace

extern "C" {

void* baorc_create(int are_points_fixed) { BA* s = new BA(); s->pointsFixed = are_points_fixed != 0; return s; }
void baorc_destroy(void* h) { delete static_cast<BA*>(h); }
void baorc_alloc(void* h, int cams, int points, int obs)
{
    BA* s = static_cast<BA*>(h);
    s->cams.resize(cams); s->pts.resize(points); s->obs.resize(obs);
}
// ref BundlerLib.cpp:261-276: Quaternionf(R).normalized() in float, cast to double, SE3Quat ctor normalises again
void baorc_set_camera(void* h, int idx, const float* pos, const float* rot_colmajor, const float* intr, int fixed)
{
    BA* s = static_cast<BA*>(h);
    Camera& c = s->cams[idx];
    float m[9];
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) m[r * 3 + cc] = rot_colmajor[cc * 3 + r];
    float q[4];
    RtoQ<float>(m, q);
    float nn = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (nn > 0.f) { float inv = std::sqrt(nn); for (int i = 0; i < 4; i++) q[i] = q[i] / inv; }
    c.T.r.x = q[0]; c.T.r.y = q[1]; c.T.r.z = q[2]; c.T.r.w = q[3];
    qnormalizeRotation(c.T.r);
    for (int i = 0; i < 3; i++) c.T.t[i] = pos[i];
    c.f = intr[2]; c.cx = intr[0]; c.cy = intr[1];
    c.fixed = fixed != 0; c.set = true;
    s->dirty = true;
}
void baorc_fix_camera(void* h, int idx, int value) { static_cast<BA*>(h)->cams[idx].fixed = value != 0; }   // does NOT dirty (ref :278-281)
void baorc_set_point(void* h, int idx, const float* xyz)
{
    BA* s = static_cast<BA*>(h);
    for (int i = 0; i < 3; i++) s->pts[idx].X[i] = xyz[i];
    s->pts[idx].set = true; s->dirty = true;
}
void baorc_set_observation(void* h, int idx, const float* uv, int cam, int pt, float info)
{
    BA* s = static_cast<BA*>(h);
    Obs& o = s->obs[idx];
    o.u = uv[0]; o.v = uv[1]; o.cam = cam; o.pt = pt; o.info = info; o.set = true; o.removed = false; o.seq = s->nextSeq++;
    s->dirty = true;
}
// ref BundlerLib.cpp:243-259, :311-350. q = (x, y, z, w). -1 leaves a pool untouched.
void baorc_alloc_tethers(void* h, int n_distance, int n_rotation, int n_transform)
{
    BA* s = static_cast<BA*>(h);
    const int n[3] = {n_distance, n_rotation, n_transform};
    for (int k = 0; k < 3; k++) if (n[k] >= 0) { s->teth[k].assign(n[k], Tether()); }
}
void baorc_set_fixed_distance(void* h, int idx, int cam1, int cam2, float distance, float weight)
{
    BA* s = static_cast<BA*>(h);
    Tether& t = s->teth[0][idx];
    t.type = 0; t.dim = 1; t.c1 = cam1; t.c2 = cam2; t.dist = distance; t.w = weight; t.set = true; t.seq = s->nextSeq++;
    s->dirty = true;
}
void baorc_set_relative_rotation(void* h, int idx, int cam1, int cam2, const float* q, float weight)
{
    BA* s = static_cast<BA*>(h);
    Tether& t = s->teth[1][idx];
    t.type = 1; t.dim = 1; t.c1 = cam1; t.c2 = cam2; t.w = weight; t.set = true; t.seq = s->nextSeq++;
    t.q.x = q[0]; t.q.y = q[1]; t.q.z = q[2]; t.q.w = q[3];                // Quaternionf::cast<double>(), not normalised
    s->dirty = true;
}
void baorc_set_relative_transform(void* h, int idx, int cam1, int cam2, const float* dpos, const float* q, float weight)
{
    BA* s = static_cast<BA*>(h);
    Tether& t = s->teth[2][idx];
    t.type = 2; t.dim = 6; t.c1 = cam1; t.c2 = cam2; t.w = weight; t.set = true; t.seq = s->nextSeq++;
    t.C.r.x = q[0]; t.C.r.y = q[1]; t.C.r.z = q[2]; t.C.r.w = q[3];
    qnormalizeRotation(t.C.r);                                             // SE3Quat(q, t) ctor
    for (int i = 0; i < 3; i++) t.C.t[i] = dpos[i];
    s->dirty = true;
}
void baorc_set_lambda(void* h, float l) { BA* s = static_cast<BA*>(h); s->iteration = 0; s->userLambdaInit = l; }
float baorc_get_lambda(void* h) { return (float)static_cast<BA*>(h)->lambda; }

// ref BundlerLib.cpp:364-447
float baorc_step(void* h, const float* huber, int n, float max_err_sq, unsigned* outliers, int cap, int* n_out)
{
    BA* s = static_cast<BA*>(h);
    for (int i = 0; i < n; i++) {
        s->huber = huber[i];
        if (!stepOnce(*s)) break;
    }
    int count = 0, m = 0;
    double error = 0;
    bool removedAny = false;
    for (int e : s->active) {
        Obs& o = s->obs[e];
        double sumSquares = o.err[0] * o.err[0] + o.err[1] * o.err[1];
        const Camera& c = s->cams[o.cam];
        // worldPose = T^-1: rotation conj(r), translation conj(r) * (-t)
        Quat rc = c.T.r; rc.x = -rc.x; rc.y = -rc.y; rc.z = -rc.z;
        Vec3 nt; for (int i = 0; i < 3; i++) nt[i] = c.T.t[i] * -1.;
        Vec3 wt = qrot(rc, nt);
        Vec3 fwdIn; fwdIn[0] = 0; fwdIn[1] = 0; fwdIn[2] = 1;
        Vec3 fwd = qrot(rc, fwdIn);
        const Vec3& X = s->pts[o.pt].X;
        double dot = (X[0] - wt[0]) * fwd[0] + (X[1] - wt[1]) * fwd[1] + (X[2] - wt[2]) * fwd[2];
        if (dot <= 0 || sumSquares > max_err_sq) {
            o.removed = true; removedAny = true;
            if (m < cap) outliers[m] = (unsigned)e;
            m++;
        } else { error += sumSquares; count++; }
    }
    if (removedAny) s->dirty = true;
    *n_out = m;
    return (float)(error / count);
}
void baorc_get_pose(void* h, int idx, float* pos, float* rot_colmajor)
{
    const Camera& c = static_cast<BA*>(h)->cams[idx];
    Quat q = c.T.r;
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
    double R[9]; qtoR(q, R);
    for (int i = 0; i < 3; i++) pos[i] = (float)c.T.t[i];
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) rot_colmajor[cc * 3 + r] = (float)R[r * 3 + cc];
}
void baorc_get_point(void* h, int idx, float* xyz)
{
    const Point& p = static_cast<BA*>(h)->pts[idx];
    for (int i = 0; i < 3; i++) xyz[i] = (float)p.X[i];
}
void baorc_get_state_f64(void* h, double* cams7, double* pts3)
{
    BA* s = static_cast<BA*>(h);
    for (size_t k = 0; k < s->cams.size(); k++) {
        const Pose& T = s->cams[k].T;
        double* o = cams7 + 7 * k;
        o[0] = T.r.x; o[1] = T.r.y; o[2] = T.r.z; o[3] = T.r.w; o[4] = T.t[0]; o[5] = T.t[1]; o[6] = T.t[2];
    }
    for (size_t i = 0; i < s->pts.size(); i++) for (int r = 0; r < 3; r++) pts3[3 * i + r] = s->pts[i].X[r];
}

} // extern "C"
