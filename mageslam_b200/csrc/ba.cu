// Bundle adjustment for sm_100a: Levenberg-Marquardt over SE(3) poses + 3-D points, Huber-weighted reprojection error,
// Schur complement, dense LDLT on the reduced camera system -- the arithmetic mage::BundlerLib delegates to g2o
// (ref Dependencies/BundlerLib/Source/BundlerLib.cpp + Dependencies/g2o), behind include/mage_b200.h.
//
// Design: FP64 end to end like the reference (no tensor cores: tcgen05 has no FP64 MMA and the
// reduced system of a local window is 48x48). The whole LM loop of a StepBundleAdjustment call -- every iteration and
// every lambda trial -- runs inside ONE persistent kernel launch, one CTA per problem, control scalars in shared memory;
// the host only builds index structure when the edge set changes and reads back a few scalars + outlier flags.
// Many independent problems are stepped by one launch (grid = problems). All reductions use a fixed order
// (no floating-point atomics), so a run is bit-reproducible.
#include "common.cuh"
#include "dense_ldlt.cuh"

#include <cooperative_groups.h>

#include <chrono>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <cfloat>
#include <cmath>
#include <limits>
#include <map>
#include <vector>

namespace mage {

struct BaCtl {
    double lambda, ni, user_lambda_init, err_sum;
    int iteration, inlier_count, stop_flag, last_ok;
    int n_flagged, pad_;
    long long lm_iters, lm_trials;
    // --- end of the head: the batched call gathers the blocks of its problems only up to here (kCtlHeadBytes)
    // final camera state of problems of up to kCtlCams cameras (q xyzw, t per camera): it comes back with the control block, so reading the
    // poses after a step costs no further copy (the tracking thread's pose-only BA reads one pose per call)
    int cams_valid, pad2_;
    double cams[7 * 16];
    long long phase_ns[32];     // cooperative kernel: time per phase seen by block 0 (diagnostics; 9.. = the dense solver's)
};
constexpr int kCtlCams = 16;
constexpr size_t kCtlHeadBytes = offsetof(BaCtl, cams_valid);
static_assert(kCtlHeadBytes % 8 == 0, "the head of BaCtl is copied as 8-byte words");

struct BaDev {
    int K, P, Ea, Kf, Pl, n, nblk, cam_parts;
    double *cam_q, *cam_t;                    // [K][4] (x y z w), [K][3]  world->camera
    const double *cam_f, *cam_cx, *cam_cy;    // [K]
    const int* cam_h;                         // [K] Hessian index of the camera or -1
    double* pt_X;                             // [P][3]
    const int *e_cam, *e_pt;                  // [Ea] active observations in insertion order
    const double *e_uv, *e_info;              // [Ea][2], [Ea]
    const int *l_pt, *l_ptr, *l_edges;        // landmarks (free points, Hessian order): point id, CSR of active edges
    const int *c_cam, *c_ptr, *c_edges;       // free cameras (Hessian order): camera id, CSR of active edges
    const int *blk_ij, *blk_ptr;              // upper blocks of the reduced system: (i1, i2), CSR into pairs
    const int2* pairs;                        // (edge of i1, edge of i2) sharing a landmark
    const int* e_l;                           // [Ea] landmark index of the edge's point or -1
    double *err, *W, *WD, *Hll, *bl, *Dinv, *db, *Hpp, *bp, *S, *bs, *x, *cam_bak, *pt_bak, *part;
    double* Hc;                               // [Ea][12] per-edge contribution to (H_ll 3x3, b_l 3)
    unsigned char* flags;                     // [Ea] 1 = outlier
    BaCtl* ctl;
    // cooperative (multi-CTA) variant: per-(block, part) partial sums of the Schur products and the grid reduction slots
    double *spart, *gred;
    int schur_parts;
    // large reduced systems (global BA): S stays in global memory and is factorised by the whole grid
    int big;
    int8_t* Zq;                               // int8 slice planes of the current panel (dense_ldlt.cuh)
    double* Ldiag;                            // factored diagonal blocks of the reduced system (dense_ldlt.cuh)
    int* dflag;                               // look-ahead counter of the dense solver (dense_ldlt.cuh)
    double* xchg;                             // [16 + 2 n] scalars / diag(Hpp) / bp exchanged between the ranks of a sharded problem (big mode only)
    int* Ez;                                  // row exponents of the slices
    // tether edges between two cameras (ref BundlerLib.cpp:24-90, :311-350), single-CTA kernel only
    int nT;
    const int4* t_def;                        // [nT] (type, cam1, cam2, error dimension)
    const double* t_meas;                     // [nT][8]: type 0 distance | type 1 q(4) | type 2 C.q(4), C.t(3); [7] = weight
    double *t_err, *t_J;                      // [nT][6], [nT][2][36] (dim x 6 row-major per vertex)
    double* ldlt_col;                         // 2 x (n + 1) doubles in shared memory for the register-blocked LDL^T (null: in-place version)
    // one-CTA-per-problem kernel: rotation matrices of the current camera state (row-major 3x3 per camera)
    double* cam_R;                            // [K][9]
    double* cam_Rold;                         // [K][9] rotation matrices of the linearisation point while a trial state is being evaluated
    // local-window fast path (fast != 0): landmarks in batches of <= kFastBE edges whose pose-landmark blocks live in shared
    // memory only; thread pair i owns the (block, part) item i and accumulates its share of the Schur products in registers
    int fast, nb, nitems;
    int pose1;                                // pose-only problem of ONE free camera (points fixed, no tethers): fused two-pass LM in k_ba_step
    const int* batch_ptr;                     // [nb + 1] landmark boundaries
    const int4* item_def;                     // [nitems] (block, part, parts of the block, 0)
    const int2* blk_items;                    // [nblk] (first item, parts)
    const int* cam_diag;                      // [Kf] index of the diagonal block of a free camera
    const int* bb_ptr;                        // [nb * nblk + 1] pair ranges per (batch, block)
    const ushort2* bpairs;                    // (edge of i1, edge of i2) relative to the batch's first edge
};

// ------------------------------------------------------------------------------------------------ small math
__device__ __forceinline__ void q_rot(const double* q, const double* p, double* r)    // Eigen _transformVector
{
    double ux = q[1] * p[2] - q[2] * p[1], uy = q[2] * p[0] - q[0] * p[2], uz = q[0] * p[1] - q[1] * p[0];
    ux += ux; uy += uy; uz += uz;
    r[0] = p[0] + q[3] * ux + (q[1] * uz - q[2] * uy);
    r[1] = p[1] + q[3] * uy + (q[2] * ux - q[0] * uz);
    r[2] = p[2] + q[3] * uz + (q[0] * uy - q[1] * ux);
}
__device__ __forceinline__ void q_to_R(const double* q, double* R)
{
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ void R_to_q(const double* m, double* q)
{
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}
__device__ __forceinline__ void q_normalize_rotation(double* q)      // ref se3quat.h:280-285
{
    if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
// ref se3quat.h:218-257 (exp) and :99-105 (operator*):  T <- exp(u) * T,  u = (omega, upsilon)
__device__ void pose_oplus(double* q, double* t, const double* u)
{
    const double om0 = u[0], om1 = u[1], om2 = u[2];
    const double theta = sqrt(om0 * om0 + om1 * om1 + om2 * om2);
    const double O[9] = {0, -om2, om1, om2, 0, -om0, -om1, om0, 0};
    double O2[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) O2[i * 3 + j] = O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j] + O[i * 3 + 2] * O[6 + j];
    double a, b, c, d;
    if (theta < 0.00001) { a = 1; b = 0.5; c = 0.5; d = 1. / 6.; }
    else { a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta); c = b; d = (theta - sin(theta)) / (theta * theta * theta); }
    double R[9], V[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { double I = (i % 4 == 0) ? 1.0 : 0.0; R[i] = I + a * O[i] + b * O2[i]; V[i] = I + c * O[i] + d * O2[i]; }
    double eq[4], et[3];
    R_to_q(R, eq);
#pragma unroll
    for (int i = 0; i < 3; i++) et[i] = V[i * 3] * u[3] + V[i * 3 + 1] * u[4] + V[i * 3 + 2] * u[5];
    q_normalize_rotation(eq);
    double rt[3];
    q_rot(eq, t, rt);
    t[0] = et[0] + rt[0]; t[1] = et[1] + rt[1]; t[2] = et[2] + rt[2];
    double r[4];
    r[3] = eq[3] * q[3] - eq[0] * q[0] - eq[1] * q[1] - eq[2] * q[2];
    r[0] = eq[3] * q[0] + eq[0] * q[3] + eq[1] * q[2] - eq[2] * q[1];
    r[1] = eq[3] * q[1] + eq[1] * q[3] + eq[2] * q[0] - eq[0] * q[2];
    r[2] = eq[3] * q[2] + eq[2] * q[3] + eq[0] * q[1] - eq[1] * q[0];
    q_normalize_rotation(r);
    q[0] = r[0]; q[1] = r[1]; q[2] = r[2]; q[3] = r[3];
}
// ref robust_kernel_impl.cpp:65-78: Huber on the squared error
__device__ __forceinline__ void huber(double e, double delta, double& rho0, double& rho1)
{
    double dsqr = delta * delta;
    if (e <= dsqr) { rho0 = e; rho1 = 1.0; }
    else { double sq = sqrt(e); rho0 = 2 * sq * delta - dsqr; rho1 = delta / sq; }
}

// fixed-order block reductions (bit-reproducible): strided partials -> shuffle tree -> warp 0 tree
__device__ double block_sum(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[32] = v;
    }
    __syncthreads();
    return sh[32];
}
__device__ double block_max(double v, double* sh)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? sh[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
        if (lane == 0) sh[32] = v;
    }
    __syncthreads();
    return sh[32];
}

// ------------------------------------------------------------------------------------------------ phases
// ref types_six_dof_expmap.h:140-147 computeError
__device__ void phase_errors(const BaDev& p, int tid, int nt)
{
    for (int e = tid; e < p.Ea; e += nt) {
        const int c = p.e_cam[e];
        double xt[3];
        q_rot(p.cam_q + 4 * c, p.pt_X + 3 * (size_t)p.e_pt[e], xt);
        xt[0] += p.cam_t[3 * c]; xt[1] += p.cam_t[3 * c + 1]; xt[2] += p.cam_t[3 * c + 2];
        const double f = p.cam_f[c];
        p.err[2 * e] = p.e_uv[2 * e] - (xt[0] / xt[2] * f + p.cam_cx[c]);
        p.err[2 * e + 1] = p.e_uv[2 * e + 1] - (xt[1] / xt[2] * f + p.cam_cy[c]);
    }
}
__device__ double tether_chi2_sum(const BaDev& p);
// ref sparse_optimizer.cpp:102-117 activeRobustChi2
__device__ double phase_chi2(const BaDev& p, double delta, double* sh)
{
    double acc = 0;
    for (int e = threadIdx.x; e < p.Ea; e += blockDim.x) {
        double e0 = p.err[2 * e], e1 = p.err[2 * e + 1], r0, r1;
        huber(p.e_info[e] * (e0 * e0 + e1 * e1), delta, r0, r1);
        acc += r0;
    }
    if (p.nT) acc += tether_chi2_sum(p);
    return block_sum(acc, sh);
}

// Jacobians of ref types_six_dof_expmap.cpp:295-331 at the current state
__device__ __forceinline__ void edge_jacobians(const BaDev& p, int e, double* Ji /*2x3*/, double* Jj /*2x6*/, bool wantPose)
{
    const int c = p.e_cam[e];
    const double* q = p.cam_q + 4 * c;
    double xt[3];
    q_rot(q, p.pt_X + 3 * (size_t)p.e_pt[e], xt);
    const double x = xt[0] + p.cam_t[3 * c], y = xt[1] + p.cam_t[3 * c + 1], z = xt[2] + p.cam_t[3 * c + 2];
    const double f = p.cam_f[c];
    double R[9];
    q_to_R(q, R);
    // one reciprocal instead of the dozen divisions of the reference's expressions (same values to rounding)
    const double iz = 1.0 / z, a = x * iz, b = y * iz, g = -(iz * f);
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
        Ji[cc] = g * (R[cc] - a * R[6 + cc]);
        Ji[3 + cc] = g * (R[3 + cc] - b * R[6 + cc]);
    }
    if (wantPose) {
        const double ab = a * b * f, izf = iz * f;
        Jj[0] = ab; Jj[1] = -(f + a * a * f); Jj[2] = b * f; Jj[3] = -izf; Jj[4] = 0; Jj[5] = a * izf;
        Jj[6] = f + b * b * f; Jj[7] = -ab; Jj[8] = -(a * f); Jj[9] = 0; Jj[10] = -izf; Jj[11] = b * izf;
    }
}

// ref block_solver.hpp:463-521 + base_binary_edge.hpp:62-134: landmark blocks H_ll, b_l and pose-landmark blocks W.
// Edge-parallel: every active edge with a free point computes its Jacobians once, stores W (6x3) and its contribution to
// (H_ll, b_l) into a per-edge slot; phase_sum_points then adds a landmark's (contiguous) slots in edge order.
__device__ void phase_build_points(const BaDev& p, double delta, int tid, int nt)
{
    for (int e = tid; e < p.Ea; e += nt) {
        if (p.e_l[e] < 0) continue;
        const int hj = p.cam_h[p.e_cam[e]];
        double Ji[6], Jj[12];
        edge_jacobians(p, e, Ji, Jj, hj >= 0);
        const double e0 = p.err[2 * e], e1 = p.err[2 * e + 1], info = p.e_info[e];
        double r0, r1;
        huber(info * (e0 * e0 + e1 * e1), delta, r0, r1);
        const double w = r1 * info, o0 = -info * e0 * r1, o1 = -info * e1 * r1;
        double* __restrict__ hc = p.Hc + 12 * (size_t)e;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            hc[9 + r] = Ji[r] * o0 + Ji[3 + r] * o1;
#pragma unroll
            for (int c = 0; c < 3; c++) hc[r * 3 + c] = w * (Ji[r] * Ji[c] + Ji[3 + r] * Ji[3 + c]);
        }
        if (hj >= 0) {
            double* __restrict__ W = p.W + 18 * (size_t)e;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) W[r * 3 + c] = w * (Jj[r] * Ji[c] + Jj[6 + r] * Ji[3 + c]);
        }
    }
}
__device__ void phase_sum_points(const BaDev& p, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        double acc[12];
#pragma unroll
        for (int i = 0; i < 12; i++) acc[i] = 0;
        const int k0 = p.l_ptr[li], k1 = p.l_ptr[li + 1];
        for (int k = k0; k < k1; k++) {
            const double* __restrict__ hc = p.Hc + 12 * (size_t)k;
#pragma unroll
            for (int i = 0; i < 12; i++) acc[i] += hc[i];
        }
#pragma unroll
        for (int i = 0; i < 9; i++) p.Hll[9 * (size_t)li + i] = acc[i];
        p.bl[3 * li] = acc[9]; p.bl[3 * li + 1] = acc[10]; p.bl[3 * li + 2] = acc[11];
    }
}

// landmark-parallel variant (one thread walks a landmark's edges): fewer instructions, used by the one-CTA-per-problem kernel
__device__ void phase_build_points_lm(const BaDev& p, double delta, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        for (int k = p.l_ptr[li]; k < p.l_ptr[li + 1]; k++) {
            const int e = k;                              // edges are stored grouped by landmark
            const int hj = p.cam_h[p.e_cam[e]];
            double Ji[6], Jj[12];
            edge_jacobians(p, e, Ji, Jj, hj >= 0);
            const double e0 = p.err[2 * e], e1 = p.err[2 * e + 1], info = p.e_info[e];
            double r0, r1;
            huber(info * (e0 * e0 + e1 * e1), delta, r0, r1);
            const double w = r1 * info, o0 = -info * e0 * r1, o1 = -info * e1 * r1;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                b[r] += Ji[r] * o0 + Ji[3 + r] * o1;
#pragma unroll
                for (int c = 0; c < 3; c++) H[r * 3 + c] += w * (Ji[r] * Ji[c] + Ji[3 + r] * Ji[3 + c]);
            }
            if (hj >= 0) {
                double* W = p.W + 18 * (size_t)e;
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) W[r * 3 + c] = w * (Jj[r] * Ji[c] + Jj[6 + r] * Ji[3 + c]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; i++) p.Hll[9 * (size_t)li + i] = H[i];
        p.bl[3 * li] = b[0]; p.bl[3 * li + 1] = b[1]; p.bl[3 * li + 2] = b[2];
    }
}

// pose blocks H_pp (diagonal 6x6) and b_p: one warp per (camera, part) slice of the camera's edge list, lanes stride the
// slice, shuffle-tree reduction, partials summed in part order by phase_finish_cams
__device__ void phase_build_cams(const BaDev& p, double delta, int warp, int nwarps, int lane)
{
    const int items = p.Kf * p.cam_parts;
    for (int it = warp; it < items; it += nwarps) {
        const int kf = it / p.cam_parts, part = it % p.cam_parts;
        const int beg = p.c_ptr[kf], end = p.c_ptr[kf + 1], len = end - beg;
        const int per = (len + p.cam_parts - 1) / p.cam_parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        double A[21], b[6];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = 0;
        for (int k = s0 + lane; k < s1; k += 32) {
            const int e = p.c_edges[k];
            double Ji[6], Jj[12];
            edge_jacobians(p, e, Ji, Jj, true);
            const double e0 = p.err[2 * e], e1 = p.err[2 * e + 1], info = p.e_info[e];
            double r0, r1;
            huber(info * (e0 * e0 + e1 * e1), delta, r0, r1);
            const double w = r1 * info, o0 = -info * e0 * r1, o1 = -info * e1 * r1;
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                b[r] += Jj[r] * o0 + Jj[6 + r] * o1;
#pragma unroll
                for (int c = r; c < 6; c++) A[idx++] += w * (Jj[r] * Jj[c] + Jj[6 + r] * Jj[6 + c]);
            }
        }
#pragma unroll
        for (int i = 0; i < 21; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) A[i] += __shfl_down_sync(0xffffffffu, A[i], o);
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) b[i] += __shfl_down_sync(0xffffffffu, b[i], o);
        if (lane == 0) {
            double* dst = p.part + (size_t)it * 27;
#pragma unroll
            for (int i = 0; i < 21; i++) dst[i] = A[i];
#pragma unroll
            for (int i = 0; i < 6; i++) dst[21 + i] = b[i];
        }
    }
}
__device__ void phase_finish_cams(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < p.Kf * 27; i += nt) {
        const int kf = i / 27, j = i % 27;
        double s = 0;
        for (int part = 0; part < p.cam_parts; part++) s += p.part[((size_t)kf * p.cam_parts + part) * 27 + j];
        if (j >= 21) p.bp[6 * kf + (j - 21)] = s;
        else {
            int r = 0, rem = j;
            while (rem >= 6 - r) { rem -= 6 - r; r++; }
            const int c = r + rem;
            p.Hpp[36 * (size_t)kf + r * 6 + c] = s;
            p.Hpp[36 * (size_t)kf + c * 6 + r] = s;
        }
    }
}

// ref optimization_algorithm_levenberg.cpp:151-165 computeLambdaInit: tau * max |diag(H)|
__device__ double phase_max_diag(const BaDev& p, double* sh)
{
    double m = 0;
    for (int i = threadIdx.x; i < p.Kf * 6; i += blockDim.x) m = fmax(m, fabs(p.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]));
    for (int i = threadIdx.x; i < p.Pl * 3; i += blockDim.x) m = fmax(m, fabs(p.Hll[9 * (size_t)(i / 3) + (i % 3) * 4]));
    return block_max(m, sh);
}

// Schur step 1 (ref block_solver.hpp:337-352): D^-1 = (H_ll + lambda I)^-1 (3x3 cofactor inverse), db = D^-1 b_l, WD = W D^-1.
// Edge-parallel: each edge inverts its landmark's block itself (a few dozen flops, cheaper than a dependent round trip) and the
// first edge of a landmark publishes D^-1 and db.
__device__ void phase_schur_points(const BaDev& p, double lambda, int tid, int nt)
{
    for (int e = tid; e < p.Ea; e += nt) {
        const int li = p.e_l[e];
        if (li < 0) continue;
        const bool hasCam = p.cam_h[p.e_cam[e]] >= 0;
        const bool first = e == p.l_ptr[li];
        if (!hasCam && !first) continue;
        double A[9], w[18];
        const double* __restrict__ Hl = p.Hll + 9 * (size_t)li;
        const double* __restrict__ W = p.W + 18 * (size_t)e;
#pragma unroll
        for (int i = 0; i < 9; i++) A[i] = Hl[i];
        if (hasCam) {
#pragma unroll
            for (int i = 0; i < 18; i++) w[i] = W[i];
        }
        A[0] += lambda; A[4] += lambda; A[8] += lambda;
        const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
        const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
        double D[9];
        D[0] = c00 * id; D[1] = (A[2] * A[7] - A[1] * A[8]) * id; D[2] = (A[1] * A[5] - A[2] * A[4]) * id;
        D[3] = c01 * id; D[4] = (A[0] * A[8] - A[2] * A[6]) * id; D[5] = (A[2] * A[3] - A[0] * A[5]) * id;
        D[6] = c02 * id; D[7] = (A[1] * A[6] - A[0] * A[7]) * id; D[8] = (A[0] * A[4] - A[1] * A[3]) * id;
        if (first) {
#pragma unroll
            for (int i = 0; i < 9; i++) p.Dinv[9 * (size_t)li + i] = D[i];
            const double b0 = p.bl[3 * li], b1 = p.bl[3 * li + 1], b2 = p.bl[3 * li + 2];
            p.db[3 * li] = D[0] * b0 + D[1] * b1 + D[2] * b2;
            p.db[3 * li + 1] = D[3] * b0 + D[4] * b1 + D[5] * b2;
            p.db[3 * li + 2] = D[6] * b0 + D[7] * b1 + D[8] * b2;
        }
        if (hasCam) {
            double* __restrict__ WD = p.WD + 18 * (size_t)e;
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) WD[r * 3 + c] = w[r * 3] * D[c] + w[r * 3 + 1] * D[3 + c] + w[r * 3 + 2] * D[6 + c];
        }
    }
    for (int i = tid; i < p.n * p.n; i += nt) p.S[i] = 0.0;
}

// Schur step 2 (ref block_solver.hpp:354-390): S_ij = Hpp_ij (+lambda) - sum_pairs WD_e1 W_e2^T per upper block (one warp per
// block, lanes stride the pair list), and the partials of coeff_i = sum_e W_e db per (camera, part)
__device__ void phase_schur_blocks(const BaDev& p, double lambda, int warp, int nwarps, int lane)
{
    for (int bi = warp; bi < p.nblk; bi += nwarps) {
        const int i1 = p.blk_ij[2 * bi], i2 = p.blk_ij[2 * bi + 1];
        double acc[36];
#pragma unroll
        for (int i = 0; i < 36; i++) acc[i] = 0;
        for (int k = p.blk_ptr[bi] + lane; k < p.blk_ptr[bi + 1]; k += 32) {
            const int2 pr = p.pairs[k];
            const double* A = p.WD + 18 * (size_t)pr.x;
            const double* B = p.W + 18 * (size_t)pr.y;
            double a[18], b[18];
#pragma unroll
            for (int i = 0; i < 18; i++) { a[i] = A[i]; b[i] = B[i]; }
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 6; c++) acc[r * 6 + c] += a[r * 3] * b[c * 3] + a[r * 3 + 1] * b[c * 3 + 1] + a[r * 3 + 2] * b[c * 3 + 2];
        }
#pragma unroll
        for (int i = 0; i < 36; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double v = -acc[r * 6 + c];
                    if (i1 == i2) v += p.Hpp[36 * (size_t)i1 + r * 6 + c] + ((r == c) ? lambda : 0.0);
                    p.S[(size_t)(6 * i1 + r) * p.n + 6 * i2 + c] = v;
                    if (i1 != i2) p.S[(size_t)(6 * i2 + c) * p.n + 6 * i1 + r] = v;
                }
        }
    }
    const int items = p.Kf * p.cam_parts;
    for (int it = warp; it < items; it += nwarps) {
        const int kf = it / p.cam_parts, part = it % p.cam_parts;
        const int beg = p.c_ptr[kf], end = p.c_ptr[kf + 1], len = end - beg;
        const int per = (len + p.cam_parts - 1) / p.cam_parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        double c6[6] = {0, 0, 0, 0, 0, 0};
        for (int k = s0 + lane; k < s1; k += 32) {
            const int e = p.c_edges[k];
            const int li = p.e_l[e];
            if (li < 0) continue;
            const double* W = p.W + 18 * (size_t)e;
            const double d0 = p.db[3 * li], d1 = p.db[3 * li + 1], d2 = p.db[3 * li + 2];
#pragma unroll
            for (int r = 0; r < 6; r++) c6[r] += W[r * 3] * d0 + W[r * 3 + 1] * d1 + W[r * 3 + 2] * d2;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c6[i] += __shfl_down_sync(0xffffffffu, c6[i], o);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 6; i++) p.part[(size_t)it * 27 + i] = c6[i];
        }
    }
}
__device__ void phase_finish_bs(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < p.n; i += nt) {
        const int kf = i / 6, j = i % 6;
        double s = 0;
        for (int part = 0; part < p.cam_parts; part++) s += p.part[((size_t)kf * p.cam_parts + part) * 27 + j];
        p.bs[i] = p.bp[i] - s;
    }
}

// Dense LDL^T of the reduced system by the whole CTA, in place in S (lower part), then the two triangular solves.
// ref linear_solver_dense.h:65-113 + Eigen LDLT: fails (returns false) when a pivot is negative (isPositive() false);
// zero pivots are tolerated like Eigen (no scaling, pseudo-inverse in the solve). x is written only on success.
// Register-blocked LDL^T for small reduced systems (n <= 16 B): the 256 threads form a 16 x 16 grid, thread (bi, bj) keeps the B x B
// block (rows B bi.., columns B bj..) of the lower triangle in registers for the whole factorisation. Per column k the owners of
// column k publish its unscaled entries (and the pivot) to a double-buffered shared vector, ONE barrier, then every thread updates its
// block -- no shared-memory read-modify-write of the matrix, which is what bounded the in-place version (0.9 us per column).
// Same arithmetic per element as the in-place loop below: S_ij -= (S_ik / d_k ... as S_ik * (1/d_k)) * S_jk, column scaled afterwards.
template <int B>
__device__ void ldlt_factor_regs(double* __restrict__ S, int n, int* s_neg, double* __restrict__ colbuf /* 2 x (n + 1) */)
{
    const int tid = threadIdx.x, bi = tid >> 4, bj = tid & 15;
    const bool lower = bi >= bj;
    double a[B][B];
#pragma unroll
    for (int r = 0; r < B; r++)
#pragma unroll
        for (int c = 0; c < B; c++) {
            const int i = B * bi + r, j = B * bj + c;
            a[r][c] = (lower && i < n && j < n) ? S[(size_t)i * n + j] : 0.0;
        }
    for (int k = 0; k < n; k++) {
        const int kb = k / B, kc = k - kb * B;
        double* col = colbuf + (k & 1) * (n + 1);
        if (bj == kb && lower) {                       // owners of column k: publish the unscaled entries of their rows (and the pivot)
#pragma unroll
            for (int r = 0; r < B; r++) {
                double v = 0.0;
#pragma unroll
                for (int c = 0; c < B; c++) if (c == kc) v = a[r][c];
                const int i = B * bi + r;
                if (i < n && i >= k) col[i] = v;
            }
        }
        __syncthreads();
        const double d = col[k];
        if (d < 0 && tid == 0) *s_neg = 1;
        const bool valid = fabs(d) > 0;
        const double inv_d = valid ? 1.0 / d : 0.0;
        if (valid && lower && bi >= kb) {
            double ai[B], aj[B];
#pragma unroll
            for (int r = 0; r < B; r++) { const int i = B * bi + r; ai[r] = (i > k && i < n) ? col[i] * inv_d : 0.0; }
#pragma unroll
            for (int c = 0; c < B; c++) { const int j = B * bj + c; aj[c] = (j > k && j < n) ? col[j] : 0.0; }
#pragma unroll
            for (int r = 0; r < B; r++)
#pragma unroll
                for (int c = 0; c < B; c++) a[r][c] -= ai[r] * aj[c];           // rows / columns <= k contribute exact zeros
            if (bj == kb) {                            // scale column k: L_ik = S_ik / d_k
#pragma unroll
                for (int r = 0; r < B; r++)
#pragma unroll
                    for (int c = 0; c < B; c++) if (c == kc && B * bi + r > k) a[r][c] *= inv_d;
            }
        }
    }
    __syncthreads();
    if (lower) {
#pragma unroll
        for (int r = 0; r < B; r++)
#pragma unroll
            for (int c = 0; c < B; c++) {
                const int i = B * bi + r, j = B * bj + c;
                if (i < n && j <= i) S[(size_t)i * n + j] = a[r][c];
            }
    }
}

__device__ bool phase_ldlt_solve(const BaDev& p, double* sh)
{
    const int n = p.n, tid = threadIdx.x, nt = blockDim.x;
    if (n == 0) return true;
    double* S = p.S;
    __shared__ int s_neg;
    if (tid == 0) s_neg = 0;
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4, nty = nt >> 4;
    if (nt == 256 && n <= 96 && p.ldlt_col) {
        if (n <= 48) ldlt_factor_regs<3>(S, n, &s_neg, p.ldlt_col);
        else if (n <= 64) ldlt_factor_regs<4>(S, n, &s_neg, p.ldlt_col);
        else ldlt_factor_regs<6>(S, n, &s_neg, p.ldlt_col);
    } else
    for (int k = 0; k < n; k++) {
        const double d = S[(size_t)k * n + k];
        if (d < 0) { if (tid == 0) s_neg = 1; }
        const bool valid = fabs(d) > 0;
        const double inv_d = valid ? 1.0 / d : 0.0;
        // trailing update with the unscaled column: S_ij -= a_ik * a_jk / d  (lower triangle, k < j <= i). No barrier is needed
        // here: the scaling of column k-1 that other threads may still be doing touches neither column k nor the trailing block.
        if (valid) {
            for (int i = k + 1 + ty; i < n; i += nty) {
                const double aik = S[(size_t)i * n + k] * inv_d;
                for (int j = k + 1 + tx; j <= i; j += 16) S[(size_t)i * n + j] -= aik * S[(size_t)j * n + k];
            }
        }
        __syncthreads();
        if (valid) for (int i = k + 1 + tid; i < n; i += nt) S[(size_t)i * n + k] *= inv_d;
    }
    __syncthreads();
    if (s_neg) return false;
    // solve L y = b (forward), D, L^T x = y (backward)
    double* y = p.bs;
    const double tol = 1.0 / DBL_MAX;
    if (n <= 64) {
        // one warp, the right-hand side in registers (lane i owns rows i and i + 32), pivots broadcast by shuffle: no CTA barriers.
        // Every y_i still receives its updates in the order k = 0, 1, ... of the loops below.
        if (tid < 32) {
            const int i0 = tid, i1 = tid + 32;
            double y0 = i0 < n ? y[i0] : 0.0, y1 = i1 < n ? y[i1] : 0.0;
            for (int k = 0; k < n; k++) {
                const double yk = __shfl_sync(0xffffffffu, k < 32 ? y0 : y1, k & 31);
                if (i0 > k && i0 < n) y0 -= S[(size_t)i0 * n + k] * yk;
                if (i1 > k && i1 < n) y1 -= S[(size_t)i1 * n + k] * yk;
            }
            if (i0 < n) { const double d = S[(size_t)i0 * n + i0]; y0 = (fabs(d) > tol) ? y0 / d : 0.0; }
            if (i1 < n) { const double d = S[(size_t)i1 * n + i1]; y1 = (fabs(d) > tol) ? y1 / d : 0.0; }
            for (int k = n - 1; k >= 0; k--) {
                const double xk = __shfl_sync(0xffffffffu, k < 32 ? y0 : y1, k & 31);
                if (i0 < k) y0 -= S[(size_t)k * n + i0] * xk;
                if (i1 < k) y1 -= S[(size_t)k * n + i1] * xk;
            }
            if (i0 < n) p.x[i0] = y0;
            if (i1 < n) p.x[i1] = y1;
        }
        __syncthreads();
        (void)sh;
        return true;
    }
    for (int k = 0; k < n; k++) {
        __syncthreads();
        const double yk = y[k];
        for (int i = k + 1 + tid; i < n; i += nt) y[i] -= S[(size_t)i * n + k] * yk;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) { const double d = S[(size_t)i * n + i]; y[i] = (fabs(d) > tol) ? y[i] / d : 0.0; }
    for (int k = n - 1; k >= 0; k--) {
        __syncthreads();
        const double xk = y[k];
        for (int i = tid; i < k; i += nt) y[i] -= S[(size_t)k * n + i] * xk;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) p.x[i] = y[i];
    (void)sh;
    return true;
}

// ref block_solver.hpp:418-444: x_l = D^-1 (b_l - W^T x_p)
__device__ void phase_backsub(const BaDev& p, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        double c0 = p.bl[3 * li], c1 = p.bl[3 * li + 1], c2 = p.bl[3 * li + 2];
        for (int k = p.l_ptr[li]; k < p.l_ptr[li + 1]; k++) {
            const int e = k;
            const int hj = p.cam_h[p.e_cam[e]];
            if (hj < 0) continue;
            const double* W = p.W + 18 * (size_t)e;
            const double* xp = p.x + 6 * hj;
#pragma unroll
            for (int r = 0; r < 6; r++) { c0 -= W[r * 3] * xp[r]; c1 -= W[r * 3 + 1] * xp[r]; c2 -= W[r * 3 + 2] * xp[r]; }
        }
        const double* D = p.Dinv + 9 * (size_t)li;
        p.x[p.n + 3 * li] = D[0] * c0 + D[1] * c1 + D[2] * c2;
        p.x[p.n + 3 * li + 1] = D[3] * c0 + D[4] * c1 + D[5] * c2;
        p.x[p.n + 3 * li + 2] = D[6] * c0 + D[7] * c1 + D[8] * c2;
    }
}
// push (ref base_vertex.h:92-94) + update (ref sparse_optimizer.cpp:433-446, oplusImpl)
__device__ void phase_backup(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < p.Kf; i += nt) {
        const int c = p.c_cam[i];
#pragma unroll
        for (int j = 0; j < 4; j++) p.cam_bak[7 * i + j] = p.cam_q[4 * c + j];
#pragma unroll
        for (int j = 0; j < 3; j++) p.cam_bak[7 * i + 4 + j] = p.cam_t[3 * c + j];
    }
    for (int i = tid; i < p.Pl * 3; i += nt) p.pt_bak[i] = p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3];
}
__device__ void phase_restore(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < p.Kf; i += nt) {
        const int c = p.c_cam[i];
#pragma unroll
        for (int j = 0; j < 4; j++) p.cam_q[4 * c + j] = p.cam_bak[7 * i + j];
#pragma unroll
        for (int j = 0; j < 3; j++) p.cam_t[3 * c + j] = p.cam_bak[7 * i + 4 + j];
    }
    for (int i = tid; i < p.Pl * 3; i += nt) p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3] = p.pt_bak[i];
}
__device__ void phase_update(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < p.Kf; i += nt) { const int c = p.c_cam[i]; pose_oplus(p.cam_q + 4 * c, p.cam_t + 3 * c, p.x + 6 * i); }
    for (int i = tid; i < p.Pl * 3; i += nt) p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3] += p.x[p.n + i];
}
// ref optimization_algorithm_levenberg.cpp:167-174 computeScale
__device__ double phase_scale(const BaDev& p, double lambda, double* sh)
{
    double acc = 0;
    const int tot = p.n + 3 * p.Pl;
    for (int j = threadIdx.x; j < tot; j += blockDim.x) {
        const double xj = p.x[j], bj = (j < p.n) ? p.bp[j] : p.bl[j - p.n];
        acc += xj * (lambda * xj + bj);
    }
    return block_sum(acc, sh);
}
// ref BundlerLib.cpp:385-427: cheirality + squared-error test per active observation, inlier error accumulation
__device__ void phase_classify(const BaDev& p, double maxErrSq, double* sh, double& errSum, int& inliers, int& flagged)
{
    double acc = 0, cnt = 0, nfl = 0;
    for (int e = threadIdx.x; e < p.Ea; e += blockDim.x) {
        const double e0 = p.err[2 * e], e1 = p.err[2 * e + 1];
        const double ss = e0 * e0 + e1 * e1;
        const int c = p.e_cam[e];
        const double qc[4] = {-p.cam_q[4 * c], -p.cam_q[4 * c + 1], -p.cam_q[4 * c + 2], p.cam_q[4 * c + 3]};
        const double nt3[3] = {-p.cam_t[3 * c], -p.cam_t[3 * c + 1], -p.cam_t[3 * c + 2]};
        double wt[3], fw[3];
        const double z[3] = {0, 0, 1};
        q_rot(qc, nt3, wt);
        q_rot(qc, z, fw);
        const double* X = p.pt_X + 3 * (size_t)p.e_pt[e];
        const double dot = (X[0] - wt[0]) * fw[0] + (X[1] - wt[1]) * fw[1] + (X[2] - wt[2]) * fw[2];
        const bool out = (dot <= 0) || (ss > maxErrSq);
        p.flags[e] = out ? 1 : 0;
        if (!out) { acc += ss; cnt += 1.0; } else nfl += 1.0;
    }
    errSum = block_sum(acc, sh);
    inliers = (int)block_sum(cnt, sh);
    flagged = (int)block_sum(nfl, sh);
}

// ------------------------------------------------------------------------------------------------ tether edges
// EdgeScaleConstraint / EdgeRotationConstraint (ref BundlerLib.cpp:24-90; BaseMultiEdge, no analytic Jacobian: g2o differentiates
// them numerically, ref core/base_multi_edge.hpp:68-118) and g2o's EdgeSE3Expmap (ref types_six_dof_expmap.h:108-127,
// .cpp:278-293). Few edges per problem (one or two per keyframe pair), so the phases below are thread-per-item loops.
__device__ __forceinline__ void q_mul(const double* a, const double* b, double* r)
{
    r[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    r[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    r[1] = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    r[2] = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
}
// ref se3quat.h:120-125 inverse, :99-105 operator* (normalises the rotation)
__device__ void pose_inv(const double* q, const double* t, double* qi, double* ti)
{
    qi[0] = -q[0]; qi[1] = -q[1]; qi[2] = -q[2]; qi[3] = q[3];
    const double nt[3] = {t[0] * -1., t[1] * -1., t[2] * -1.};
    q_rot(qi, nt, ti);
}
__device__ void pose_mul(const double* qa, const double* ta, const double* qb, const double* tb, double* q, double* t)
{
    double rt[3];
    q_rot(qa, tb, rt);
    t[0] = ta[0] + rt[0]; t[1] = ta[1] + rt[1]; t[2] = ta[2] + rt[2];
    q_mul(qa, qb, q);
    q_normalize_rotation(q);
}
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
// ref se3quat.h:171-210 SE3Quat::log -> (omega, upsilon)
__device__ void pose_log(const double* q, const double* t, double* res)
{
    double R[9];
    q_to_R(q, R);
    const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
    const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    double om[3], c2;
    if (fabs(d) > 0.99999) { om[0] = 0.5 * dR[0]; om[1] = 0.5 * dR[1]; om[2] = 0.5 * dR[2]; c2 = 1. / 12.; }
    else {
        const double theta = acos(d);
        const double k = theta / (2 * sqrt(1 - d * d));
        om[0] = k * dR[0]; om[1] = k * dR[1]; om[2] = k * dR[2];
        c2 = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
    }
    const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
    double O2[9];
    mat3_mul(O, O, O2);
#pragma unroll
    for (int i = 0; i < 3; i++) {
        res[i] = om[i];
        double acc = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double v = ((i == j) ? 1.0 : 0.0) - 0.5 * O[i * 3 + j] + c2 * O2[i * 3 + j];
            acc = (j == 0) ? v * t[0] : acc + v * t[j];
        }
        res[3 + i] = acc;
    }
}
// ref se3quat.h:217-226 adj(): [R 0; skew(t) R, R] (row-major 6x6), scaled by sgn
__device__ void pose_adj(const double* q, const double* t, double sgn, double* A)
{
    double R[9], SR[9];
    q_to_R(q, R);
    const double Sk[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    mat3_mul(Sk, R, SR);
#pragma unroll
    for (int i = 0; i < 36; i++) A[i] = 0.0 * sgn;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) { A[r * 6 + c] = sgn * R[r * 3 + c]; A[(3 + r) * 6 + 3 + c] = sgn * R[r * 3 + c]; A[(3 + r) * 6 + c] = sgn * SR[r * 3 + c]; }
}
// computeError of the three edge types at explicit poses of their two cameras
__device__ void tether_error(int type, const double* m, const double* q1, const double* t1, const double* q2, const double* t2, double* err)
{
    const double w = m[7];
    if (type == 0) {
        const double dx = t2[0] - t1[0], dy = t2[1] - t1[1], dz = t2[2] - t1[2];
        err[0] = (m[0] - sqrt(dx * dx + dy * dy + dz * dz)) * w;
    } else if (type == 1) {
        // (v1^-1 * v2).rotation().angularDistance(measurement): d = rel * conj(meas); 2 atan2(|d.vec|, |d.w|)
        const double c1[4] = {-q1[0], -q1[1], -q1[2], q1[3]};
        double rel[4], d[4];
        q_mul(c1, q2, rel);
        q_normalize_rotation(rel);
        const double mc[4] = {-m[0], -m[1], -m[2], m[3]};
        q_mul(rel, mc, d);
        err[0] = 2 * atan2(sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), fabs(d[3])) * w;
    } else {
        // log(v2^-1 * C * v1)
        double qi[4], ti[3], qa[4], ta[3], qe[4], te[3];
        pose_inv(q2, t2, qi, ti);
        pose_mul(qi, ti, m, m + 4, qa, ta);
        pose_mul(qa, ta, q1, t1, qe, te);
        pose_log(qe, te, err);
    }
}
__device__ __noinline__ void phase_tether_errors(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < p.nT; i += nt) {
        const int4 d = p.t_def[i];
        tether_error(d.x, p.t_meas + 8 * (size_t)i, p.cam_q + 4 * d.y, p.cam_t + 3 * d.y, p.cam_q + 4 * d.z, p.cam_t + 3 * d.z, p.t_err + 6 * (size_t)i);
    }
}
// chi2 = e^T Omega e (no robust kernel): Omega = I for the two scalar edges (the weight multiplies the error), w I for type 2
__device__ __noinline__ double tether_chi2_sum(const BaDev& p)      // this thread's share of the tether edges
{
    double acc = 0;
    for (int i = threadIdx.x; i < p.nT; i += blockDim.x) {
        const int4 d = p.t_def[i];
        const double om = d.x == 2 ? p.t_meas[8 * (size_t)i + 7] : 1.0;
        double ss = 0;
        for (int k = 0; k < d.w; k++) { const double e = p.t_err[6 * (size_t)i + k]; ss += e * (om * e); }
        acc += ss;
    }
    return acc;
}
// linearizeOplus: one thread per (edge, vertex, tangent direction). Scalar edges: central differences with delta = 1e-9 through
// oplus on the vertex (push / oplus / computeError / pop twice); type 2: the analytic adjoints (thread (v, d) = (0, 0) only).
__device__ __noinline__ void phase_tether_linearize(const BaDev& p, int tid, int nt)
{
    for (int it = tid; it < p.nT * 12; it += nt) {
        const int i = it / 12, v = (it % 12) / 6, dd = it % 6;
        const int4 d = p.t_def[i];
        const double* m = p.t_meas + 8 * (size_t)i;
        double* J = p.t_J + 72 * (size_t)i;
        const double *q1 = p.cam_q + 4 * d.y, *t1 = p.cam_t + 3 * d.y, *q2 = p.cam_q + 4 * d.z, *t2 = p.cam_t + 3 * d.z;
        if (d.x == 2) {
            if (v != 0 || dd != 0) continue;
            // Xi = (Tj^-1 * Tij).adj(),  Xj = -(Ti^-1 * Tij^-1).adj()
            double qi[4], ti[3], qa[4], ta[3], qc[4], tc[3];
            pose_inv(q2, t2, qi, ti);
            pose_mul(qi, ti, m, m + 4, qa, ta);
            pose_adj(qa, ta, 1.0, J);
            pose_inv(m, m + 4, qc, tc);
            pose_inv(q1, t1, qi, ti);
            pose_mul(qi, ti, qc, tc, qa, ta);
            pose_adj(qa, ta, -1.0, J + 36);
            continue;
        }
        const int cam = v == 0 ? d.y : d.z;
        if (p.cam_h[cam] < 0) continue;                         // fixed vertex: no Jacobian (ref base_multi_edge.hpp:81-83)
        const double delta = 1e-9, scalar = 1 / (2 * delta);
        double ep[1], em[1];
#pragma unroll
        for (int sgn = 0; sgn < 2; sgn++) {
            double u[6] = {0, 0, 0, 0, 0, 0};
            u[dd] = sgn == 0 ? delta : -delta;
            double q[4] = {p.cam_q[4 * cam], p.cam_q[4 * cam + 1], p.cam_q[4 * cam + 2], p.cam_q[4 * cam + 3]};
            double t[3] = {p.cam_t[3 * cam], p.cam_t[3 * cam + 1], p.cam_t[3 * cam + 2]};
            pose_oplus(q, t, u);
            tether_error(d.x, m, v == 0 ? q : q1, v == 0 ? t : t1, v == 0 ? q2 : q, v == 0 ? t2 : t, sgn == 0 ? ep : em);
        }
        J[36 * v + dd] = scalar * (ep[0] - em[0]);              // 1 x 6 row
    }
}
// constructQuadraticForm, diagonal part (ref base_multi_edge.hpp:155-176, base_binary_edge.hpp:75-103): one thread per free
// camera walks the edges in order: H_ii += A^T Omega A, b_i += A^T (-Omega e)
__device__ __noinline__ void phase_tether_accumulate(const BaDev& p, int tid, int nt)
{
    for (int kf = tid; kf < p.Kf; kf += nt) {
        const int cam = p.c_cam[kf];
        for (int i = 0; i < p.nT; i++) {
            const int4 d = p.t_def[i];
            if (d.y != cam && d.z != cam) continue;
            const int v = d.y == cam ? 0 : 1;
            const double om = d.x == 2 ? p.t_meas[8 * (size_t)i + 7] : 1.0;
            const double* A = p.t_J + 72 * (size_t)i + 36 * v;
            const double* e = p.t_err + 6 * (size_t)i;
            for (int r = 0; r < 6; r++) {
                double bb = 0;
                for (int k = 0; k < d.w; k++) bb += A[k * 6 + r] * (-(om * e[k]));
                p.bp[6 * kf + r] += bb;
                for (int c = 0; c < 6; c++) {
                    double hh = 0;
                    for (int k = 0; k < d.w; k++) hh += (A[k * 6 + r] * om) * A[k * 6 + c];
                    p.Hpp[36 * (size_t)kf + r * 6 + c] += hh;
                }
            }
        }
    }
}
// off-diagonal part: H_ij += A^T Omega B goes straight into the assembled reduced system (warp 0, edges in order)
__device__ __noinline__ void phase_tether_offdiag(const BaDev& p, int warp, int lane)
{
    if (warp != 0) return;
    for (int i = 0; i < p.nT; i++) {
        const int4 d = p.t_def[i];
        const int h1 = p.cam_h[d.y], h2 = p.cam_h[d.z];
        if (h1 >= 0 && h2 >= 0) {
            const double om = d.x == 2 ? p.t_meas[8 * (size_t)i + 7] : 1.0;
            const double* A = p.t_J + 72 * (size_t)i;
            const double* B = A + 36;
            for (int el = lane; el < 36; el += 32) {
                const int r = el / 6, c = el % 6;
                double hh = 0;
                for (int k = 0; k < d.w; k++) hh += (A[k * 6 + r] * om) * B[k * 6 + c];
                p.S[(size_t)(6 * h1 + r) * p.n + 6 * h2 + c] += hh;
                p.S[(size_t)(6 * h2 + c) * p.n + 6 * h1 + r] += hh;
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------ one-CTA fast path
// Phases of k_ba_step (one CTA per problem, many problems per launch). Same algorithm as above, arranged so that a thread's
// dependent chain is short: the projection of an edge is evaluated once per state (a = x/z, b = y/z, 1/z kept in a compact
// per-edge record, one reciprocal instead of g2o's dozen divisions), every Jacobian is rebuilt from that record and the camera's
// rotation matrix in shared memory, landmark phases run one thread per landmark with the landmark block in registers, and the
// Schur products give every lane half of a 6x6 block so that the accumulators stay in registers without spilling.
constexpr int kRec = 4;                     // doubles per edge record: a, b, 1/z, (pad) -- 32-byte aligned
constexpr int kWDs = 20;                    // doubles per WD record in this path: rows 0..2 | pad | rows 3..5 | pad (two 16-byte aligned halves)

__device__ void f_cam_R(const BaDev& p, int tid, int nt)
{
    for (int c = tid; c < p.K; c += nt) q_to_R(p.cam_q + 4 * c, p.cam_R + 9 * c);
}
// computeError of every active edge at the current state (ref types_six_dof_expmap.h:140-147) + the edge record
__device__ void f_project(const BaDev& p, int tid, int nt)
{
#pragma unroll 2
    for (int e = tid; e < p.Ea; e += nt) {
        const int c = p.e_cam[e];
        const double* __restrict__ R = p.cam_R + 9 * c;
        const double* __restrict__ X = p.pt_X + 3 * (size_t)p.e_pt[e];
        const double X0 = X[0], X1 = X[1], X2 = X[2];
        const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + p.cam_t[3 * c];
        const double y = R[3] * X0 + R[4] * X1 + R[5] * X2 + p.cam_t[3 * c + 1];
        const double z = R[6] * X0 + R[7] * X1 + R[8] * X2 + p.cam_t[3 * c + 2];
        const double iz = 1.0 / z, a = x * iz, b = y * iz, f = p.cam_f[c];
        const double2 uv = *reinterpret_cast<const double2*>(p.e_uv + 2 * (size_t)e);
        *reinterpret_cast<double2*>(p.err + 2 * (size_t)e) = make_double2(uv.x - (a * f + p.cam_cx[c]), uv.y - (b * f + p.cam_cy[c]));
        double2* rec = reinterpret_cast<double2*>(p.Hc + kRec * (size_t)e);
        rec[0] = make_double2(a, b);
        rec[1] = make_double2(iz, 0.0);
    }
}
// Jacobians of ref types_six_dof_expmap.cpp:295-331 from the edge record: Ji = -(f/z) [[1 0 -a], [0 1 -b]] R  (2x3),
// Jj = f [[ab, -(1+a^2), b, -1/z, 0, a/z], [1+b^2, -ab, -a, 0, -1/z, b/z]]  (2x6, rotation first)
__device__ __forceinline__ void f_point_jac(const double* __restrict__ R, double a, double b, double iz, double f, double* J0, double* J1)
{
    const double g = -(iz * f);
#pragma unroll
    for (int c = 0; c < 3; c++) { J0[c] = g * (R[c] - a * R[6 + c]); J1[c] = g * (R[3 + c] - b * R[6 + c]); }
}
__device__ __forceinline__ void f_pose_jac(double a, double b, double iz, double f, double* P0, double* P1)
{
    const double ab = a * b * f, izf = iz * f;
    P0[0] = ab; P0[1] = -(f + a * a * f); P0[2] = b * f; P0[3] = -izf; P0[4] = 0.0; P0[5] = a * izf;
    P1[0] = f + b * b * f; P1[1] = -ab; P1[2] = -(a * f); P1[3] = 0.0; P1[4] = -izf; P1[5] = b * izf;
}
// landmark blocks H_ll, b_l and pose-landmark blocks W (ref block_solver.hpp:463-521, base_binary_edge.hpp:62-134)
__device__ void f_build_points(const BaDev& p, double delta, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        double H00 = 0, H01 = 0, H02 = 0, H11 = 0, H12 = 0, H22 = 0, b0 = 0, b1 = 0, b2 = 0;
        const int k0 = p.l_ptr[li], k1 = p.l_ptr[li + 1];
#pragma unroll 2
        for (int e = k0; e < k1; e++) {                  // edges are stored grouped by landmark
            const int c = p.e_cam[e];
            const double2* rec = reinterpret_cast<const double2*>(p.Hc + kRec * (size_t)e);
            const double2 ab = rec[0];
            const double iz = rec[1].x, f = p.cam_f[c];
            const double2 er = *reinterpret_cast<const double2*>(p.err + 2 * (size_t)e);
            const double info = p.e_info[e];
            double r0, r1;
            huber(info * (er.x * er.x + er.y * er.y), delta, r0, r1);
            const double w = r1 * info, o0 = -info * er.x * r1, o1 = -info * er.y * r1;
            double J0[3], J1[3];
            f_point_jac(p.cam_R + 9 * c, ab.x, ab.y, iz, f, J0, J1);
            const double w0[3] = {w * J0[0], w * J0[1], w * J0[2]}, w1[3] = {w * J1[0], w * J1[1], w * J1[2]};
            H00 += w0[0] * J0[0] + w1[0] * J1[0]; H01 += w0[0] * J0[1] + w1[0] * J1[1]; H02 += w0[0] * J0[2] + w1[0] * J1[2];
            H11 += w0[1] * J0[1] + w1[1] * J1[1]; H12 += w0[1] * J0[2] + w1[1] * J1[2]; H22 += w0[2] * J0[2] + w1[2] * J1[2];
            b0 += J0[0] * o0 + J1[0] * o1; b1 += J0[1] * o0 + J1[1] * o1; b2 += J0[2] * o0 + J1[2] * o1;
            if (p.cam_h[c] >= 0) {
                double P0[6], P1[6];
                f_pose_jac(ab.x, ab.y, iz, f, P0, P1);
                double Wv[18];
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int cc = 0; cc < 3; cc++) Wv[r * 3 + cc] = P0[r] * w0[cc] + P1[r] * w1[cc];
                double2* W2 = reinterpret_cast<double2*>(p.W + 18 * (size_t)e);
#pragma unroll
                for (int i = 0; i < 9; i++) W2[i] = make_double2(Wv[2 * i], Wv[2 * i + 1]);
            }
        }
        double* Hl = p.Hll + 9 * (size_t)li;
        Hl[0] = H00; Hl[1] = H01; Hl[2] = H02; Hl[3] = H01; Hl[4] = H11; Hl[5] = H12; Hl[6] = H02; Hl[7] = H12; Hl[8] = H22;
        p.bl[3 * li] = b0; p.bl[3 * li + 1] = b1; p.bl[3 * li + 2] = b2;
    }
}
// pose blocks H_pp (upper 21) and b_p: one warp per (camera, part) slice, lanes stride the slice (same partials as phase_build_cams)
__device__ void f_build_cams(const BaDev& p, double delta, int warp, int nwarps, int lane)
{
    const int items = p.Kf * p.cam_parts;
    for (int it = warp; it < items; it += nwarps) {
        const int kf = it / p.cam_parts, part = it % p.cam_parts;
        const int beg = p.c_ptr[kf], end = p.c_ptr[kf + 1], len = end - beg;
        const int per = (len + p.cam_parts - 1) / p.cam_parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        const double f = p.cam_f[p.c_cam[kf]];
        double A[21], b[6];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = 0;
        for (int k = s0 + lane; k < s1; k += 32) {
            const int e = p.c_edges[k];
            const double2* rec = reinterpret_cast<const double2*>(p.Hc + kRec * (size_t)e);
            const double2 ab = rec[0];
            const double iz = rec[1].x;
            const double2 er = *reinterpret_cast<const double2*>(p.err + 2 * (size_t)e);
            const double info = p.e_info[e];
            double r0, r1;
            huber(info * (er.x * er.x + er.y * er.y), delta, r0, r1);
            const double w = r1 * info, o0 = -info * er.x * r1, o1 = -info * er.y * r1;
            double P0[6], P1[6];
            f_pose_jac(ab.x, ab.y, iz, f, P0, P1);
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                b[r] += P0[r] * o0 + P1[r] * o1;
                const double a0 = w * P0[r], a1 = w * P1[r];
#pragma unroll
                for (int c = r; c < 6; c++) A[idx++] += a0 * P0[c] + a1 * P1[c];
            }
        }
#pragma unroll
        for (int i = 0; i < 21; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) A[i] += __shfl_down_sync(0xffffffffu, A[i], o);
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) b[i] += __shfl_down_sync(0xffffffffu, b[i], o);
        if (lane == 0) {
            double* dst = p.part + (size_t)it * 27;
#pragma unroll
            for (int i = 0; i < 21; i++) dst[i] = A[i];
#pragma unroll
            for (int i = 0; i < 6; i++) dst[21 + i] = b[i];
        }
    }
}
// Schur step 1 (ref block_solver.hpp:337-352), one thread per landmark: D^-1 = (H_ll + lambda I)^-1 once, db = D^-1 b_l, then
// WD = W D^-1 for each of its edges with a free camera (stored as two 16-byte aligned halves of 3 rows)
__device__ void f_schur_points(const BaDev& p, double lambda, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        const double* __restrict__ Hl = p.Hll + 9 * (size_t)li;
        const double A0 = Hl[0] + lambda, A1 = Hl[1], A2 = Hl[2], A3 = Hl[3], A4 = Hl[4] + lambda, A5 = Hl[5], A6 = Hl[6], A7 = Hl[7], A8 = Hl[8] + lambda;
        const double c00 = A4 * A8 - A5 * A7, c01 = A5 * A6 - A3 * A8, c02 = A3 * A7 - A4 * A6;
        const double id = 1.0 / (A0 * c00 + A1 * c01 + A2 * c02);
        double D[9];
        D[0] = c00 * id; D[1] = (A2 * A7 - A1 * A8) * id; D[2] = (A1 * A5 - A2 * A4) * id;
        D[3] = c01 * id; D[4] = (A0 * A8 - A2 * A6) * id; D[5] = (A2 * A3 - A0 * A5) * id;
        D[6] = c02 * id; D[7] = (A1 * A6 - A0 * A7) * id; D[8] = (A0 * A4 - A1 * A3) * id;
#pragma unroll
        for (int i = 0; i < 9; i++) p.Dinv[9 * (size_t)li + i] = D[i];
        const double b0 = p.bl[3 * li], b1 = p.bl[3 * li + 1], b2 = p.bl[3 * li + 2];
        p.db[3 * li] = D[0] * b0 + D[1] * b1 + D[2] * b2;
        p.db[3 * li + 1] = D[3] * b0 + D[4] * b1 + D[5] * b2;
        p.db[3 * li + 2] = D[6] * b0 + D[7] * b1 + D[8] * b2;
        const int k0 = p.l_ptr[li], k1 = p.l_ptr[li + 1];
#pragma unroll 2
        for (int e = k0; e < k1; e++) {
            if (p.cam_h[p.e_cam[e]] < 0) continue;
            const double2* W2 = reinterpret_cast<const double2*>(p.W + 18 * (size_t)e);
            double w[18];
#pragma unroll
            for (int i = 0; i < 9; i++) { const double2 v = W2[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
            double o[kWDs];
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) o[(r / 3) * 10 + (r % 3) * 3 + c] = w[r * 3] * D[c] + w[r * 3 + 1] * D[3 + c] + w[r * 3 + 2] * D[6 + c];
            o[9] = 0.0; o[19] = 0.0;
            double2* O2 = reinterpret_cast<double2*>(p.WD + kWDs * (size_t)e);
#pragma unroll
            for (int i = 0; i < kWDs / 2; i++) O2[i] = make_double2(o[2 * i], o[2 * i + 1]);
        }
    }
    for (int i = tid; i < p.n * p.n; i += nt) p.S[i] = 0.0;
}
// Schur step 2 (ref block_solver.hpp:354-390): one warp per upper block, a lane pair per (edge, edge) product -- the even lane
// accumulates rows 0..2 of the 6x6 block, the odd lane rows 3..5 (18 accumulators each) -- then a parity-preserving shuffle tree
__device__ void f_schur_blocks(const BaDev& p, double lambda, int warp, int nwarps, int lane)
{
    const int half = lane & 1;
    for (int bi = warp; bi < p.nblk; bi += nwarps) {
        const int i1 = p.blk_ij[2 * bi], i2 = p.blk_ij[2 * bi + 1];
        double acc[18];
#pragma unroll
        for (int i = 0; i < 18; i++) acc[i] = 0;
        const int end = p.blk_ptr[bi + 1];
        for (int k = p.blk_ptr[bi] + (lane >> 1); k < end; k += 16) {
            const int2 pr = p.pairs[k];
            const double2* __restrict__ A2 = reinterpret_cast<const double2*>(p.WD + kWDs * (size_t)pr.x + 10 * half);
            const double2* __restrict__ B2 = reinterpret_cast<const double2*>(p.W + 18 * (size_t)pr.y);
            double a[10], b[18];
#pragma unroll
            for (int i = 0; i < 5; i++) { const double2 v = A2[i]; a[2 * i] = v.x; a[2 * i + 1] = v.y; }
#pragma unroll
            for (int i = 0; i < 9; i++) { const double2 v = B2[i]; b[2 * i] = v.x; b[2 * i + 1] = v.y; }
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 6; c++) acc[r * 6 + c] += a[r * 3] * b[c * 3] + a[r * 3 + 1] * b[c * 3 + 1] + a[r * 3 + 2] * b[c * 3 + 2];
        }
#pragma unroll
        for (int i = 0; i < 18; i++)
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
        if (lane < 2) {
#pragma unroll
            for (int rr = 0; rr < 3; rr++)
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const int r = 3 * half + rr;
                    double v = -acc[rr * 6 + c];
                    if (i1 == i2) v += p.Hpp[36 * (size_t)i1 + r * 6 + c] + ((r == c) ? lambda : 0.0);
                    p.S[(size_t)(6 * i1 + r) * p.n + 6 * i2 + c] = v;
                    if (i1 != i2) p.S[(size_t)(6 * i2 + c) * p.n + 6 * i1 + r] = v;
                }
        }
    }
    const int items = p.Kf * p.cam_parts;
    for (int it = warp; it < items; it += nwarps) {
        const int kf = it / p.cam_parts, part = it % p.cam_parts;
        const int beg = p.c_ptr[kf], end = p.c_ptr[kf + 1], len = end - beg;
        const int per = (len + p.cam_parts - 1) / p.cam_parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        double c6[6] = {0, 0, 0, 0, 0, 0};
        for (int k = s0 + lane; k < s1; k += 32) {
            const int e = p.c_edges[k];
            const int li = p.e_l[e];
            if (li < 0) continue;
            const double2* W2 = reinterpret_cast<const double2*>(p.W + 18 * (size_t)e);
            double w[18];
#pragma unroll
            for (int i = 0; i < 9; i++) { const double2 v = W2[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
            const double d0 = p.db[3 * li], d1 = p.db[3 * li + 1], d2 = p.db[3 * li + 2];
#pragma unroll
            for (int r = 0; r < 6; r++) c6[r] += w[r * 3] * d0 + w[r * 3 + 1] * d1 + w[r * 3 + 2] * d2;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c6[i] += __shfl_down_sync(0xffffffffu, c6[i], o);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 6; i++) p.part[(size_t)it * 27 + i] = c6[i];
        }
    }
}
// ref block_solver.hpp:418-444: x_l = D^-1 (b_l - W^T x_p), one thread per landmark
__device__ void f_backsub(const BaDev& p, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        double c0 = p.bl[3 * li], c1 = p.bl[3 * li + 1], c2 = p.bl[3 * li + 2];
        const int k0 = p.l_ptr[li], k1 = p.l_ptr[li + 1];
#pragma unroll 2
        for (int e = k0; e < k1; e++) {
            const int hj = p.cam_h[p.e_cam[e]];
            if (hj < 0) continue;
            const double2* W2 = reinterpret_cast<const double2*>(p.W + 18 * (size_t)e);
            double w[18];
#pragma unroll
            for (int i = 0; i < 9; i++) { const double2 v = W2[i]; w[2 * i] = v.x; w[2 * i + 1] = v.y; }
            const double* xp = p.x + 6 * hj;
#pragma unroll
            for (int r = 0; r < 6; r++) { const double xr = xp[r]; c0 -= w[r * 3] * xr; c1 -= w[r * 3 + 1] * xr; c2 -= w[r * 3 + 2] * xr; }
        }
        const double* D = p.Dinv + 9 * (size_t)li;
        p.x[p.n + 3 * li] = D[0] * c0 + D[1] * c1 + D[2] * c2;
        p.x[p.n + 3 * li + 1] = D[3] * c0 + D[4] * c1 + D[5] * c2;
        p.x[p.n + 3 * li + 2] = D[6] * c0 + D[7] * c1 + D[8] * c2;
    }
}


// ------------------------------------------------------------------------------------------------ local-window fast path
// For windows with <= kFastMaxKf free cameras the 6x3 pose-landmark blocks W and W D^-1 are never written to global memory:
// a trial walks the landmarks in batches of <= kFastBE edges, rebuilds the batch's blocks from the 32-byte edge records into
// shared memory, and thread pair i -- owner of one (reduced-system block, part) item for the whole trial -- adds the batch's
// products to 18 register accumulators per thread. One fixed-order reduction over the parts at the end: bit-reproducible, no
// atomics, and the per-problem traffic drops from ~13 MB to ~3 MB per LM iteration (what bounds many problems in flight).
constexpr int kBaThreadsC = 256;             // = kBaThreads (declared below)
constexpr int kFastBE = 512;                // edges per batch
constexpr int kFastBL = 128;                // landmarks per batch (their D^-1 and D^-1 b are staged in shared memory, 9 doubles each)
constexpr int kFastItems = 128;             // (block, part) items; thread t and t + 128 own the left / right 3 columns of item t's 6x6 block
constexpr int kFastMaxKf = 10;
constexpr int kFastRec = 18;                // doubles per staged edge: a b | 1/z jdb0 | jdb1 - | w J D^-1 (2x3) | w J (2x3)
constexpr int kFastRed = 21;                // doubles per thread in the final reduction: 18 products + 3 coefficients
constexpr int kFastLD = 9;                  // D^-1 (6 unique) + D^-1 b_l (3) per staged landmark
constexpr size_t kFastRegionBytes = 82 * 1024;
static_assert(sizeof(double) * (kFastBE * kFastRec + kFastBL * kFastLD) <= kFastRegionBytes, "batch buffers must fit the region");
static_assert(sizeof(double) * ((6 * kFastMaxKf) * (6 * kFastMaxKf + 1) + 2 * kFastItems * kFastRed) <= kFastRegionBytes, "S + reduction slots must fit the region");

// computeError at the current state + the edge records; returns this thread's share of the robust chi2 (edge order = g_chi2's)
__device__ double g_project(const BaDev& p, double* __restrict__ rec, double* __restrict__ err, double delta, int tid, int nt)
{
    double acc = 0;
#pragma unroll 4
    for (int e = tid; e < p.Ea; e += nt) {
        const int c = p.e_cam[e];
        const double* __restrict__ R = p.cam_R + 9 * c;
        const double* __restrict__ X = p.pt_X + 3 * (size_t)p.e_pt[e];
        const double X0 = X[0], X1 = X[1], X2 = X[2];
        const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + p.cam_t[3 * c];
        const double y = R[3] * X0 + R[4] * X1 + R[5] * X2 + p.cam_t[3 * c + 1];
        const double z = R[6] * X0 + R[7] * X1 + R[8] * X2 + p.cam_t[3 * c + 2];
        const double iz = 1.0 / z, a = x * iz, b = y * iz, f = p.cam_f[c];
        const double2 uv = *reinterpret_cast<const double2*>(p.e_uv + 2 * (size_t)e);
        const double e0 = uv.x - (a * f + p.cam_cx[c]), e1 = uv.y - (b * f + p.cam_cy[c]);
        *reinterpret_cast<double2*>(err + 2 * (size_t)e) = make_double2(e0, e1);
        double2* r2 = reinterpret_cast<double2*>(rec + kRec * (size_t)e);
        r2[0] = make_double2(a, b);
        r2[1] = make_double2(iz, 0.0);
        double r0, r1;
        huber(p.e_info[e] * (e0 * e0 + e1 * e1), delta, r0, r1);
        acc += r0;
    }
    return acc;
}
__device__ double g_chi2(const BaDev& p, const double* __restrict__ err, double delta, double* sh)
{
    double acc = 0;
    for (int e = threadIdx.x; e < p.Ea; e += blockDim.x) {
        const double2 er = *reinterpret_cast<const double2*>(err + 2 * (size_t)e);
        double r0, r1;
        huber(p.e_info[e] * (er.x * er.x + er.y * er.y), delta, r0, r1);
        acc += r0;
    }
    if (p.nT) acc += tether_chi2_sum(p);
    return block_sum(acc, sh);
}
// landmark blocks; the robust weight w of every edge goes into its record for the later phases
__device__ void g_build_points(const BaDev& p, double* __restrict__ rec, const double* __restrict__ err, double delta, int tid, int nt)
{
    for (int li = tid; li < p.Pl; li += nt) {
        double H00 = 0, H01 = 0, H02 = 0, H11 = 0, H12 = 0, H22 = 0, b0 = 0, b1 = 0, b2 = 0;
        const int k0 = p.l_ptr[li], k1 = p.l_ptr[li + 1];
#pragma unroll 2
        for (int e = k0; e < k1; e++) {
            const int c = p.e_cam[e];
            double2* r2 = reinterpret_cast<double2*>(rec + kRec * (size_t)e);
            const double2 ab = r2[0];
            const double iz = r2[1].x, f = p.cam_f[c];
            const double2 er = *reinterpret_cast<const double2*>(err + 2 * (size_t)e);
            const double info = p.e_info[e];
            double r0, r1;
            huber(info * (er.x * er.x + er.y * er.y), delta, r0, r1);
            const double w = r1 * info, o0 = -(w * er.x), o1 = -(w * er.y);
            r2[1] = make_double2(iz, w);
            double J0[3], J1[3];
            f_point_jac(p.cam_R + 9 * c, ab.x, ab.y, iz, f, J0, J1);
            const double w0[3] = {w * J0[0], w * J0[1], w * J0[2]}, w1[3] = {w * J1[0], w * J1[1], w * J1[2]};
            H00 += w0[0] * J0[0] + w1[0] * J1[0]; H01 += w0[0] * J0[1] + w1[0] * J1[1]; H02 += w0[0] * J0[2] + w1[0] * J1[2];
            H11 += w0[1] * J0[1] + w1[1] * J1[1]; H12 += w0[1] * J0[2] + w1[1] * J1[2]; H22 += w0[2] * J0[2] + w1[2] * J1[2];
            b0 += J0[0] * o0 + J1[0] * o1; b1 += J0[1] * o0 + J1[1] * o1; b2 += J0[2] * o0 + J1[2] * o1;
        }
        double* Hl = p.Hll + 9 * (size_t)li;
        Hl[0] = H00; Hl[1] = H01; Hl[2] = H02; Hl[3] = H01; Hl[4] = H11; Hl[5] = H12; Hl[6] = H02; Hl[7] = H12; Hl[8] = H22;
        p.bl[3 * li] = b0; p.bl[3 * li + 1] = b1; p.bl[3 * li + 2] = b2;
    }
}
__device__ void g_build_cams(const BaDev& p, const double* __restrict__ rec, const double* __restrict__ err, int parts, int warp, int nwarps, int lane)
{
    const int items = p.Kf * parts;
    for (int it = warp; it < items; it += nwarps) {
        const int kf = it / parts, part = it % parts;
        const int beg = p.c_ptr[kf], end = p.c_ptr[kf + 1], len = end - beg;
        const int per = (len + parts - 1) / parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        const double f = p.cam_f[p.c_cam[kf]];
        double A[21], b[6];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = 0;
#pragma unroll
        for (int i = 0; i < 6; i++) b[i] = 0;
#pragma unroll 2
        for (int k = s0 + lane; k < s1; k += 32) {
            const int e = p.c_edges[k];
            const double2* r2 = reinterpret_cast<const double2*>(rec + kRec * (size_t)e);
            const double2 ab = r2[0], zw = r2[1];
            const double2 er = *reinterpret_cast<const double2*>(err + 2 * (size_t)e);
            const double w = zw.y, o0 = -(w * er.x), o1 = -(w * er.y);
            double P0[6], P1[6];
            f_pose_jac(ab.x, ab.y, zw.x, f, P0, P1);
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                b[r] += P0[r] * o0 + P1[r] * o1;
                const double a0 = w * P0[r], a1 = w * P1[r];
#pragma unroll
                for (int c = r; c < 6; c++) A[idx++] += a0 * P0[c] + a1 * P1[c];
            }
        }
#pragma unroll
        for (int i = 0; i < 21; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) A[i] += __shfl_down_sync(0xffffffffu, A[i], o);
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) b[i] += __shfl_down_sync(0xffffffffu, b[i], o);
        if (lane == 0) {
            double* dst = p.part + (size_t)it * 27;
#pragma unroll
            for (int i = 0; i < 21; i++) dst[i] = A[i];
#pragma unroll
            for (int i = 0; i < 6; i++) dst[21 + i] = b[i];
        }
    }
}
__device__ void g_finish_cams(const BaDev& p, int parts, int tid, int nt)
{
    for (int i = tid; i < p.Kf * 27; i += nt) {
        const int kf = i / 27, j = i % 27;
        double s = 0;
        for (int part = 0; part < parts; part++) s += p.part[((size_t)kf * parts + part) * 27 + j];
        if (j >= 21) p.bp[6 * kf + (j - 21)] = s;
        else {
            int r = 0, rem = j;
            while (rem >= 6 - r) { rem -= 6 - r; r++; }
            const int c = r + rem;
            p.Hpp[36 * (size_t)kf + r * 6 + c] = s;
            p.Hpp[36 * (size_t)kf + c * 6 + r] = s;
        }
    }
}
// W = w Jj^T Ji of one edge from its record (6x3 row-major)
__device__ __forceinline__ void g_edge_W(const BaDev& p, const double* __restrict__ camR, const double* __restrict__ rec, int e, int c, double* W)
{
    const double2* r2 = reinterpret_cast<const double2*>(rec + kRec * (size_t)e);
    const double2 ab = r2[0], zw = r2[1];
    const double f = p.cam_f[c];
    double J0[3], J1[3], P0[6], P1[6];
    f_point_jac(camR + 9 * c, ab.x, ab.y, zw.x, f, J0, J1);
    f_pose_jac(ab.x, ab.y, zw.x, f, P0, P1);
    const double w0[3] = {zw.y * J0[0], zw.y * J0[1], zw.y * J0[2]}, w1[3] = {zw.y * J1[0], zw.y * J1[1], zw.y * J1[2]};
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) W[r * 3 + cc] = P0[r] * w0[cc] + P1[r] * w1[cc];
}
// Schur complement of one trial (ref block_solver.hpp:337-390) + assembly of the reduced system and its right-hand side
// into shared memory. region = the CTA's kFastRegionBytes scratch (batch buffer, then S | bs | reduction slots).
// W = w Pj^T J is rank 2 (P = 2x6 pose Jacobian, J = 2x3 point Jacobian), so  (W_i D^-1) W_j^T = P_i^T [ (w_i J_i D^-1)(w_j J_j)^T ] P_j:
// an edge is staged as (a, b, 1/z) -- P is rebuilt from those -- plus the two 2x3 factors, 144 bytes instead of two 6x3 blocks.
__device__ void g_schur(const BaDev& p, const double* __restrict__ rec, double lambda, double* region)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    double* sD = region + (size_t)kFastBE * kFastRec;                  // [kFastBL][kFastLD] behind the staged edges
    const int item = tid & (kFastItems - 1), half = tid >> 7;          // warp-uniform half: columns 3*half .. 3*half+2 (and rows, for the coefficients)
    int blk = -1, part = 0, nparts = 1;
    bool diag = false;
    double f1 = 0, f2 = 0;
    if (item < p.nitems) {
        const int4 d = p.item_def[item];
        blk = d.x; part = d.y; nparts = d.z;
        const int i1 = p.blk_ij[2 * blk], i2 = p.blk_ij[2 * blk + 1];
        diag = i1 == i2;
        f1 = p.cam_f[p.c_cam[i1]]; f2 = p.cam_f[p.c_cam[i2]];
    }
    double acc[18], cacc[3] = {0, 0, 0};
#pragma unroll
    for (int i = 0; i < 18; i++) acc[i] = 0;
    for (int b = 0; b < p.nb; b++) {
        const int lm0 = p.batch_ptr[b], lm1 = p.batch_ptr[b + 1];
        const int e0 = p.l_ptr[lm0], ne = p.l_ptr[lm1] - e0;
        int q0 = 0, qn = 0;
        if (blk >= 0) { q0 = p.bb_ptr[b * p.nblk + blk]; qn = p.bb_ptr[b * p.nblk + blk + 1] - q0; }
        __syncthreads();                                     // the previous batch has been consumed
        // the edges' own data first (two per thread, independent loads in flight under the landmark phase)
        int ec[kFastBE / kBaThreadsC], el[kFastBE / kBaThreadsC];
        double2 eab[kFastBE / kBaThreadsC], ezw[kFastBE / kBaThreadsC];
#pragma unroll
        for (int u = 0; u < kFastBE / kBaThreadsC; u++) {
            const int t = tid + u * kBaThreadsC;
            ec[u] = -1; el[u] = 0;
            if (t < ne) {
                const int e = e0 + t;
                ec[u] = p.e_cam[e]; el[u] = p.e_l[e] - lm0;
                const double2* r2 = reinterpret_cast<const double2*>(rec + kRec * (size_t)e);
                eab[u] = r2[0]; ezw[u] = r2[1];
            }
        }
        // one thread per landmark of the batch: D^-1 = (H_ll + lambda I)^-1 and D^-1 b_l, to shared memory (and to global for the back-substitution)
        if (tid < lm1 - lm0) {
            const int li = lm0 + tid;
            const double* __restrict__ Hl = p.Hll + 9 * (size_t)li;
            const double A0 = Hl[0] + lambda, A1 = Hl[1], A2 = Hl[2], A4 = Hl[4] + lambda, A5 = Hl[5], A8 = Hl[8] + lambda;      // symmetric
            const double c00 = A4 * A8 - A5 * A5, c01 = A5 * A2 - A1 * A8, c02 = A1 * A5 - A4 * A2;
            const double id = 1.0 / (A0 * c00 + A1 * c01 + A2 * c02);
            const double D0 = c00 * id, D1 = c01 * id, D2 = c02 * id, D4 = (A0 * A8 - A2 * A2) * id, D5 = (A2 * A1 - A0 * A5) * id, D8 = (A0 * A4 - A1 * A1) * id;
            const double b0 = p.bl[3 * li], b1 = p.bl[3 * li + 1], b2 = p.bl[3 * li + 2];
            const double db0 = D0 * b0 + D1 * b1 + D2 * b2, db1 = D1 * b0 + D4 * b1 + D5 * b2, db2 = D2 * b0 + D5 * b1 + D8 * b2;
            double* Dg = p.Dinv + 9 * (size_t)li;
            Dg[0] = D0; Dg[1] = D1; Dg[2] = D2; Dg[3] = D1; Dg[4] = D4; Dg[5] = D5; Dg[6] = D2; Dg[7] = D5; Dg[8] = D8;
            p.db[3 * li] = db0; p.db[3 * li + 1] = db1; p.db[3 * li + 2] = db2;
            double* sd = sD + kFastLD * tid;
            sd[0] = D0; sd[1] = D1; sd[2] = D2; sd[3] = D4; sd[4] = D5; sd[5] = D8; sd[6] = db0; sd[7] = db1; sd[8] = db2;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kFastBE / kBaThreadsC; u++) {
            const int c = ec[u];
            if (c >= 0 && p.cam_h[c] >= 0) {
                const int t = tid + u * kBaThreadsC;
                const double* sd = sD + kFastLD * el[u];
                const double D0 = sd[0], D1 = sd[1], D2 = sd[2], D4 = sd[3], D5 = sd[4], D8 = sd[5], db0 = sd[6], db1 = sd[7], db2 = sd[8];
                const double2 ab = eab[u], zw = ezw[u];
                double J0[3], J1[3];
                f_point_jac(p.cam_R + 9 * c, ab.x, ab.y, zw.x, p.cam_f[c], J0, J1);
#pragma unroll
                for (int k = 0; k < 3; k++) { J0[k] *= zw.y; J1[k] *= zw.y; }                      // w J
                double2* o = reinterpret_cast<double2*>(region + kFastRec * t);
                o[0] = ab;
                o[1] = make_double2(zw.x, J0[0] * db0 + J0[1] * db1 + J0[2] * db2);
                o[2] = make_double2(J1[0] * db0 + J1[1] * db1 + J1[2] * db2, 0.0);
                const double g00 = J0[0] * D0 + J0[1] * D1 + J0[2] * D2, g01 = J0[0] * D1 + J0[1] * D4 + J0[2] * D5, g02 = J0[0] * D2 + J0[1] * D5 + J0[2] * D8;
                const double g10 = J1[0] * D0 + J1[1] * D1 + J1[2] * D2, g11 = J1[0] * D1 + J1[1] * D4 + J1[2] * D5, g12 = J1[0] * D2 + J1[1] * D5 + J1[2] * D8;
                o[3] = make_double2(g00, g01); o[4] = make_double2(g02, g10); o[5] = make_double2(g11, g12);
                o[6] = make_double2(J0[0], J0[1]); o[7] = make_double2(J0[2], J1[0]); o[8] = make_double2(J1[1], J1[2]);
            }
        }
        ushort2 pr = make_ushort2(0, 0);                     // the first pair of this thread: fetched before the barrier, not after it
        const int kbeg = q0 + (part * qn) / nparts, kend = q0 + ((part + 1) * qn) / nparts;
        if (blk >= 0 && kbeg < kend) pr = p.bpairs[kbeg];
        __syncthreads();
        if (blk >= 0) {
            for (int k = kbeg; k < kend; k++) {
                const double2* __restrict__ A2 = reinterpret_cast<const double2*>(region + kFastRec * pr.x);
                const double2* __restrict__ B2 = reinterpret_cast<const double2*>(region + kFastRec * pr.y);
                if (k + 1 < kend) pr = p.bpairs[k + 1];          // next pair's indices under this pair's arithmetic
                const double2 ia = A2[0], iz = A2[1], g0 = A2[3], g1 = A2[4], g2 = A2[5];
                const double2 ja = B2[0], jz = B2[1], h0 = B2[6], h1 = B2[7], h2 = B2[8];
                // M = (w_i J_i D^-1)(w_j J_j)^T
                const double M00 = g0.x * h0.x + g0.y * h0.y + g1.x * h1.x, M01 = g0.x * h1.y + g0.y * h2.x + g1.x * h2.y;
                const double M10 = g1.y * h0.x + g2.x * h0.y + g2.y * h1.x, M11 = g1.y * h1.y + g2.x * h2.x + g2.y * h2.y;
                // T = M P_j[:, 3*half ..]
                double T0[3], T1[3];
                if (half == 0) {
                    const double q00 = ja.x * ja.y * f2, q01 = -(f2 + ja.x * ja.x * f2), q02 = ja.y * f2;
                    const double q10 = f2 + ja.y * ja.y * f2, q12 = -(ja.x * f2);
                    T0[0] = M00 * q00 + M01 * q10; T0[1] = M00 * q01 - M01 * q00; T0[2] = M00 * q02 + M01 * q12;
                    T1[0] = M10 * q00 + M11 * q10; T1[1] = M10 * q01 - M11 * q00; T1[2] = M10 * q02 + M11 * q12;
                } else {
                    const double zf = jz.x * f2;
                    T0[0] = -(zf * M00); T0[1] = -(zf * M01); T0[2] = zf * (ja.x * M00 + ja.y * M01);
                    T1[0] = -(zf * M10); T1[1] = -(zf * M11); T1[2] = zf * (ja.x * M10 + ja.y * M11);
                }
                double P0[6], P1[6];
                f_pose_jac(ia.x, ia.y, iz.x, f1, P0, P1);
#pragma unroll
                for (int r = 0; r < 6; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++) acc[r * 3 + c] += P0[r] * T0[c] + P1[r] * T1[c];
                if (diag & (half == 0)) {
                    const double j1 = A2[2].x;
#pragma unroll
                    for (int r = 0; r < 3; r++) cacc[r] += P0[r] * iz.y + P1[r] * j1;
                } else if (diag) {
                    const double j1 = A2[2].x;
#pragma unroll
                    for (int r = 0; r < 3; r++) cacc[r] += P0[3 + r] * iz.y + P1[3 + r] * j1;
                }
            }
        }
    }
    __syncthreads();
    const int n = p.n;
    double* S = region;
    double* bs = S + (size_t)n * n;
    double* red = bs + n;
    {
        double* dst = red + (size_t)(item * 2 + half) * kFastRed;
#pragma unroll
        for (int i = 0; i < 18; i++) dst[i] = acc[i];
#pragma unroll
        for (int i = 0; i < 3; i++) dst[18 + i] = cacc[i];
    }
    for (int i = tid; i < n * n; i += nt) S[i] = 0.0;
    __syncthreads();
    for (int o = tid; o < p.nblk * 36; o += nt) {
        const int bi = o / 36, r = (o % 36) / 6, c = o % 6;
        const int2 it = p.blk_items[bi];
        double s = 0;
        for (int q = 0; q < it.y; q++) s += red[((it.x + q) * 2 + c / 3) * kFastRed + r * 3 + (c % 3)];
        const int i1 = p.blk_ij[2 * bi], i2 = p.blk_ij[2 * bi + 1];
        double v = -s;
        if (i1 == i2) v += p.Hpp[36 * (size_t)i1 + r * 6 + c] + ((r == c) ? lambda : 0.0);
        S[(size_t)(6 * i1 + r) * n + 6 * i2 + c] = v;
        if (i1 != i2) S[(size_t)(6 * i2 + c) * n + 6 * i1 + r] = v;
    }
    for (int i = tid; i < n; i += nt) {
        const int kf = i / 6, r = i % 6;
        const int2 it = p.blk_items[p.cam_diag[kf]];
        double s = 0;
        for (int q = 0; q < it.y; q++) s += red[((it.x + q) * 2 + r / 3) * kFastRed + 18 + (r % 3)];
        bs[i] = p.bp[i] - s;
    }
}
// One pass per trial after the reduced system has been solved and the cameras have been updated (their old rotations are in
// cam_Rold): per landmark  x_l = D^-1 (b_l - W^T x_p)  (ref block_solver.hpp:418-444; W rebuilt from the records at the
// linearisation point), push + oplus of the point (ref base_vertex.h:92-94, types_sba.h:149-153), then computeError of its edges at
// the trial state into the trial records / errors. Returns this thread's shares of the robust chi2 and of computeScale's sum.
__device__ void g_backsub_update_project(const BaDev& p, const double* __restrict__ recCur, double* __restrict__ recTry, double* __restrict__ errTry,
                                         double lambda, double delta, bool solved, int tid, int nt, double& chiPart, double& scalePart)
{
    double chi = 0, sc = 0;
    for (int j = tid; j < p.n; j += nt) { const double xj = p.x[j]; sc += xj * (lambda * xj + p.bp[j]); }
    for (int li = tid; li < p.Pl; li += nt) {
        const int k0 = p.l_ptr[li], k1 = p.l_ptr[li + 1];
        const double bl0 = p.bl[3 * li], bl1 = p.bl[3 * li + 1], bl2 = p.bl[3 * li + 2];
        double x0, x1, x2;
        if (solved) {
            double c0 = bl0, c1 = bl1, c2 = bl2;
#pragma unroll 2
            for (int e = k0; e < k1; e++) {
                const int c = p.e_cam[e], hj = p.cam_h[c];
                if (hj < 0) continue;
                double W[18];
                g_edge_W(p, p.cam_Rold, recCur, e, c, W);
                const double* xp = p.x + 6 * hj;
#pragma unroll
                for (int r = 0; r < 6; r++) { const double xr = xp[r]; c0 -= W[r * 3] * xr; c1 -= W[r * 3 + 1] * xr; c2 -= W[r * 3 + 2] * xr; }
            }
            const double* D = p.Dinv + 9 * (size_t)li;
            x0 = D[0] * c0 + D[1] * c1 + D[2] * c2; x1 = D[3] * c0 + D[4] * c1 + D[5] * c2; x2 = D[6] * c0 + D[7] * c1 + D[8] * c2;
            p.x[p.n + 3 * li] = x0; p.x[p.n + 3 * li + 1] = x1; p.x[p.n + 3 * li + 2] = x2;
        } else {                                   // failed factorisation: g2o updates with the stale increment (ref block_solver.hpp:396-400)
            x0 = p.x[p.n + 3 * li]; x1 = p.x[p.n + 3 * li + 1]; x2 = p.x[p.n + 3 * li + 2];
        }
        double* X = p.pt_X + 3 * (size_t)p.l_pt[li];
        const double o0 = X[0], o1 = X[1], o2 = X[2];
        p.pt_bak[3 * li] = o0; p.pt_bak[3 * li + 1] = o1; p.pt_bak[3 * li + 2] = o2;
        const double X0 = o0 + x0, X1 = o1 + x1, X2 = o2 + x2;
        X[0] = X0; X[1] = X1; X[2] = X2;
        sc += x0 * (lambda * x0 + bl0); sc += x1 * (lambda * x1 + bl1); sc += x2 * (lambda * x2 + bl2);
#pragma unroll 2
        for (int e = k0; e < k1; e++) {
            const int c = p.e_cam[e];
            const double* __restrict__ R = p.cam_R + 9 * c;
            const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + p.cam_t[3 * c];
            const double y = R[3] * X0 + R[4] * X1 + R[5] * X2 + p.cam_t[3 * c + 1];
            const double z = R[6] * X0 + R[7] * X1 + R[8] * X2 + p.cam_t[3 * c + 2];
            const double iz = 1.0 / z, a = x * iz, b = y * iz, f = p.cam_f[c];
            const double2 uv = *reinterpret_cast<const double2*>(p.e_uv + 2 * (size_t)e);
            const double e0 = uv.x - (a * f + p.cam_cx[c]), e1 = uv.y - (b * f + p.cam_cy[c]);
            *reinterpret_cast<double2*>(errTry + 2 * (size_t)e) = make_double2(e0, e1);
            double2* r2 = reinterpret_cast<double2*>(recTry + kRec * (size_t)e);
            r2[0] = make_double2(a, b);
            r2[1] = make_double2(iz, 0.0);
            double r0, r1;
            huber(p.e_info[e] * (e0 * e0 + e1 * e1), delta, r0, r1);
            chi += r0;
        }
    }
    chiPart = chi; scalePart = sc;
}
// push + oplus of the free cameras (ref base_vertex.h:92-94, types_six_dof_expmap.h:98-101); the rotation matrices of the
// linearisation point are kept for the back-substitution
__device__ void g_update_cams(const BaDev& p, int tid, int nt)
{
    for (int i = tid; i < 9 * p.K; i += nt) p.cam_Rold[i] = p.cam_R[i];
    for (int i = tid; i < p.Kf; i += nt) {
        const int c = p.c_cam[i];
#pragma unroll
        for (int j = 0; j < 4; j++) p.cam_bak[7 * i + j] = p.cam_q[4 * c + j];
#pragma unroll
        for (int j = 0; j < 3; j++) p.cam_bak[7 * i + 4 + j] = p.cam_t[3 * c + j];
        pose_oplus(p.cam_q + 4 * c, p.cam_t + 3 * c, p.x + 6 * i);
    }
}

// ------------------------------------------------------------------------------------------------ the persistent LM kernel
// One CTA per problem runs StepBundleAdjustment's whole loop: for each Huber width one g2o LM iteration
// (ref optimization_algorithm_levenberg.cpp:57-149, up to 10 lambda trials), then the outlier classification.
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define PH(i) do { if (blockIdx.x == 0 && tid == 0) { long long _n = gtimer(); ctl->phase_ns[i] += _n - t_last; t_last = _n; } } while (0)
constexpr int kBaThreads = 256;

// Shared-memory residency: the reduced system S (n x n) + its right-hand side and the camera state (pose, intrinsics,
// Hessian index) live in dynamic shared memory whenever they fit -- the dense LDL^T and every per-edge camera lookup then run
// at shared-memory latency instead of L2 latency. Global memory stays the source of truth across launches.
constexpr int kBaMaxSmemCams = 1024;
constexpr size_t kCoopSmemMax = 220 * 1024;        // dynamic shared memory of the cooperative kernel (one CTA per SM)

__host__ __device__ inline size_t ba_smem_need_S(int n) { return sizeof(double) * ((size_t)n * n + n); }
__host__ __device__ inline size_t ba_smem_need_cams(int K) { return sizeof(double) * 10 * (size_t)K + sizeof(int) * (size_t)K; }
__host__ __device__ inline size_t ba_smem_need_cams_R(int K) { return sizeof(double) * 28 * (size_t)K + sizeof(int) * (size_t)K; }     // + rotation matrices (current, linearisation point)

// ---- pose-only problems of one free camera (ref Tracking/TrackLocalMap.cpp:421-501 OptimizeCameraPose: ArePointsFixed, the current frame's
// pose against 50 - 400 fixed map points, 3 then 4 LM iterations, twice per frame on the tracking thread). The reduced system IS the
// camera's 6 x 6 block, so an iteration is two sweeps over the edges -- linearise (errors, robust chi2, H and b in one pass, 28 sums per
// thread, fixed-order reduction) and the trial state's chi2 -- around a 6 x 6 LDL^T by one thread; the general path's fifteen phases with a
// barrier and a round trip to global memory each cost 19 us per iteration on 300 edges, this one 6.5 (ncu, cold caches).
__device__ __forceinline__ void h_project(const double* __restrict__ R, const double* __restrict__ t, const double* __restrict__ X, double& a, double& b, double& iz)
{
    const double X0 = X[0], X1 = X[1], X2 = X[2];
    const double x = R[0] * X0 + R[1] * X1 + R[2] * X2 + t[0];
    const double y = R[3] * X0 + R[4] * X1 + R[5] * X2 + t[1];
    const double z = R[6] * X0 + R[7] * X1 + R[8] * X2 + t[2];
    iz = 1.0 / z; a = x * iz; b = y * iz;
}
// errors of all edges at the current pose (stored: the outlier pass reads the last computed ones) and this thread's share of the robust chi2
__device__ double h_errors_chi2(const BaDev& p, int c, double delta, int tid, int nt)
{
    const double* __restrict__ R = p.cam_R + 9 * c;
    const double* __restrict__ t = p.cam_t + 3 * c;
    const double f = p.cam_f[c], cx = p.cam_cx[c], cy = p.cam_cy[c];
    double acc = 0;
    for (int e = tid; e < p.Ea; e += nt) {
        double a, b, iz;
        h_project(R, t, p.pt_X + 3 * (size_t)p.e_pt[e], a, b, iz);
        const double2 uv = *reinterpret_cast<const double2*>(p.e_uv + 2 * (size_t)e);
        const double e0 = uv.x - (a * f + cx), e1 = uv.y - (b * f + cy);
        *reinterpret_cast<double2*>(p.err + 2 * (size_t)e) = make_double2(e0, e1);
        double r0, r1;
        huber(p.e_info[e] * (e0 * e0 + e1 * e1), delta, r0, r1);
        acc += r0;
    }
    return acc;
}
// linearisation at the current pose: errors, robust chi2, the upper triangle of H (21) and b (6) -> red[0..27] (shared), summed in a
// fixed order (lanes by shuffle tree, warps in sequence); same per-edge arithmetic as f_project + f_build_cams
__device__ void h_linearize(const BaDev& p, int c, double delta, double* __restrict__ part /* [warps][28] */, double* __restrict__ red /* [28] */)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const double* __restrict__ R = p.cam_R + 9 * c;
    const double* __restrict__ t = p.cam_t + 3 * c;
    const double f = p.cam_f[c], cx = p.cam_cx[c], cy = p.cam_cy[c];
    double A[28];
#pragma unroll
    for (int i = 0; i < 28; i++) A[i] = 0;
    for (int e = tid; e < p.Ea; e += nt) {
        double a, b, iz;
        h_project(R, t, p.pt_X + 3 * (size_t)p.e_pt[e], a, b, iz);
        const double2 uv = *reinterpret_cast<const double2*>(p.e_uv + 2 * (size_t)e);
        const double e0 = uv.x - (a * f + cx), e1 = uv.y - (b * f + cy);
        *reinterpret_cast<double2*>(p.err + 2 * (size_t)e) = make_double2(e0, e1);
        const double info = p.e_info[e];
        double r0, r1;
        huber(info * (e0 * e0 + e1 * e1), delta, r0, r1);
        A[27] += r0;
        const double w = r1 * info, o0 = -info * e0 * r1, o1 = -info * e1 * r1;
        double P0[6], P1[6];
        f_pose_jac(a, b, iz, f, P0, P1);
        int idx = 0;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            A[21 + r] += P0[r] * o0 + P1[r] * o1;
            const double a0 = w * P0[r], a1 = w * P1[r];
#pragma unroll
            for (int cc = r; cc < 6; cc++) A[idx++] += a0 * P0[cc] + a1 * P1[cc];
        }
    }
#pragma unroll
    for (int i = 0; i < 28; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) A[i] += __shfl_down_sync(0xffffffffu, A[i], o);
        if (lane == 0) part[warp * 28 + i] = A[i];
    }
    __syncthreads();
    if (tid < 28) { double s = 0; for (int w = 0; w < nw; w++) s += part[w * 28 + tid]; red[tid] = s; }
    __syncthreads();
}
// (H + lambda I) x = b for the 6 x 6 block by one thread: LDL^T with the pivot rules of phase_ldlt_solve (a negative pivot fails like
// Eigen's isPositive(), a zero pivot is skipped), x written only on success
__device__ bool h_solve6(const double* __restrict__ red, double lambda, double* __restrict__ x)
{
    double S[6][6], y[6];
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int cc = r; cc < 6; cc++) { const double v = red[idx++]; S[r][cc] = v; S[cc][r] = v; }
#pragma unroll
    for (int r = 0; r < 6; r++) { S[r][r] += lambda; y[r] = red[21 + r]; }
    bool neg = false;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const double d = S[k][k];
        if (d < 0) neg = true;
        const bool valid = fabs(d) > 0;
        const double inv_d = valid ? 1.0 / d : 0.0;
        if (valid) {
#pragma unroll
            for (int i = k + 1; i < 6; i++) {
                const double aik = S[i][k] * inv_d;
#pragma unroll
                for (int j = k + 1; j <= i; j++) S[i][j] -= aik * S[j][k];
            }
#pragma unroll
            for (int i = k + 1; i < 6; i++) S[i][k] *= inv_d;
        }
    }
    if (neg) return false;
    const double tol = 1.0 / DBL_MAX;
#pragma unroll
    for (int k = 0; k < 6; k++)
#pragma unroll
        for (int i = k + 1; i < 6; i++) y[i] -= S[i][k] * y[k];
#pragma unroll
    for (int i = 0; i < 6; i++) y[i] = (fabs(S[i][i]) > tol) ? y[i] / S[i][i] : 0.0;
#pragma unroll
    for (int k = 5; k >= 0; k--)
#pragma unroll
        for (int i = 0; i < k; i++) y[i] -= S[k][i] * y[k];
#pragma unroll
    for (int i = 0; i < 6; i++) x[i] = y[i];
    return true;
}

// POSE1 = true compiles the pose-only path of one free camera in (its own kernel: the registers of its 28 running sums must not weigh on
// the batched local-window path, which lives at 128 registers for two CTAs per SM)
template <bool POSE1>
__global__ void __launch_bounds__(kBaThreads, POSE1 ? 1 : 2) k_ba_step_t(const BaDev* __restrict__ probs, const float* __restrict__ huberW, int nIters, float maxErrSq,
                                                                          unsigned dynBytes)
{
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ double sh[33];
    __shared__ double s_ldlt_col[2 * 97];
    __shared__ double s_lambda, s_ni, s_rho;
    __shared__ int s_accept, s_stop;
    __shared__ BaDev s_p;
    __shared__ double *g_cam_q, *g_cam_t;       // global home of the camera state (written back at the end)
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
    // set-up by the whole CTA: the problem descriptor and the camera state are fetched with one load per thread (one thread copying them
    // word after word was 29 % of the samples of a pose-only call under ncu's cold caches; with warm caches the call's time is unchanged)
    __shared__ const double *g_cam_f, *g_cam_cx, *g_cam_cy;
    __shared__ const int* g_cam_h;
    {
        static_assert(sizeof(BaDev) % 8 == 0, "BaDev is copied as 8-byte words");
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(probs + blockIdx.x);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&s_p);
        for (int i = tid; i < (int)(sizeof(BaDev) / 8); i += nt) dst[i] = src[i];
    }
    __syncthreads();
    if (tid == 0) {
        s_p.ldlt_col = s_ldlt_col;
        g_cam_q = nullptr; g_cam_t = nullptr;
        size_t off = 0;
        if (s_p.fast && kFastRegionBytes + ba_smem_need_cams_R(s_p.K) > dynBytes) s_p.fast = 0;      // cannot happen (host sizes the launch)
        if (s_p.fast) {
            s_p.S = reinterpret_cast<double*>(dyn); s_p.bs = s_p.S + (size_t)s_p.n * s_p.n;
            off = kFastRegionBytes;
        } else if (s_p.n > 0 && ba_smem_need_S(s_p.n) <= dynBytes) {
            s_p.S = reinterpret_cast<double*>(dyn); s_p.bs = s_p.S + (size_t)s_p.n * s_p.n;
            off = ba_smem_need_S(s_p.n);
        }
        if (s_p.K <= kBaMaxSmemCams && off + ba_smem_need_cams_R(s_p.K) <= dynBytes) {
            double* c = reinterpret_cast<double*>(dyn + off);
            g_cam_q = s_p.cam_q; g_cam_t = s_p.cam_t;
            g_cam_f = s_p.cam_f; g_cam_cx = s_p.cam_cx; g_cam_cy = s_p.cam_cy; g_cam_h = s_p.cam_h;
            s_p.cam_R = c + 10 * s_p.K; s_p.cam_Rold = c + 19 * s_p.K;
            s_p.cam_q = c; s_p.cam_t = c + 4 * s_p.K; s_p.cam_f = c + 7 * s_p.K; s_p.cam_cx = c + 8 * s_p.K; s_p.cam_cy = c + 9 * s_p.K;
            s_p.cam_h = reinterpret_cast<int*>(c + 28 * s_p.K);
        }
    }
    __syncthreads();
    if (g_cam_q) {                                             // the camera state into shared memory, all threads
        const int K = s_p.K;
        double* sq = s_p.cam_q; double* st = s_p.cam_t;
        double* sf = const_cast<double*>(s_p.cam_f); double* sx = const_cast<double*>(s_p.cam_cx); double* sy = const_cast<double*>(s_p.cam_cy);
        int* shh = const_cast<int*>(s_p.cam_h);
        for (int i = tid; i < 4 * K; i += nt) sq[i] = g_cam_q[i];
        for (int i = tid; i < 3 * K; i += nt) st[i] = g_cam_t[i];
        for (int i = tid; i < K; i += nt) { sf[i] = g_cam_f[i]; sx[i] = g_cam_cx[i]; sy[i] = g_cam_cy[i]; shh[i] = g_cam_h[i]; }
    }
    __syncthreads();
    const BaDev& p = s_p;
    BaCtl* ctl = p.ctl;
    if (tid == 0) { s_lambda = ctl->lambda; s_ni = ctl->ni; s_stop = 0; }
    __syncthreads();
    int iteration = ctl->iteration;
    long long trials = 0, iters = 0;

    long long t_last = (blockIdx.x == 0 && tid == 0) ? gtimer() : 0;     // %globaltimer reads serialise across the chip: one thread only
    f_cam_R(p, tid, nt);
    __syncthreads();
    if (p.fast) {
        // local-window fast path: edge records + errors are double-buffered (current state | trial state) so that a rejected
        // trial costs no re-projection and the outlier pass still sees the last COMPUTED errors, like g2o
        double* recCur = p.Hc; double* recTry = p.Hc + kRec * (size_t)p.Ea;
        double* errCur = p.err; double* errTry = p.Hc + 2 * kRec * (size_t)p.Ea;
        const double* errLast = errCur;
        double* region = reinterpret_cast<double*>(dyn);
        bool haveState = false;
        double chiDelta = -1.0, currentChi = 0;         // currentChi is valid for the Huber width chiDelta at the current state
        const int camParts = max(1, nw / max(p.Kf, 1));
        for (int it = 0; it < nIters; it++) {
            if (s_stop) break;
            const double delta = (double)huberW[it];
            if (!haveState) {
                double part = g_project(p, recCur, errCur, delta, tid, nt);
                if (p.nT) { phase_tether_errors(p, tid, nt); __syncthreads(); part += tether_chi2_sum(p); }
                currentChi = block_sum(part, sh);
                chiDelta = delta;
                haveState = true; errLast = errCur;
            } else if (delta != chiDelta) {
                currentChi = g_chi2(p, errCur, delta, sh);
                chiDelta = delta;
            }
            PH(0);
            g_build_points(p, recCur, errCur, delta, tid, nt);
            if (p.nT) phase_tether_linearize(p, tid, nt);
            __syncthreads();
            g_build_cams(p, recCur, errCur, camParts, warp, nw, lane);
            __syncthreads();
            g_finish_cams(p, camParts, tid, nt);
            __syncthreads();
            if (p.nT) { phase_tether_accumulate(p, tid, nt); __syncthreads(); }
            PH(1);
            if (iteration == 0) {
                const double md = phase_max_diag(p, sh);
                if (tid == 0) { s_lambda = ctl->user_lambda_init > 0 ? ctl->user_lambda_init : 1e-5 * md; s_ni = 2; }
                __syncthreads();
            }
            double rho = 0;
            int qmax = 0;
            bool lambdaFinite = true;
            do {
                const double lambda = s_lambda;
                g_schur(p, recCur, lambda, region);
                __syncthreads();
                if (p.nT) { phase_tether_offdiag(p, warp, lane); __syncthreads(); }
                PH(3);
                const bool ok2 = phase_ldlt_solve(p, sh);
                __syncthreads();
                PH(4);
                g_update_cams(p, tid, nt);
                __syncthreads();
                f_cam_R(p, tid, nt);
                __syncthreads();
                PH(6);
                double part, scalePart;
                g_backsub_update_project(p, recCur, recTry, errTry, lambda, delta, ok2, tid, nt, part, scalePart);
                if (p.nT) { phase_tether_errors(p, tid, nt); __syncthreads(); part += tether_chi2_sum(p); }
                errLast = errTry;
                double tempChi = block_sum(part, sh);
                if (!ok2) tempChi = DBL_MAX;
                const double scale = block_sum(scalePart, sh) + 1e-3;
                PH(7);
                if (tid == 0) {
                    double r = (currentChi - tempChi) / scale;
                    s_rho = r;
                    if (r > 0 && isfinite(tempChi)) {
                        double alpha = 1. - pow((2 * r - 1), 3.0);
                        alpha = fmin(alpha, 2. / 3.);
                        s_lambda = lambda * fmax(1. / 3., alpha);
                        s_ni = 2;
                        s_accept = 1;
                    } else {
                        s_lambda = lambda * s_ni;
                        s_ni = s_ni * 2;
                        s_accept = 0;
                    }
                }
                __syncthreads();
                rho = s_rho;
                if (s_accept) {
                    currentChi = tempChi;
                    double* t1 = recCur; recCur = recTry; recTry = t1;
                    double* t2 = errCur; errCur = errTry; errTry = t2;
                } else {
                    phase_restore(p, tid, nt);
                    __syncthreads();
                    f_cam_R(p, tid, nt);
                    if (p.nT) haveState = false;          // the tether errors are single-buffered: recompute at the restored state
                    __syncthreads();
                    if (!isfinite(s_lambda)) { lambdaFinite = false; trials++; break; }
                }
                qmax++;
                trials++;
            } while (rho < 0 && qmax < 10);
            iteration++;
            iters++;
            if (qmax == 10 || rho == 0 || !lambdaFinite) { if (tid == 0) s_stop = 1; }     // Terminate => Step() == false => break
            __syncthreads();
        }
        if (errLast != p.err) {
            for (int i = tid; i < 2 * p.Ea; i += nt) p.err[i] = errLast[i];
            __syncthreads();
        }
    } else if (POSE1 && p.pose1) {
        __shared__ double s_hpart[(kBaThreads / 32) * 28], s_hred[28];
        __shared__ int s_ok;
        const int c = p.c_cam[0];
        for (int it = 0; it < nIters; it++) {
            if (s_stop) break;
            const double delta = (double)huberW[it];
            h_linearize(p, c, delta, s_hpart, s_hred);
            double currentChi = s_hred[27];
            PH(1);
            double rho = 0;
            int qmax = 0;
            bool lambdaFinite = true;
            do {
                if (tid == 0) {
                    if (iteration == 0 && qmax == 0) {       // ref computeLambdaInit: tau * max |diag(H)|
                        double md = 0;
                        int idx = 0;
                        for (int r = 0; r < 6; r++) { md = fmax(md, fabs(s_hred[idx])); idx += 6 - r; }
                        s_lambda = ctl->user_lambda_init > 0 ? ctl->user_lambda_init : 1e-5 * md; s_ni = 2;
                    }
                    double x6[6];
                    const bool ok = h_solve6(s_hred, s_lambda, x6);
                    if (ok) for (int i = 0; i < 6; i++) p.x[i] = x6[i];
                    else for (int i = 0; i < 6; i++) x6[i] = p.x[i];          // a failed factorisation leaves the previous increment, like g2o
                    for (int j = 0; j < 4; j++) p.cam_bak[j] = p.cam_q[4 * c + j];
                    for (int j = 0; j < 3; j++) p.cam_bak[4 + j] = p.cam_t[3 * c + j];
                    pose_oplus(p.cam_q + 4 * c, p.cam_t + 3 * c, x6);
                    q_to_R(p.cam_q + 4 * c, p.cam_R + 9 * c);
                    double sc = 0;
                    for (int j = 0; j < 6; j++) sc += x6[j] * (s_lambda * x6[j] + s_hred[21 + j]);
                    s_rho = sc;                                 // the scale until the decision below turns it into rho
                    s_ok = ok ? 1 : 0;
                }
                __syncthreads();
                PH(4);
                const double lambda = s_lambda;
                double tempChi = block_sum(h_errors_chi2(p, c, delta, tid, nt), sh);
                if (!s_ok) tempChi = DBL_MAX;
                PH(7);
                if (tid == 0) {
                    const double scale = s_rho + 1e-3;
                    double r = (currentChi - tempChi) / scale;
                    s_rho = r;
                    if (r > 0 && isfinite(tempChi)) {
                        double alpha = 1. - pow((2 * r - 1), 3.0);
                        alpha = fmin(alpha, 2. / 3.);
                        s_lambda = lambda * fmax(1. / 3., alpha);
                        s_ni = 2;
                        s_accept = 1;
                    } else {
                        s_lambda = lambda * s_ni;
                        s_ni = s_ni * 2;
                        s_accept = 0;
                        for (int j = 0; j < 4; j++) p.cam_q[4 * c + j] = p.cam_bak[j];
                        for (int j = 0; j < 3; j++) p.cam_t[3 * c + j] = p.cam_bak[4 + j];
                        q_to_R(p.cam_q + 4 * c, p.cam_R + 9 * c);
                    }
                }
                __syncthreads();
                rho = s_rho;
                if (s_accept) currentChi = tempChi;
                else if (!isfinite(s_lambda)) { lambdaFinite = false; trials++; break; }
                qmax++;
                trials++;
            } while (rho < 0 && qmax < 10);
            iteration++;
            iters++;
            if (qmax == 10 || rho == 0 || !lambdaFinite) { if (tid == 0) s_stop = 1; }     // Terminate => Step() == false => break
            __syncthreads();
        }
    } else {
    bool errValid = false;              // p.err / the edge records describe the current state (true after an accepted trial)
    for (int it = 0; it < nIters; it++) {
        if (s_stop) break;
        const double delta = (double)huberW[it];
        if (!errValid) {
            f_project(p, tid, nt);
            if (p.nT) phase_tether_errors(p, tid, nt);
            __syncthreads();
        }
        double currentChi = phase_chi2(p, delta, sh);
        PH(0);
        f_build_points(p, delta, tid, nt);
        f_build_cams(p, delta, warp, nw, lane);
        if (p.nT) phase_tether_linearize(p, tid, nt);
        __syncthreads();
        phase_finish_cams(p, tid, nt);
        __syncthreads();
        if (p.nT) { phase_tether_accumulate(p, tid, nt); __syncthreads(); }
        PH(1);
        if (iteration == 0) {
            const double md = phase_max_diag(p, sh);
            if (tid == 0) { s_lambda = ctl->user_lambda_init > 0 ? ctl->user_lambda_init : 1e-5 * md; s_ni = 2; }
            __syncthreads();
        }
        double rho = 0;
        int qmax = 0;
        bool lambdaFinite = true;
        do {
            const double lambda = s_lambda;
            phase_backup(p, tid, nt);
            f_schur_points(p, lambda, tid, nt);
            __syncthreads();
            PH(2);
            f_schur_blocks(p, lambda, warp, nw, lane);
            __syncthreads();
            phase_finish_bs(p, tid, nt);
            if (p.nT) phase_tether_offdiag(p, warp, lane);
            __syncthreads();
            PH(3);
            const bool ok2 = phase_ldlt_solve(p, sh);
            __syncthreads();
            PH(4);
            if (ok2) f_backsub(p, tid, nt);
            __syncthreads();
            phase_update(p, tid, nt);
            __syncthreads();
            f_cam_R(p, tid, nt);
            __syncthreads();
            PH(6);
            f_project(p, tid, nt);
            if (p.nT) phase_tether_errors(p, tid, nt);
            __syncthreads();
            double tempChi = phase_chi2(p, delta, sh);
            if (!ok2) tempChi = DBL_MAX;
            const double scale = phase_scale(p, lambda, sh) + 1e-3;
            PH(7);
            if (tid == 0) {
                double r = (currentChi - tempChi) / scale;
                s_rho = r;
                if (r > 0 && isfinite(tempChi)) {
                    double alpha = 1. - pow((2 * r - 1), 3.0);
                    alpha = fmin(alpha, 2. / 3.);
                    s_lambda = lambda * fmax(1. / 3., alpha);
                    s_ni = 2;
                    s_accept = 1;
                } else {
                    s_lambda = lambda * s_ni;
                    s_ni = s_ni * 2;
                    s_accept = 0;
                }
            }
            __syncthreads();
            rho = s_rho;
            if (s_accept) { currentChi = tempChi; errValid = true; }
            else {
                // the records keep describing the REJECTED state (the outlier pass reads the last computed errors, like g2o);
                // the Jacobians of further trials come from W, which was built before the trial
                errValid = false;
                phase_restore(p, tid, nt);
                __syncthreads();
                f_cam_R(p, tid, nt);
                __syncthreads();
                if (!isfinite(s_lambda)) { lambdaFinite = false; trials++; break; }
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < 10);
        iteration++;
        iters++;
        if (qmax == 10 || rho == 0 || !lambdaFinite) { if (tid == 0) s_stop = 1; }     // Terminate => Step() == false => break
        __syncthreads();
    }
    }
    double errSum; int inl, nfl;
    phase_classify(p, (double)maxErrSq, sh, errSum, inl, nfl);
    if (g_cam_q) {
        for (int i = tid; i < 4 * p.K; i += nt) g_cam_q[i] = p.cam_q[i];
        for (int i = tid; i < 3 * p.K; i += nt) g_cam_t[i] = p.cam_t[i];
    }
    if (p.K <= kCtlCams) for (int i = tid; i < 7 * p.K; i += nt) ctl->cams[i] = (i % 7 < 4) ? p.cam_q[4 * (i / 7) + i % 7] : p.cam_t[3 * (i / 7) + i % 7 - 4];
    if (tid == 0) {
        ctl->lambda = s_lambda; ctl->ni = s_ni; ctl->iteration = iteration;
        ctl->err_sum = errSum; ctl->inlier_count = inl; ctl->stop_flag = s_stop; ctl->n_flagged = nfl;
        ctl->lm_iters += iters; ctl->lm_trials += trials;
        ctl->cams_valid = p.K <= kCtlCams ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------ cooperative variant
// Same algorithm, ONE problem spread over a cooperative grid (latency path of a single StepBundleAdjustment call): phases are
// partitioned over all threads of the grid and separated by grid-wide barriers; reductions go through fixed-order per-block
// partials (bit-reproducible for a given grid size); the camera state is replicated in every block's shared memory and updated
// redundantly, so no broadcast is needed; block 0 assembles and factorises the reduced system in its shared memory.
namespace cg = cooperative_groups;
constexpr int kCoopThreads = 256;
constexpr int kCoopMaxBlocks = 148;
constexpr int kCoopRedVals = 2;
constexpr int kSchurParts = 8;              // partial-product slices per reduced-system block (compile time: unrolled sums)

template <int NV>
__device__ void grid_sum(cg::grid_group& grid, double (&v)[NV], const BaDev& p, int& seq, double* sh, double* sh_out)
{
    double* slot = p.gred + (size_t)(seq & 1) * kCoopRedVals * kCoopMaxBlocks;
    seq++;
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const double s = block_sum(v[j], sh);
        if (threadIdx.x == 0) slot[j * kCoopMaxBlocks + blockIdx.x] = s;
    }
    grid.sync();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NV; j++) {
            double t = 0;
            for (unsigned bI = 0; bI < gridDim.x; bI++) t += slot[j * kCoopMaxBlocks + bI];
            sh_out[j] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; j++) v[j] = sh_out[j];
    __syncthreads();
}

// per-(block, part) partial products  sum_pairs WD_e1 W_e2^T  (one warp per item), LANE PER PAIR: a lane fetches the two 144-byte
// records of its pair with nine 16-byte loads each, multiplies the whole 6 x 6 block out of registers (108 FMAs for 18 loads) and keeps
// its own 36 sums; the 32 lanes' sums meet in a fixed order through shared memory once per item. Before, the warp shared ONE pair
// (lane = block element, twelve 8-byte loads for six FMAs, four pairs in flight per warp at 8 warps per SM): pure load latency and
// issue slots, 1.3 ms of the 4.3 ms global-BA step and 50 us of a local-window iteration.
constexpr size_t kSchurStageBytes = (size_t)(kCoopThreads / 32) * 32 * 36 * sizeof(double);      // 72 KB per CTA: [warp][lane][36]

__device__ __forceinline__ double lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }

__device__ void phase_schur_products_parts(const BaDev& p, int warp, int nwarps, int lane, uint32_t stage_cta)
{
    const int items = p.nblk * p.schur_parts, parts = p.schur_parts;
    const uint32_t red = stage_cta + (uint32_t)(threadIdx.x >> 5) * (32 * 36 * 8);
    const int2* __restrict__ pairs = p.pairs;
    const double* __restrict__ WDp = p.WD;
    const double* __restrict__ Wp = p.W;
    const int* __restrict__ blk_ptr = p.blk_ptr;
    for (int it = warp; it < items; it += nwarps) {
        const int bi = it / parts, part = it - bi * parts;
        const int beg = blk_ptr[bi], end = blk_ptr[bi + 1], per = (end - beg + parts - 1) / parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        double acc[36];
#pragma unroll
        for (int e = 0; e < 36; e++) acc[e] = 0.0;
        int k = s0 + lane;
        int2 pr = k < s1 ? pairs[k] : make_int2(0, 0);
        while (k < s1) {
            const int kn = k + 32;
            const int2 prn = kn < s1 ? pairs[kn] : make_int2(0, 0);   // the next pair's indices travel with this pair's records
            const double2* __restrict__ A2 = reinterpret_cast<const double2*>(WDp + 18 * (size_t)pr.x);
            const double2* __restrict__ B2 = reinterpret_cast<const double2*>(Wp + 18 * (size_t)pr.y);
            double a[18], b[18];
#pragma unroll
            for (int q = 0; q < 9; q++) { const double2 va = A2[q], vb = B2[q]; a[2 * q] = va.x; a[2 * q + 1] = va.y; b[2 * q] = vb.x; b[2 * q + 1] = vb.y; }
#pragma unroll
            for (int r = 0; r < 6; r++)
#pragma unroll
                for (int c = 0; c < 6; c++)
                    acc[6 * r + c] = fma(a[3 * r + 2], b[3 * c + 2], fma(a[3 * r + 1], b[3 * c + 1], fma(a[3 * r], b[3 * c], acc[6 * r + c])));
            k = kn; pr = prn;
        }
        // lane l's sums -> row l of the warp's shared tile; element e of the block = the column sum, lanes in order
#pragma unroll
        for (int e = 0; e < 36; e++) sts_f64(red + (uint32_t)(lane * 36 + e) * 8, acc[e]);
        __syncwarp();
        double s0v = 0, s1v = 0;
#pragma unroll 8
        for (int l = 0; l < 32; l++) s0v += lds_f64(red + (uint32_t)(l * 36 + lane) * 8);
        if (lane < 4) {
#pragma unroll 8
            for (int l = 0; l < 32; l++) s1v += lds_f64(red + (uint32_t)(l * 36 + 32 + lane) * 8);
        }
        double* dst = p.spart + (size_t)it * 36;
        dst[lane] = s0v;
        if (lane < 4) dst[32 + lane] = s1v;
        __syncwarp();
    }
}
// per-(camera, part) partials of coeff_i = sum_e W_e db (same slots as the single-CTA kernel)
__device__ void phase_coeff_parts(const BaDev& p, int warp, int nwarps, int lane)
{
    const int items = p.Kf * p.cam_parts;
    for (int it = warp; it < items; it += nwarps) {
        const int kf = it / p.cam_parts, part = it % p.cam_parts;
        const int beg = p.c_ptr[kf], end = p.c_ptr[kf + 1], len = end - beg;
        const int per = (len + p.cam_parts - 1) / p.cam_parts;
        const int s0 = beg + part * per, s1 = min(s0 + per, end);
        double c6[6] = {0, 0, 0, 0, 0, 0};
        for (int k = s0 + lane; k < s1; k += 32) {
            const int e = p.c_edges[k];
            const int li = p.e_l[e];
            if (li < 0) continue;
            const double* W = p.W + 18 * (size_t)e;
            const double d0 = p.db[3 * li], d1 = p.db[3 * li + 1], d2 = p.db[3 * li + 2];
#pragma unroll
            for (int r = 0; r < 6; r++) c6[r] += W[r * 3] * d0 + W[r * 3 + 1] * d1 + W[r * 3 + 2] * d2;
        }
#pragma unroll
        for (int i = 0; i < 6; i++)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c6[i] += __shfl_down_sync(0xffffffffu, c6[i], o);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 6; i++) p.part[(size_t)it * 27 + i] = c6[i];
        }
    }
}
// block 0: S = Hpp (+lambda) - sum of partial products, mirrored, in shared memory; bs = bp - coeff
__device__ void phase_assemble_reduced(const BaDev& p, double lambda)
{
    const int tid = threadIdx.x, nt = blockDim.x, n = p.n;
    for (int i = tid; i < n * n; i += nt) p.S[i] = 0.0;
    __syncthreads();
    for (int item = tid; item < p.nblk * 36; item += nt) {
        const int bi = item / 36, rc = item % 36, r = rc / 6, c = rc % 6;
        const int i1 = p.blk_ij[2 * bi], i2 = p.blk_ij[2 * bi + 1];
        double acc = 0;
        const double* __restrict__ sp = p.spart + (size_t)bi * kSchurParts * 36 + rc;
#pragma unroll
        for (int part = 0; part < kSchurParts; part++) acc += sp[part * 36];
        double v = -acc;
        if (i1 == i2) v += p.Hpp[36 * (size_t)i1 + rc] + ((r == c) ? lambda : 0.0);
        p.S[(size_t)(6 * i1 + r) * n + 6 * i2 + c] = v;
        if (i1 != i2) p.S[(size_t)(6 * i2 + c) * n + 6 * i1 + r] = v;
    }
    phase_finish_bs(p, tid, nt);
    __syncthreads();
}

// ---- large reduced systems ------------------------------------------------------------------------------------------
// S = Hpp (+lambda) - partial products, written (lower part + mirror) into the global n x n matrix zeroed by phase_schur_points
__device__ void phase_assemble_big(const BaDev& p, double lambda, int gtid, int gnt)
{
    const int n = p.n;
    for (int item = gtid; item < p.nblk * 36; item += gnt) {
        const int bi = item / 36, rc = item % 36, r = rc / 6, c = rc % 6;
        const int i1 = p.blk_ij[2 * bi], i2 = p.blk_ij[2 * bi + 1];
        double acc = 0;
        const double* __restrict__ sp = p.spart + (size_t)bi * kSchurParts * 36 + rc;
#pragma unroll
        for (int part = 0; part < kSchurParts; part++) acc += sp[part * 36];
        double v = -acc;
        if (i1 == i2) v += p.Hpp[36 * (size_t)i1 + rc] + ((r == c) ? lambda : 0.0);
        p.S[(size_t)(6 * i1 + r) * n + 6 * i2 + c] = v;
        if (i1 != i2) p.S[(size_t)(6 * i2 + c) * n + 6 * i1 + r] = v;
    }
}

__global__ void __launch_bounds__(kCoopThreads) k_ba_step_coop(const BaDev* __restrict__ prob, const float* __restrict__ huberW, int nIters, float maxErrSq,
                                                                unsigned dynBytes)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ double sh[33];
    __shared__ double sh_out[4];
    __shared__ double s_lambda, s_ni, s_rho;
    __shared__ int s_accept, s_stop;
    __shared__ BaDev s_p;
    __shared__ double *g_cam_q, *g_cam_t;
    __shared__ int s_cam_global;                                // the camera state stayed in global memory: one copy for the whole grid
    __shared__ uint32_t s_stage;                                // record stage of the Schur products (kSchurStageBytes), shared-window address
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int gtid = blockIdx.x * nt + tid, gnt = gridDim.x * nt, gwarp = gtid >> 5, gnw = gnt >> 5;
    __shared__ double s_ldlt_col[2 * 97];
    if (tid == 0) {
        s_p = prob[0];
        s_p.ldlt_col = s_ldlt_col;
        // big mode: S / bs stay in global memory, the head of dyn belongs to the dense solver (and, before it runs, to the record
        // stage of the Schur products); small mode: S and bs, then the stage. The camera state follows when it fits.
        size_t cam_off = dense::kSmemBytes;
        s_stage = (uint32_t)__cvta_generic_to_shared(dyn);
        if (!s_p.big) {
            s_p.S = reinterpret_cast<double*>(dyn); s_p.bs = s_p.S + (size_t)s_p.n * s_p.n;
            s_stage = (uint32_t)__cvta_generic_to_shared(dyn + ba_smem_need_S(s_p.n));
            cam_off = ba_smem_need_S(s_p.n) + kSchurStageBytes;
        }
        g_cam_q = s_p.cam_q; g_cam_t = s_p.cam_t;
        s_cam_global = 1;
        if (dynBytes >= cam_off + ba_smem_need_cams(s_p.K)) {
            s_cam_global = 0;   // the camera state is staged in shared memory when it fits beside the solver
            double* c = reinterpret_cast<double*>(dyn + cam_off);
            const double *gf = s_p.cam_f, *gx = s_p.cam_cx, *gy = s_p.cam_cy; const int* gh = s_p.cam_h;
            double* sq = c; double* st = c + 4 * s_p.K; double* sf = c + 7 * s_p.K; double* sx = c + 8 * s_p.K; double* sy = c + 9 * s_p.K;
            int* shh = reinterpret_cast<int*>(c + 10 * s_p.K);
            for (int k = 0; k < s_p.K; k++) {
                for (int j = 0; j < 4; j++) sq[4 * k + j] = g_cam_q[4 * k + j];
                for (int j = 0; j < 3; j++) st[3 * k + j] = g_cam_t[3 * k + j];
                sf[k] = gf[k]; sx[k] = gx[k]; sy[k] = gy[k]; shh[k] = gh[k];
            }
            s_p.cam_q = sq; s_p.cam_t = st; s_p.cam_f = sf; s_p.cam_cx = sx; s_p.cam_cy = sy; s_p.cam_h = shh;
        }
    }
    __syncthreads();
    const BaDev& p = s_p;
    BaCtl* ctl = p.ctl;
    if (tid == 0) { s_lambda = ctl->lambda; s_ni = ctl->ni; s_stop = 0; }
    __syncthreads();
    int iteration = ctl->iteration, seq = 0;
    long long trials = 0, iters = 0;
    long long t_last = (blockIdx.x == 0 && tid == 0) ? gtimer() : 0;

    // errors of this thread's edges + robust chi2 partial
    auto errors_chi2 = [&](double delta) {
        double acc = 0;
        for (int e = gtid; e < p.Ea; e += gnt) {
            const int c = p.e_cam[e];
            double xt[3];
            q_rot(p.cam_q + 4 * c, p.pt_X + 3 * (size_t)p.e_pt[e], xt);
            xt[0] += p.cam_t[3 * c]; xt[1] += p.cam_t[3 * c + 1]; xt[2] += p.cam_t[3 * c + 2];
            const double f = p.cam_f[c];
            const double e0 = p.e_uv[2 * e] - (xt[0] / xt[2] * f + p.cam_cx[c]), e1 = p.e_uv[2 * e + 1] - (xt[1] / xt[2] * f + p.cam_cy[c]);
            p.err[2 * e] = e0; p.err[2 * e + 1] = e1;
            double r0, r1;
            huber(p.e_info[e] * (e0 * e0 + e1 * e1), delta, r0, r1);
            acc += r0;
        }
        return acc;
    };

    for (int it = 0; it < nIters; it++) {
        if (s_stop) break;
        const double delta = (double)huberW[it];
        double red[2];
        red[0] = errors_chi2(delta); red[1] = 0;
        grid_sum<1>(grid, reinterpret_cast<double(&)[1]>(red[0]), p, seq, sh, sh_out);
        double currentChi = red[0];
        PH(0);
        phase_build_points(p, delta, gtid, gnt);
        phase_build_cams(p, delta, gwarp, gnw, lane);
        grid.sync();
        PH(1);
        phase_sum_points(p, gtid, gnt);
        phase_finish_cams(p, gtid, gnt);
        grid.sync();
        if (iteration == 0) {
            double m = 0;
            for (int i = gtid; i < p.Kf * 6; i += gnt) m = fmax(m, fabs(p.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]));
            for (int i = gtid; i < p.Pl * 3; i += gnt) m = fmax(m, fabs(p.Hll[9 * (size_t)(i / 3) + (i % 3) * 4]));
            const double bm = block_max(m, sh);
            double* slot = p.gred + (size_t)(seq & 1) * kCoopRedVals * kCoopMaxBlocks; seq++;
            if (tid == 0) slot[blockIdx.x] = bm;
            grid.sync();
            if (tid == 0) {
                double md = 0;
                for (unsigned bI = 0; bI < gridDim.x; bI++) md = fmax(md, slot[bI]);
                s_lambda = ctl->user_lambda_init > 0 ? ctl->user_lambda_init : 1e-5 * md; s_ni = 2;
            }
            __syncthreads();
        }
        double rho = 0;
        int qmax = 0;
        bool lambdaFinite = true;
        do {
            const double lambda = s_lambda;
            // push: points partitioned over the grid, cameras per block (replicated state)
            for (int i = tid; i < p.Kf; i += nt) {
                const int c = p.c_cam[i];
                for (int j = 0; j < 4; j++) p.cam_bak[7 * i + j] = p.cam_q[4 * c + j];      // every block writes the same values
                for (int j = 0; j < 3; j++) p.cam_bak[7 * i + 4 + j] = p.cam_t[3 * c + j];
            }
            for (int i = gtid; i < p.Pl * 3; i += gnt) p.pt_bak[i] = p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3];
            phase_schur_points(p, lambda, gtid, gnt);          // also zeroes p.S (block-local shared memory)
            grid.sync();
            PH(2);
            phase_schur_products_parts(p, gwarp, gnw, lane, s_stage);
            phase_coeff_parts(p, gwarp, gnw, lane);
            grid.sync();
            PH(3);
            if (!p.big) {
                if (blockIdx.x == 0) {
                    phase_assemble_reduced(p, lambda);
                    PH(8);
                    const bool ok = phase_ldlt_solve(p, sh);
                    if (tid == 0) ctl->last_ok = ok ? 1 : 0;
                }
            } else {
                phase_assemble_big(p, lambda, gtid, gnt);
                if (blockIdx.x == 0) { phase_finish_bs(p, tid, nt); if (tid == 0) ctl->last_ok = 1; }
                grid.sync();
                PH(8);
                const dense::Scratch dsc = {p.Zq, p.Ez, p.Ldiag, p.dflag, p.bs, ctl->phase_ns + 9};
                dense::ldlt_grid(grid, p.S, p.n, dsc, dyn, &ctl->last_ok);           // factorisation + forward substitution of bs
                grid.sync();
                if (*reinterpret_cast<volatile int*>(&ctl->last_ok)) {                                  // uniform: written before the grid barrier
                    dense::solve_back_grid(grid, p.S, p.n, dsc, dyn);
                    for (int i = gtid; i < p.n; i += gnt) p.x[i] = p.bs[i];
                }
            }
            PH(4);
            grid.sync();
            PH(5);
            const bool ok2 = *reinterpret_cast<volatile int*>(&ctl->last_ok) != 0;
            if (ok2) phase_backsub(p, gtid, gnt);
            __syncthreads();
            // update: a landmark's increment was written by the same thread that applies it; cameras are replicated
            for (int li = gtid; li < p.Pl; li += gnt)
                for (int r = 0; r < 3; r++) p.pt_X[3 * (size_t)p.l_pt[li] + r] += p.x[p.n + 3 * li + r];
            if (s_cam_global) { for (int i = gtid; i < p.Kf; i += gnt) { const int c = p.c_cam[i]; pose_oplus(p.cam_q + 4 * c, p.cam_t + 3 * c, p.x + 6 * i); } }
            else for (int i = tid; i < p.Kf; i += nt) { const int c = p.c_cam[i]; pose_oplus(p.cam_q + 4 * c, p.cam_t + 3 * c, p.x + 6 * i); }
            grid.sync();
            PH(6);
            red[0] = errors_chi2(delta);
            red[1] = 0;
            {
                const int tot = p.n + 3 * p.Pl;
                for (int j = gtid; j < tot; j += gnt) {
                    const double xj = p.x[j], bj = (j < p.n) ? p.bp[j] : p.bl[j - p.n];
                    red[1] += xj * (lambda * xj + bj);
                }
            }
            grid_sum<2>(grid, red, p, seq, sh, sh_out);
            PH(7);
            double tempChi = red[0];
            if (!ok2) tempChi = DBL_MAX;
            const double scale = red[1] + 1e-3;
            if (tid == 0) {
                double r = (currentChi - tempChi) / scale;
                s_rho = r;
                if (r > 0 && isfinite(tempChi)) {
                    double alpha = 1. - pow((2 * r - 1), 3.0);
                    alpha = fmin(alpha, 2. / 3.);
                    s_lambda = lambda * fmax(1. / 3., alpha);
                    s_ni = 2;
                    s_accept = 1;
                } else {
                    s_lambda = lambda * s_ni;
                    s_ni = s_ni * 2;
                    s_accept = 0;
                }
            }
            __syncthreads();
            rho = s_rho;
            if (s_accept) currentChi = tempChi;
            else {
                for (int i = tid; i < p.Kf; i += nt) {
                    const int c = p.c_cam[i];
                    for (int j = 0; j < 4; j++) p.cam_q[4 * c + j] = p.cam_bak[7 * i + j];
                    for (int j = 0; j < 3; j++) p.cam_t[3 * c + j] = p.cam_bak[7 * i + 4 + j];
                }
                for (int i = gtid; i < p.Pl * 3; i += gnt) p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3] = p.pt_bak[i];
                grid.sync();
                if (!isfinite(s_lambda)) { lambdaFinite = false; trials++; break; }
            }
            qmax++;
            trials++;
        } while (rho < 0 && qmax < 10);
        iteration++;
        iters++;
        if (qmax == 10 || rho == 0 || !lambdaFinite) { if (tid == 0) s_stop = 1; }
        __syncthreads();
    }
    // outlier classification (ref BundlerLib.cpp:385-427), partitioned over the grid
    {
        double red[2] = {0, 0};
        for (int e = gtid; e < p.Ea; e += gnt) {
            const double e0 = p.err[2 * e], e1 = p.err[2 * e + 1];
            const double ss = e0 * e0 + e1 * e1;
            const int c = p.e_cam[e];
            const double qc[4] = {-p.cam_q[4 * c], -p.cam_q[4 * c + 1], -p.cam_q[4 * c + 2], p.cam_q[4 * c + 3]};
            const double nt3[3] = {-p.cam_t[3 * c], -p.cam_t[3 * c + 1], -p.cam_t[3 * c + 2]};
            double wt[3], fw[3];
            const double z[3] = {0, 0, 1};
            q_rot(qc, nt3, wt);
            q_rot(qc, z, fw);
            const double* X = p.pt_X + 3 * (size_t)p.e_pt[e];
            const double dot = (X[0] - wt[0]) * fw[0] + (X[1] - wt[1]) * fw[1] + (X[2] - wt[2]) * fw[2];
            const bool out = (dot <= 0) || (ss > (double)maxErrSq);
            p.flags[e] = out ? 1 : 0;
            if (!out) { red[0] += ss; red[1] += 1.0; }
        }
        grid_sum<2>(grid, red, p, seq, sh, sh_out);
        if (blockIdx.x == 0) {
            if (!s_cam_global) {
                for (int i = tid; i < 4 * p.K; i += nt) g_cam_q[i] = p.cam_q[i];
                for (int i = tid; i < 3 * p.K; i += nt) g_cam_t[i] = p.cam_t[i];
            }
            if (p.K <= kCtlCams) for (int i = tid; i < 7 * p.K; i += nt) ctl->cams[i] = (i % 7 < 4) ? p.cam_q[4 * (i / 7) + i % 7] : p.cam_t[3 * (i / 7) + i % 7 - 4];
            if (tid == 0) {
                ctl->lambda = s_lambda; ctl->ni = s_ni; ctl->iteration = iteration;
                ctl->err_sum = red[0]; ctl->inlier_count = (int)red[1]; ctl->stop_flag = s_stop; ctl->n_flagged = p.Ea - (int)red[1];
                ctl->lm_iters += iters; ctl->lm_trials += trials;
                ctl->cams_valid = p.K <= kCtlCams ? 1 : 0;
            }
        }
    }
}

// copies every problem's control block into one contiguous table (one D2H copy for a whole batch)
// (a warp per problem, the block copied as 8-byte words)
__global__ void k_ba_gather_ctl(const BaDev* __restrict__ probs, int n, BaCtl* __restrict__ out)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(probs[i].ctl);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(out) + kCtlHeadBytes * i);      // heads only, packed
    for (int w = lane; w < (int)(kCtlHeadBytes / 8); w += 32) dst[w] = src[w];
}

} // namespace mage

// ====================================================================================================================
// Host side
// ====================================================================================================================
using namespace mage;

namespace {

template <class T> void R_to_q_host(const T* m, T* q)          // Eigen quaternion-from-matrix
{
    T t = m[0] + m[4] + m[8];
    if (t > T(0)) {
        t = std::sqrt(t + T(1)); q[3] = T(0.5) * t; t = T(0.5) / t;
        q[0] = (m[7] - m[5]) * t; q[1] = (m[2] - m[6]) * t; q[2] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + T(1));
        q[i] = T(0.5) * t; t = T(0.5) / t;
        q[3] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}

struct HostObs { double u, v, info; int cam, pt; bool set, removed; long seq; };
// tether edge: type 0 fixed distance, 1 relative rotation, 2 relative transform; m = measurement (+ weight in m[7])
struct HostTether { int type = -1, c1 = -1, c2 = -1; double m[8] = {0, 0, 0, 0, 0, 0, 0, 0}; bool set = false; long seq = -1; };

} // namespace

struct mage_ba_s {
    bool points_fixed = false;
    int K = 0, P = 0, E = 0;
    // host mirrors of the problem definition (state is authoritative on the device once uploaded)
    std::vector<double> cam_q, cam_t, cam_f, cam_cx, cam_cy, pt_X;
    std::vector<char> cam_fixed, cam_set, pt_set;
    std::vector<HostObs> obs;
    std::vector<HostTether> teth[3];    // the three constraint pools of BundlerLib.h:41-48
    long next_seq = 0;
    bool dirty = true, useless = false, state_uploaded = false, host_state_valid = true;
    bool host_cams_valid = false;      // the camera part of the host mirror came back with the control block of the last step
    bool defer_sync = false;           // mage_ba_step: uploads, kernel and read-back share h->stream, so nothing waits for the uploads on the host
    std::vector<PinnedStage> pending;  // pinned staging buffers of copies still in flight on h->stream (returned to the pool after its next synchronisation)
    float* stage_huber = nullptr;      // room for the Huber widths in the newest pending staging buffer
    double user_lambda_init = 0, lambda = -1;
    int iteration_reset = 1;           // SetCurrentLambda / InitializeOptimization reset m_iteration to 0
    std::vector<int> active;           // active observation ids in insertion order
    // device
    DeviceArena state, work;
    BaDev dev{};
    BaDev* d_dev = nullptr;
    BaCtl* d_ctl = nullptr;
    float* d_huber = nullptr; int huber_cap = 0;
    BaDev* d_table = nullptr; int table_cap = 0;      // descriptor table of mage_ba_step_many (owned by the lead handle)
    BaCtl* d_ctl_table = nullptr; int ctl_cap = 0;    // gathered control blocks of a batch
    std::vector<BaCtl> h_ctl_table;
    cudaStream_t stream = nullptr;
    int64_t stats[4] = {0, 0, 0, 0};
    BaCtl h_ctl{};                      // host copies of the last call's control block / outlier flags
    std::vector<unsigned char> h_flags;
    std::vector<unsigned> last_outliers;     // observation indices removed by the last step (also after mage_ba_step_many)
    int coop_blocks = 0;               // > 0: cooperative launch available, grid size to use
    int coop_blocks_max = 0;           // co-resident limit (large problems use the whole chip)
};

// after h->stream has been synchronised: the copies that read the handle's pending staging buffers are done
static void ba_release_pending(mage_ba_t h)
{
    for (auto& st : h->pending) stage_release(st);
    h->pending.clear();
    h->stage_huber = nullptr;
}

// dynamic shared memory for one problem: reduced system + camera state when they fit in 200 KB, else whatever subset fits
static size_t ba_dyn_smem(int n, int K, int fast)
{
    const size_t limit = 200 * 1024;
    size_t need = 0;
    if (fast) need = kFastRegionBytes;
    else if (n > 0 && ba_smem_need_S(n) <= limit) need = ba_smem_need_S(n);
    if (K <= kBaMaxSmemCams && need + ba_smem_need_cams_R(K) <= limit) need += ba_smem_need_cams_R(K);
    return need;
}

static int ba_upload_state(mage_ba_t h)
{
    // camera / point state, intrinsics and the control block live in one arena sized at first upload (pools are allocated once,
    // ref :198-230) and go up in ONE copy from a pinned staging buffer laid out like the arena (seven copies from pageable vectors cost
    // 40 us of driver calls per new instance -- a third of a pose-only call)
    DeviceArena& A = h->state;
    MAGE_REQUIRE(!A.base, MAGE_ERR_INVALID, "ba_upload_state: state already uploaded");
    A.pooled = true;
    const size_t oq = A.reserve(sizeof(double) * 4 * h->K), ot = A.reserve(sizeof(double) * 3 * h->K);
    const size_t of = A.reserve(sizeof(double) * h->K), ox = A.reserve(sizeof(double) * h->K), oy = A.reserve(sizeof(double) * h->K);
    const size_t op = A.reserve(sizeof(double) * 3 * h->P);
    const size_t oc = A.reserve(sizeof(BaCtl)), up_end = A.reserve(0, 256), od = A.reserve(sizeof(BaDev));
    MAGE_CUDA_TRY(A.commit());
    h->dev.cam_q = A.at<double>(oq); h->dev.cam_t = A.at<double>(ot);
    h->dev.cam_f = A.at<double>(of); h->dev.cam_cx = A.at<double>(ox); h->dev.cam_cy = A.at<double>(oy);
    h->dev.pt_X = A.at<double>(op);
    h->d_ctl = A.at<BaCtl>(oc); h->d_dev = A.at<BaDev>(od);
    PinnedStage stage = stage_acquire(up_end);
    MAGE_REQUIRE(stage.p, MAGE_ERR_CUDA, "ba_upload_state: no pinned staging memory");
    memset(stage.p, 0, up_end);
    memcpy(stage.p + oq, h->cam_q.data(), sizeof(double) * 4 * h->K); memcpy(stage.p + ot, h->cam_t.data(), sizeof(double) * 3 * h->K);
    memcpy(stage.p + of, h->cam_f.data(), sizeof(double) * h->K); memcpy(stage.p + ox, h->cam_cx.data(), sizeof(double) * h->K);
    memcpy(stage.p + oy, h->cam_cy.data(), sizeof(double) * h->K);
    memcpy(stage.p + op, h->pt_X.data(), sizeof(double) * 3 * h->P);
    BaCtl c{}; c.lambda = -1; c.ni = 2;
    c.user_lambda_init = h->user_lambda_init;                 // iteration 0 and the user lambda of a new instance go up with the block itself
    h->iteration_reset = 0;
    memcpy(stage.p + oc, &c, sizeof(c));
    h->pending.push_back(stage);
    MAGE_CUDA_TRY(cudaMemcpyAsync(A.base, stage.p, up_end, cudaMemcpyHostToDevice, h->stream));
    h->state_uploaded = true;
    return MAGE_OK;
}

// ref sparse_optimizer.cpp:208-272 initializeOptimization + :168-192 buildIndexMapping + block_solver.hpp:103-256
// buildStructure, done on the host whenever the edge set changed (BundlerLib.cpp:156-166); uploads the index structure.
static int ba_build_structure(mage_ba_t h)
{
    const bool timing = getenv("MAGE_BA_TIMING") != nullptr;
    auto tnow = [] { return std::chrono::steady_clock::now(); };
    auto t_begin = tnow();
    auto mark = [&](const char* what) { if (timing) { auto t = tnow(); fprintf(stderr, "[ba_build_structure] %-22s %8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_begin).count()); t_begin = t; } };
    h->active.clear();
    std::vector<std::pair<long, int>> order;
    for (int e = 0; e < h->E; e++) {
        const HostObs& o = h->obs[e];
        if (!o.set || o.removed) continue;
        if (h->points_fixed && h->cam_fixed[o.cam]) continue;               // allVerticesFixed
        order.push_back({o.seq, e});
    }
    if (!std::is_sorted(order.begin(), order.end())) std::sort(order.begin(), order.end());      // EdgeIDCompare = insertion order (callers usually set the observations in that order already)
    for (auto& pr : order) h->active.push_back(pr.second);
    const int Ea = (int)h->active.size();
    std::vector<char> camA(h->K, 0), ptA(h->P, 0);
    for (int e : h->active) { camA[h->obs[e].cam] = 1; ptA[h->obs[e].pt] = 1; }
    // tether edges: active unless both cameras are fixed (allVerticesFixed); they make their cameras active vertices
    std::vector<const HostTether*> tact;
    for (auto& pool : h->teth) for (const HostTether& t : pool) {
        if (!t.set || (h->cam_fixed[t.c1] && h->cam_fixed[t.c2])) continue;
        tact.push_back(&t); camA[t.c1] = 1; camA[t.c2] = 1;
    }
    std::sort(tact.begin(), tact.end(), [](const HostTether* a, const HostTether* b) { return a->seq < b->seq; });
    const int nT = (int)tact.size();
    std::vector<int> cam_h(h->K, -1), pt_l(h->P, -1), c_cam, l_pt;
    for (int k = 0; k < h->K; k++) if (camA[k] && !h->cam_fixed[k]) { cam_h[k] = (int)c_cam.size(); c_cam.push_back(k); }
    std::vector<int> bal_batch_ptr;                 // landmark boundaries of the balanced batches (fast path), empty otherwise
    if (!h->points_fixed) {     // point vertex ids count down (ref BundlerLib.cpp:210-218): Hessian order = descending index
        for (int i = h->P - 1; i >= 0; i--) if (ptA[i]) l_pt.push_back(i);
        const int Kf0 = (int)c_cam.size();
        if (Kf0 >= 1 && Kf0 <= kFastMaxKf && !l_pt.empty() && !getenv("MAGE_BA_NO_FAST") && getenv("MAGE_BA_BALANCE")) {
            // Opt-in (MAGE_BA_BALANCE=1): measured +0.4 % on the batched kernel for ~0.7 ms of host time per structure build, which a
            // one-shot local-BA window never earns back. The landmark order is free (it only fixes summation orders), so landmarks are dealt into the
            // shared-memory batches such that every reduced-system block gets about the same number of (edge, edge) products in every
            // batch -- the threads of a warp own different blocks and run the longest list of the warp, batch by batch.
            std::vector<std::vector<int>> pc(h->P);
            std::vector<int> pe(h->P, 0);
            for (int e : h->active) { const HostObs& o = h->obs[e]; pe[o.pt]++; if (cam_h[o.cam] >= 0) pc[o.pt].push_back(cam_h[o.cam]); }
            long tot_e = 0; int max_e = 0;
            for (int pt : l_pt) { tot_e += pe[pt]; max_e = std::max(max_e, pe[pt]); }
            if (max_e <= kFastBE) {
                int nbat = (int)std::max<long>(1, std::max((tot_e + kFastBE - 1) / kFastBE, ((long)l_pt.size() + kFastBL - 1) / kFastBL));
                std::vector<std::vector<int>> load(nbat, std::vector<int>(Kf0 * Kf0, 0)), members(nbat);
                std::vector<int> edges(nbat, 0);
                for (int pt : l_pt) {
                    int best = -1; long best_cost = 0;
                    for (int b = 0; b < (int)members.size(); b++) {
                        if (edges[b] + pe[pt] > kFastBE || (int)members[b].size() >= kFastBL) continue;
                        long cost = 0;
                        for (int i : pc[pt]) for (int j : pc[pt]) if (i <= j) cost += load[b][i * Kf0 + j];
                        cost = cost * 4096 + edges[b];
                        if (best < 0 || cost < best_cost) { best = b; best_cost = cost; }
                    }
                    if (best < 0) { best = (int)members.size(); members.emplace_back(); load.emplace_back(Kf0 * Kf0, 0); edges.push_back(0); }
                    members[best].push_back(pt); edges[best] += pe[pt];
                    for (int i : pc[pt]) for (int j : pc[pt]) if (i <= j) load[best][i * Kf0 + j]++;
                }
                l_pt.clear();
                bal_batch_ptr.push_back(0);
                for (auto& m : members) { if (m.empty()) continue; l_pt.insert(l_pt.end(), m.begin(), m.end()); bal_batch_ptr.push_back((int)l_pt.size()); }
            }
        }
        for (int i = 0; i < (int)l_pt.size(); i++) pt_l[l_pt[i]] = i;
    }
    // device edge order: grouped by landmark (so a landmark's edges are contiguous and l_edges is the identity), insertion
    // order inside a group; the reference's activeEdges order (insertion) is restored on the host when outliers are reported
    if (!h->points_fixed) {       // stable counting sort by landmark (every active edge's point is a landmark when points are free)
        std::vector<int> start(l_pt.size() + 1, 0), sorted(h->active.size());
        for (int e : h->active) start[pt_l[h->obs[e].pt] + 1]++;
        for (size_t i = 0; i < l_pt.size(); i++) start[i + 1] += start[i];
        for (int e : h->active) sorted[start[pt_l[h->obs[e].pt]]++] = e;
        h->active.swap(sorted);
    }
    mark("order + balance");
    const int Kf = (int)c_cam.size(), Pl = (int)l_pt.size(), n = 6 * Kf;
    h->useless = (Kf + Pl) == 0;
    h->dirty = false;
    h->iteration_reset = 1;
    h->stats[3]++;
    if (h->useless) { h->dev.Ea = 0; h->dev.nT = 0; return MAGE_OK; }
    std::vector<int4> t_def(nT);
    std::vector<double> t_meas(8 * (size_t)nT);
    for (int i = 0; i < nT; i++) {
        t_def[i] = make_int4(tact[i]->type, tact[i]->c1, tact[i]->c2, tact[i]->type == 2 ? 6 : 1);
        for (int k = 0; k < 8; k++) t_meas[8 * (size_t)i + k] = tact[i]->m[k];
    }

    std::vector<int> e_cam(Ea), e_pt(Ea), e_l(Ea);
    std::vector<double> e_uv(2 * (size_t)Ea), e_info(Ea);
    std::vector<int> l_ptr(Pl + 1, 0), c_ptr(Kf + 1, 0);
    for (int a = 0; a < Ea; a++) {
        const HostObs& o = h->obs[h->active[a]];
        e_cam[a] = o.cam; e_pt[a] = o.pt; e_l[a] = pt_l[o.pt];
        e_uv[2 * a] = o.u; e_uv[2 * a + 1] = o.v; e_info[a] = o.info;
        if (pt_l[o.pt] >= 0) l_ptr[pt_l[o.pt] + 1]++;
        if (cam_h[o.cam] >= 0) c_ptr[cam_h[o.cam] + 1]++;
    }
    for (int i = 0; i < Pl; i++) l_ptr[i + 1] += l_ptr[i];
    for (int i = 0; i < Kf; i++) c_ptr[i + 1] += c_ptr[i];
    std::vector<int> l_edges(l_ptr[Pl]), c_edges(c_ptr[Kf]), lf(l_ptr.begin(), l_ptr.end() - 1), cf(c_ptr.begin(), c_ptr.end() - 1);
    for (int a = 0; a < Ea; a++) {
        if (e_l[a] >= 0) l_edges[lf[e_l[a]]++] = a;
        if (cam_h[e_cam[a]] >= 0) c_edges[cf[cam_h[e_cam[a]]]++] = a;
    }
    // upper blocks of the reduced system and their (edge, edge) pair lists; every diagonal block exists. Blocks in (i1, i2)
    // lexicographic order, pairs of a block in landmark order -- two counting passes over a Kf x Kf table (a std::map keyed by the
    // block when the table would be unreasonably large)
    std::vector<int> blk_ij, blk_ptr(1, 0);
    std::vector<int2> pairs;
    std::vector<int> le_h(l_edges.size());                      // reduced-system index of the camera behind each landmark-edge slot (-1: fixed)
    for (size_t k = 0; k < l_edges.size(); k++) le_h[k] = cam_h[e_cam[l_edges[k]]];
    auto for_each_pair = [&](auto&& fn) {
        for (int li = 0; li < Pl; li++) {
            const int kb = l_ptr[li], ke = l_ptr[li + 1];
            for (int k1 = kb; k1 < ke; k1++) {
                const int i1 = le_h[k1];
                if (i1 < 0) continue;
                const int a1 = l_edges[k1];
                for (int k2 = kb; k2 < ke; k2++) {
                    const int i2 = le_h[k2];
                    if (i2 < i1) continue;                              // (also skips fixed cameras: -1 < i1)
                    fn(i1, i2, a1, l_edges[k2]);
                }
            }
        }
    };
    if ((size_t)Kf * Kf <= (size_t)1 << 22) {
        std::vector<int> cnt((size_t)Kf * Kf, 0), slot((size_t)Kf * Kf, -1);
        for_each_pair([&](int i1, int i2, int, int) { cnt[(size_t)i1 * Kf + i2]++; });
        for (int i1 = 0; i1 < Kf; i1++)
            for (int i2 = i1; i2 < Kf; i2++) {
                const size_t q = (size_t)i1 * Kf + i2;
                if (cnt[q] == 0 && i1 != i2) continue;
                slot[q] = blk_ptr.back();
                blk_ij.push_back(i1); blk_ij.push_back(i2);
                blk_ptr.push_back(blk_ptr.back() + cnt[q]);
            }
        pairs.resize((size_t)blk_ptr.back());
        for_each_pair([&](int i1, int i2, int a1, int a2) { pairs[(size_t)slot[(size_t)i1 * Kf + i2]++] = make_int2(a1, a2); });
    } else {
        std::map<std::pair<int, int>, std::vector<int2>> blocks;
        for (int i = 0; i < Kf; i++) blocks[{i, i}];
        for_each_pair([&](int i1, int i2, int a1, int a2) { blocks[{i1, i2}].push_back(make_int2(a1, a2)); });
        for (auto& kv : blocks) {
            blk_ij.push_back(kv.first.first); blk_ij.push_back(kv.first.second);
            pairs.insert(pairs.end(), kv.second.begin(), kv.second.end());
            blk_ptr.push_back((int)pairs.size());
        }
    }
    mark("csr + pair lists");
    const int nblk = (int)blk_ij.size() / 2;
    // local-window fast path (see g_schur): batches of landmarks, (block, part) items, pair lists per (batch, block)
    int fast = 0, nb = 0, nitems = 0;
    std::vector<int> batch_ptr, cam_diag(std::max(Kf, 1), 0), bb_ptr;
    std::vector<int4> item_def;
    std::vector<int2> blk_items(std::max(nblk, 1));
    std::vector<ushort2> bpairs;
    if (!h->points_fixed && Kf >= 1 && Kf <= kFastMaxKf && Pl >= 1 && nblk <= kFastItems && !getenv("MAGE_BA_NO_FAST")) {
        fast = 1;
        if (!bal_batch_ptr.empty()) batch_ptr = bal_batch_ptr;
        else {
        batch_ptr.push_back(0);
        for (int li = 0, e_in = 0, l_in = 0; li < Pl; li++) {
            const int ne = l_ptr[li + 1] - l_ptr[li];
            if (ne > kFastBE) { fast = 0; break; }
            if (e_in + ne > kFastBE || l_in == kFastBL) { batch_ptr.push_back(li); e_in = 0; l_in = 0; }
            e_in += ne; l_in++;
        }
        batch_ptr.push_back(Pl);
        }
    }
    if (fast) {
        nb = (int)batch_ptr.size() - 1;
        // parts per block: start with one each, then hand the spare thread pairs to the block with the longest share
        std::vector<int> parts(nblk, 1), cnt(nblk);
        for (int b = 0; b < nblk; b++) cnt[b] = blk_ptr[b + 1] - blk_ptr[b];
        for (int spare = kFastItems - nblk; spare > 0; spare--) {
            int best = 0;
            for (int b = 1; b < nblk; b++) if ((long)cnt[b] * parts[best] > (long)cnt[best] * parts[b]) best = b;
            parts[best]++;
        }
        for (int b = 0; b < nblk; b++) {
            blk_items[b] = make_int2((int)item_def.size(), parts[b]);
            for (int q = 0; q < parts[b]; q++) item_def.push_back(make_int4(b, q, parts[b], 0));
            if (blk_ij[2 * b] == blk_ij[2 * b + 1]) cam_diag[blk_ij[2 * b]] = b;
        }
        nitems = (int)item_def.size();
        // pairs regrouped by (batch, block); inside a group the landmark order of the block's pair list is kept
        std::vector<int> lm_batch(Pl);
        for (int b = 0; b < nb; b++) for (int li = batch_ptr[b]; li < batch_ptr[b + 1]; li++) lm_batch[li] = b;
        // counting sort of the pairs by (batch, block): sizes, exclusive scan, fill (the order inside a group is the block's pair order)
        bb_ptr.assign((size_t)nb * nblk + 1, 0);
        for (int b = 0; b < nblk; b++)
            for (int k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) bb_ptr[(size_t)lm_batch[e_l[pairs[k].x]] * nblk + b + 1]++;
        for (size_t q = 0; q < (size_t)nb * nblk; q++) bb_ptr[q + 1] += bb_ptr[q];
        bpairs.resize(pairs.size());
        std::vector<int> fill(bb_ptr.begin(), bb_ptr.end() - 1);
        for (int b = 0; b < nblk; b++)
            for (int k = blk_ptr[b]; k < blk_ptr[b + 1]; k++) {
                const int2 pr = pairs[k];
                const int bt = lm_batch[e_l[pr.x]], e0 = l_ptr[batch_ptr[bt]];
                bpairs[fill[(size_t)bt * nblk + b]++] = make_ushort2((unsigned short)(pr.x - e0), (unsigned short)(pr.y - e0));
            }
    }
    const int cam_parts = std::max(1, std::min(16, 256 / std::max(Kf, 1)));     // (camera, part) reduction items
    const int schur_parts = kSchurParts;

    mark("fast-path tables");
    DeviceArena& W = h->work;
    W.release(); W = DeviceArena(); W.pooled = true;        // the handle's calls are synchronous: nothing in flight uses the old arena
    auto rI = [&](size_t cnt) { return W.reserve(sizeof(int) * std::max<size_t>(cnt, 1)); };
    auto rD = [&](size_t cnt) { return W.reserve(sizeof(double) * std::max<size_t>(cnt, 1)); };
    size_t o_camh = rI(h->K), o_ecam = rI(Ea), o_ept = rI(Ea), o_el = rI(Ea), o_euv = rD(2 * (size_t)Ea), o_einfo = rD(Ea);
    size_t o_lpt = rI(Pl), o_lptr = rI(Pl + 1), o_ledges = rI(l_edges.size());
    size_t o_ccam = rI(Kf), o_cptr = rI(Kf + 1), o_cedges = rI(c_edges.size());
    size_t o_bij = rI(blk_ij.size()), o_bptr = rI(blk_ptr.size()), o_pairs = W.reserve(sizeof(int2) * std::max<size_t>(pairs.size(), 1));
    size_t o_bptr2 = rI(batch_ptr.size()), o_idef = W.reserve(sizeof(int4) * std::max<size_t>(item_def.size(), 1)), o_bitems = W.reserve(sizeof(int2) * blk_items.size());
    size_t o_cdiag = rI(cam_diag.size()), o_bbptr = rI(bb_ptr.size()), o_bpairs = W.reserve(sizeof(ushort2) * std::max<size_t>(bpairs.size(), 1));
    size_t o_tdef = W.reserve(sizeof(int4) * std::max(nT, 1)), o_tmeas = rD(8 * (size_t)nT);
    const size_t upload_end = W.reserve(0, 256);               // everything above is host-built and goes up in ONE copy; the work arrays follow
    size_t o_err = rD(2 * (size_t)Ea), o_W = rD(18 * (size_t)Ea), o_WD = rD(20 * (size_t)Ea), o_Hll = rD(9 * (size_t)Pl), o_bl = rD(3 * (size_t)Pl);
    size_t o_Dinv = rD(9 * (size_t)Pl), o_db = rD(3 * (size_t)Pl), o_Hpp = rD(36 * (size_t)Kf), o_bp = rD(n), o_S = rD((size_t)n * n), o_bs = rD(n);
    size_t o_x = rD(n + 3 * (size_t)Pl), o_cbak = rD(7 * (size_t)Kf), o_pbak = rD(3 * (size_t)Pl), o_part = rD((size_t)Kf * cam_parts * 27);
    size_t o_flags = W.reserve(std::max(Ea, 1));
    size_t o_Hc = rD(12 * (size_t)Ea), o_camR = rD(18 * (size_t)h->K);
    size_t o_spart = rD((size_t)nblk * schur_parts * 36), o_gred = rD(2 * (size_t)kCoopRedVals * kCoopMaxBlocks);
    const int big = ba_smem_need_S(n) > 56 * 1024 ? 1 : 0;             // reduced system too large for one CTA's shared memory
    size_t o_Zq = W.reserve(big ? dense::scratch_zq_bytes(n) : 16, 1024), o_Ez = rI(big ? dense::scratch_ez_count(n) : 1);
    size_t o_Ldiag = W.reserve(big ? dense::scratch_ldiag_bytes(n) : 16, 256), o_dflag = rI(1), o_xchg = rD(big ? 16 + 2 * (size_t)n : 1);
    size_t o_terr = rD(6 * (size_t)nT), o_tJ = rD(72 * (size_t)nT);
    MAGE_CUDA_TRY(W.commit());
    mark("cudaMalloc");
    MAGE_CUDA_TRY(cudaMemsetAsync(W.base + upload_end, 0, W.size - upload_end, h->stream));
    // the host-built tables are packed into one pinned staging buffer (same offsets as in the arena) and go up in one copy: twenty
    // separate copies from pageable vectors cost 0.2 ms of driver calls per window
    const size_t st_dev = align_up(std::max<size_t>(upload_end, 256), 256), st_huber = st_dev + align_up(sizeof(BaDev), 256);      // the problem descriptor and the Huber widths ride along
    PinnedStage stage = stage_acquire(st_huber + 64 * sizeof(float));
    MAGE_REQUIRE(stage.p, MAGE_ERR_CUDA, "ba_build_structure: no pinned staging memory");
    memset(stage.p, 0, upload_end);
    auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
        if (bytes) memcpy(stage.p + off, src, bytes);
        return cudaSuccess;
    };
    MAGE_CUDA_TRY(up(o_camh, cam_h.data(), sizeof(int) * h->K));
    MAGE_CUDA_TRY(up(o_ecam, e_cam.data(), sizeof(int) * Ea)); MAGE_CUDA_TRY(up(o_ept, e_pt.data(), sizeof(int) * Ea));
    MAGE_CUDA_TRY(up(o_el, e_l.data(), sizeof(int) * Ea));
    MAGE_CUDA_TRY(up(o_euv, e_uv.data(), sizeof(double) * 2 * Ea)); MAGE_CUDA_TRY(up(o_einfo, e_info.data(), sizeof(double) * Ea));
    MAGE_CUDA_TRY(up(o_lpt, l_pt.data(), sizeof(int) * Pl)); MAGE_CUDA_TRY(up(o_lptr, l_ptr.data(), sizeof(int) * (Pl + 1)));
    MAGE_CUDA_TRY(up(o_ledges, l_edges.data(), sizeof(int) * l_edges.size()));
    MAGE_CUDA_TRY(up(o_ccam, c_cam.data(), sizeof(int) * Kf)); MAGE_CUDA_TRY(up(o_cptr, c_ptr.data(), sizeof(int) * (Kf + 1)));
    MAGE_CUDA_TRY(up(o_cedges, c_edges.data(), sizeof(int) * c_edges.size()));
    MAGE_CUDA_TRY(up(o_bij, blk_ij.data(), sizeof(int) * blk_ij.size())); MAGE_CUDA_TRY(up(o_bptr, blk_ptr.data(), sizeof(int) * blk_ptr.size()));
    MAGE_CUDA_TRY(up(o_pairs, pairs.data(), sizeof(int2) * pairs.size()));
    MAGE_CUDA_TRY(up(o_tdef, t_def.data(), sizeof(int4) * nT)); MAGE_CUDA_TRY(up(o_tmeas, t_meas.data(), sizeof(double) * 8 * nT));
    if (fast) {
        MAGE_CUDA_TRY(up(o_bptr2, batch_ptr.data(), sizeof(int) * batch_ptr.size())); MAGE_CUDA_TRY(up(o_idef, item_def.data(), sizeof(int4) * item_def.size()));
        MAGE_CUDA_TRY(up(o_bitems, blk_items.data(), sizeof(int2) * blk_items.size())); MAGE_CUDA_TRY(up(o_cdiag, cam_diag.data(), sizeof(int) * cam_diag.size()));
        MAGE_CUDA_TRY(up(o_bbptr, bb_ptr.data(), sizeof(int) * bb_ptr.size())); MAGE_CUDA_TRY(up(o_bpairs, bpairs.data(), sizeof(ushort2) * bpairs.size()));
    }
    BaDev& d = h->dev;
    d.K = h->K; d.P = h->P; d.Ea = Ea; d.Kf = Kf; d.Pl = Pl; d.n = n; d.nblk = nblk; d.cam_parts = cam_parts;
    d.cam_h = W.at<int>(o_camh); d.e_cam = W.at<int>(o_ecam); d.e_pt = W.at<int>(o_ept); d.e_l = W.at<int>(o_el);
    d.e_uv = W.at<double>(o_euv); d.e_info = W.at<double>(o_einfo);
    d.l_pt = W.at<int>(o_lpt); d.l_ptr = W.at<int>(o_lptr); d.l_edges = W.at<int>(o_ledges);
    d.c_cam = W.at<int>(o_ccam); d.c_ptr = W.at<int>(o_cptr); d.c_edges = W.at<int>(o_cedges);
    d.blk_ij = W.at<int>(o_bij); d.blk_ptr = W.at<int>(o_bptr); d.pairs = W.at<int2>(o_pairs);
    d.err = W.at<double>(o_err); d.W = W.at<double>(o_W); d.WD = W.at<double>(o_WD); d.Hll = W.at<double>(o_Hll); d.bl = W.at<double>(o_bl);
    d.Dinv = W.at<double>(o_Dinv); d.db = W.at<double>(o_db); d.Hpp = W.at<double>(o_Hpp); d.bp = W.at<double>(o_bp); d.S = W.at<double>(o_S);
    d.bs = W.at<double>(o_bs); d.x = W.at<double>(o_x); d.cam_bak = W.at<double>(o_cbak); d.pt_bak = W.at<double>(o_pbak); d.part = W.at<double>(o_part);
    d.flags = W.at<unsigned char>(o_flags);
    d.Hc = W.at<double>(o_Hc); d.cam_R = W.at<double>(o_camR); d.cam_Rold = d.cam_R + 9 * (size_t)h->K;
    d.fast = fast; d.nb = nb; d.nitems = nitems;
    d.pose1 = (h->points_fixed && Kf == 1 && Pl == 0 && nT == 0 && !getenv("MAGE_BA_NO_POSE1")) ? 1 : 0;
    d.batch_ptr = W.at<int>(o_bptr2); d.item_def = W.at<int4>(o_idef); d.blk_items = W.at<int2>(o_bitems);
    d.cam_diag = W.at<int>(o_cdiag); d.bb_ptr = W.at<int>(o_bbptr); d.bpairs = W.at<ushort2>(o_bpairs);
    d.ctl = h->d_ctl;
    d.spart = W.at<double>(o_spart); d.gred = W.at<double>(o_gred); d.schur_parts = schur_parts;
    d.big = big; d.Zq = W.at<int8_t>(o_Zq); d.Ez = W.at<int>(o_Ez); d.Ldiag = W.at<double>(o_Ldiag); d.dflag = W.at<int>(o_dflag); d.xchg = W.at<double>(o_xchg);
    d.nT = nT; d.t_def = W.at<int4>(o_tdef); d.t_meas = W.at<double>(o_tmeas); d.t_err = W.at<double>(o_terr); d.t_J = W.at<double>(o_tJ);
    memcpy(stage.p + st_dev, &d, sizeof(BaDev));
    cudaError_t eup = upload_end ? cudaMemcpyAsync(W.base, stage.p, upload_end, cudaMemcpyHostToDevice, h->stream) : cudaSuccess;
    if (eup == cudaSuccess) eup = cudaMemcpyAsync(h->d_dev, stage.p + st_dev, sizeof(BaDev), cudaMemcpyHostToDevice, h->stream);
    mark("enqueue uploads");
    h->pending.push_back(stage);
    h->stage_huber = reinterpret_cast<float*>(stage.p + st_huber);
    // the staging buffers go back to the pool once the copies have been consumed: here, unless the caller keeps everything on h->stream
    // and synchronises it itself at the end of the call (mage_ba_step)
    if (!h->defer_sync) {
        if (eup == cudaSuccess) eup = cudaStreamSynchronize(h->stream);
        ba_release_pending(h);
    }
    MAGE_CUDA_TRY(eup);
    mark("sync");
    return MAGE_OK;
}

// kernel attributes and the co-resident grid of the cooperative (multi-CTA) variant: set up once per device, not per instance (the
// tracking thread makes a new instance per pose-only call: the four driver queries cost 25 us each time)
struct DevSetup { std::once_flag once; bool ok = false; int coop_blocks_max = 0; };
static const DevSetup& ba_device_setup()
{
    static DevSetup g_setup[64];
    int dev = 0;
    cudaGetDevice(&dev);
    DevSetup& ds = g_setup[dev & 63];
    std::call_once(ds.once, [&] {
        ds.ok = cudaFuncSetAttribute(k_ba_step_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess &&
                cudaFuncSetAttribute(k_ba_step_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
        int coop = 0, sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (coop && cudaFuncSetAttribute(k_ba_step_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCoopSmemMax) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ba_step_coop, kCoopThreads, kCoopSmemMax) == cudaSuccess && per_sm > 0)
        {
            ds.coop_blocks_max = std::min(kCoopMaxBlocks, sms * per_sm);
        }
        cudaGetLastError();
    });
    return ds;
}

extern "C" int mage_ba_create(int are_points_fixed, mage_ba_t* out)
{
    MAGE_REQUIRE(out, MAGE_ERR_INVALID, "mage_ba_create: null argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: bundle adjustment has no CPU fallback"); return MAGE_ERR_CUDA; }
    mage_ba_s* h = new mage_ba_s();
    h->points_fixed = are_points_fixed != 0;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete h; return MAGE_ERR_CUDA; }
    const DevSetup& ds = ba_device_setup();
    if (!ds.ok) { set_error("cudaFuncSetAttribute failed"); cudaStreamDestroy(h->stream); delete h; return MAGE_ERR_CUDA; }
    {
        const char* env = getenv("MAGE_BA_COOP_BLOCKS");          // read per instance: the tools sweep it inside one process
        const int want = env ? atoi(env) : 32;
        if (want > 0 && ds.coop_blocks_max > 0) { h->coop_blocks_max = ds.coop_blocks_max; h->coop_blocks = std::min(want, ds.coop_blocks_max); }
    }
    *out = h;
    return MAGE_OK;
}

extern "C" void mage_ba_destroy(mage_ba_t h)
{
    if (!h) return;
    cudaStreamSynchronize(h->stream);
    ba_release_pending(h);
    h->state.release(); h->work.release();
    if (h->d_huber) cudaFreeAsync(h->d_huber, h->stream);
    if (h->d_table) cudaFree(h->d_table);
    if (h->d_ctl_table) cudaFree(h->d_ctl_table);
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int mage_ba_alloc_cameras(mage_ba_t h, int count)
{
    MAGE_REQUIRE(h && count >= 0, MAGE_ERR_INVALID, "mage_ba_alloc_cameras: bad argument");
    MAGE_REQUIRE(h->K == 0 && !h->state_uploaded, MAGE_ERR_INVALID, "can only allocate once");      // ref :200 assert
    h->K = count;
    h->cam_q.assign(4 * (size_t)count, 0.0); h->cam_t.assign(3 * (size_t)count, 0.0);
    for (int k = 0; k < count; k++) h->cam_q[4 * k + 3] = 1.0;
    h->cam_f.assign(count, 1.0); h->cam_cx.assign(count, 0.0); h->cam_cy.assign(count, 0.0);
    h->cam_fixed.assign(count, 0); h->cam_set.assign(count, 0);
    return MAGE_OK;
}
extern "C" int mage_ba_alloc_points(mage_ba_t h, int count)
{
    MAGE_REQUIRE(h && count >= 0, MAGE_ERR_INVALID, "mage_ba_alloc_points: bad argument");
    MAGE_REQUIRE(h->P == 0 && !h->state_uploaded, MAGE_ERR_INVALID, "can only allocate once");
    h->P = count; h->pt_X.assign(3 * (size_t)count, 0.0); h->pt_set.assign(count, 0);
    return MAGE_OK;
}
extern "C" int mage_ba_alloc_observations(mage_ba_t h, int count)
{
    MAGE_REQUIRE(h && count >= 0, MAGE_ERR_INVALID, "mage_ba_alloc_observations: bad argument");
    MAGE_REQUIRE(h->E == 0, MAGE_ERR_INVALID, "can only allocate once");
    h->E = count; h->obs.assign(count, HostObs{0, 0, 0, -1, -1, false, false, -1});
    return MAGE_OK;
}

// ref BundlerLib.cpp:261-276: estimate = SE3Quat(Quaternionf(R).normalized() -> double, t -> double); f = intrinsics[2]
extern "C" int mage_ba_set_camera(mage_ba_t h, int idx, const float* pos, const float* rot, const float* intr, int is_fixed)
{
    MAGE_REQUIRE(h && pos && rot && intr && idx >= 0 && idx < h->K, MAGE_ERR_INVALID, "mage_ba_set_camera: bad argument (idx %d of %d)", idx, h ? h->K : 0);
    MAGE_REQUIRE(!h->state_uploaded, MAGE_ERR_UNSUPPORTED, "cameras cannot be re-set after the first step");
    float m[9], q[4];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r * 3 + c] = rot[c * 3 + r];
    R_to_q_host<float>(m, q);
    float nn = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (nn > 0.f) { float nrm = std::sqrt(nn); for (int i = 0; i < 4; i++) q[i] = q[i] / nrm; }
    double qd[4] = {q[0], q[1], q[2], q[3]};
    if (qd[3] < 0) for (int i = 0; i < 4; i++) qd[i] = -qd[i];
    double n = std::sqrt(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
    for (int i = 0; i < 4; i++) h->cam_q[4 * idx + i] = qd[i] / n;
    for (int i = 0; i < 3; i++) h->cam_t[3 * idx + i] = pos[i];
    h->cam_f[idx] = intr[2]; h->cam_cx[idx] = intr[0]; h->cam_cy[idx] = intr[1];
    h->cam_fixed[idx] = is_fixed != 0; h->cam_set[idx] = 1;
    h->dirty = true;
    return MAGE_OK;
}
extern "C" int mage_ba_fix_camera(mage_ba_t h, int idx, int value)
{
    MAGE_REQUIRE(h && idx >= 0 && idx < h->K, MAGE_ERR_INVALID, "mage_ba_fix_camera: bad index");
    if ((h->cam_fixed[idx] != 0) != (value != 0)) { h->cam_fixed[idx] = value != 0; h->dirty = true; }
    return MAGE_OK;
}
extern "C" int mage_ba_set_point(mage_ba_t h, int idx, const float* xyz)
{
    MAGE_REQUIRE(h && xyz && idx >= 0 && idx < h->P, MAGE_ERR_INVALID, "mage_ba_set_point: bad argument");
    MAGE_REQUIRE(!h->state_uploaded, MAGE_ERR_UNSUPPORTED, "points cannot be re-set after the first step");
    for (int i = 0; i < 3; i++) h->pt_X[3 * (size_t)idx + i] = xyz[i];
    h->pt_set[idx] = 1; h->dirty = true;
    return MAGE_OK;
}
extern "C" int mage_ba_set_observation(mage_ba_t h, int idx, const float* uv, int cam, int pt, float info)
{
    MAGE_REQUIRE(h && uv && idx >= 0 && idx < h->E && cam >= 0 && cam < h->K && pt >= 0 && pt < h->P, MAGE_ERR_INVALID,
                 "mage_ba_set_observation: bad argument (idx %d cam %d pt %d)", idx, cam, pt);
    HostObs& o = h->obs[idx];
    o.u = uv[0]; o.v = uv[1]; o.info = info; o.cam = cam; o.pt = pt; o.set = true; o.removed = false; o.seq = h->next_seq++;
    h->dirty = true;
    return MAGE_OK;
}
extern "C" int mage_ba_set_cameras_bulk(mage_ba_t h, int n, const float* pos, const float* rot, const float* intr, const int32_t* fixed)
{
    MAGE_REQUIRE(h && pos && rot && intr && fixed && n <= h->K, MAGE_ERR_INVALID, "mage_ba_set_cameras_bulk: bad argument");
    for (int k = 0; k < n; k++) { int rc = mage_ba_set_camera(h, k, pos + 3 * k, rot + 9 * k, intr + 4 * k, fixed[k]); if (rc) return rc; }
    return MAGE_OK;
}
extern "C" int mage_ba_set_points_bulk(mage_ba_t h, int n, const float* xyz)
{
    MAGE_REQUIRE(h && xyz && n <= h->P, MAGE_ERR_INVALID, "mage_ba_set_points_bulk: bad argument");
    for (int i = 0; i < n; i++) { int rc = mage_ba_set_point(h, i, xyz + 3 * (size_t)i); if (rc) return rc; }
    return MAGE_OK;
}
extern "C" int mage_ba_set_observations_bulk(mage_ba_t h, int n, const float* uv, const int32_t* cam, const int32_t* pt, const float* info)
{
    MAGE_REQUIRE(h && uv && cam && pt && info && n <= h->E, MAGE_ERR_INVALID, "mage_ba_set_observations_bulk: bad argument");
    for (int e = 0; e < n; e++) { int rc = mage_ba_set_observation(h, e, uv + 2 * (size_t)e, cam[e], pt[e], info[e]); if (rc) return rc; }
    return MAGE_OK;
}
// ---- tether edges between two cameras (ref BundlerLib.h:41-48, BundlerLib.cpp:243-259 pools, :311-350 setters) ----------
static int ba_alloc_tethers(mage_ba_t h, int pool, int count, const char* who)
{
    MAGE_REQUIRE(h && count >= 0, MAGE_ERR_INVALID, "%s: bad argument", who);
    h->teth[pool].assign(count, HostTether());
    h->dirty = true;                     // edges set before a re-allocation are gone: the device structure is rebuilt at the next step
    return MAGE_OK;
}
extern "C" int mage_ba_alloc_fixed_distance_constraints(mage_ba_t h, int count) { return ba_alloc_tethers(h, 0, count, "mage_ba_alloc_fixed_distance_constraints"); }
extern "C" int mage_ba_alloc_relative_rotation_constraints(mage_ba_t h, int count) { return ba_alloc_tethers(h, 1, count, "mage_ba_alloc_relative_rotation_constraints"); }
extern "C" int mage_ba_alloc_relative_transform_constraints(mage_ba_t h, int count) { return ba_alloc_tethers(h, 2, count, "mage_ba_alloc_relative_transform_constraints"); }

static int ba_tether_slot(mage_ba_t h, int pool, int idx, int cam1, int cam2, HostTether** out, const char* who)
{
    MAGE_REQUIRE(h && idx >= 0 && idx < (int)h->teth[pool].size() && cam1 >= 0 && cam1 < h->K && cam2 >= 0 && cam2 < h->K && cam1 != cam2,
                 MAGE_ERR_INVALID, "%s: bad argument (idx %d cam1 %d cam2 %d)", who, idx, cam1, cam2);
    HostTether& t = h->teth[pool][idx];
    t = HostTether();
    t.type = pool; t.c1 = cam1; t.c2 = cam2; t.set = true; t.seq = h->next_seq++;
    h->dirty = true;
    *out = &t;
    return MAGE_OK;
}
// error = (distance - |t2 - t1|) * weight on the view-transform translations (ref :24-55, :311-322)
extern "C" int mage_ba_set_fixed_distance_constraint(mage_ba_t h, int idx, int cam1, int cam2, float distance, float weight)
{
    HostTether* t;
    int rc = ba_tether_slot(h, 0, idx, cam1, cam2, &t, "mage_ba_set_fixed_distance_constraint");
    if (rc) return rc;
    t->m[0] = distance; t->m[7] = weight;
    return MAGE_OK;
}
// error = angularDistance((T1^-1 T2).rotation(), deltaRotation) * weight; the quaternion is used as given (ref :57-90, :324-336)
extern "C" int mage_ba_set_relative_rotation_constraint(mage_ba_t h, int idx, int cam1, int cam2, const float* q_xyzw, float weight)
{
    MAGE_REQUIRE(q_xyzw, MAGE_ERR_INVALID, "mage_ba_set_relative_rotation_constraint: null quaternion");
    HostTether* t;
    int rc = ba_tether_slot(h, 1, idx, cam1, cam2, &t, "mage_ba_set_relative_rotation_constraint");
    if (rc) return rc;
    for (int i = 0; i < 4; i++) t->m[i] = q_xyzw[i];
    t->m[7] = weight;
    return MAGE_OK;
}
// g2o EdgeSE3Expmap: error = log(T2^-1 C T1), C = SE3Quat(deltaRotation, deltaPosition) (normalised), information = weight I (ref :338-350)
extern "C" int mage_ba_set_relative_transform_constraint(mage_ba_t h, int idx, int cam1, int cam2, const float* delta_position,
                                                         const float* q_xyzw, float weight)
{
    MAGE_REQUIRE(q_xyzw && delta_position, MAGE_ERR_INVALID, "mage_ba_set_relative_transform_constraint: null argument");
    HostTether* t;
    int rc = ba_tether_slot(h, 2, idx, cam1, cam2, &t, "mage_ba_set_relative_transform_constraint");
    if (rc) return rc;
    double q[4] = {q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]};
    if (q[3] < 0) for (int i = 0; i < 4; i++) q[i] = -q[i];
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) t->m[i] = q[i] / n;
    for (int i = 0; i < 3; i++) t->m[4 + i] = delta_position[i];
    t->m[7] = weight;
    return MAGE_OK;
}

// ref BundlerLib.cpp:123-130 + :352-355: resets the iteration counter to 0 and sets the user lambda
extern "C" int mage_ba_set_lambda(mage_ba_t h, float l)
{
    MAGE_REQUIRE(h, MAGE_ERR_INVALID, "null handle");
    h->user_lambda_init = l; h->iteration_reset = 1;
    return MAGE_OK;
}
extern "C" int mage_ba_get_lambda(mage_ba_t h, float* l)
{
    MAGE_REQUIRE(h && l, MAGE_ERR_INVALID, "null argument");
    *l = (float)h->lambda;
    return MAGE_OK;
}

static int ba_prepare(mage_ba_t h, const float* huber, int n_iters, bool upload_huber = true)
{
    if (!h->state_uploaded) { int rc = ba_upload_state(h); if (rc) return rc; }
    if (h->dirty) { int rc = ba_build_structure(h); if (rc) return rc; }
    if (h->useless) return MAGE_OK;
    if (upload_huber && n_iters > h->huber_cap) {                      // stream-ordered pool: no cudaMalloc latency per new window
        if (h->d_huber) cudaFreeAsync(h->d_huber, h->stream);
        h->huber_cap = std::max(16, n_iters);
        MAGE_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&h->d_huber), sizeof(float) * h->huber_cap, h->stream));
    }
    if (upload_huber && n_iters) {
        const float* src = huber;
        if (h->stage_huber && n_iters <= 64) { memcpy(h->stage_huber, huber, sizeof(float) * n_iters); src = h->stage_huber; }      // pinned: no staging by the driver
        MAGE_CUDA_TRY(cudaMemcpyAsync(h->d_huber, src, sizeof(float) * n_iters, cudaMemcpyHostToDevice, h->stream));
    }
    if (h->iteration_reset) {
        // m_iteration = 0 (and the user lambda) take effect at the next solve()
        struct { double user; } u{h->user_lambda_init};
        MAGE_CUDA_TRY(cudaMemcpyAsync(&h->d_ctl->user_lambda_init, &u.user, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        int zero = 0;
        MAGE_CUDA_TRY(cudaMemcpyAsync(&h->d_ctl->iteration, &zero, sizeof(int), cudaMemcpyHostToDevice, h->stream));
        MAGE_CUDA_TRY(cudaStreamSynchronize(h->stream));
        ba_release_pending(h);
        h->iteration_reset = 0;
    }
    return MAGE_OK;
}

// read-back of one call: the control block first; the per-edge outlier flags only when the kernel flagged something
static int ba_fetch_flags(mage_ba_t h, cudaStream_t s)
{
    if (h->h_ctl.n_flagged > 0) {
        h->h_flags.assign(std::max(h->dev.Ea, 1), 0);
        MAGE_CUDA_TRY(cudaMemcpyAsync(h->h_flags.data(), h->dev.flags, h->dev.Ea, cudaMemcpyDeviceToHost, s));
        MAGE_CUDA_TRY(cudaStreamSynchronize(s));
    }
    return MAGE_OK;
}
static int ba_finish(mage_ba_t h, unsigned int* outliers, int cap, int* n_out, float* mean, bool have_ctl = false, cudaStream_t s = nullptr)
{
    *n_out = 0;
    if (h->useless) { *mean = std::numeric_limits<float>::quiet_NaN(); return MAGE_OK; }
    if (!have_ctl) {
        s = h->stream;
        MAGE_CUDA_TRY(cudaMemcpyAsync(&h->h_ctl, h->d_ctl, sizeof(BaCtl), cudaMemcpyDeviceToHost, s));
        MAGE_CUDA_TRY(cudaStreamSynchronize(s));
        ba_release_pending(h);
    }
    { int rc = ba_fetch_flags(h, s); if (rc) return rc; }
    const BaCtl& c = h->h_ctl;
    const std::vector<unsigned char>& flags = h->h_flags;
    h->lambda = c.lambda;
    h->stats[0] = c.lm_iters; h->stats[1] = c.lm_trials;
    h->host_state_valid = false;
    h->host_cams_valid = c.cams_valid != 0 && h->K <= kCtlCams;
    if (h->host_cams_valid) {
        for (int k = 0; k < h->K; k++) {
            for (int j = 0; j < 4; j++) h->cam_q[4 * k + j] = c.cams[7 * k + j];
            for (int j = 0; j < 3; j++) h->cam_t[3 * k + j] = c.cams[7 * k + 4 + j];
        }
        if (h->points_fixed) h->host_state_valid = true;      // the points never move: the host mirror is complete again
    }
    std::vector<std::pair<long, int>> flagged;                          // (insertion sequence, observation) of every removed edge
    if (c.n_flagged > 0)                                                // nothing to scan (and no flag read-back) when the kernel flagged no edge
        for (int a = 0; a < h->dev.Ea; a++)
            if (flags[a]) flagged.push_back({h->obs[h->active[a]].seq, h->active[a]});
    std::sort(flagged.begin(), flagged.end());                          // the reference walks activeEdges() in insertion order (:387)
    int m = 0;
    h->last_outliers.clear();
    for (auto& fl : flagged) {
        const int e = fl.second;
        h->obs[e].removed = true; h->dirty = true;                     // removeEdge => m_dirty (ref :112-116, :435-441)
        if (outliers && m < cap) outliers[m] = (unsigned)e;
        h->last_outliers.push_back((unsigned)e);
        m++;
    }
    *n_out = m;
    *mean = (float)(c.err_sum / (double)c.inlier_count);               // NaN when no inlier, like the reference's 0/0
    return MAGE_OK;
}

extern "C" int mage_ba_step(mage_ba_t h, const float* huber, int n_iters, float max_err_sq, unsigned int* outliers, int cap,
                            int* n_outliers, float* mean_sq_error)
{
    MAGE_REQUIRE(h && n_outliers && mean_sq_error && (huber || n_iters == 0) && n_iters >= 0, MAGE_ERR_INVALID, "mage_ba_step: bad argument");
    h->defer_sync = true;
    int rc = ba_prepare(h, huber, n_iters);
    h->defer_sync = false;
    if (rc) return rc;
    if (h->useless && !h->pending.empty()) { MAGE_CUDA_TRY(cudaStreamSynchronize(h->stream)); ba_release_pending(h); }
    if (!h->useless) {
        // dynamic shared memory: the reduced system (small problems) or the dense solver's staging (large ones), then the camera state
        // when it still fits -- otherwise the cameras stay in global memory
        const size_t coop_head = h->dev.big ? dense::kSmemBytes : ba_smem_need_S(h->dev.n) + kSchurStageBytes;
        const bool cams_fit = h->dev.K <= kBaMaxSmemCams && coop_head + ba_smem_need_cams(h->dev.K) <= kCoopSmemMax;
        const size_t coop_smem = coop_head + (cams_fit ? ba_smem_need_cams(h->dev.K) : 0);
        // tether edges are handled by the single-CTA kernel only (windows with tethers are stereo / IMU local BA, never the global size)
        // (stereo / IMU windows; the one-CTA kernel keeps reduced systems of up to 158 unknowns in its shared memory -- beyond that a window with
        // tethers is not a case the reference produces)
        const bool one_cta_fits = ba_smem_need_S(h->dev.n) <= 200 * 1024;
        MAGE_REQUIRE(!(h->dev.nT > 0 && !one_cta_fits), MAGE_ERR_UNSUPPORTED, "tether edges are not supported on problems with %d pose unknowns", h->dev.n);
        const bool use_coop = h->dev.nT == 0 && h->coop_blocks > 1 && (cams_fit || h->dev.big) && coop_smem <= kCoopSmemMax && (h->dev.Ea >= 1024 || h->dev.big);
        MAGE_REQUIRE(use_coop || !h->dev.big || one_cta_fits, MAGE_ERR_UNSUPPORTED, "reduced camera system of %d unknowns needs the cooperative kernel (not available)", h->dev.n);
        // small systems: enough CTAs that every (reduced-system block, part) item of the Schur products gets its own warp in ONE round
        // (288 items on 32 CTAs = 256 warps ran a second, nearly empty round)
        int coop_grid = h->dev.big ? h->coop_blocks_max : h->coop_blocks;
        if (!h->dev.big && !getenv("MAGE_BA_COOP_BLOCKS"))
            coop_grid = std::min(h->coop_blocks_max, std::max(coop_grid, div_up(h->dev.nblk * h->dev.schur_parts, kCoopThreads / 32)));
        ProfScope ps(PROF_BA_STEP, h->stream);
        if (use_coop) {
            const BaDev* d_dev = h->d_dev; const float* d_hub = h->d_huber; unsigned dynb = (unsigned)coop_smem;
            void* args[] = {(void*)&d_dev, (void*)&d_hub, (void*)&n_iters, (void*)&max_err_sq, (void*)&dynb};
            MAGE_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_ba_step_coop, dim3(coop_grid), dim3(kCoopThreads), args, coop_smem, h->stream));
        } else {
            const unsigned dyn1 = (unsigned)ba_dyn_smem(h->dev.n, h->dev.K, h->dev.fast);
            if (h->dev.pose1) k_ba_step_t<true><<<1, kBaThreads, dyn1, h->stream>>>(h->d_dev, h->d_huber, n_iters, max_err_sq, dyn1);
            else k_ba_step_t<false><<<1, kBaThreads, dyn1, h->stream>>>(h->d_dev, h->d_huber, n_iters, max_err_sq, dyn1);
        }
        MAGE_CUDA_TRY(cudaGetLastError());
        h->stats[2]++;
    }
    return ba_finish(h, outliers, cap, n_outliers, mean_sq_error);
}

// the observations the last StepBundleAdjustment / step_many call removed from this problem, in the reference's order
extern "C" int mage_ba_last_outliers(mage_ba_t h, unsigned int* outliers, int capacity, int* n_outliers)
{
    MAGE_REQUIRE(h && n_outliers && (outliers || capacity == 0), MAGE_ERR_INVALID, "mage_ba_last_outliers: bad argument");
    *n_outliers = (int)h->last_outliers.size();
    for (int i = 0; i < *n_outliers && i < capacity; i++) outliers[i] = h->last_outliers[i];
    return MAGE_OK;
}

extern "C" int mage_ba_last_outlier_counts(mage_ba_t* hs, int n, int* counts)
{
    MAGE_REQUIRE(hs && counts && n >= 0, MAGE_ERR_INVALID, "mage_ba_last_outlier_counts: bad argument");
    for (int i = 0; i < n; i++) counts[i] = hs[i] ? (int)hs[i]->last_outliers.size() : 0;
    return MAGE_OK;
}

extern "C" int mage_ba_step_many(mage_ba_t* hs, int n, const float* huber, int n_iters, float max_err_sq, float* means)
{
    MAGE_REQUIRE(hs && n >= 1 && means && (huber || n_iters == 0), MAGE_ERR_INVALID, "mage_ba_step_many: bad argument");
    {   // a handle listed twice would be stepped by two CTAs at once
        std::vector<mage_ba_t> seen(hs, hs + n);
        std::sort(seen.begin(), seen.end());
        MAGE_REQUIRE(std::adjacent_find(seen.begin(), seen.end()) == seen.end(), MAGE_ERR_INVALID, "mage_ba_step_many: the same handle is listed twice");
    }
    // gather the per-problem descriptors into one table and step every problem with ONE launch (grid = problems)
    std::vector<BaDev> table;
    std::vector<int> live;
    {   // fresh windows: the state upload and the structure build are host work per handle (0.8 ms for a local window) -- one host
        // thread per handle, up to eight at a time (the handles share nothing; each worker binds to the caller's device)
        std::vector<int> todo;
        for (int i = 0; i < n; i++) if (hs[i] && (hs[i]->dirty || !hs[i]->state_uploaded)) todo.push_back(i);
        // (a box usually runs one process per GPU: leave each its share of the host cores)
        int ndev = 1;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); ndev = 1; }
        const int share = std::max(2, (int)std::thread::hardware_concurrency() / ndev);
        const int nthreads = getenv("MAGE_BA_SERIAL_PREPARE") ? 1 : std::min<int>({(int)todo.size(), 8, share});
        if (nthreads > 1) {
            int dev = 0;
            cudaGetDevice(&dev);
            std::atomic<int> next{0};
            std::vector<int> rcs(todo.size(), MAGE_OK);
            std::vector<std::string> msgs(todo.size());
            std::vector<std::thread> pool;
            for (int t = 0; t < nthreads; t++)
                pool.emplace_back([&]() {
                    cudaSetDevice(dev);
                    for (int k; (k = next.fetch_add(1)) < (int)todo.size();) {
                        rcs[k] = ba_prepare(hs[todo[k]], huber, n_iters, false);
                        if (rcs[k]) msgs[k] = mage_last_error();          // the message lives in the worker's thread-local buffer
                    }
                });
            for (auto& th : pool) th.join();
            for (size_t k = 0; k < todo.size(); k++)
                if (rcs[k]) { set_error("%s", msgs[k].c_str()); return rcs[k]; }
        }
    }
    for (int i = 0; i < n; i++) {
        int rc = ba_prepare(hs[i], huber, n_iters, false);        // the Huber table is uploaded once, on the lead handle
        if (rc) return rc;
        if (!hs[i]->useless) {
            MAGE_REQUIRE(!hs[i]->dev.big, MAGE_ERR_UNSUPPORTED, "problem %d has a large reduced system: step it with mage_ba_step (cooperative kernel)", i);
            table.push_back(hs[i]->dev); live.push_back(i);
        }
    }
    mage_ba_t lead = live.empty() ? hs[0] : hs[live[0]];
    if (!live.empty()) {
        if (n_iters > lead->huber_cap) {
            if (lead->d_huber) cudaFreeAsync(lead->d_huber, lead->stream);
            lead->huber_cap = std::max(16, n_iters);
            MAGE_CUDA_TRY(cudaMallocAsync(reinterpret_cast<void**>(&lead->d_huber), sizeof(float) * lead->huber_cap, lead->stream));
        }
        if (n_iters) MAGE_CUDA_TRY(cudaMemcpyAsync(lead->d_huber, huber, sizeof(float) * n_iters, cudaMemcpyHostToDevice, lead->stream));
    }
    size_t dyn = 0;
    for (auto& d : table) dyn = std::max(dyn, ba_dyn_smem(d.n, d.K, d.fast));
    if (!table.empty()) {
        if ((int)table.size() > lead->table_cap) {
            if (lead->d_table) cudaFree(lead->d_table);
            lead->table_cap = (int)table.size();
            MAGE_CUDA_TRY(cudaMalloc(&lead->d_table, sizeof(BaDev) * table.size()));
        }
        BaDev* d_table = lead->d_table;
        cudaError_t e = cudaMemcpyAsync(d_table, table.data(), sizeof(BaDev) * table.size(), cudaMemcpyHostToDevice, lead->stream);
        if (e == cudaSuccess) {
            { ProfScope ps(PROF_BA_STEP, lead->stream); k_ba_step_t<false><<<(unsigned)table.size(), kBaThreads, dyn, lead->stream>>>(d_table, lead->d_huber, n_iters, max_err_sq, (unsigned)dyn); }
            e = cudaGetLastError();
        }
        MAGE_CUDA_TRY(e);
        lead->stats[2]++;
    }
    if (!table.empty()) {
        const int nl = (int)table.size();
        if (nl > lead->ctl_cap) {
            if (lead->d_ctl_table) cudaFree(lead->d_ctl_table);
            lead->ctl_cap = nl;
            MAGE_CUDA_TRY(cudaMalloc(&lead->d_ctl_table, sizeof(BaCtl) * nl));
        }
        lead->h_ctl_table.resize(nl);
        k_ba_gather_ctl<<<div_up(nl, 4), 128, 0, lead->stream>>>(lead->d_table, nl, lead->d_ctl_table);
        PinnedStage st = stage_acquire(kCtlHeadBytes * nl);          // pinned: the read-back is one DMA, not a staged pageable copy
        void* dst = st.p ? static_cast<void*>(st.p) : static_cast<void*>(lead->h_ctl_table.data());
        cudaError_t e = cudaMemcpyAsync(dst, lead->d_ctl_table, kCtlHeadBytes * nl, cudaMemcpyDeviceToHost, lead->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(lead->stream);
        if (e == cudaSuccess)
            for (int k = 0; k < nl; k++) {       // the camera states stay on the device: the getters fetch them on demand
                memcpy(&hs[live[k]]->h_ctl, static_cast<const uint8_t*>(dst) + kCtlHeadBytes * k, kCtlHeadBytes);
                hs[live[k]]->h_ctl.cams_valid = 0;
            }
        stage_release(st);
        MAGE_CUDA_TRY(e);
    }
    for (int i = 0; i < n; i++) {
        int nout = 0;
        int rc = ba_finish(hs[i], nullptr, 0, &nout, &means[i], true, lead->stream);
        if (rc) return rc;
    }
    return MAGE_OK;
}

static int ba_sync_host_state(mage_ba_t h)
{
    if (h->host_state_valid || !h->state_uploaded) return MAGE_OK;
    if (!h->host_cams_valid) {
        MAGE_CUDA_TRY(cudaMemcpyAsync(h->cam_q.data(), h->dev.cam_q, sizeof(double) * 4 * h->K, cudaMemcpyDeviceToHost, h->stream));
        MAGE_CUDA_TRY(cudaMemcpyAsync(h->cam_t.data(), h->dev.cam_t, sizeof(double) * 3 * h->K, cudaMemcpyDeviceToHost, h->stream));
    }
    MAGE_CUDA_TRY(cudaMemcpyAsync(h->pt_X.data(), h->dev.pt_X, sizeof(double) * 3 * h->P, cudaMemcpyDeviceToHost, h->stream));
    MAGE_CUDA_TRY(cudaStreamSynchronize(h->stream));
    ba_release_pending(h);
    h->host_state_valid = true;
    return MAGE_OK;
}

// ref BundlerLib.cpp:457-465: t -> float, normalized quaternion -> rotation matrix -> float (column-major)
extern "C" int mage_ba_get_pose(mage_ba_t h, int idx, float* pos, float* rot)
{
    MAGE_REQUIRE(h && pos && rot && idx >= 0 && idx < h->K, MAGE_ERR_INVALID, "mage_ba_get_pose: bad argument");
    int rc = ba_sync_host_state(h);
    if (rc) return rc;
    double q[4] = {h->cam_q[4 * idx], h->cam_q[4 * idx + 1], h->cam_q[4 * idx + 2], h->cam_q[4 * idx + 3]};
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] /= n;
    double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    for (int i = 0; i < 3; i++) pos[i] = (float)h->cam_t[3 * idx + i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rot[c * 3 + r] = (float)R[r * 3 + c];
    return MAGE_OK;
}
extern "C" int mage_ba_get_point(mage_ba_t h, int idx, float* xyz)
{
    MAGE_REQUIRE(h && xyz && idx >= 0 && idx < h->P, MAGE_ERR_INVALID, "mage_ba_get_point: bad argument");
    int rc = ba_sync_host_state(h);
    if (rc) return rc;
    for (int i = 0; i < 3; i++) xyz[i] = (float)h->pt_X[3 * (size_t)idx + i];
    return MAGE_OK;
}
extern "C" int mage_ba_get_poses_bulk(mage_ba_t h, float* pos, float* rot)
{
    MAGE_REQUIRE(h && pos && rot, MAGE_ERR_INVALID, "null argument");
    for (int k = 0; k < h->K; k++) { int rc = mage_ba_get_pose(h, k, pos + 3 * k, rot + 9 * k); if (rc) return rc; }
    return MAGE_OK;
}
extern "C" int mage_ba_get_points_bulk(mage_ba_t h, float* xyz)
{
    MAGE_REQUIRE(h && xyz, MAGE_ERR_INVALID, "null argument");
    for (int i = 0; i < h->P; i++) { int rc = mage_ba_get_point(h, i, xyz + 3 * (size_t)i); if (rc) return rc; }
    return MAGE_OK;
}
extern "C" int mage_ba_get_state_f64(mage_ba_t h, double* cams7, double* pts3)
{
    MAGE_REQUIRE(h && cams7 && pts3, MAGE_ERR_INVALID, "null argument");
    int rc = ba_sync_host_state(h);
    if (rc) return rc;
    for (int k = 0; k < h->K; k++) {
        for (int i = 0; i < 4; i++) cams7[7 * k + i] = h->cam_q[4 * k + i];
        for (int i = 0; i < 3; i++) cams7[7 * k + 4 + i] = h->cam_t[3 * k + i];
    }
    std::copy(h->pt_X.begin(), h->pt_X.end(), pts3);
    return MAGE_OK;
}
extern "C" int mage_ba_debug_phase_ns(mage_ba_t h, long long out[16])
{
    MAGE_REQUIRE(h && out && h->d_ctl, MAGE_ERR_INVALID, "null argument");
    BaCtl c;
    MAGE_CUDA_TRY(cudaMemcpy(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; i++) out[i] = c.phase_ns[i];
    return MAGE_OK;
}

// ---- TrackLocalMap::OptimizeCameraPose in one call (ref Tracking/TrackLocalMap.cpp:421-501). The reference builds
// a BundlerLib with ArePointsFixed for every call -- one camera, the frame's matched map points, one observation each --, runs ONE
// StepBundleAdjustment(numIterations x huberWidth, maxOutlierErrorSquared, outliers) and reads pose 0 back; it does so twice per frame on
// the tracking thread. Through the handle interface that is an instance, two arenas, a structure build and five copies per call
// (250 us for 30 us of kernel); here the whole problem is packed into ONE pinned buffer laid out like the device arena of a cached
// context, goes up in one copy, runs the same kernel as the handle path (k_ba_step_t<true>: same arithmetic, same results) and comes
// back -- control block, camera state and outlier flags -- in one copy.
namespace {
struct PoseCtx { cudaStream_t stream = nullptr; uint8_t* d = nullptr; uint8_t* h = nullptr; size_t cap = 0; int dev = -1; };
std::mutex g_pose_mu;
std::vector<PoseCtx> g_pose_free;

struct PoseLayout {
    size_t dev, ctl, flags, huber, cam, x, idx, e_cam, e_pt, pt, uv, info, upload_end, err, bak, R, total;
    explicit PoseLayout(int n)
    {
        const size_t N = (size_t)std::max(n, 1);
        size_t o = 0;
        auto take = [&](size_t bytes, size_t align = 16) { o = align_up(o, align); const size_t at = o; o += bytes; return at; };
        dev = take(sizeof(BaDev), 256);
        ctl = take(sizeof(BaCtl), 256); flags = take(N, 8);            // read back together: [ctl, flags + n)
        huber = take(64 * sizeof(float));
        cam = take(10 * sizeof(double));                               // q(4) t(3) f cx cy
        x = take(6 * sizeof(double));
        idx = take(2 * sizeof(int));                                   // cam_h[0] = 0, c_cam[0] = 0
        e_cam = take(N * sizeof(int)); e_pt = take(N * sizeof(int));
        pt = take(3 * N * sizeof(double)); uv = take(2 * N * sizeof(double)); info = take(N * sizeof(double));
        upload_end = take(0, 256);
        err = take(2 * N * sizeof(double)); bak = take(7 * sizeof(double)); R = take(18 * sizeof(double));
        total = take(0, 256);
    }
};

void pose_ctx_release(PoseCtx c)
{
    std::lock_guard<std::mutex> lk(g_pose_mu);
    if (g_pose_free.size() < 8) { g_pose_free.push_back(c); return; }
    cudaFree(c.d); cudaFreeHost(c.h); cudaStreamDestroy(c.stream);
}
bool pose_ctx_acquire(size_t bytes, PoseCtx& c)
{
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lk(g_pose_mu);
        for (size_t i = 0; i < g_pose_free.size(); i++)
            if (g_pose_free[i].dev == dev) { c = g_pose_free[i]; g_pose_free.erase(g_pose_free.begin() + i); break; }
    }
    if (!c.stream && cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) return false;
    c.dev = dev;
    if (c.cap < bytes) {
        if (c.d) cudaFree(c.d);
        if (c.h) cudaFreeHost(c.h);
        c.d = nullptr; c.h = nullptr; c.cap = 0;
        const size_t cap = std::max<size_t>(bytes + bytes / 2, 256 * 1024);
        if (cudaMalloc(&c.d, cap) != cudaSuccess || cudaHostAlloc(reinterpret_cast<void**>(&c.h), cap, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            if (c.d) cudaFree(c.d);
            if (c.h) cudaFreeHost(c.h);
            cudaStreamDestroy(c.stream);
            c = PoseCtx();
            return false;
        }
        c.cap = cap;
    }
    return true;
}
// q (x y z w, double) and t of a view transform from the float position / column-major rotation the reference hands over
// (ref BundlerLib.cpp:261-276, same steps as mage_ba_set_camera)
void pose_to_qt(const float* pos, const float* rot, double* q, double* t)
{
    float m[9], qf[4];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) m[r * 3 + c] = rot[c * 3 + r];
    R_to_q_host<float>(m, qf);
    const float nn = qf[0] * qf[0] + qf[1] * qf[1] + qf[2] * qf[2] + qf[3] * qf[3];
    if (nn > 0.f) { const float nrm = std::sqrt(nn); for (int i = 0; i < 4; i++) qf[i] = qf[i] / nrm; }
    double qd[4] = {qf[0], qf[1], qf[2], qf[3]};
    if (qd[3] < 0) for (int i = 0; i < 4; i++) qd[i] = -qd[i];
    const double n = std::sqrt(qd[0] * qd[0] + qd[1] * qd[1] + qd[2] * qd[2] + qd[3] * qd[3]);
    for (int i = 0; i < 4; i++) q[i] = qd[i] / n;
    for (int i = 0; i < 3; i++) t[i] = pos[i];
}
// ref BundlerLib.cpp:457-465 (same steps as mage_ba_get_pose)
void qt_to_pose(const double* qin, const double* t, float* pos, float* rot)
{
    double q[4] = {qin[0], qin[1], qin[2], qin[3]};
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] /= n;
    const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
    const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0], tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    const double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    for (int i = 0; i < 3; i++) pos[i] = (float)t[i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) rot[c * 3 + r] = (float)R[r * 3 + c];
}
} // namespace

extern "C" int mage_optimize_camera_pose(const float* position, const float* rotation, const float* intrinsics, int n, const float* map_points,
                                         const float* projections, const float* information, int n_iters, float huber_width,
                                         float max_outlier_err_sq, float* out_position, float* out_rotation, unsigned int* outliers, int cap,
                                         int* n_outliers, float* mean_sq_error)
{
    MAGE_REQUIRE(position && rotation && intrinsics && out_position && out_rotation && n_outliers && n >= 0 && n_iters >= 0 && n_iters <= 64 &&
                 (n == 0 || (map_points && projections && information)) && (outliers || cap == 0),
                 MAGE_ERR_INVALID, "mage_optimize_camera_pose: bad argument (n %d, n_iters %d of at most 64)", n, n_iters);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: bundle adjustment has no CPU fallback"); return MAGE_ERR_CUDA; }
    *n_outliers = 0;
    double q[4], t[3];
    pose_to_qt(position, rotation, q, t);
    if (n == 0) {                                               // no edge: the optimizer has nothing to do (the handle path returns NaN too)
        qt_to_pose(q, t, out_position, out_rotation);
        if (mean_sq_error) *mean_sq_error = std::numeric_limits<float>::quiet_NaN();
        return MAGE_OK;
    }
    MAGE_REQUIRE(ba_device_setup().ok, MAGE_ERR_CUDA, "cudaFuncSetAttribute failed");
    const PoseLayout L(n);
    PoseCtx c;
    MAGE_REQUIRE(pose_ctx_acquire(L.total, c), MAGE_ERR_CUDA, "mage_optimize_camera_pose: no memory for a problem of %d points", n);
    uint8_t* hb = c.h;
    memset(hb, 0, L.upload_end);
    BaDev d{};
    d.K = 1; d.P = n; d.Ea = n; d.Kf = 1; d.Pl = 0; d.n = 6; d.nblk = 0; d.cam_parts = 1; d.pose1 = 1;
    double* dcam = reinterpret_cast<double*>(c.d + L.cam);
    d.cam_q = dcam; d.cam_t = dcam + 4; d.cam_f = dcam + 7; d.cam_cx = dcam + 8; d.cam_cy = dcam + 9;
    d.cam_h = reinterpret_cast<int*>(c.d + L.idx); d.c_cam = reinterpret_cast<int*>(c.d + L.idx) + 1;
    d.pt_X = reinterpret_cast<double*>(c.d + L.pt);
    d.e_cam = reinterpret_cast<int*>(c.d + L.e_cam); d.e_pt = reinterpret_cast<int*>(c.d + L.e_pt);
    d.e_uv = reinterpret_cast<double*>(c.d + L.uv); d.e_info = reinterpret_cast<double*>(c.d + L.info);
    d.err = reinterpret_cast<double*>(c.d + L.err); d.x = reinterpret_cast<double*>(c.d + L.x); d.cam_bak = reinterpret_cast<double*>(c.d + L.bak);
    d.cam_R = reinterpret_cast<double*>(c.d + L.R); d.cam_Rold = d.cam_R + 9;
    d.flags = c.d + L.flags; d.ctl = reinterpret_cast<BaCtl*>(c.d + L.ctl);
    memcpy(hb + L.dev, &d, sizeof(BaDev));
    BaCtl ctl{}; ctl.lambda = -1; ctl.ni = 2;                   // a new instance: iteration 0, no user lambda
    memcpy(hb + L.ctl, &ctl, sizeof(BaCtl));
    float* hub = reinterpret_cast<float*>(hb + L.huber);
    for (int i = 0; i < n_iters; i++) hub[i] = huber_width;
    double* hc = reinterpret_cast<double*>(hb + L.cam);
    for (int i = 0; i < 4; i++) hc[i] = q[i];
    for (int i = 0; i < 3; i++) hc[4 + i] = t[i];
    hc[7] = intrinsics[2]; hc[8] = intrinsics[0]; hc[9] = intrinsics[1];
    int* hept = reinterpret_cast<int*>(hb + L.e_pt);
    double* hpt = reinterpret_cast<double*>(hb + L.pt); double* huv = reinterpret_cast<double*>(hb + L.uv); double* hin = reinterpret_cast<double*>(hb + L.info);
    for (int i = 0; i < n; i++) {
        hept[i] = i;
        hpt[3 * (size_t)i] = map_points[3 * (size_t)i]; hpt[3 * (size_t)i + 1] = map_points[3 * (size_t)i + 1]; hpt[3 * (size_t)i + 2] = map_points[3 * (size_t)i + 2];
        huv[2 * (size_t)i] = projections[2 * (size_t)i]; huv[2 * (size_t)i + 1] = projections[2 * (size_t)i + 1];
        hin[i] = information[i];
    }
    const unsigned dyn = (unsigned)ba_dyn_smem(6, 1, 0);
    cudaError_t e = cudaMemcpyAsync(c.d, hb, L.upload_end, cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) {
        ProfScope ps(PROF_BA_STEP, c.stream);
        k_ba_step_t<true><<<1, kBaThreads, dyn, c.stream>>>(reinterpret_cast<const BaDev*>(c.d + L.dev), reinterpret_cast<const float*>(c.d + L.huber), n_iters,
                                                           max_outlier_err_sq, dyn);
        e = cudaGetLastError();
    }
    const size_t back = L.flags + (size_t)n - L.ctl;            // control block (with the camera state) + the outlier flags
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb + L.ctl, c.d + L.ctl, back, cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) { set_error("mage_optimize_camera_pose: %s", cudaGetErrorString(e)); pose_ctx_release(c); return MAGE_ERR_CUDA; }
    const BaCtl* rc = reinterpret_cast<const BaCtl*>(hb + L.ctl);
    qt_to_pose(rc->cams, rc->cams + 4, out_position, out_rotation);
    int m = 0;
    if (rc->n_flagged > 0) {
        const uint8_t* fl = hb + L.flags;
        for (int i = 0; i < n; i++)
            if (fl[i]) { if (m < cap) outliers[m] = (unsigned)i; m++; }
    }
    *n_outliers = m;
    if (mean_sq_error) *mean_sq_error = (float)(rc->err_sum / (double)rc->inlier_count);
    pose_ctx_release(c);
    return MAGE_OK;
}

// ---- sharded global bundle adjustment (SURVEY 8(e) / 8(f) "next": observations partitioned by landmark over the ranks, ONE exchange step
// per lambda trial: the all-reduce of the reduced camera system). Every rank holds all cameras and its own landmarks; the LM loop of
// k_ba_step_coop is cut at the points where the ranks must agree, the host (mageslam_b200/sharded.py) runs NCCL all-reduces on the
// device buffers in between and takes the accept / reject decision like g2o does (ref optimization_algorithm_levenberg.cpp:57-174):
//   LINEARIZE  errors, robust chi2 (partial), landmark blocks, this rank's part of Hpp and bp           -> xchg: chi2, max diag(Hll), diag(Hpp_r), bp_r
//   SCHUR      backups, D^-1 (lambda), Schur products, S_r = Hpp_r - sum W D^-1 W^T (no damping yet), bs_r   -> S, bs (summed over the ranks)
//   SOLVE      S += lambda I, dense solve (every rank the same), own landmarks' increments, update, chi2 of the trial, x^T (lambda x + b) (partial)
//   RESTORE    rejected trial: state back from the backups;   FINISH  sum of squared errors and edge count for the returned mean
enum { kShardLinearize = 1, kShardSchur = 2, kShardSolve = 3, kShardRestore = 4, kShardFinish = 5 };

__device__ double shard_errors_chi2(const BaDev& p, double delta, int gtid, int gnt)
{
    double acc = 0;
    for (int e = gtid; e < p.Ea; e += gnt) {
        const int c = p.e_cam[e];
        double xt[3];
        q_rot(p.cam_q + 4 * c, p.pt_X + 3 * (size_t)p.e_pt[e], xt);
        xt[0] += p.cam_t[3 * c]; xt[1] += p.cam_t[3 * c + 1]; xt[2] += p.cam_t[3 * c + 2];
        const double f = p.cam_f[c];
        const double e0 = p.e_uv[2 * e] - (xt[0] / xt[2] * f + p.cam_cx[c]), e1 = p.e_uv[2 * e + 1] - (xt[1] / xt[2] * f + p.cam_cy[c]);
        p.err[2 * e] = e0; p.err[2 * e + 1] = e1;
        double r0, r1;
        huber(p.e_info[e] * (e0 * e0 + e1 * e1), delta, r0, r1);
        acc += r0;
    }
    return acc;
}

__global__ void __launch_bounds__(kCoopThreads) k_ba_shard_stage(const BaDev* __restrict__ prob, int stage, double delta, double lambda, int lead)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ double sh[33];
    __shared__ double sh_out[4];
    __shared__ BaDev s_p;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int gtid = blockIdx.x * nt + tid, gnt = gridDim.x * nt, gwarp = gtid >> 5, gnw = gnt >> 5;
    if (tid == 0) s_p = prob[0];                               // the camera state stays in global memory: it lives across the stage launches
    __syncthreads();
    const BaDev& p = s_p;
    BaCtl* ctl = p.ctl;
    int seq = 0;
    double red[2];
    if (stage == kShardLinearize) {
        red[0] = shard_errors_chi2(p, delta, gtid, gnt);
        grid_sum<1>(grid, reinterpret_cast<double(&)[1]>(red[0]), p, seq, sh, sh_out);
        if (gtid == 0) p.xchg[0] = red[0];
        phase_build_points(p, delta, gtid, gnt);
        phase_build_cams(p, delta, gwarp, gnw, lane);
        grid.sync();
        phase_sum_points(p, gtid, gnt);
        phase_finish_cams(p, gtid, gnt);
        grid.sync();
        double m = 0;
        for (int i = gtid; i < p.Pl * 3; i += gnt) m = fmax(m, fabs(p.Hll[9 * (size_t)(i / 3) + (i % 3) * 4]));
        const double bm = block_max(m, sh);
        double* slot = p.gred + (size_t)(seq & 1) * kCoopRedVals * kCoopMaxBlocks; seq++;
        if (tid == 0) slot[blockIdx.x] = bm;
        grid.sync();
        if (gtid == 0) { double md = 0; for (unsigned bI = 0; bI < gridDim.x; bI++) md = fmax(md, slot[bI]); p.xchg[1] = md; }
        for (int i = gtid; i < p.n; i += gnt) { p.xchg[16 + i] = p.Hpp[36 * (size_t)(i / 6) + (i % 6) * 7]; p.xchg[16 + p.n + i] = p.bp[i]; }
    } else if (stage == kShardSchur) {
        for (int i = tid; i < p.Kf; i += nt) {
            const int c = p.c_cam[i];
            for (int j = 0; j < 4; j++) p.cam_bak[7 * i + j] = p.cam_q[4 * c + j];      // every block writes the same values
            for (int j = 0; j < 3; j++) p.cam_bak[7 * i + 4 + j] = p.cam_t[3 * c + j];
        }
        for (int i = gtid; i < p.Pl * 3; i += gnt) p.pt_bak[i] = p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3];
        phase_schur_points(p, lambda, gtid, gnt);              // also zeroes p.S
        grid.sync();
        phase_schur_products_parts(p, gwarp, gnw, lane, (uint32_t)__cvta_generic_to_shared(dyn));
        phase_coeff_parts(p, gwarp, gnw, lane);
        grid.sync();
        phase_assemble_big(p, 0.0, gtid, gnt);                 // this rank's Hpp_r - products; the damping goes on after the all-reduce
        if (blockIdx.x == 0) phase_finish_bs(p, tid, nt);      // bs_r = bp_r - coeff_r
    } else if (stage == kShardSolve) {
        for (int i = gtid; i < p.n; i += gnt) p.S[(size_t)i * p.n + i] += lambda;
        if (gtid == 0) ctl->last_ok = 1;
        grid.sync();
        const dense::Scratch dsc = {p.Zq, p.Ez, p.Ldiag, p.dflag, p.bs, nullptr};
        dense::ldlt_grid(grid, p.S, p.n, dsc, dyn, &ctl->last_ok);
        grid.sync();
        const bool ok = *reinterpret_cast<volatile int*>(&ctl->last_ok) != 0;
        if (ok) {
            dense::solve_back_grid(grid, p.S, p.n, dsc, dyn);
            for (int i = gtid; i < p.n; i += gnt) p.x[i] = p.bs[i];
        }
        grid.sync();
        if (ok) phase_backsub(p, gtid, gnt);
        __syncthreads();
        for (int li = gtid; li < p.Pl; li += gnt)
            for (int r = 0; r < 3; r++) p.pt_X[3 * (size_t)p.l_pt[li] + r] += p.x[p.n + 3 * li + r];
        for (int i = gtid; i < p.Kf; i += gnt) { const int c = p.c_cam[i]; pose_oplus(p.cam_q + 4 * c, p.cam_t + 3 * c, p.x + 6 * i); }
        grid.sync();
        red[0] = shard_errors_chi2(p, delta, gtid, gnt);
        red[1] = 0;
        for (int j = gtid; j < 3 * p.Pl; j += gnt) { const double xj = p.x[p.n + j]; red[1] += xj * (lambda * xj + p.bl[j]); }
        if (lead) for (int j = gtid; j < p.n; j += gnt) { const double xj = p.x[j]; red[1] += xj * (lambda * xj + p.xchg[16 + p.n + j]); }      // bp summed over the ranks
        grid_sum<2>(grid, red, p, seq, sh, sh_out);
        if (gtid == 0) { p.xchg[0] = ok ? red[0] : DBL_MAX; p.xchg[1] = red[1]; p.xchg[2] = ok ? 1.0 : 0.0; }
    } else if (stage == kShardRestore) {
        for (int i = gtid; i < p.Kf; i += gnt) {
            const int c = p.c_cam[i];
            for (int j = 0; j < 4; j++) p.cam_q[4 * c + j] = p.cam_bak[7 * i + j];
            for (int j = 0; j < 3; j++) p.cam_t[3 * c + j] = p.cam_bak[7 * i + 4 + j];
        }
        for (int i = gtid; i < p.Pl * 3; i += gnt) p.pt_X[3 * (size_t)p.l_pt[i / 3] + i % 3] = p.pt_bak[i];
    } else if (stage == kShardFinish) {
        shard_errors_chi2(p, delta, gtid, gnt);                // the errors of the state as it stands (a step may end on a restored trial)
        grid.sync();
        red[0] = 0; red[1] = 0;
        for (int e = gtid; e < p.Ea; e += gnt) { red[0] += p.err[2 * e] * p.err[2 * e] + p.err[2 * e + 1] * p.err[2 * e + 1]; red[1] += 1.0; }
        grid_sum<2>(grid, red, p, seq, sh, sh_out);
        if (gtid == 0) { p.xchg[0] = red[0]; p.xchg[1] = red[1]; }
    }
}

// Prepares a shard (state upload, structure build) and hands out the device buffers the ranks exchange: the reduced system S (n x n),
// its right-hand side bs (n) and the small exchange vector (16 + 2 n doubles, see k_ba_shard_stage).
extern "C" int mage_ba_shard_prepare(mage_ba_t h, int* n, double** d_S, double** d_bs, double** d_xchg)
{
    MAGE_REQUIRE(h && n && d_S && d_bs && d_xchg, MAGE_ERR_INVALID, "mage_ba_shard_prepare: null argument");
    int rc = ba_prepare(h, nullptr, 0);
    if (rc) return rc;
    MAGE_REQUIRE(!h->useless && h->dev.big && h->dev.nT == 0 && !h->points_fixed, MAGE_ERR_UNSUPPORTED,
                 "sharding is for global problems (a reduced system too large for one CTA's shared memory, free points, no tether edges)");
    MAGE_REQUIRE(h->coop_blocks_max > 1, MAGE_ERR_UNSUPPORTED, "cooperative launch not available");
    static bool attr_set = false;
    if (!attr_set) { MAGE_CUDA_TRY(cudaFuncSetAttribute(k_ba_shard_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense::kSmemBytes)); attr_set = true; }
    *n = h->dev.n; *d_S = h->dev.S; *d_bs = h->dev.bs; *d_xchg = h->dev.xchg;
    return MAGE_OK;
}

extern "C" int mage_ba_shard_stage(mage_ba_t h, int stage, double delta, double lambda, int lead)
{
    MAGE_REQUIRE(h && stage >= kShardLinearize && stage <= kShardFinish && !h->dirty && h->dev.big, MAGE_ERR_INVALID, "mage_ba_shard_stage: bad argument (prepare first)");
    const BaDev* d_dev = h->d_dev;
    void* args[] = {(void*)&d_dev, (void*)&stage, (void*)&delta, (void*)&lambda, (void*)&lead};
    MAGE_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_ba_shard_stage, dim3(h->coop_blocks_max), dim3(kCoopThreads), args, dense::kSmemBytes, h->stream));
    MAGE_CUDA_TRY(cudaStreamSynchronize(h->stream));
    ba_release_pending(h);
    if (stage == kShardSolve || stage == kShardRestore) { h->host_state_valid = false; h->host_cams_valid = false; }      // the getters read the device state back
    h->stats[2]++;
    return MAGE_OK;
}

// diagnostics: a work array of the built problem by name index (0 x, 1 Hpp, 2 bp, 3 bl, 4 Hll, 5 bs, 6 S) copied to the host
extern "C" int mage_ba_debug_get_array(mage_ba_t h, int which, double* out, long long capacity, long long* count)
{
    MAGE_REQUIRE(h && out && count && !h->dirty && !h->useless, MAGE_ERR_INVALID, "mage_ba_debug_get_array: bad argument");
    const BaDev& d = h->dev;
    const double* src = nullptr; long long n = 0;
    switch (which) {
        case 0: src = d.x; n = d.n + 3LL * d.Pl; break;
        case 1: src = d.Hpp; n = 36LL * d.Kf; break;
        case 2: src = d.bp; n = d.n; break;
        case 3: src = d.bl; n = 3LL * d.Pl; break;
        case 4: src = d.Hll; n = 9LL * d.Pl; break;
        case 5: src = d.bs; n = d.n; break;
        case 6: src = d.S; n = d.big ? (long long)d.n * d.n : 0; break;
        default: break;
    }
    MAGE_REQUIRE(src && n > 0 && n <= capacity, MAGE_ERR_INVALID, "mage_ba_debug_get_array: unknown array or buffer too small (%lld needed)", n);
    MAGE_CUDA_TRY(cudaStreamSynchronize(h->stream));
    MAGE_CUDA_TRY(cudaMemcpy(out, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    *count = n;
    return MAGE_OK;
}

// ---- the dense solver on its own (test entry): factor a symmetric positive definite matrix and solve one right-hand side ------------
__global__ void __launch_bounds__(kCoopThreads) k_dense_debug(double* S, int n, double* y, dense::Scratch sc, int* ok)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char dyn[];
    dense::ldlt_grid(grid, S, n, sc, dyn, ok);
    grid.sync();
    dense::export_diag_blocks(S, n, sc);
    if (*reinterpret_cast<volatile int*>(ok)) dense::solve_back_grid(grid, S, n, sc, dyn);
}

// A (n x n, row-major, symmetric; the lower triangle is read) and b on the host -> x, optionally the factor (L strictly below the
// diagonal, D on it; the upper triangle is returned as it was passed in) and the phase timers of dense_ldlt.cuh.
extern "C" int mage_dense_debug_solve(int n, const double* A, const double* b, double* x, double* factor, int* positive, long long phase_ns[16])
{
    MAGE_REQUIRE(n >= 1 && A && b && x && positive, MAGE_ERR_INVALID, "mage_dense_debug_solve: bad argument");
    MAGE_REQUIRE(n % 2 == 0, MAGE_ERR_UNSUPPORTED, "n must be even (a reduced camera system has 6 unknowns per camera; rows are read as 16-byte pairs)");
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    MAGE_REQUIRE(coop, MAGE_ERR_UNSUPPORTED, "cooperative launch not available");
    MAGE_CUDA_TRY(cudaFuncSetAttribute(k_dense_debug, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dense::kSmemBytes));
    MAGE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_dense_debug, kCoopThreads, dense::kSmemBytes));
    MAGE_REQUIRE(per_sm >= 1, MAGE_ERR_UNSUPPORTED, "the dense solver does not fit an SM");
    DeviceArena W;
    const size_t o_S = W.reserve(sizeof(double) * (size_t)n * n), o_y = W.reserve(sizeof(double) * n), o_zq = W.reserve(dense::scratch_zq_bytes(n), 1024);
    const size_t o_ez = W.reserve(sizeof(int) * dense::scratch_ez_count(n)), o_ok = W.reserve(sizeof(int)), o_ns = W.reserve(sizeof(long long) * 16);
    const size_t o_ld = W.reserve(dense::scratch_ldiag_bytes(n), 256), o_fl = W.reserve(sizeof(int));
    MAGE_CUDA_TRY(W.commit());
    cudaError_t e = cudaMemset(W.base, 0, W.size);
    if (e == cudaSuccess) e = cudaMemcpy(W.base + o_S, A, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(W.base + o_y, b, sizeof(double) * n, cudaMemcpyHostToDevice);
    const int one = 1;
    long long kernel_ns = 0;
    if (e == cudaSuccess) e = cudaMemcpy(W.base + o_ok, &one, sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        double* dS = W.at<double>(o_S); double* dy = W.at<double>(o_y); int* dok = W.at<int>(o_ok);
        // MAGE_DENSE_NO_TIMERS=1: no in-kernel phase timers (reading %globaltimer costs the timed thread about a microsecond each time)
        dense::Scratch sc = {W.at<int8_t>(o_zq), W.at<int>(o_ez), W.at<double>(o_ld), W.at<int>(o_fl), dy, getenv("MAGE_DENSE_NO_TIMERS") ? nullptr : W.at<long long>(o_ns)};
        int nn = n;
        void* args[] = {(void*)&dS, (void*)&nn, (void*)&dy, (void*)&sc, (void*)&dok};
        cudaEvent_t ev0, ev1;
        cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        cudaEventRecord(ev0, 0);
        e = cudaLaunchCooperativeKernel((const void*)k_dense_debug, dim3(std::min(kCoopMaxBlocks, sms * per_sm)), dim3(kCoopThreads), args, dense::kSmemBytes, 0);
        cudaEventRecord(ev1, 0);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        float ms = 0;
        if (e == cudaSuccess) cudaEventElapsedTime(&ms, ev0, ev1);
        kernel_ns = (long long)(ms * 1e6);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    if (e == cudaSuccess) e = cudaMemcpy(x, W.base + o_y, sizeof(double) * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(positive, W.base + o_ok, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && factor) e = cudaMemcpy(factor, W.base + o_S, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && phase_ns) e = cudaMemcpy(phase_ns, W.base + o_ns, sizeof(long long) * 16, cudaMemcpyDeviceToHost);
    if (phase_ns) phase_ns[15] = kernel_ns;                            // the whole kernel by CUDA events
    W.release();
    if (e != cudaSuccess) { set_error("mage_dense_debug_solve: %s", cudaGetErrorString(e)); return MAGE_ERR_CUDA; }
    return MAGE_OK;
}

extern "C" int mage_ba_get_stats(mage_ba_t h, int64_t stats[4])
{
    MAGE_REQUIRE(h && stats, MAGE_ERR_INVALID, "null argument");
    for (int i = 0; i < 4; i++) stats[i] = h->stats[i];
    return MAGE_OK;
}
