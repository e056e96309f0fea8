// Grid-wide dense LDL^T of the reduced camera system with the forward substitution folded in, and the backward substitution
// (global bundle adjustment).
//
// Replaces g2o's LinearSolverDense::solve (ref Dependencies/g2o/g2o/solvers/dense/linear_solver_dense.h:65-113: Eigen::LDLT of the
// n x n reduced system, n = 6 x free cameras, 2988 for the 500-key-frame tier problem) inside the cooperative LM kernel of ba.cu.
// Right-looking blocked factorisation without pivoting (the damped system is SPD; a non-positive pivot => "not positive"), panels of
// NB = 128 columns, two grid barriers per panel:
//   (a) the 128 x 128 diagonal block is factored in shared memory (32-wide sub-blocks: one warp factors the 32 x 32 block with a row
//       per lane, one thread per row below solves its 32 entries, all threads apply the rank-32 update). The first block by every
//       CTA (same arithmetic, same result); every later one by the LAST CTA of the grid while the others run the trailing update of
//       the panel before it (look-ahead: the two update tiles that cover the block open the chunks of two workers and are counted
//       in a flag the factor CTA waits on), published to global memory and copied into shared memory by everybody;
//   (b) a warp per (up to three) trailing rows solves Y = A_ik L_kk^-T against the block in shared memory, stores L_ik = Y D^-1
//       and cuts Z = Y D^-1/2 -- the row of the Cholesky factor, so that the update below is -Z Z^T with ONE operand -- into
//       SL = 8 signed 7-bit slices per value relative to the row's largest exponent (exact: z = 2^E sum_s q_s 2^(-6-7s) + a remainder
//       below 2^(E-56)), written as int8 planes in the 128-byte-swizzled K-major layout the tensor core reads. Row n of the matrix
//       is the right-hand side of the linear system: the last row of the bordered factor is D^-1 L^-1 b, the forward substitution;
//   (c) the trailing update A_ij -= sum_m Z_im Z_jm runs on the 5th-generation tensor cores as exact integer arithmetic (the Ozaki
//       scheme): for every slice pair (s, t) with s + t = d <= 7 one chain of tcgen05.mma kind::i8 (M = 128 rows x N = 64 columns x
//       K = 32, four per pair) accumulates q_s(i) . q_t(j) in s32 in tensor memory, one accumulator per d (|sum| <= 8 * 128 * 64 * 64
//       = 2^22); the epilogue reads the eight accumulators back (tcgen05.ld), combines them in FP64 with the weights 2^(-12-7d),
//       scales by 2^(E_i + E_j) and subtracts from A in place. The products dropped (s + t >= 8) are below 2^-46 of the row maxima's
//       product -- the rounding error an FP64 dot product of 128 terms carries anyway. Operand planes come in by cp.async.bulk (the
//       TMA engine's linear mode), slice by slice, so the first MMAs start while the later slices are still in flight.
// The backward substitution takes one grid barrier per 128 unknowns (every CTA solves the diagonal triangle itself).
// FP64 CUDA cores keep everything that is not a dense contraction: the diagonal blocks, the panel solves, the substitution.
#pragma once


#include <cooperative_groups.h>

#include <cfloat>
#include <cstdint>

namespace mage {
namespace dense {

namespace cg = cooperative_groups;

constexpr int NB = 128;                       // panel width = K of the tensor-core update
constexpr int SL = 8;                         // int8 slices per operand value (7 bits each)
constexpr int TM = 128, TN = 64;              // update tile: rows x columns
constexpr int PLANE = TM * NB;                // bytes of one (row tile, slice) plane in the canonical layout
constexpr int BPLANE = TN * NB;               // bytes of the 64-row operand of one slice
constexpr int kThreads = 256;
constexpr int LTP = NB + 2;                   // row pitch (doubles) of the transposed diagonal block in shared memory: 16-byte rows
constexpr int kMinExp = 1023 - 400;           // rows whose largest |z| is below 2^-400 contribute nothing
constexpr size_t kSmemUpdate = (size_t)SL * PLANE + (size_t)SL * BPLANE + TN * sizeof(double);
constexpr size_t kSmemTrsm = sizeof(double) * ((size_t)NB * LTP + 2 * NB + 32 * 96 + 128 + 32 * 64);
constexpr int kPack = NB * LTP + 2 * NB;      // doubles of one factored diagonal block as it is published: Lt, then 1 / D, then 1 / sqrt(D)
constexpr size_t kSmemBytes = (kSmemUpdate > kSmemTrsm ? kSmemUpdate : kSmemTrsm) + 1024;     // + slack to align the base to 1 KB
// instruction descriptor (kind::i8): D = s32, A and B signed 8 bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr uint32_t kSBO = 1024;               // operand planes are K-major rows of 128 bytes under the 128-byte swizzle: stride of an 8-row group

// global scratch of one problem (allocated with it, zero-initialised)
struct Scratch {
    int8_t* Zq;       // [row tiles of 128][SL][PLANE] over n + 1 rows; rows > n are never written and stay zero
    int* Ez;          // [row tiles * 128] exponent E of the row: z = z' 2^E, |z'| < 1
    double* Ldiag;    // [panels][kPack] the factored diagonal blocks (column-major, pitch LTP) with 1 / D and 1 / sqrt(D) behind each
    int* flag;        // counts the update tiles that cover the next diagonal blocks (zero at kernel start; see ldlt_grid)
    double* ytmp;     // [n] right-hand side in, solution out (16-byte aligned; may be the caller's vector)
    long long* ns;    // optional [16] phase timers written by CTA 0: diag, panel, update, barriers, solve, ... (null: off)
};

// Rounds a pointer into the dynamic shared memory up to 1 KB by pointer arithmetic (a round trip through an integer would make
// every access behind it a generic load / store instead of LDS / STS).
__device__ __forceinline__ uint8_t* align_1k(uint8_t* p) { return p + ((1024u - ((uint32_t)__cvta_generic_to_shared(p) & 1023u)) & 1023u); }
__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ uint64_t smem_desc(uint32_t a)
{
    // start address, leading-dimension offset (unused under a swizzle: 1), stride offset, descriptor version 1, layout SWIZZLE_128B
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(kSBO >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate), "r"(0u) : "memory");
}
// the same with the descriptors given by their low words (start address, LBO); the high word is the constant of smem_desc
__device__ __forceinline__ void mma_i8_lo(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t accumulate)
{
    constexpr uint32_t hi = (uint32_t)(((uint64_t)(kSBO >> 4) << 32 | (1ull << 46) | (2ull << 61)) >> 32);
    asm volatile("{\n .reg .pred p;\n .reg .b64 da, db;\n setp.ne.b32 p, %4, 0;\n mov.b64 da, {%1, %6};\n mov.b64 db, {%2, %6};\n"
                 " tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, {%5, %5, %5, %5}, p;\n}"
                 ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(kIdesc), "r"(accumulate), "r"(0u), "r"(hi) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_inval(uint32_t bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Waits for the phase with the given parity. A protocol mistake must not hang the cooperative grid for ever: after about four
// seconds the wait gives up and raises *fail (the caller reports "solver failed"; the results of the launch are discarded).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* fail)
{
    long long t0 = 0;
    for (unsigned spin = 0;; spin++) {
        uint32_t done;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 0xFFFFu) == 0xFFFFu) {
            const long long t = gtime();
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ll) { *fail = 1; return; }
        }
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                 "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 columns of this warp's 32 accumulator rows; the registers are valid after tmem_ld16_wait (which takes them as in-out operands
// so that no use can be scheduled above the wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&v)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]),
                   "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}
// 1 / d for a normal positive d (a pivot; anything else is reported as "not positive" by the caller): hardware seed, third-order
// step and one Newton step -- the sequence __drcp_rn starts with, without its special-case tail (70 instructions per use)
__device__ __forceinline__ double rcp_pos(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
}
__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }      // -1022 <= e <= 1023

// ---- (a) diagonal block, factored by EVERY CTA redundantly (same arithmetic, same result): the panel rows below need the factored
// block in shared memory anyway, so this costs no extra traffic and saves a grid barrier plus a round trip through global memory.
// The block lives column-major in shared memory, Lt[c * LTP + r] = A(r, c) for r >= c (LTP = 130: 16-byte aligned columns, and a
// quarter-warp storing 16 bytes each at a stride of LTP doubles covers all 32 banks once). Two-level blocking, sub-blocks of 32:
//   1. warp 0 factors the 32 x 32 block on the diagonal in registers (lane = row, column entries exchanged by shuffle);
//   2. one thread per row below solves its 32 entries against that block (L entries are warp-wide broadcasts), keeps the unscaled
//      row Y in a k-major side buffer and stores L = Y D^-1;
//   3. all threads subtract Y L^T from the rest of the block in 4 x 4 register blocks.
// Rows / columns >= nb (last block of the matrix) are padded as identity.
constexpr int YP = 96;                        // rows below a 32-wide sub-panel: at most 96

__device__ __forceinline__ void load_diag_block(const double* __restrict__ A, int ld, int nb, double* __restrict__ Lt)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = warp; t < 10; t += kThreads / 32) {                // the ten 32 x 32 tiles of the lower triangle
        const int tr = t < 1 ? 0 : t < 3 ? 1 : t < 6 ? 2 : 3, tc = t - tr * (tr + 1) / 2;
        if (32 * tc >= nb && tr != tc) continue;
        const int c = 32 * tc + lane;
        double v[32];
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const int i = 32 * tr + r;
            v[r] = (i < nb && c <= i) ? __ldcg(A + (size_t)i * ld + c) : (i == c ? 1.0 : 0.0);          // coalesced along the row
        }
        double* dst = Lt + c * LTP + 32 * tr;
#pragma unroll
        for (int r = 0; r < 32; r += 2) *reinterpret_cast<double2*>(dst + r) = make_double2(v[r], v[r + 1]);
    }
}

__device__ void factor_diag_smem(double* __restrict__ Lt, double* __restrict__ invD, double* __restrict__ rsD, double* __restrict__ Yt, int nb, int* s_bad, long long* ns, bool timed_cta)
{
    double* colb = Yt + 32 * YP;                                // [2][64] column exchange of the sub-block factorisation, upper halves zero
    double* Lsub = colb + 128;                                  // [32][64] the factored 32 x 32 sub-block, column-major, upper halves zero
    if (threadIdx.x < 128) colb[threadIdx.x] = 0.0;
    for (int i = threadIdx.x; i < 32 * 64; i += kThreads) Lsub[i] = 0.0;
    __syncthreads();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool timed = ns && timed_cta && tid == 0;             // reading %globaltimer is slow and serialises across the chip: one thread only
    long long tm = timed ? gtime() : 0;
    auto lap = [&](int slot) { if (timed) { const long long t = gtime(); ns[slot] += t - tm; tm = t; } };
    for (int j0 = 0; j0 < nb; j0 += 32) {
        if (warp == 0) {                                            // 1. the 32 x 32 block on the diagonal
            // lane = row, the row's 32 entries in registers. Per column: the column goes through a double-buffered shared vector that
            // every lane reads back as broadcasts (the pivot first: its reciprocal is the critical path), then one rank-1 update.
            // The registers are a WINDOW that starts at the current column and moves on by four columns per pass of a rolled loop:
            // fully unrolled, the 32 columns are 120 KB of straight-line code and the warp spent most of its time waiting for
            // instructions (ncu: stall_no_inst 57 %). Entries shifted in behind column 31 are zeros and stay unused.
            double a[32];
#pragma unroll
            for (int c = 0; c < 32; c++) a[c] = c <= lane ? Lt[(j0 + c) * LTP + j0 + lane] : 0.0;
            bool bad = false;
            double mine = 1.0;
            // software pipeline: column k + 1 is published (and its pivot's reciprocal started) right after the first update of
            // column k has produced it, so the reciprocal's latency runs under the other 30 updates of column k
            colb[lane] = a[0];
            __syncwarp();
            double d = colb[0];
            double inv = fabs(d) > 0 ? rcp_pos(d) : 0.0;
#pragma unroll 1
            for (int kw = 0; kw < 32; kw += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int k = kw + u;
                    const double ck = a[u];
                    const double* cb = colb + (u & 1) * 64;      // [2][64]: the upper halves stay zero (reads of a window past column 31)
                    double* cbn = colb + ((u + 1) & 1) * 64;
                    double cj[32];                                // A(k + t - u, k), t > u
                    if ((u + 1) & 1) cj[u + 1] = cb[kw + u + 1];
#pragma unroll
                    for (int t = (u + 2) & ~1; t < 32; t += 2) {
                        const double2 c2 = *reinterpret_cast<const double2*>(cb + kw + t);
                        cj[t] = c2.x; cj[t + 1] = c2.y;
                    }
                    if (!(d > 0)) bad = true;
                    const double l = ck * inv;
                    a[u + 1] -= l * cj[u + 1];                     // A(i, k + 1) -= L(i, k) A(k + 1, k): the next column
                    cbn[lane] = a[u + 1];
                    __syncwarp();
                    const double dn = cbn[k + 1];                  // (past column 31: the zero half)
                    const double invn = fabs(dn) > 0 ? rcp_pos(dn) : 0.0;
#pragma unroll
                    for (int t = u + 2; t < 32; t++) a[t] -= l * cj[t];          // A(i, j) -= L(i, k) A(j, k)
                    if (lane == k) mine = d;
                    if (lane >= k) Lt[(j0 + k) * LTP + j0 + lane] = lane == k ? d : l;
                    Lsub[k * 64 + lane] = lane > k ? l : 0.0;        // compact copy for the row solves below: column k, zeros behind row 31
                    d = dn; inv = invn;
                }
#pragma unroll
                for (int t = 0; t < 28; t++) a[t] = a[t + 4];
                a[28] = 0.0; a[29] = 0.0; a[30] = 0.0; a[31] = 0.0;
            }
            invD[j0 + lane] = fabs(mine) > 0 ? 1.0 / mine : 0.0;
            rsD[j0 + lane] = mine > 0 ? 1.0 / sqrt(mine) : 0.0;
            if (bad && lane == 0) *s_bad = 1;
        }
        __syncthreads();
        lap(5);
        const int below = nb - j0 - 32;                             // rows under the sub-block (<= 96)
        if (below <= 0) break;
        const int below4 = (below + 3) & ~3;                        // padded rows are zero and stay zero
        if (tid < below4) {                                         // 2. Y L_sub^T = A  ->  Y, L = Y D^-1
            // thread = row; the same moving register window as above (rolled loop, four columns per pass)
            const int row = j0 + 32 + tid;
            double a[32];
#pragma unroll
            for (int c = 0; c < 32; c++) a[c] = Lt[(j0 + c) * LTP + row];
#pragma unroll 1
            for (int kw = 0; kw < 32; kw += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int c = kw + u;
                    const double yc = a[u];
                    const double* lc = Lsub + c * 64 + kw;              // L_sub(kw + t, c), the same address for every thread; t > u (zeros past row 31)
                    if ((u + 1) & 1) a[u + 1] -= yc * lc[u + 1];
#pragma unroll
                    for (int t = (u + 2) & ~1; t < 32; t += 2) {
                        const double2 l2 = *reinterpret_cast<const double2*>(lc + t);
                        a[t] -= yc * l2.x; a[t + 1] -= yc * l2.y;
                    }
                    Yt[c * YP + tid] = yc;
                    Lt[(j0 + c) * LTP + row] = yc * invD[j0 + c];
                }
#pragma unroll
                for (int t = 0; t < 28; t++) a[t] = a[t + 4];
                a[28] = 0.0; a[29] = 0.0; a[30] = 0.0; a[31] = 0.0;
            }
        }
        __syncthreads();
        lap(6);
        {                                                           // 3. A(i, j) -= sum_k Y(i, k) L(j, k) on the remaining lower triangle
            const int nq = below4 >> 2, nblk = nq * (nq + 1) / 2, o = j0 + 32;
            for (int idx = tid; idx < nblk; idx += kThreads) {
                int bi = (int)((sqrtf(8.f * (float)idx + 1.f) - 1.f) * 0.5f);
                while (bi * (bi + 1) / 2 > idx) bi--;
                while ((bi + 1) * (bi + 2) / 2 <= idx) bi++;
                const int bj = idx - bi * (bi + 1) / 2;
                double acc[4][4];
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) acc[r][c] = 0.0;
#pragma unroll 8
                for (int k = 0; k < 32; k++) {
                    const double2 y0 = *reinterpret_cast<const double2*>(Yt + k * YP + 4 * bi), y1 = *reinterpret_cast<const double2*>(Yt + k * YP + 4 * bi + 2);
                    const double2 l0 = *reinterpret_cast<const double2*>(Lt + (j0 + k) * LTP + o + 4 * bj), l1 = *reinterpret_cast<const double2*>(Lt + (j0 + k) * LTP + o + 4 * bj + 2);
                    const double y[4] = {y0.x, y0.y, y1.x, y1.y}, l[4] = {l0.x, l0.y, l1.x, l1.y};
#pragma unroll
                    for (int r = 0; r < 4; r++)
#pragma unroll
                        for (int c = 0; c < 4; c++) acc[r][c] = fma(y[r], l[c], acc[r][c]);
                }
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    double* dst = Lt + (o + 4 * bj + c) * LTP + o + 4 * bi;
                    double2 v0 = *reinterpret_cast<double2*>(dst), v1 = *reinterpret_cast<double2*>(dst + 2);
                    v0.x -= acc[0][c]; v0.y -= acc[1][c]; v1.x -= acc[2][c]; v1.y -= acc[3][c];
                    *reinterpret_cast<double2*>(dst) = v0; *reinterpret_cast<double2*>(dst + 2) = v1;
                }
            }
        }
        __syncthreads();
        lap(7);
    }
}

// ---- (b) panel rows. The CTA holds Lt[c][r] = L_kk(r, c) (r > c), 1 / D and 1 / sqrt(D) in shared memory; a warp owns R rows at a
// time (their dependency chains interleave, the loads of L are shared), lane l the columns 4l .. 4l + 3 of each. Forward
// substitution over the 128 columns: the finished y_c is broadcast by shuffle, every lane updates the columns it owns right of c.
// Row n is the right-hand side of the linear system (see ldlt_grid).
template <int R>
__device__ __forceinline__ void panel_rows_R(double* __restrict__ S, int n, int ntot, int k0, int nb, bool quantise, const Scratch& sc, const double* __restrict__ Lt,
                                             const double* __restrict__ invD, const double* __restrict__ rsD, int first, int gnw, int lane)
{
    double a[R][4];
    double* arow[R];
    int rows[R];
#pragma unroll
    for (int i = 0; i < R; i++) {
        rows[i] = first + i * gnw;
        const bool act = rows[i] < ntot;
        arow[i] = (rows[i] < n ? S + (size_t)rows[i] * n : sc.ytmp) + k0 + 4 * lane;
        if (act && nb == NB) {
            const double2 v0 = *reinterpret_cast<const double2*>(arow[i]), v1 = *reinterpret_cast<const double2*>(arow[i] + 2);
            a[i][0] = v0.x; a[i][1] = v0.y; a[i][2] = v1.x; a[i][3] = v1.y;
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) a[i][j] = (act && 4 * lane + j < nb) ? arow[i][j] : 0.0;
        }
    }
    const int ncb = (nb + 3) >> 2;
#pragma unroll 1
    for (int cb = 0; cb < ncb; cb++) {                           // columns 4 cb .. 4 cb + 3 are owned by lane cb
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int c = 4 * cb + j;
            double yc[R];
#pragma unroll
            for (int i = 0; i < R; i++) yc[i] = __shfl_sync(0xffffffffu, a[i][j], cb);
            if (lane >= cb) {
                const double2 l0 = *reinterpret_cast<const double2*>(Lt + c * LTP + 4 * lane), l1 = *reinterpret_cast<const double2*>(Lt + c * LTP + 4 * lane + 2);
                if (lane > cb) {
#pragma unroll
                    for (int i = 0; i < R; i++) { a[i][0] -= yc[i] * l0.x; a[i][1] -= yc[i] * l0.y; a[i][2] -= yc[i] * l1.x; a[i][3] -= yc[i] * l1.y; }
                } else {                                        // the owner of c: only its columns right of c
#pragma unroll
                    for (int i = 0; i < R; i++) {
                        if (j < 1) a[i][1] -= yc[i] * l0.y;
                        if (j < 2) a[i][2] -= yc[i] * l1.x;
                        if (j < 3) a[i][3] -= yc[i] * l1.y;
                    }
                }
            }
        }
    }
    // L = Y D^-1 back into the matrix, Z = Y D^-1/2 into int8 slices
#pragma unroll
    for (int i = 0; i < R; i++) {
        if (rows[i] >= ntot) continue;                          // warp-uniform
        const int row = rows[i];
        double z[4];
        int eb = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int m = 4 * lane + j;
            z[j] = a[i][j] * rsD[m];
            a[i][j] = a[i][j] * invD[m];
            eb = max(eb, (__double2hiint(z[j]) >> 20) & 0x7ff);
        }
        if (nb == NB) {
            *reinterpret_cast<double2*>(arow[i]) = make_double2(a[i][0], a[i][1]);
            *reinterpret_cast<double2*>(arow[i] + 2) = make_double2(a[i][2], a[i][3]);
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (4 * lane + j < nb) arow[i][j] = a[i][j];
        }
        if (!quantise) continue;
        eb = (int)__reduce_max_sync(0xffffffffu, (unsigned)eb);
        const bool live = eb >= kMinExp && eb < 0x7ff;
        const int E = live ? eb - 1023 + 1 : 0;                  // z = z' 2^E with |z'| < 1
        if (lane == 0) sc.Ez[row] = E;
        const double down = live ? pow2(-E) : 0.0;
        double t[4];
#pragma unroll
        for (int j = 0; j < 4; j++) t[j] = z[j] * down * 64.0;
        // plane = 128 rows x 128 bytes (K = the panel's 128 columns); 16-byte chunk c of row r sits at chunk c ^ (r & 7): the
        // 128-byte swizzle the tensor core reads conflict-free (the unswizzled core-matrix layout ran the MMAs at 40 % of this rate)
        int8_t* dst = sc.Zq + (size_t)(row >> 7) * SL * PLANE + (row & 127) * 128 + (((lane >> 2) ^ (row & 7)) << 4) + 4 * (lane & 3);
#pragma unroll
        for (int s = 0; s < SL; s++) {
            uint32_t w = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // round to nearest by adding 1.5 * 2^52: the integer is the sum's low word, its FP64 value the sum minus the constant
                // (the F2I / I2F instructions run at a tenth of the FP64 add rate)
                const double m = t[j] + 6755399441055744.0;
                const int q = __double2loint(m);
                t[j] = (t[j] - (m - 6755399441055744.0)) * 128.0;
                w |= ((uint32_t)q & 0xffu) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(dst + (size_t)s * PLANE) = w;
        }
    }
}

__device__ void panel_rows(double* __restrict__ S, int n, int ntot, int k0, int nb, const Scratch& sc, const double* __restrict__ Lt, const double* __restrict__ invD,
                           const double* __restrict__ rsD, int gwarp, int gnw, int lane)
{
    const int r0 = k0 + nb, m = ntot - r0;
    const bool quantise = r0 < n;                               // the last panel has nothing to update
    if (m <= gnw) { if (r0 + gwarp < ntot) panel_rows_R<1>(S, n, ntot, k0, nb, quantise, sc, Lt, invD, rsD, r0 + gwarp, gnw, lane); }
    else if (m <= 2 * gnw) { if (r0 + gwarp < ntot) panel_rows_R<2>(S, n, ntot, k0, nb, quantise, sc, Lt, invD, rsD, r0 + gwarp, gnw, lane); }
    else for (int first = r0 + gwarp; first < ntot; first += 3 * gnw) panel_rows_R<3>(S, n, ntot, k0, nb, quantise, sc, Lt, invD, rsD, first, gnw, lane);
}

// shared state of the tensor-core update of one CTA
struct UpdateShared {
    unsigned long long full[SL];       // slice s of the operands has landed (transaction bytes)
    unsigned long long accfull[SL];    // accumulator d is complete (tcgen05.commit)
    unsigned long long tempty[SL];     // accumulator d has been read back by the 128 epilogue threads
    unsigned long long smemfree;       // every MMA of the tile has read its operands
    uint32_t tmem;
    int fail;
};

__device__ __forceinline__ int tiles_of_row(int it, int jt0, int nt64) { return min(2 * it + 1, nt64 - 1) - jt0 + 1; }

// ---- (c) A_ij -= Z_i . Z_j over the trailing tiles (rows >= r0; 128 x 64 tiles of the lower triangle), contiguous chunks per CTA.
// warp 0 lane 0 loads, warp 1 lane 0 issues the MMAs, warps 4-7 read the accumulators back and update A. `done` counts the tiles
// this CTA has processed since the barriers were initialised (their phase parities).
__device__ void update_tiles(double* __restrict__ S, int n, int ntot, int r0, const Scratch& sc, uint8_t* smem, UpdateShared& us, int& done, int worker, int nworkers)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt128 = (ntot + TM - 1) / TM, nt64 = (n + TN - 1) / TN, it0 = r0 / TM, jt0 = r0 / TN;
    int total = 0;
    for (int it = it0; it < nt128; it++) total += tiles_of_row(it, jt0, nt64);
    const int chunk = (total + nworkers - 1) / nworkers;
    // contiguous chunks of the tile list (a row tile's operand planes are reused along the chunk), except that the first two tiles
    // -- they cover the next diagonal block, which the factor CTA is waiting for -- open the chunks of two different workers:
    // worker 0 takes tiles 0, 2 .. chunk, worker 1 takes 1, chunk + 1 .. 2 chunk - 1, worker w >= 2 takes [w chunk, (w + 1) chunk)
    const bool split = chunk >= 2 && nworkers >= 2;
    int first, lo, hi;                                          // the worker's tiles: `first` (if >= 0), then [lo, hi)
    if (split && worker == 0) { first = 0; lo = 2; hi = min(total, chunk + 1); }
    else if (split && worker == 1) { first = 1; lo = chunk + 1; hi = min(total, 2 * chunk); }
    else { first = -1; lo = min(total, worker * chunk); hi = min(total, lo + chunk); }
    const int mine = (first >= 0 ? 1 : 0) + max(hi - lo, 0);
    if (mine == 0) return;
    const uint32_t sA = saddr(smem), sB = sA + SL * PLANE;
    double* colscale = reinterpret_cast<double*>(smem + (size_t)SL * PLANE + (size_t)SL * BPLANE);
    const uint32_t tmem = us.tmem;
    int last_it = -1;
    for (int i = 0; i < mine; i++) {
        const int tile = first >= 0 ? (i == 0 ? first : lo + i - 1) : lo + i;
        int it = it0, off = tile;                               // row tile / column tile of the linear tile index
        while (off >= tiles_of_row(it, jt0, nt64)) { off -= tiles_of_row(it, jt0, nt64); it++; }
        const int jt = jt0 + off;
        const uint32_t par = (uint32_t)done & 1u, prev = par ^ 1u;
        if (warp == 0) {
            if (lane == 0) {
                if (done > 0) mbar_wait(saddr(&us.smemfree), prev, &us.fail);
                const bool loadA = it != last_it;
                const int8_t* gA = sc.Zq + (size_t)it * SL * PLANE;
                const int8_t* gB = sc.Zq + (size_t)(jt >> 1) * SL * PLANE + (jt & 1) * BPLANE;
#pragma unroll 1
                for (int s = 0; s < SL; s++) {
                    const uint32_t bar = saddr(&us.full[s]);
                    mbar_expect_tx(bar, (uint32_t)BPLANE + (loadA ? (uint32_t)PLANE : 0u));
                    if (loadA) bulk_load(sA + s * PLANE, gA + (size_t)s * PLANE, PLANE, bar);
                    bulk_load(sB + s * BPLANE, gB + (size_t)s * PLANE, BPLANE, bar);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const bool timedm = sc.ns && blockIdx.x == 0;
                const uint32_t descA_lo = (uint32_t)smem_desc(sA), descB_lo = (uint32_t)smem_desc(sB);
                long long m0 = timedm ? gtime() : 0, tf = 0, te = 0;
#pragma unroll 1
                for (int d = 0; d < SL; d++) {
                    mbar_wait(saddr(&us.full[d]), par, &us.fail);
                    if (timedm) { const long long t = gtime(); tf += t - m0; m0 = t; }
                    if (done > 0) mbar_wait(saddr(&us.tempty[d]), prev, &us.fail);
                    if (timedm) { const long long t = gtime(); te += t - m0; m0 = t; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // one thread issues the 144 MMAs of a tile: the descriptors are a constant high word and a low word that moves
                    // by a constant per slice / K step (built from scratch each time, the issue loop -- not the tensor core -- set the pace)
#pragma unroll 1
                    for (int s = 0; s <= d; s++) {
                        const uint32_t la = descA_lo + (uint32_t)s * (PLANE >> 4), lb = descB_lo + (uint32_t)(d - s) * (BPLANE >> 4);
                        mma_i8_lo(tmem + (uint32_t)(TN * d), la, lb, s > 0 ? 1u : 0u);
                        mma_i8_lo(tmem + (uint32_t)(TN * d), la + 2, lb + 2, 1u);
                        mma_i8_lo(tmem + (uint32_t)(TN * d), la + 4, lb + 4, 1u);
                        mma_i8_lo(tmem + (uint32_t)(TN * d), la + 6, lb + 6, 1u);
                    }
                    mma_commit(saddr(&us.accfull[d]));
                }
                mma_commit(saddr(&us.smemfree));
                if (timedm) { sc.ns[11] += tf; sc.ns[12] += te; sc.ns[13] += 1; }
            }
        } else if (warp >= 4) {
            const int row = TM * it + 32 * (warp - 4) + lane;
            asm volatile("bar.sync 1, 128;" ::: "memory");           // the previous tile's column scales are no longer read
            if (tid - 128 < TN) { const int j = TN * jt + tid - 128; colscale[tid - 128] = j < n ? pow2(sc.Ez[j]) : 0.0; }
            const double rowscale = row < ntot ? pow2(sc.Ez[row] - 12) : 0.0;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const bool timed = sc.ns && blockIdx.x == 0 && tid == 128;
            long long tw = 0, ts = 0, t0 = 0;
            double* out = (row < n ? S + (size_t)row * n : sc.ytmp) + TN * jt;          // row n = the right-hand side
            const int ncol = row < ntot ? min(TN, min(n, row + 1) - TN * jt) : 0;   // columns j <= row (and < n) of this tile
            const uint32_t tbase = tmem + ((uint32_t)(32 * (warp - 4)) << 16);
            // Two passes over the column halves of the tile. The 32 values of S a pass updates are fetched into registers when the
            // pass before it ends (the first pass: before the accumulators are waited for), so their latency -- a warp reads 32
            // different rows -- hides behind the combination of the eight accumulators; those are released to the next tile's MMAs
            // as the second pass reads them.
            double2 pre[16];
            auto fetch = [&](int half) {
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int c = 32 * half + 2 * q;
                    if (c + 1 < ncol) pre[q] = *reinterpret_cast<const double2*>(out + c);
                    else pre[q] = make_double2(c < ncol ? out[c] : 0.0, 0.0);
                }
            };
            fetch(0);
#pragma unroll 1
            for (int half = 0; half < 2; half++) {
                double acc[32];
#pragma unroll
                for (int c = 0; c < 32; c++) acc[c] = 0.0;
                uint32_t v[2][16];
                if (timed) t0 = gtime();
                mbar_wait(saddr(&us.accfull[0]), par, &us.fail);
                if (timed) tw += gtime() - t0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld16_issue(tbase + (uint32_t)(32 * half), v[0]);
                // 8 accumulators x 2 quarters of 16 columns; the load of the next quarter is in flight while one is combined.
                // s32 -> f64 without the conversion instruction (a tenth of the DFMA rate): 2^52 + 2^31 + v as a bit pattern, minus the constant.
#pragma unroll 1
                for (int d = 0; d < SL; d++) {
                    const double wd = pow2(-7 * d);
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        tmem_ld16_wait(v[q]);
                        if (q == 0) tmem_ld16_issue(tbase + (uint32_t)(TN * d + 32 * half + 16), v[1]);
                        else {
                            if (half == 1) {                      // accumulator d is in registers: the next tile may overwrite it
                                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                                mbar_arrive(saddr(&us.tempty[d]));
                            }
                            if (d + 1 < SL) {
                                if (timed) t0 = gtime();
                                mbar_wait(saddr(&us.accfull[d + 1]), par, &us.fail);      // second pass: passed long ago
                                if (timed) tw += gtime() - t0;
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                                tmem_ld16_issue(tbase + (uint32_t)(TN * (d + 1) + 32 * half), v[0]);
                            }
                        }
#pragma unroll
                        for (int c = 0; c < 16; c++) {
                            const double x = __hiloint2double(0x43300000, (int)(v[q][c] ^ 0x80000000u)) - 4503601774854144.0;
                            acc[16 * q + c] = fma(x, wd, acc[16 * q + c]);
                        }
                    }
                }
                if (timed) t0 = gtime();
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int c = 32 * half + 2 * q;
                    double2 o = pre[q];
                    o.x -= acc[2 * q] * rowscale * colscale[c]; o.y -= acc[2 * q + 1] * rowscale * colscale[c + 1];
                    if (c + 1 < ncol) *reinterpret_cast<double2*>(out + c) = o;
                    else if (c < ncol) out[c] = o.x;
                }
                if (half == 0) fetch(1);
                if (timed) ts += gtime() - t0;
            }
            if (timed) { sc.ns[8] += tw; sc.ns[10] += ts; }
            if (it == it0 && jt <= jt0 + 1) {                      // a tile of the next diagonal block is in memory: tell the factor CTA
                __threadfence();
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (tid == 128) atomicAdd(sc.flag, 1);
            }
        }
        last_it = it;
        done++;
    }
}

// In-place factorisation of the symmetric n x n matrix S (row-major, lower triangle read and written) together with the forward
// substitution of one right-hand side: sc.ytmp is treated as row n of the matrix (L of the bordered matrix [S b; b^T .] has
// D^-1 L^-1 b as its last row), so it runs through the same panel solves and tensor-core updates as every other row and only the
// backward substitution (solve_back_grid) remains. Result: strict lower part of S <- L below the diagonal blocks, sc.Ldiag <- the
// factored 128 x 128 diagonal blocks (column-major, pitch LTP; L below the diagonal, D on it), sc.ytmp <- D^-1 L^-1 b. *okflag is
// cleared when a pivot is not positive (or the tensor-core protocol timed out). Every thread of the cooperative grid calls this with
// the same arguments; `smem` = kSmemBytes of dynamic shared memory. Two grid barriers per panel.
__device__ void ldlt_grid(cg::grid_group& grid, double* __restrict__ S, int n, const Scratch& sc, uint8_t* smem_raw, int* okflag)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int gwarp = (int)blockIdx.x * (nt >> 5) + warp, gnw = (int)gridDim.x * (nt >> 5);
    const int ntot = n + 1;
    uint8_t* smem = align_1k(smem_raw);
    __shared__ UpdateShared us;
    __shared__ int s_bad;
    const bool timed0 = sc.ns && blockIdx.x == 0 && tid == 0;   // reading %globaltimer is slow and serialises across the chip: one thread only
    long long t_mark = timed0 ? gtime() : 0;
    auto lap = [&](int slot) { if (timed0) { const long long t = gtime(); sc.ns[slot] += t - t_mark; t_mark = t; } };
    const bool need_update = n > NB;
    if (need_update) {
        if (warp == 0) {
            // no relinquish_alloc_permit: the LM loop calls this once per trial inside one launch (allocating again after the permit
            // was given up traps); the kernel runs one CTA per SM, nobody else waits for tensor memory
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr(&us.tmem)), "r"(512u) : "memory");
        }
        if (tid == 0) {
            for (int s = 0; s < SL; s++) { mbar_init(saddr(&us.full[s]), 1); mbar_init(saddr(&us.accfull[s]), 1); mbar_init(saddr(&us.tempty[s]), 128); }
            mbar_init(saddr(&us.smemfree), 1);
            us.fail = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (need_update) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    double* Lt = reinterpret_cast<double*>(smem);
    double* invD = Lt + NB * LTP;
    double* rsD = invD + NB;
    double* Yt = rsD + NB;
    int done = 0;
    // Look-ahead: the last CTA of the grid takes no update tiles. While the others run the trailing update of panel k it waits for
    // the (at most two) tiles that cover the NEXT diagonal block -- the first ones in worker 0's chunk; their writers count them
    // in sc.flag --, factors that block and publishes it; after the grid barrier everybody copies the factored block into shared
    // memory instead of factoring it. Only the first diagonal block is factored by every CTA (nothing to overlap it with).
    const bool lookahead = gridDim.x >= 8;
    const bool factor_cta = lookahead && blockIdx.x == gridDim.x - 1;
    const int nworkers = lookahead ? (int)gridDim.x - 1 : (int)gridDim.x;
    const int flag0 = lookahead && tid == 0 ? *reinterpret_cast<volatile int*>(sc.flag) : 0;      // the counter runs on across the calls of one launch
    int expected = 0;
    auto publish = [&](int kb) {                                 // Lt, 1 / D, 1 / sqrt(D) of the block just factored -> sc.Ldiag[kb]
        double2* dst = reinterpret_cast<double2*>(sc.Ldiag + (size_t)kb * kPack);
        const double2* src = reinterpret_cast<const double2*>(Lt);
        for (int i = tid; i < kPack / 2; i += nt) dst[i] = src[i];
        if (tid == 0 && s_bad) *okflag = 0;
    };
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int nb = min(NB, n - k0);
        if (k0 == 0 || !lookahead) {                             // (a) every CTA factors the block (same arithmetic, same result)
            load_diag_block(S + (size_t)k0 * n + k0, n, nb, Lt);
            __syncthreads();
            factor_diag_smem(Lt, invD, rsD, Yt, nb, &s_bad, sc.ns, blockIdx.x == 0);
            if (blockIdx.x == 0) publish(k0 / NB);
        } else {                                                 // (a') factored during the previous update phase
            const double2* src = reinterpret_cast<const double2*>(sc.Ldiag + (size_t)(k0 / NB) * kPack);
            const uint32_t dst = saddr(Lt);
            for (int i = tid; i < kPack / 2; i += nt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + i) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
        }
        lap(0);
        panel_rows(S, n, ntot, k0, nb, sc, Lt, invD, rsD, gwarp, gnw, lane);      // (b)
        __threadfence();
        lap(1);
        const int r0 = k0 + nb;
        if (r0 >= n) break;
        grid.sync();
        lap(3);
        if (factor_cta) {
            const int nbn = min(NB, n - r0), nt64 = (n + TN - 1) / TN, jt0 = r0 / TN;
            expected += min(jt0 + 1, nt64 - 1) - jt0 + 1;         // tiles (row tile r0 / 128, column tiles jt0, jt0 + 1) of this update
            const bool timedf = sc.ns && tid == 0;
            long long f0 = timedf ? gtime() : 0;
            if (tid == 0) {
                int seen;
                do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(sc.flag) : "memory"); } while (seen - flag0 < expected);
            }
            __syncthreads();
            __threadfence();
            if (timedf) { const long long t = gtime(); sc.ns[14] += t - f0; f0 = t; }
            load_diag_block(S + (size_t)r0 * n + r0, n, nbn, Lt);
            __syncthreads();
            factor_diag_smem(Lt, invD, rsD, Yt, nbn, &s_bad, sc.ns, true);
            publish(r0 / NB);
            if (timedf) sc.ns[9] += gtime() - f0;
        } else {
            asm volatile("fence.proxy.async;" ::: "memory");      // the planes other CTAs stored become visible to this CTA's bulk copies
            update_tiles(S, n, ntot, r0, sc, smem, us, done, (int)blockIdx.x, nworkers);          // (c)
        }
        __threadfence();
        lap(2);
        grid.sync();
        lap(3);
    }
    if (need_update) {
        __syncthreads();
        if (tid == 0 && us.fail) *okflag = 0;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(us.tmem), "r"(512u) : "memory");
        if (tid == 0) {                                          // the next call initialises the barriers again
            for (int s = 0; s < SL; s++) { mbar_inval(saddr(&us.full[s])); mbar_inval(saddr(&us.accfull[s])); mbar_inval(saddr(&us.tempty[s])); }
            mbar_inval(saddr(&us.smemfree));
        }
        __syncthreads();
    }
}

// Copies the factored diagonal blocks from sc.Ldiag into S (L below the diagonal, D on it): only the test entry wants the factor in
// one piece. Call after a grid barrier behind ldlt_grid.
__device__ void export_diag_blocks(double* __restrict__ S, int n, const Scratch& sc)
{
    const int gtid = (int)(blockIdx.x * blockDim.x + threadIdx.x), gnt = (int)(gridDim.x * blockDim.x);
    const int nblk = (n + NB - 1) / NB;
    for (int i = gtid; i < nblk * NB * NB; i += gnt) {
        const int kb = i / (NB * NB), r = (i / NB) % NB, c = i % NB;
        if (c <= r && kb * NB + r < n) S[(size_t)(kb * NB + r) * n + kb * NB + c] = sc.Ldiag[(size_t)kb * kPack + c * LTP + r];
    }
}

// Solves L^T x = y in place (y = sc.ytmp, as ldlt_grid leaves it), 128 unknowns per step from the last block up, ONE grid barrier per
// step: every CTA solves the triangle on the diagonal itself (same arithmetic, same result -- cheaper than a barrier and a round trip
// through global memory) -- per 32-wide sub-block one warp with one unknown per lane, the column of L in registers, solved values
// broadcast by shuffle, then the sub-block's contribution to the earlier unknowns of the block -- and then subtracts the block's
// contribution from its share of all earlier unknowns (32 columns x 8 row groups per CTA pass, partial sums combined in a fixed
// order). The next diagonal block travels into shared memory by cp.async meanwhile.
__device__ void solve_back_grid(cg::grid_group& grid, const double* __restrict__ S, int n, const Scratch& sc, uint8_t* smem_raw)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    double* y = sc.ytmp;
    double* T = reinterpret_cast<double*>(align_1k(smem_raw));   // the diagonal block, Lt[c][r]
    double* xs = T + NB * LTP;                                   // [NB] unknowns of the block
    double* red = xs + NB;                                       // [8][32] partial sums
    const bool timed0 = sc.ns && blockIdx.x == 0 && tid == 0;
    const long long t_mark = timed0 ? gtime() : 0;
    auto fetch_block = [&](int k0) {                             // Ldiag[k0 / NB] -> T, asynchronously
        const int nb = min(NB, n - k0), cnt = ((nb + 31) & ~31) * LTP / 2;
        const double2* src = reinterpret_cast<const double2*>(sc.Ldiag + (size_t)(k0 / NB) * kPack);
        const uint32_t dst = saddr(T);
        for (int i = tid; i < cnt; i += nt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * i), "l"(src + i) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int klast = ((n - 1) / NB) * NB;
    fetch_block(klast);
    for (int k0 = klast; k0 >= 0; k0 -= NB) {
        const int nb = min(NB, n - k0);
        // the block solved in the previous step goes back into y only now: every CTA read its right-hand side from there at the
        // top of that step, and all of them have passed a grid barrier since
        if (blockIdx.x == 0 && k0 < klast && tid < min(NB, n - (k0 + NB))) y[k0 + NB + tid] = xs[tid];
        __syncthreads();
        if (tid < NB) xs[tid] = tid < nb ? y[k0 + tid] : 0.0;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            for (int j0 = (nb - 1) & ~31; j0 >= 0; j0 -= 32) {
                double col[32];                                  // L(j0 + r, j0 + lane), r > lane
#pragma unroll
                for (int r = 0; r < 32; r += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(T + (j0 + lane) * LTP + j0 + r);
                    col[r] = v.x; col[r + 1] = v.y;
                }
                double a = xs[j0 + lane];
#pragma unroll
                for (int r = 31; r >= 1; r--) {
                    const double xr = __shfl_sync(0xffffffffu, a, r);
                    if (lane < r) a -= xr * col[r];
                }
                xs[j0 + lane] = a;
                __syncwarp();
                for (int c = lane; c < j0; c += 32) {            // earlier unknowns of this block
                    double acc = 0;
#pragma unroll
                    for (int r = 0; r < 32; r += 2) {
                        const double2 v = *reinterpret_cast<const double2*>(T + c * LTP + j0 + r);
                        acc = fma(v.x, xs[j0 + r], acc); acc = fma(v.y, xs[j0 + r + 1], acc);
                    }
                    xs[c] -= acc;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (k0 == 0) { if (blockIdx.x == 0 && tid < nb) y[tid] = xs[tid]; break; }      // nobody reads y[0 .. nb) any more
        fetch_block(k0 - NB);
        {
            const int cl = tid & 31, rg = tid >> 5;              // 32 columns x 8 groups of 16 rows
            for (int g = (int)blockIdx.x; 32 * g < k0; g += (int)gridDim.x) {
                const int i = 32 * g + cl;                        // k0 is a multiple of 128: i < k0
                double acc = 0;
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    const int j = 16 * rg + jj;
                    if (j < nb) acc = fma(S[(size_t)(k0 + j) * n + i], xs[j], acc);
                }
                red[32 * rg + cl] = acc;
                __syncthreads();
                if (rg == 0) {
                    double sum = 0;
#pragma unroll
                    for (int q = 0; q < 8; q++) sum += red[32 * q + cl];
                    y[i] -= sum;
                }
                __syncthreads();
            }
        }
        __threadfence();
        grid.sync();
    }
    __threadfence();
    grid.sync();                                                 // the first block's unknowns (written by CTA 0) are visible to every CTA
    if (timed0) sc.ns[4] += gtime() - t_mark;
}

inline size_t scratch_zq_bytes(int n) { return (size_t)((n + 1 + TM - 1) / TM) * SL * PLANE; }
inline size_t scratch_ez_count(int n) { return (size_t)((n + 1 + TM - 1) / TM) * TM; }
inline size_t scratch_ldiag_bytes(int n) { return (size_t)((n + NB - 1) / NB) * kPack * sizeof(double); }

} // namespace dense
} // namespace mage
