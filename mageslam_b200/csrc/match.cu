// Brute-force two-way Hamming matcher for sm_100a.
//
// Replaces Match() (ref Core/MAGESLAM/Source/Tracking/FeatureMatcher.cpp:61-190: two cv::BFMatcher::radiusMatch passes,
// min best/second-best difference, cross-check) and GetDescriptorDistance (ref :453-504). Closed form in SURVEY A.5:
//   for each query q: C = {t : H(q,t) <= maxHamming}; d1 <= d2 the two smallest; best(q) = argmin iff C != {} and
//   (|C| == 1 or d2 - d1 >= minDiff); emit (a, best(a), d1) in ascending a iff best_B(best_A(a)) == a.
// The operands (2 x 64 KB at 2000 descriptors) live in shared memory / L2 -- compute bound, not HBM bound. One launch computes both
// directions of every pair of a batch. Two kernels produce the same packed best / second-best statistics: k_match_dir evaluates the
// distance matrix on the tcgen05 tensor cores as an int8 dot product (default, maxHamming <= 64), k_match_dir_popc with xor + popc
// over 8 x 32-bit words behind a 96-bit filter (wider radii, or MAGE_MATCH_POPC=1).
//
// The reference sizes its backward lookup by the number of non-empty rows but indexes it by query index (:123-137,
// :158) -- an out-of-range access whenever some B descriptor has no candidate. The lookup is sized nB here (the intent).
// Ties for best resolve to the lowest index; with minDiff >= 1 (all shipped settings) a tie is rejected anyway.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

namespace mage {

struct MatchJob {
    const uint8_t* desc[2];     // [0] = A, [1] = B, 32 bytes per descriptor, 4-byte aligned
    const uint8_t* mask[2];     // nullable, one byte per descriptor
    const int* count_ptr[2];    // nullable: device-resident counts (device batched variant)
    int count[2];
    int cap;                    // clamp for device counts
};

constexpr int kChunk = 256;          // train descriptors staged per iteration
constexpr int kQPerWarp = 4;
constexpr int kWarps = 8;
constexpr int kQPerBlock = kQPerWarp * kWarps;
constexpr unsigned kNoKey = 0xFFFFFFFFu;

__device__ __forceinline__ int job_count(const MatchJob& j, int side)
{
    int n = j.count_ptr[side] ? *j.count_ptr[side] : j.count[side];
    return max(0, min(n, j.cap));
}

// The distance matrix is evaluated ONCE for both directions. Each (a, b) first goes through a 96-bit filter -- the first three
// words with a carry-save adder (ones = x0^x1^x2, twos = maj(x0,x1,x2); popc(ones) + 2 popc(twos) is the exact 96-bit
// distance, a lower bound of the full one) -- which rejects all but ~1e-4 of random pairs at maxHamming 30 with 2 POPC.
// Survivors get the exact 256-bit distance; if it is within maxHamming the pair updates four packed (distance << 16 | index)
// slots with integer atomics: best/second of row a (A -> B) and best/second of column b (B -> A). "second" receives the
// loser of every best update, which yields exactly the second smallest key whatever the arrival order, so the result is
// deterministic. stats layout: [pair][4][stride] = fwd best, fwd second, bwd best, bwd second.
__global__ void __launch_bounds__(kWarps * 32) k_match_dir_popc(const MatchJob* __restrict__ jobs, unsigned* __restrict__ stats, int stride, int max_hamming)
{
    __shared__ uint32_t tw[8][kChunk];          // train words, transposed: conflict-free lane-strided reads
    __shared__ uint32_t tx[kChunk];             // t0 ^ t1 ^ t2 per train
    __shared__ uint32_t t01s[kChunk];           // t0 ^ t1 per train
    __shared__ uint8_t tvalid[kChunk];
    __shared__ uint32_t qw[kQPerBlock][8];
    const int pair = blockIdx.y;
    const MatchJob& job = jobs[pair];
    const int nQ = job_count(job, 0), nT = job_count(job, 1);
    const int q0 = blockIdx.x * kQPerBlock;
    if (q0 >= nQ) return;
    const uint32_t* Q = reinterpret_cast<const uint32_t*>(job.desc[0]);
    const uint4* T4 = reinterpret_cast<const uint4*>(job.desc[1]);
    const uint8_t* mQ = job.mask[0];
    const uint8_t* mT = job.mask[1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned* fbest = stats + ((size_t)pair * 4 + 0) * stride;
    unsigned* fsecond = stats + ((size_t)pair * 4 + 1) * stride;
    unsigned* bbest = stats + ((size_t)pair * 4 + 2) * stride;
    unsigned* bsecond = stats + ((size_t)pair * 4 + 3) * stride;

    for (int i = threadIdx.x; i < kQPerBlock * 8; i += blockDim.x) {
        int q = q0 + (i >> 3);
        qw[i >> 3][i & 7] = (q < nQ) ? Q[(size_t)q * 8 + (i & 7)] : 0u;
    }
    __syncthreads();
    uint32_t q[kQPerWarp][3], qx[kQPerWarp], q01[kQPerWarp];
    bool qok[kQPerWarp];
#pragma unroll
    for (int k = 0; k < kQPerWarp; k++) {
        const int qi = q0 + warp * kQPerWarp + k;
        q[k][0] = qw[warp * kQPerWarp + k][0]; q[k][1] = qw[warp * kQPerWarp + k][1]; q[k][2] = qw[warp * kQPerWarp + k][2];
        qx[k] = q[k][0] ^ q[k][1] ^ q[k][2];
        q01[k] = q[k][0] ^ q[k][1];
        qok[k] = qi < nQ && (!mQ || mQ[qi]);
    }
    for (int c0 = 0; c0 < nT; c0 += kChunk) {
        __syncthreads();
        // 16-byte coalesced global loads; thread i -> descriptor i/2, half i%2
        for (int i = threadIdx.x; i < kChunk * 2; i += blockDim.x) {
            const int tl = i >> 1, half = i & 1, t = c0 + tl;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (t < nT) v = T4[(size_t)t * 2 + half];
            tw[half * 4 + 0][tl] = v.x; tw[half * 4 + 1][tl] = v.y; tw[half * 4 + 2][tl] = v.z; tw[half * 4 + 3][tl] = v.w;
            if (half == 0) { tx[tl] = v.x ^ v.y ^ v.z; t01s[tl] = v.x ^ v.y; tvalid[tl] = (t < nT && (!mT || mT[t])) ? 1 : 0; }
        }
        __syncthreads();
        // filter pass: branch-free over the lane's 8 trains x 4 queries (32 pairs), survivors recorded in a bit mask
        static_assert(kChunk / 32 * kQPerWarp == 32, "one hit bit per (train, query) evaluation");
        unsigned hits = 0;
#pragma unroll
        for (int i = 0; i < kChunk / 32; i++) {
            const int ti = lane + 32 * i;
            const uint32_t t0 = tw[0][ti], t2 = tw[2][ti], t01 = t01s[ti], txx = tx[ti];
#pragma unroll
            for (int k = 0; k < kQPerWarp; k++) {
                // maj(x0, x1, x2) = (x0 != x1) ? x2 : x0, with x0 ^ x1 = (q0 ^ q1) ^ (t0 ^ t1) from precomputed halves:
                // 4 LOP3 for the carry word + 1 for the sum word (xor/popc saturate the half-rate ALU pipe; every LOP3 counts)
                const uint32_t u = q01[k] ^ t01, x0 = q[k][0] ^ t0, x2 = q[k][2] ^ t2;
                const uint32_t twos = (u & x2) | (~u & x0);
                // the sign of (maxHamming - d96) is shifted in with one funnel shift (ALU pipe) and the subtraction is left to
                // IMADs (FMA pipe): xor/popc saturate the half-rate ALU pipe, so a compare + select there costs real time
                int slack;
                asm("mad.lo.s32 %0, %1, -2, %2;" : "=r"(slack) : "r"(__popc(twos)), "r"(max_hamming));
                asm("mad.lo.s32 %0, %1, -1, %2;" : "=r"(slack) : "r"(__popc(qx[k] ^ txx)), "r"(slack));
                hits = __funnelshift_l((unsigned)slack, hits, 1);           // reject bit of evaluation n ends up at bit 31 - n
            }
        }
        hits = ~hits;
        // survivors (about 1e-4 of random pairs, plus the true matches): exact distance, then the four packed updates
        while (hits) {
            const int bit = __ffs(hits) - 1;
            hits &= hits - 1;
            const int n = 31 - bit, i = n / kQPerWarp, k = n % kQPerWarp;
            const int ti = lane + 32 * i;
            if (!tvalid[ti] || !qok[k]) continue;
            const int ql = warp * kQPerWarp + k;
            int d = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) d += __popc(qw[ql][w] ^ tw[w][ti]);
            if (d > max_hamming) continue;                          // radiusMatch keeps d <= maxDistance only
            const unsigned tidx = (unsigned)(c0 + ti), qidx = (unsigned)(q0 + ql);
            const unsigned kf = ((unsigned)d << 16) | tidx, kb = ((unsigned)d << 16) | qidx;
            unsigned old = atomicMin(&fbest[qidx], kf);
            atomicMin(&fsecond[qidx], max(old, kf));
            old = atomicMin(&bbest[tidx], kb);
            atomicMin(&bsecond[tidx], max(old, kb));
        }
    }
}

// ---- tcgen05 path (the default for maxHamming <= kUmmaMaxHamming) ------------------------------------------------------------
// With the descriptor bits spread to signed bytes (+1 for a set bit, -1 for a clear one) the Hamming distance is a dot product,
//     <a, b> = 256 - 2 H(a, b),
// so the 2000 x 2000 distance matrix of a frame pair is an int8 GEMM with K = 256: 5th-generation tensor cores
// (tcgen05.mma kind::i8, M = 128 queries x N = 256 trains x K = 32 per instruction, operands in shared memory in the canonical
// K-major no-swizzle layout, s32 accumulators in tensor memory) take it off the POPC pipe that bounds k_match_dir_popc.
// The accumulators come back with tcgen05.ld (thread = query row, 32 train columns at a time); a running maximum against
// T = 256 - 2 maxHamming finds the columns inside the radius -- their exact distance is (256 - acc) / 2 -- and those
// (about 1 per query) make the same four packed atomic updates as above, so the statistics and hence the matches are identical.
// CTA = 128 queries (A tile expanded once, 32 KB) x a walk over the train tiles of 128 (B tile 32 KB, double-buffered): the bit -> byte
// expansion of tile j + 1 and the read-out of tile j - 1 run while the tensor core works on tile j; the accumulators alternate
// between the two halves of the CTA's 256 TMEM columns. Masked or absent descriptors expand to zero bytes: acc = 0 < T.
constexpr int kUmmaM = 128, kUmmaN = 128;
constexpr int kUmmaThreads = 512;                                  // 16 working warps (the expansion and the read-out are latency bound with fewer) + 1 issuing warp
constexpr int kUmmaMaxHamming = 64;                               // T must stay positive; wider radii take the popc kernel
constexpr uint32_t kUmmaLBO = 128, kUmmaSBO = 256;                // K-chunk (16 B) stride, 8-row group stride inside one K = 32 step
constexpr uint32_t kUmmaStepA = kUmmaM * 32, kUmmaStepB = kUmmaN * 32;      // bytes per K = 32 step
constexpr uint32_t kUmmaBytesA = kUmmaM * 256, kUmmaBytesB = kUmmaN * 256;
constexpr size_t kUmmaSmemBytes = kUmmaBytesA + 2 * (size_t)kUmmaBytesB + 1024;
// instruction descriptor (kind::i8): D = s32, A and B signed 8 bit, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kUmmaIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kUmmaN >> 3) << 17) | ((uint32_t)(kUmmaM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: start address, leading (K chunk) and stride (row group) byte offsets in 16-byte units, version 1
// (Blackwell), no swizzle
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kUmmaLBO >> 4) << 16) | ((uint64_t)(kUmmaSBO >> 4) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kUmmaIdesc), "r"(accumulate), "r"(0u) : "memory");
}

// bounded wait on an mbarrier phase: a descriptor mistake must end in a trap, not in a hung device. The bound is TIME (ten seconds of
// %globaltimer, looked at every 64 K polls), not a poll count: under time-slicing, a debugger or a profiler's replay a correct kernel
// may poll for a long while, and a trap poisons the whole context.
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity)
{
    long long t0 = 0;
    for (unsigned spin = 0;; spin++) {
        uint32_t done;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 0xFFFFu) == 0xFFFFu) {
            long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 10000000000ll) __trap();
        }
    }
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

// A pair inside the radius makes four packed updates, two of which need the value an atomic returns: ~2 L2 round trips. With one CTA
// of 8 warps per SM nothing hides that latency, so the read-out only appends (query row, train, distance) to a shared-memory list and
// the whole CTA drains the list at the end, one entry per thread, all round trips in flight together.
constexpr int kUmmaListCap = 2048;
__device__ __forceinline__ void match_update(unsigned qidx, unsigned tidx, unsigned d, unsigned* fbest, unsigned* fsecond, unsigned* bbest, unsigned* bsecond)
{
    const unsigned kf = (d << 16) | tidx, kb = (d << 16) | qidx;
    unsigned old = atomicMin(&fbest[qidx], kf);
    atomicMin(&fsecond[qidx], max(old, kf));
    old = atomicMin(&bbest[tidx], kb);
    atomicMin(&bsecond[tidx], max(old, kb));
}
__device__ __noinline__ void match_update_direct(unsigned qidx, unsigned tidx, unsigned d, unsigned* fbest, unsigned* fsecond, unsigned* bbest, unsigned* bsecond)
{
    match_update(qidx, tidx, d, fbest, fsecond, bbest, bsecond);
}

// NW consecutive descriptor words (starting at word first_chunk / 2) -> their 2 NW K chunks of signed bytes at `row` of an operand tile.
// Any fixed assignment of descriptor bits to K positions gives the same dot product as long as both operands use it, so a group of
// four K positions takes bits b, b + 8, b + 16, b + 24 of one word: (w >> b) & 0x01010101 is already the 0/1 byte form (one shift, one
// AND), and 0xFFFFFFFF - 0xFE t turns 1 -> 0x01 (+1), 0 -> 0xFF (-1) in every byte with one IMAD. An absent row is all zero bytes.
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& c)       // 32-bit shared-window address: no generic-pointer arithmetic
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w) : "memory");
}

template <int NW>
__device__ __forceinline__ void umma_expand(uint32_t tile, uint32_t kstep_bytes, int row, int first_chunk, const uint32_t (&w)[NW], bool present)
{
    const uint32_t dst = tile + (row >> 3) * kUmmaSBO + (row & 7) * 16 + (first_chunk >> 1) * kstep_bytes;       // first_chunk is even
    if (!present) {
#pragma unroll
        for (int q = 0; q < 2 * NW; q++) sts128(dst + (q >> 1) * kstep_bytes + (q & 1) * kUmmaLBO, make_uint4(0u, 0u, 0u, 0u));
        return;
    }
#pragma unroll
    for (int q = 0; q < 2 * NW; q++) {                             // chunk first_chunk + q = K positions 16 (first_chunk + q) .. + 15
        const uint32_t x = w[q >> 1] >> (4 * (q & 1));
        uint4 c;
        c.x = 0xFFFFFFFFu - (x & 0x01010101u) * 0xFEu;
        c.y = 0xFFFFFFFFu - ((x >> 1) & 0x01010101u) * 0xFEu;
        c.z = 0xFFFFFFFFu - ((x >> 2) & 0x01010101u) * 0xFEu;
        c.w = 0xFFFFFFFFu - ((x >> 3) & 0x01010101u) * 0xFEu;
        sts128(dst + (q >> 1) * kstep_bytes + (q & 1) * kUmmaLBO, c);
    }
}

__global__ void __launch_bounds__(kUmmaThreads + 32, 2) k_match_dir(const MatchJob* __restrict__ jobs, unsigned* __restrict__ stats, int stride, int max_hamming)
{
    extern __shared__ uint8_t umma_smem_raw[];
    __shared__ __align__(8) uint64_t mbar[2], full[2];           // MMA of a tile done / B buffer of a tile expanded (512 arrivals)
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t hit_list[kUmmaListCap];                    // row | train << 7 | distance << 23
    __shared__ int hit_count;
    const int pair = blockIdx.y;
    const MatchJob& job = jobs[pair];
    const int nQ = job_count(job, 0), nT = job_count(job, 1);
    const int q0 = blockIdx.x * kUmmaM;
    const int ntiles = (nT + kUmmaN - 1) / kUmmaN;
    if (q0 >= nQ || (int)blockIdx.z >= ntiles) return;
    const uint32_t sA = (smem_addr(umma_smem_raw) + 1023u) & ~1023u;         // shared-window addresses
    const uint32_t sB = sA + kUmmaBytesA;                          // two buffers of kUmmaBytesB
    const uint4* Q4 = reinterpret_cast<const uint4*>(job.desc[0]);
    const uint4* T4 = reinterpret_cast<const uint4*>(job.desc[1]);
    const uint8_t* mQ = job.mask[0];
    const uint8_t* mT = job.mask[1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned* fbest = stats + ((size_t)pair * 4 + 0) * stride;
    unsigned* fsecond = stats + ((size_t)pair * 4 + 1) * stride;
    unsigned* bbest = stats + ((size_t)pair * 4 + 2) * stride;
    unsigned* bsecond = stats + ((size_t)pair * 4 + 3) * stride;
    const int thr = 256 - 2 * max_hamming;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        hit_count = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&mbar[0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&mbar[1])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&full[0])), "r"(kUmmaThreads) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&full[1])), "r"(kUmmaThreads) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (tid < 2 * kUmmaM) {                                        // A tile: this CTA's queries, expanded once (thread = row, half)
        const int row = tid & (kUmmaM - 1), h = tid / kUmmaM, qi = q0 + row;
        const bool present = qi < nQ && (!mQ || mQ[qi]);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (present) v = __ldg(Q4 + (size_t)qi * 2 + h);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        umma_expand<4>(sA, kUmmaStepA, row, 8 * h, w, present);
    }
    // read-out of the tile issued in iteration it_done: warp w owns TMEM lanes 32 (w % 4) .. + 31 (queries) and, by w / 4, 32 of the
    // 128 train columns
    auto read_out = [&](int it_done, int tile) {
        const int buf = it_done & 1;
        mbar_wait_bounded(smem_addr(&mbar[buf]), (uint32_t)(it_done >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = 32 * (warp & 3) + lane;
        const unsigned qidx = (unsigned)(q0 + row);
        const int col0 = (warp >> 2) * 32;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(buf * kUmmaN + col0), v);
        // maxima of the four groups of 8 columns; a group is scanned only when some lane of the warp has a hit in it
#pragma unroll
        for (int gq = 0; gq < 4; gq++) {
            const uint32_t* u = v + 8 * gq;
            int m = __vimax3_s32((int)u[0], (int)u[1], (int)u[2]);
            m = __vimax3_s32(m, (int)u[3], (int)u[4]);
            m = __vimax3_s32(m, (int)u[5], (int)u[6]);
            m = max(m, (int)u[7]);
            if (m < thr) continue;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if ((int)u[i] < thr) continue;
                const unsigned tidx = (unsigned)(tile * kUmmaN + col0 + 8 * gq + i), d = (256u - u[i]) >> 1;
                const int slot = atomicAdd(&hit_count, 1);
                if (slot < kUmmaListCap) hit_list[slot] = (unsigned)row | (tidx << 7) | (d << 23);
                else match_update_direct(qidx, tidx, d, fbest, fsecond, bbest, bsecond);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    // Warps 0-15 expand and read out; warp 16 only issues the MMAs (the issue blocks while the tensor pipe's queue is full and must not
    // hold up a warp that has other work). The B tile and the accumulators are double-buffered, so the tensor core works on tile j
    // while tile j + 1 is expanded and tile j - 1 read out. No CTA-wide barrier in the loop: every working thread arrives on full[buf]
    // once it has expanded its part of the buffer (and, by program order, read its part of the accumulator half out); the issuing
    // thread waits for the 512 arrivals, the working threads wait for the MMA's commit only when they read that tile out. Two such CTAs share an SM (97 KB of shared memory and 256 TMEM columns each) to hide the
    // latencies of the expansion and the read-out.
    const bool mma_warp = warp == kUmmaThreads / 32;
    const int brow = tid & (kUmmaN - 1), bquarter = (tid / kUmmaN) & 3;           // B tile: thread = (train row, quarter descriptor)
    auto load_quarter = [&](int tile, bool& present) {
        const int ti = tile * kUmmaN + brow;
        present = ti < nT && (!mT || mT[ti]);
        return present ? __ldg(reinterpret_cast<const uint2*>(T4 + (size_t)ti * 2) + bquarter) : make_uint2(0u, 0u);
    };
    int it = 0, prev_tile = -1;
    bool present = false;
    uint2 v = make_uint2(0u, 0u);
    if (!mma_warp) v = load_quarter(blockIdx.z, present);
    for (int tile = blockIdx.z; tile < ntiles; tile += gridDim.z, it++) {
        const int buf = it & 1;
        if (!mma_warp) {
            // the MMA that last read this buffer (iteration it - 2) was waited for by every thread in read_out(it - 2)
            const uint32_t w[2] = {v.x, v.y};
            umma_expand<2>(sB + (uint32_t)buf * kUmmaBytesB, kUmmaStepB, brow, 4 * bquarter, w, present);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the tensor core's reads
        }
        if (mma_warp) {
            if (lane == 0) {
                // every working thread has expanded its part of B[buf] and read its part of accumulator half `buf` out (tile it - 2)
                mbar_wait_bounded(smem_addr(&full[buf]), (uint32_t)(it >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = sA, b0 = sB + (uint32_t)buf * kUmmaBytesB;
#pragma unroll
                for (int ks = 0; ks < 8; ks++)
                    umma_i8(tmem + (uint32_t)(buf * kUmmaN), umma_desc(a0 + ks * kUmmaStepA), umma_desc(b0 + ks * kUmmaStepB), ks > 0 ? 1u : 0u);
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&mbar[buf])) : "memory");
            }
            __syncwarp();
        } else {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(&full[buf])) : "memory");
            if (tile + (int)gridDim.z < ntiles) v = load_quarter(tile + gridDim.z, present);     // in flight during the read-out
            if (it > 0) read_out(it - 1, prev_tile);
        }
        prev_tile = tile;
    }
    if (!mma_warp && it > 0) read_out(it - 1, prev_tile);
    __syncthreads();
    for (int i = tid, n = min(hit_count, kUmmaListCap); i < n; i += blockDim.x) {
        const unsigned e = hit_list[i];
        match_update((unsigned)q0 + (e & 127u), (e >> 7) & 0xFFFFu, e >> 23, fbest, fsecond, bbest, bsecond);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// grid for either matcher kernel; the tcgen05 one splits the train tiles over blockIdx.z until there are enough CTAs to fill the GPU
// (a single 2000 x 2000 pair is only 16 query blocks)
static cudaError_t launch_match_dir(const MatchJob* d_jobs, unsigned* best, int max_desc, int max_q, int max_t, int n_pairs, int max_hamming, cudaStream_t s)
{
    static const bool force_popc = getenv("MAGE_MATCH_POPC") != nullptr;
    if (force_popc || max_hamming < 0 || max_hamming > kUmmaMaxHamming) {
        dim3 grid(div_up(max_q, kQPerBlock), n_pairs);
        k_match_dir_popc<<<grid, kWarps * 32, 0, s>>>(d_jobs, best, max_desc, max_hamming);
        return cudaSuccess;
    }
    const int gx = std::max(1, div_up(max_q, kUmmaM)), tiles = std::max(1, div_up(max_t, kUmmaN));
    const int gz = std::min(tiles, std::max(1, div_up(2 * 148, gx * n_pairs)));
    dim3 grid(gx, n_pairs, gz);
    k_match_dir<<<grid, kUmmaThreads + 32, kUmmaSmemBytes, s>>>(d_jobs, best, max_desc, max_hamming);
    return cudaSuccess;
}

// min-difference test (ref :130-135 / :149-154) on the packed stats, then cross-check (ref :158) and ordered emission
__device__ __forceinline__ unsigned resolve_best(unsigned best, unsigned second, int min_diff)
{
    if (best == kNoKey) return kNoKey;
    if (second != kNoKey && (int)(second >> 16) - (int)(best >> 16) < min_diff) return kNoKey;
    return best;
}

constexpr int kEmitThreads = 1024;      // one CTA per pair walks the queries in rounds of two dependent global round trips each: 2 rounds for 2000, not 8
__global__ void __launch_bounds__(kEmitThreads) k_match_emit(const MatchJob* __restrict__ jobs, const unsigned* __restrict__ stats, int stride, int min_diff,
                                                    mage_dmatch* __restrict__ out, int out_cap, int* __restrict__ out_count)
{
    __shared__ int warp_sum[kEmitThreads / 32];
    __shared__ int base;
    const int pair = blockIdx.x;
    const MatchJob& job = jobs[pair];
    const int nA = job_count(job, 0), nB = job_count(job, 1);
    const unsigned* fbest = stats + ((size_t)pair * 4 + 0) * stride;
    const unsigned* fsecond = stats + ((size_t)pair * 4 + 1) * stride;
    const unsigned* bbest = stats + ((size_t)pair * 4 + 2) * stride;
    const unsigned* bsecond = stats + ((size_t)pair * 4 + 3) * stride;
    mage_dmatch* o = out + (size_t)pair * out_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int a0 = 0; a0 < nA; a0 += blockDim.x) {
        const int a = a0 + threadIdx.x;
        bool ok = false; unsigned key = kNoKey;
        if (a < nA) {
            key = resolve_best(fbest[a], fsecond[a], min_diff);
            if (key != kNoKey) {
                const int t = key & 0xFFFF;
                if (t < nB) { const unsigned bk = resolve_best(bbest[t], bsecond[t], min_diff); ok = bk != kNoKey && (int)(bk & 0xFFFF) == a; }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_sum[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; w++) off += warp_sum[w];
        off += __popc(m & ((1u << lane) - 1));
        if (ok && off < out_cap) { mage_dmatch dm; dm.query_idx = a; dm.train_idx = (int)(key & 0xFFFF); dm.distance = (float)(key >> 16); o[off] = dm; }
        __syncthreads();
        if (threadIdx.x == 0) { int tot = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += warp_sum[w]; base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_count[pair] = min(base, out_cap);
}

// IndexedMatch (ref Tracking/FeatureMatcher.cpp:192-268): the candidate sets come from the vocabulary (BoW node lists, the
// caller's business) instead of the whole image, as CSR lists. One warp per query descriptor and direction (blockIdx.y:
// 0 = A -> B over a2b, 1 = B -> A over b2a); lanes stride over the list with the exact 256-bit distance. TrackMatch
// (ref :28-56) keeps the true smallest and second smallest distance below maxHamming, the best being the FIRST list entry
// with the smallest distance: key = distance << 22 | list position. The stats land in the layout k_match_emit resolves
// (min-difference test, B -> A best must point back to a, ascending-a emission) -- the reference's second loop evaluates
// B -> A only for matched b's, but its result for a given b does not depend on which a asked.
__global__ void __launch_bounds__(256) k_indexed_best(const MatchJob* __restrict__ jobs, const int* __restrict__ off0, const int* __restrict__ cand0,
                                                      const int* __restrict__ off1, const int* __restrict__ cand1, unsigned* __restrict__ stats,
                                                      int stride, int max_hamming)
{
    const MatchJob& job = jobs[0];
    const int dir = blockIdx.y;
    const int nQ = job_count(job, dir), nT = job_count(job, 1 - dir);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * (blockDim.x >> 5) + warp;
    if (q >= nQ) return;
    const uint8_t* mQ = job.mask[dir];
    const uint8_t* mT = job.mask[1 - dir];
    if (mQ && !mQ[q]) return;                                   // stats stay kNoKey
    const uint4* Q4 = reinterpret_cast<const uint4*>(job.desc[dir]) + (size_t)q * 2;
    const uint4* T4 = reinterpret_cast<const uint4*>(job.desc[1 - dir]);
    const uint4 qa = __ldg(Q4), qb = __ldg(Q4 + 1);
    const int* off = dir == 0 ? off0 : off1;
    const int* cand = dir == 0 ? cand0 : cand1;
    const int beg = off[q], end = off[q + 1];
    const unsigned inf = (unsigned)(max_hamming + 1);
    unsigned bd = inf, sd = inf, bpos = 0, bidx = 0;
    for (int j = beg + lane; j < end; j += 32) {
        const int t = cand[j];
        if (t < 0 || t >= nT || (mT && !mT[t])) continue;
        const uint4 ta = __ldg(T4 + (size_t)t * 2), tb = __ldg(T4 + (size_t)t * 2 + 1);
        const unsigned d = __popc(qa.x ^ ta.x) + __popc(qa.y ^ ta.y) + __popc(qa.z ^ ta.z) + __popc(qa.w ^ ta.w) +
                           __popc(qb.x ^ tb.x) + __popc(qb.y ^ tb.y) + __popc(qb.z ^ tb.z) + __popc(qb.w ^ tb.w);
        if (d < bd) { sd = bd; bd = d; bpos = (unsigned)(j - beg); bidx = (unsigned)t; }
        else if (d < sd) sd = d;
    }
    unsigned key = (bd << 22) | min(bpos, 0x3FFFFFu);          // lists longer than 4M entries are rejected on the host
    unsigned kmin = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    if ((kmin >> 22) >= inf) return;                            // no candidate below maxHamming
    const bool winner = key == kmin && bd < inf;
    unsigned second = winner ? sd : bd;                         // everyone else's best competes for second place
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) second = min(second, __shfl_xor_sync(0xffffffffu, second, o));
    if (winner) {
        stats[((size_t)(2 * dir) + 0) * stride + q] = (bd << 16) | bidx;
        stats[((size_t)(2 * dir) + 1) * stride + q] = second >= inf ? kNoKey : (second << 16);
    }
}

__global__ void k_desc_distance(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, int n, int* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) d += __popc(a[(size_t)i * 8 + w] ^ b[(size_t)i * 8 + w]);
    out[i] = d;
}

} // namespace mage

using namespace mage;

struct mage_matcher_s {
    int max_desc = 0, max_pairs = 0;
    MatchJob* h_jobs = nullptr;     // pinned
    MatchJob* d_jobs = nullptr;
    unsigned* d_best = nullptr;     // [max_pairs][4][max_desc]: fwd best, fwd second, bwd best, bwd second
    uint8_t* d_stage = nullptr;     // single-pair host API staging: A, B, masks, matches, count
    mage_dmatch* h_out = nullptr;   // pinned staging for results (+ the count behind them)
    uint8_t* h_in = nullptr;        // pinned staging of the single-pair host API's descriptors (A | B, laid out like d_stage)
    int* h_count = nullptr;
    cudaStream_t own_stream = nullptr;
    // memo of the last device-batched job table (a video stream re-submits the same slot pairs every batch)
    const uint8_t* memo_desc = nullptr; const int* memo_counts = nullptr; size_t memo_stride = 0;
    std::vector<int> memo_a, memo_b;
};

extern "C" int mage_matcher_create(int max_descriptors, int max_pairs, mage_matcher_t* out)
{
    MAGE_REQUIRE(out && max_descriptors >= 1 && max_descriptors <= 65535 && max_pairs >= 1, MAGE_ERR_INVALID,
                 "mage_matcher_create: max_descriptors must be 1..65535 and max_pairs >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: the matcher has no CPU fallback"); return MAGE_ERR_CUDA; }
    mage_matcher_s* m = new mage_matcher_s();
    m->max_desc = max_descriptors; m->max_pairs = max_pairs;
    size_t stage = (size_t)max_descriptors * (32 + 32 + 1 + 1) + sizeof(mage_dmatch) * (size_t)max_descriptors + 64 + 256;
    cudaError_t e = cudaMallocHost(&m->h_jobs, sizeof(MatchJob) * max_pairs);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_jobs, sizeof(MatchJob) * max_pairs);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_best, sizeof(unsigned) * 4 * (size_t)max_pairs * max_descriptors);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_stage, stage);
    if (e == cudaSuccess) e = cudaMallocHost(&m->h_out, sizeof(mage_dmatch) * (size_t)max_descriptors + sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost(&m->h_in, (size_t)max_descriptors * 64);
    if (e == cudaSuccess) e = cudaMallocHost(&m->h_count, sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_match_dir, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmemBytes);      // per device
    if (e != cudaSuccess) { set_error("mage_matcher_create: %s", cudaGetErrorString(e)); mage_matcher_destroy(m); return MAGE_ERR_CUDA; }
    *out = m;
    return MAGE_OK;
}

extern "C" void mage_matcher_destroy(mage_matcher_t m)
{
    if (!m) return;
    if (m->h_jobs) cudaFreeHost(m->h_jobs);
    if (m->d_jobs) cudaFree(m->d_jobs);
    if (m->d_best) cudaFree(m->d_best);
    if (m->d_stage) cudaFree(m->d_stage);
    if (m->h_out) cudaFreeHost(m->h_out);
    if (m->h_in) cudaFreeHost(m->h_in);
    if (m->h_count) cudaFreeHost(m->h_count);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

static int match_launch(mage_matcher_t m, int n_pairs, int max_q, int max_t, int max_hamming, int min_diff, mage_dmatch* d_out, int out_cap,
                        int* d_counts, cudaStream_t s)
{
    MAGE_CUDA_TRY(cudaMemcpyAsync(m->d_jobs, m->h_jobs, sizeof(MatchJob) * n_pairs, cudaMemcpyHostToDevice, s));
    MAGE_CUDA_TRY(cudaMemsetAsync(m->d_best, 0xFF, sizeof(unsigned) * 4 * (size_t)n_pairs * m->max_desc, s));
    { ProfScope ps(PROF_MATCH_DIR, s); MAGE_CUDA_TRY(launch_match_dir(m->d_jobs, m->d_best, m->max_desc, max_q, max_t, n_pairs, max_hamming, s)); }
    { ProfScope ps(PROF_MATCH_EMIT, s); k_match_emit<<<n_pairs, kEmitThreads, 0, s>>>(m->d_jobs, m->d_best, m->max_desc, min_diff, d_out, out_cap, d_counts); }
    MAGE_CUDA_TRY(cudaGetLastError());
    return MAGE_OK;
}

extern "C" int mage_match_bf(mage_matcher_t m, const uint8_t* descA, int nA, const uint8_t* maskA, const uint8_t* descB, int nB,
                             const uint8_t* maskB, int max_hamming, int min_diff, mage_dmatch* out, int* count, void* stream)
{
    MAGE_REQUIRE(m && count && out, MAGE_ERR_INVALID, "mage_match_bf: null argument");
    MAGE_REQUIRE(nA >= 0 && nB >= 0 && nA <= m->max_desc && nB <= m->max_desc, MAGE_ERR_INVALID, "descriptor count exceeds matcher capacity %d", m->max_desc);
    *count = 0;
    // ref :72-77: nothing to match when either (masked) side is empty
    if (nA == 0 || nB == 0) return MAGE_OK;
    MAGE_REQUIRE(descA && descB, MAGE_ERR_INVALID, "mage_match_bf: null descriptors");
    cudaStream_t s = stream ? (cudaStream_t)stream : m->own_stream;
    const size_t N = (size_t)m->max_desc;
    uint8_t* dA = m->d_stage; uint8_t* dB = dA + N * 32; uint8_t* dmA = dB + N * 32; uint8_t* dmB = dmA + N;
    mage_dmatch* dOut = reinterpret_cast<mage_dmatch*>(m->d_stage + align_up(N * 66, 256));
    int* dCnt = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(dOut) + sizeof(mage_dmatch) * N);
    // both descriptor sets through ONE copy from pinned staging laid out like the device buffer (two copies from the caller's pageable
    // arrays are two staged transfers of the driver), the matches and their count back in one copy likewise
    memcpy(m->h_in, descA, (size_t)nA * 32);
    memcpy(m->h_in + N * 32, descB, (size_t)nB * 32);
    MAGE_CUDA_TRY(cudaMemcpyAsync(dA, m->h_in, N * 32 + (size_t)nB * 32, cudaMemcpyHostToDevice, s));
    if (maskA) MAGE_CUDA_TRY(cudaMemcpyAsync(dmA, maskA, nA, cudaMemcpyHostToDevice, s));
    if (maskB) MAGE_CUDA_TRY(cudaMemcpyAsync(dmB, maskB, nB, cudaMemcpyHostToDevice, s));
    m->memo_desc = nullptr; m->memo_a.clear(); m->memo_b.clear();
    MatchJob& j = m->h_jobs[0];
    j.desc[0] = dA; j.desc[1] = dB; j.mask[0] = maskA ? dmA : nullptr; j.mask[1] = maskB ? dmB : nullptr;
    j.count_ptr[0] = j.count_ptr[1] = nullptr; j.count[0] = nA; j.count[1] = nB; j.cap = m->max_desc;
    int rc = match_launch(m, 1, nA, nB, max_hamming, min_diff, dOut, m->max_desc, dCnt, s);
    if (rc != MAGE_OK) return rc;
    MAGE_CUDA_TRY(cudaMemcpyAsync(m->h_out, dOut, sizeof(mage_dmatch) * N + sizeof(int), cudaMemcpyDeviceToHost, s));      // [matches | count]
    MAGE_CUDA_TRY(cudaStreamSynchronize(s));
    *count = *reinterpret_cast<const int*>(reinterpret_cast<const uint8_t*>(m->h_out) + sizeof(mage_dmatch) * N);
    memcpy(out, m->h_out, sizeof(mage_dmatch) * (size_t)(*count));
    return MAGE_OK;
}

// Registers a table of device-resident pairs once (uploaded synchronously); mage_match_run_jobs then launches any
// sub-range of it without touching the host table again (a video stream re-submits the same slot pairs every batch).
extern "C" int mage_matcher_set_jobs_device(mage_matcher_t m, const uint8_t* d_desc, const int* d_counts, size_t slot_stride,
                                            const int* a_index, const int* b_index, int n_pairs)
{
    MAGE_REQUIRE(m && d_desc && d_counts && a_index && b_index, MAGE_ERR_INVALID, "mage_matcher_set_jobs_device: null argument");
    MAGE_REQUIRE(n_pairs >= 1 && n_pairs <= m->max_pairs, MAGE_ERR_INVALID, "n_pairs %d exceeds matcher capacity %d", n_pairs, m->max_pairs);
    MAGE_REQUIRE(slot_stride % 16 == 0 && ((uintptr_t)d_desc % 16) == 0, MAGE_ERR_INVALID, "d_desc and slot_stride must be 16-byte aligned");
    const bool same = m->memo_desc == d_desc && m->memo_counts == d_counts && m->memo_stride == slot_stride &&
                      (int)m->memo_a.size() == n_pairs && std::equal(a_index, a_index + n_pairs, m->memo_a.begin()) &&
                      std::equal(b_index, b_index + n_pairs, m->memo_b.begin());
    if (same) return MAGE_OK;
    MAGE_CUDA_TRY(cudaDeviceSynchronize());          // the table may still be in use by earlier launches
    m->memo_desc = d_desc; m->memo_counts = d_counts; m->memo_stride = slot_stride;
    m->memo_a.assign(a_index, a_index + n_pairs); m->memo_b.assign(b_index, b_index + n_pairs);
    for (int p = 0; p < n_pairs; p++) {
        MatchJob& j = m->h_jobs[p];
        j.desc[0] = d_desc + (size_t)a_index[p] * slot_stride; j.desc[1] = d_desc + (size_t)b_index[p] * slot_stride;
        j.mask[0] = j.mask[1] = nullptr;
        j.count_ptr[0] = d_counts + a_index[p]; j.count_ptr[1] = d_counts + b_index[p];
        j.count[0] = j.count[1] = 0; j.cap = (int)std::min<size_t>(m->max_desc, slot_stride / 32);
    }
    MAGE_CUDA_TRY(cudaMemcpy(m->d_jobs, m->h_jobs, sizeof(MatchJob) * n_pairs, cudaMemcpyHostToDevice));
    return MAGE_OK;
}

extern "C" int mage_match_run_jobs(mage_matcher_t m, int first_pair, int n_pairs, int max_hamming, int min_diff, mage_dmatch* d_matches,
                                   int capacity, int* d_match_counts, void* stream)
{
    MAGE_REQUIRE(m && d_matches && d_match_counts && first_pair >= 0 && n_pairs >= 1 && first_pair + n_pairs <= (int)m->memo_a.size(),
                 MAGE_ERR_INVALID, "mage_match_run_jobs: pair range outside the registered table");
    cudaStream_t s = (cudaStream_t)stream;
    const int cap = m->h_jobs[first_pair].cap;
    unsigned* best = m->d_best + (size_t)first_pair * 4 * m->max_desc;
    MAGE_CUDA_TRY(cudaMemsetAsync(best, 0xFF, sizeof(unsigned) * 4 * (size_t)n_pairs * m->max_desc, s));
    { ProfScope ps(PROF_MATCH_DIR, s); MAGE_CUDA_TRY(launch_match_dir(m->d_jobs + first_pair, best, m->max_desc, cap, cap, n_pairs, max_hamming, s)); }
    { ProfScope ps(PROF_MATCH_EMIT, s); k_match_emit<<<n_pairs, kEmitThreads, 0, s>>>(m->d_jobs + first_pair, best, m->max_desc, min_diff, d_matches, capacity, d_match_counts); }
    MAGE_CUDA_TRY(cudaGetLastError());
    return MAGE_OK;
}

extern "C" int mage_match_bf_device(mage_matcher_t m, const uint8_t* d_desc, const int* d_counts, size_t slot_stride,
                                    const int* a_index, const int* b_index, int n_pairs, int max_hamming, int min_diff,
                                    mage_dmatch* d_matches, int capacity, int* d_match_counts, void* stream)
{
    MAGE_REQUIRE(capacity >= 1, MAGE_ERR_INVALID, "capacity must be >= 1");
    int rc = mage_matcher_set_jobs_device(m, d_desc, d_counts, slot_stride, a_index, b_index, n_pairs);
    if (rc != MAGE_OK) return rc;
    return mage_match_run_jobs(m, 0, n_pairs, max_hamming, min_diff, d_matches, capacity, d_match_counts, stream);
}

extern "C" int mage_indexed_match(mage_matcher_t m, const uint8_t* descA, int nA, const uint8_t* maskA, const uint8_t* descB, int nB,
                                  const uint8_t* maskB, const int* a2b_offsets, const int* a2b_candidates, const int* b2a_offsets,
                                  const int* b2a_candidates, int max_hamming, int min_diff, mage_dmatch* out, int* count, void* stream)
{
    MAGE_REQUIRE(m && count && out, MAGE_ERR_INVALID, "mage_indexed_match: null argument");
    MAGE_REQUIRE(nA >= 0 && nB >= 0 && nA <= m->max_desc && nB <= m->max_desc, MAGE_ERR_INVALID, "descriptor count exceeds matcher capacity %d", m->max_desc);
    *count = 0;
    if (nA == 0 || nB == 0) return MAGE_OK;                     // ref :208: imageAMaskCount == 0 || imageBMaskCount == 0
    MAGE_REQUIRE(descA && descB && a2b_offsets && b2a_offsets, MAGE_ERR_INVALID, "mage_indexed_match: null descriptors / offsets");
    MAGE_REQUIRE(max_hamming >= 0 && max_hamming <= 256, MAGE_ERR_INVALID, "mage_indexed_match: max_hamming outside 0..256");
    const int nab = a2b_offsets[nA], nba = b2a_offsets[nB];
    MAGE_REQUIRE(a2b_offsets[0] == 0 && b2a_offsets[0] == 0 && nab >= 0 && nba >= 0, MAGE_ERR_INVALID, "mage_indexed_match: malformed CSR offsets");
    for (int i = 0; i < nA; i++) MAGE_REQUIRE(a2b_offsets[i + 1] >= a2b_offsets[i] && a2b_offsets[i + 1] - a2b_offsets[i] < (1 << 22), MAGE_ERR_INVALID, "mage_indexed_match: a2b offsets must ascend, lists < 4M entries");
    for (int i = 0; i < nB; i++) MAGE_REQUIRE(b2a_offsets[i + 1] >= b2a_offsets[i] && b2a_offsets[i + 1] - b2a_offsets[i] < (1 << 22), MAGE_ERR_INVALID, "mage_indexed_match: b2a offsets must ascend, lists < 4M entries");
    MAGE_REQUIRE((nab == 0 || a2b_candidates) && (nba == 0 || b2a_candidates), MAGE_ERR_INVALID, "mage_indexed_match: null candidate list");
    cudaStream_t s = stream ? (cudaStream_t)stream : m->own_stream;
    const size_t N = (size_t)m->max_desc;
    uint8_t* dA = m->d_stage; uint8_t* dB = dA + N * 32; uint8_t* dmA = dB + N * 32; uint8_t* dmB = dmA + N;
    mage_dmatch* dOut = reinterpret_cast<mage_dmatch*>(m->d_stage + align_up(N * 66, 256));
    int* dCnt = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(dOut) + sizeof(mage_dmatch) * N);
    // candidate lists are per call and of arbitrary size: stream-ordered scratch
    int* d_csr = nullptr;
    const size_t n_csr = (size_t)(nA + 1) + (size_t)(nB + 1) + (size_t)nab + (size_t)nba;
    MAGE_CUDA_TRY(cudaMallocAsync(&d_csr, sizeof(int) * n_csr, s));
    int* d_off0 = d_csr; int* d_off1 = d_off0 + nA + 1; int* d_c0 = d_off1 + nB + 1; int* d_c1 = d_c0 + nab;
    cudaError_t e = cudaMemcpyAsync(dA, descA, (size_t)nA * 32, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dB, descB, (size_t)nB * 32, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && maskA) e = cudaMemcpyAsync(dmA, maskA, nA, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && maskB) e = cudaMemcpyAsync(dmB, maskB, nB, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off0, a2b_offsets, sizeof(int) * (nA + 1), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_off1, b2a_offsets, sizeof(int) * (nB + 1), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && nab) e = cudaMemcpyAsync(d_c0, a2b_candidates, sizeof(int) * nab, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && nba) e = cudaMemcpyAsync(d_c1, b2a_candidates, sizeof(int) * nba, cudaMemcpyHostToDevice, s);
    m->memo_desc = nullptr; m->memo_a.clear(); m->memo_b.clear();
    MatchJob& j = m->h_jobs[0];
    j.desc[0] = dA; j.desc[1] = dB; j.mask[0] = maskA ? dmA : nullptr; j.mask[1] = maskB ? dmB : nullptr;
    j.count_ptr[0] = j.count_ptr[1] = nullptr; j.count[0] = nA; j.count[1] = nB; j.cap = m->max_desc;
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_jobs, m->h_jobs, sizeof(MatchJob), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_best, 0xFF, sizeof(unsigned) * 4 * (size_t)m->max_desc, s);
    if (e == cudaSuccess) {
        const int nmax = nA > nB ? nA : nB;
        k_indexed_best<<<dim3(div_up(nmax, 8), 2), 256, 0, s>>>(m->d_jobs, d_off0, d_c0, d_off1, d_c1, m->d_best, m->max_desc, max_hamming);
        k_match_emit<<<1, kEmitThreads, 0, s>>>(m->d_jobs, m->d_best, m->max_desc, min_diff, dOut, m->max_desc, dCnt);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->h_count, dCnt, sizeof(int), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->h_out, dOut, sizeof(mage_dmatch) * (size_t)nA, cudaMemcpyDeviceToHost, s);
    cudaFreeAsync(d_csr, s);
    cudaError_t e2 = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = e2;
    MAGE_CUDA_TRY(e);
    *count = *m->h_count;
    memcpy(out, m->h_out, sizeof(mage_dmatch) * (size_t)(*count));
    return MAGE_OK;
}

extern "C" int mage_descriptor_distance_device(const uint8_t* d_a, const uint8_t* d_b, int n, int* d_out, void* stream)
{
    MAGE_REQUIRE(d_a && d_b && d_out && n >= 0, MAGE_ERR_INVALID, "mage_descriptor_distance_device: bad argument");
    if (n == 0) return MAGE_OK;
    k_desc_distance<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t*>(d_a), reinterpret_cast<const uint32_t*>(d_b), n, d_out);
    MAGE_CUDA_TRY(cudaGetLastError());
    return MAGE_OK;
}
