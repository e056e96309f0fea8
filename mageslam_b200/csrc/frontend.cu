// Per-frame front-end for a video stream: ORB extract of a batch of frames + brute-force match of every frame against its
// predecessor, the call sequence of the reference's tracking front-end (ref Tasks/ImageAnalyzer.cpp:119 -> Process,
// Tracking/MapInitialization.cpp:585 -> Match) in batched form. Built only on the C ABI of orb.cu / match.cu.
//
// Host variant: frames come from (pinned) host memory in chunks; chunk k+1 is uploaded on a copy stream while chunk k is
// processed on the compute stream and chunk k-1's results are downloaded on a third stream (events order the hand-offs).
// Device variant: frames already in HBM, results stay in HBM, fully asynchronous.
#include "common.cuh"

#include <algorithm>
#include <vector>

using namespace mage;

struct mage_frontend_s {
    mage_orb_t orb = nullptr;
    mage_matcher_t matcher = nullptr;
    int width = 0, height = 0, pitch = 0, batch = 0, chunk = 0, cap = 0, max_hamming = 30, min_diff = 1;
    // device slots: slot 0 = last frame of the previous call, slots 1..batch = frames of an even call, batch+1..2*batch = of an odd call
    // (two result sets, so the compute of call j+1 never waits for the download of call j)
    int cur = 0;                                   // result set of the most recent call
    uint8_t* d_images = nullptr;
    mage_keypoint* d_kps = nullptr;
    uint8_t* d_desc = nullptr;
    int* d_counts = nullptr;
    mage_dmatch* d_matches = nullptr;
    int* d_match_counts = nullptr;
    cudaStream_t s_copy = nullptr, s_compute = nullptr, s_match = nullptr, s_out = nullptr;
    // the matcher runs on its own stream: Match(call j) only needs the descriptors of call j, so it executes under the extraction of
    // the next chunk / call (different pipes: popcount vs packed min/max). ev_x = extracted, ev_m = matched (per chunk, per result set)
    cudaEvent_t ev_dev_x[2] = {nullptr, nullptr}, ev_dev_m[2] = {nullptr, nullptr};       // device-resident path, per result set
    bool dev_used[2] = {false, false};
    long long dev_calls = 0;
    // two calls can be in flight (mage_frontend_submit / _wait): call j stages its frames in image buffer j % 2; per chunk,
    // ev_in = uploaded, ev_done = computed, ev_out = results delivered to the caller's host buffers
    struct Flight { std::vector<cudaEvent_t> ev_in, ev_done, ev_match, ev_out; cudaEvent_t ev_all_matched = nullptr; int nch = 0; bool pending = false; };
    Flight flight[2];
    long long submitted = 0, waited = 0;
    cudaEvent_t ev_prev = nullptr;
    std::vector<int> a_idx, b_idx;
    bool has_prev = false;
};

extern "C" void mage_frontend_destroy(mage_frontend_s* f);

extern "C" int mage_frontend_create(const mage_orb_params* p, int width, int height, int batch, int chunk, int max_hamming, int min_diff,
                                    mage_frontend_s** out)
{
    MAGE_REQUIRE(p && out && batch >= 1, MAGE_ERR_INVALID, "mage_frontend_create: bad argument");
    if (chunk <= 0 || chunk > batch) chunk = batch;
    mage_frontend_s* f = new mage_frontend_s();
    f->width = width; f->height = height; f->batch = batch; f->chunk = chunk; f->max_hamming = max_hamming; f->min_diff = min_diff;
    int rc = mage_orb_create(p, width, height, chunk, &f->orb);
    if (rc != MAGE_OK) { delete f; return rc; }
    int nf[16]; int sum = 0;
    mage_orb_level_info(f->orb, nullptr, nullptr, nullptr, nf);
    for (unsigned l = 0; l < p->nlevels; l++) sum += nf[l];
    f->cap = std::max((int)p->nfeatures, sum);
    rc = mage_matcher_create(std::min(f->cap, 65535), 2 * batch, &f->matcher);
    if (rc != MAGE_OK) { mage_frontend_destroy(f); return rc; }
    const size_t B1 = 2 * (size_t)batch + 1;
    f->pitch = (int)align_up((size_t)width, 16);          // staging rows: 16-byte aligned so every frame base is too
    cudaError_t e = cudaMalloc(&f->d_images, (size_t)f->pitch * height * batch * 2);       // double-buffered staging
    if (e == cudaSuccess) e = cudaMalloc(&f->d_kps, sizeof(mage_keypoint) * f->cap * B1);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_desc, (size_t)32 * f->cap * B1);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_counts, sizeof(int) * B1);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_matches, sizeof(mage_dmatch) * f->cap * (size_t)batch * 2);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_match_counts, sizeof(int) * batch * 2);
    if (e == cudaSuccess) e = cudaMemset(f->d_counts, 0, sizeof(int) * B1);
    if (e == cudaSuccess) e = cudaMemset(f->d_match_counts, 0, sizeof(int) * batch * 2);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->s_copy, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->s_compute, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->s_out, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->s_match, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&f->ev_dev_x[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_dev_m[i], cudaEventDisableTiming);
    }
    const int nchunks = div_up(batch, chunk);
    for (auto& fl : f->flight) {
        fl.ev_in.assign(nchunks, nullptr); fl.ev_done.assign(nchunks, nullptr); fl.ev_match.assign(nchunks, nullptr); fl.ev_out.assign(nchunks, nullptr);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fl.ev_all_matched, cudaEventDisableTiming);
        for (int i = 0; i < nchunks && e == cudaSuccess; i++) {
            e = cudaEventCreateWithFlags(&fl.ev_in[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fl.ev_done[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fl.ev_out[i], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fl.ev_match[i], cudaEventDisableTiming);
        }
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_prev, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("mage_frontend_create: %s", cudaGetErrorString(e)); mage_frontend_destroy(f); return MAGE_ERR_CUDA; }
    f->a_idx.resize(2 * batch); f->b_idx.resize(2 * batch);
    for (int s = 0; s < 2; s++)
        for (int i = 0; i < batch; i++) {                                               // frame i (query) vs frame i-1 (train); slot 0 precedes frame 0
            f->a_idx[s * batch + i] = s * batch + i + 1;
            f->b_idx[s * batch + i] = i == 0 ? 0 : s * batch + i;
        }
    rc = mage_matcher_set_jobs_device(f->matcher, f->d_desc, f->d_counts, (size_t)32 * f->cap, f->a_idx.data(), f->b_idx.data(), 2 * batch);
    if (rc != MAGE_OK) { mage_frontend_destroy(f); return rc; }
    *out = f;
    return MAGE_OK;
}

extern "C" void mage_frontend_destroy(mage_frontend_s* f)
{
    if (!f) return;
    cudaDeviceSynchronize();
    if (f->orb) mage_orb_destroy(f->orb);
    if (f->matcher) mage_matcher_destroy(f->matcher);
    cudaFree(f->d_images); cudaFree(f->d_kps); cudaFree(f->d_desc); cudaFree(f->d_counts); cudaFree(f->d_matches); cudaFree(f->d_match_counts);
    for (auto& fl : f->flight) {
        for (auto e : fl.ev_in) if (e) cudaEventDestroy(e);
        for (auto e : fl.ev_done) if (e) cudaEventDestroy(e);
        for (auto e : fl.ev_out) if (e) cudaEventDestroy(e);
        for (auto e : fl.ev_match) if (e) cudaEventDestroy(e);
        if (fl.ev_all_matched) cudaEventDestroy(fl.ev_all_matched);
    }
    for (int i = 0; i < 2; i++) { if (f->ev_dev_x[i]) cudaEventDestroy(f->ev_dev_x[i]); if (f->ev_dev_m[i]) cudaEventDestroy(f->ev_dev_m[i]); }
    if (f->s_match) cudaStreamDestroy(f->s_match);
    if (f->ev_prev) cudaEventDestroy(f->ev_prev);
    if (f->s_copy) cudaStreamDestroy(f->s_copy);
    if (f->s_compute) cudaStreamDestroy(f->s_compute);
    if (f->s_out) cudaStreamDestroy(f->s_out);
    delete f;
}

// new sequence: forget the previous frame (its slot gets an empty descriptor set => no matches for the next first frame)
extern "C" int mage_frontend_reset(mage_frontend_s* f)
{
    MAGE_REQUIRE(f, MAGE_ERR_INVALID, "null handle");
    MAGE_CUDA_TRY(cudaDeviceSynchronize());
    MAGE_CUDA_TRY(cudaMemset(f->d_counts, 0, sizeof(int)));
    f->has_prev = false;
    for (auto& fl : f->flight) fl.pending = false;
    f->waited = f->submitted;
    f->dev_used[0] = f->dev_used[1] = false;
    return MAGE_OK;
}

// extract of frames [c0, c1) whose pixels are at d_img (frame stride fs, row stride st) into result set `set`, on stream s
static int frontend_extract(mage_frontend_s* f, int set, const uint8_t* d_img, int st, size_t fs, int c0, int c1, cudaStream_t s)
{
    const size_t cap = (size_t)f->cap;
    const int base = set * f->batch;               // first result slot (minus one) and first match job of this result set
    return mage_orb_extract_device(f->orb, d_img, c1 - c0, f->width, f->height, st, fs, f->d_kps + cap * (base + c0 + 1),
                                   f->d_desc + 32 * cap * (base + c0 + 1), f->cap, f->d_counts + base + c0 + 1, s);
}
// Match(frame i, frame i-1) for the frames [c0, c1) of result set `set`, on stream s
static int frontend_match(mage_frontend_s* f, int set, int c0, int c1, cudaStream_t s)
{
    const size_t cap = (size_t)f->cap;
    const int base = set * f->batch;
    return mage_match_run_jobs(f->matcher, base + c0, c1 - c0, f->max_hamming, f->min_diff, f->d_matches + cap * (base + c0), f->cap,
                               f->d_match_counts + base + c0, s);
}

// keep the last frame of this call as "previous" for the next one. Runs on the MATCH stream, after the call's matches (which read
// the old slot 0) and before the next call's (which read the new one).
static int frontend_roll(mage_frontend_s* f, int set, int n, cudaStream_t s)
{
    const size_t cap = (size_t)f->cap, last = (size_t)set * f->batch + n;
    MAGE_CUDA_TRY(cudaMemcpyAsync(f->d_desc, f->d_desc + 32 * cap * last, 32 * cap, cudaMemcpyDeviceToDevice, s));
    MAGE_CUDA_TRY(cudaMemcpyAsync(f->d_kps, f->d_kps + cap * last, sizeof(mage_keypoint) * cap, cudaMemcpyDeviceToDevice, s));
    MAGE_CUDA_TRY(cudaMemcpyAsync(f->d_counts, f->d_counts + last, sizeof(int), cudaMemcpyDeviceToDevice, s));
    f->has_prev = true;
    f->cur = set;
    return MAGE_OK;
}

// Device-resident frames. The extraction is enqueued on the caller's stream, the matches on the handle's match stream (ordered by
// events), alternating between the two result sets, so Match(call j) overlaps the extraction of call j+1. The results of a call are
// complete once mage_frontend_join(f, stream) has been enqueued and `stream` has reached it (or after a device synchronisation).
extern "C" int mage_frontend_process_device(mage_frontend_s* f, const uint8_t* d_images, int n, int stride, size_t frame_stride, void* stream)
{
    MAGE_REQUIRE(f && d_images && n >= 1 && n <= f->batch, MAGE_ERR_INVALID, "mage_frontend_process_device: bad argument");
    cudaStream_t s = (cudaStream_t)stream;       // NULL = the default stream, like any CUDA API
    const int set = (int)(f->dev_calls & 1);
    cudaStream_t sm = prof_enabled() ? s : f->s_match;       // per-kernel timing (mage_profile_*) keeps everything on one stream
    if (f->dev_used[set]) MAGE_CUDA_TRY(cudaStreamWaitEvent(s, f->ev_dev_m[set], 0));      // the matches of two calls ago read this result set
    for (int c0 = 0; c0 < n; c0 += f->chunk) {
        int c1 = std::min(n, c0 + f->chunk);
        int rc = frontend_extract(f, set, d_images + (size_t)c0 * frame_stride, stride, frame_stride, c0, c1, s);
        if (rc != MAGE_OK) return rc;
        MAGE_CUDA_TRY(cudaEventRecord(f->ev_dev_x[set], s));
        MAGE_CUDA_TRY(cudaStreamWaitEvent(sm, f->ev_dev_x[set], 0));
        rc = frontend_match(f, set, c0, c1, sm);
        if (rc != MAGE_OK) return rc;
    }
    int rc = frontend_roll(f, set, n, sm);
    if (rc != MAGE_OK) return rc;
    MAGE_CUDA_TRY(cudaEventRecord(f->ev_dev_m[set], sm));
    f->dev_used[set] = true;
    f->dev_calls++;
    return MAGE_OK;
}

// makes `stream` wait for everything mage_frontend_process_device has enqueued so far (the matches run on another stream)
extern "C" int mage_frontend_join(mage_frontend_s* f, void* stream)
{
    MAGE_REQUIRE(f, MAGE_ERR_INVALID, "null handle");
    for (int i = 0; i < 2; i++) if (f->dev_used[i]) MAGE_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, f->ev_dev_m[i], 0));
    return MAGE_OK;
}

extern "C" int mage_frontend_device_buffers(mage_frontend_s* f, mage_keypoint** d_kps, uint8_t** d_desc, int** d_counts, mage_dmatch** d_matches,
                                            int** d_match_counts, int* capacity)
{
    MAGE_REQUIRE(f, MAGE_ERR_INVALID, "null handle");
    const size_t cap = (size_t)f->cap, base = (size_t)f->cur * f->batch;            // the result set of the most recent call
    if (d_kps) *d_kps = f->d_kps + cap * (base + 1); if (d_desc) *d_desc = f->d_desc + 32 * cap * (base + 1); if (d_counts) *d_counts = f->d_counts + base + 1;
    if (d_matches) *d_matches = f->d_matches + cap * base; if (d_match_counts) *d_match_counts = f->d_match_counts + base; if (capacity) *capacity = f->cap;
    return MAGE_OK;
}

// Asynchronous host variant: enqueue upload -> compute -> download of one call and return. Call j uses staging buffer j % 2, so
// its upload overlaps the compute of call j-1 and the download of call j-1 overlaps its compute; the hazards are per chunk:
//   upload(j, k)  waits for compute(j-2, k)  (same staging buffer)
//   compute(j, k) waits for download(j-2, k) (same device result set; there are two)
static int frontend_submit(mage_frontend_s* f, const uint8_t* images, int n, int stride, size_t frame_stride, mage_keypoint* kps,
                           uint8_t* desc, int* counts, mage_dmatch* matches, int* match_counts)
{
    const size_t cap = (size_t)f->cap, W = (size_t)f->width, H = (size_t)f->height, PT = (size_t)f->pitch;
    const int nch = div_up(n, f->chunk);
    const int b = (int)(f->submitted & 1);
    mage_frontend_s::Flight& me = f->flight[b];
    uint8_t* stage = f->d_images + PT * H * (size_t)f->batch * b;
    for (int k = 0; k < nch; k++) {
        const int c0 = k * f->chunk, c1 = std::min(n, c0 + f->chunk);
        if (me.pending && k < me.nch) MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_copy, me.ev_done[k], 0));            // call j-2 read this chunk's frames
        if ((size_t)stride == PT && frame_stride == PT * H)
            MAGE_CUDA_TRY(cudaMemcpyAsync(stage + PT * H * c0, images + frame_stride * c0, PT * H * (c1 - c0), cudaMemcpyHostToDevice, f->s_copy));
        else
            for (int i = c0; i < c1; i++)
                MAGE_CUDA_TRY(cudaMemcpy2DAsync(stage + PT * H * i, PT, images + frame_stride * i, stride, W, H, cudaMemcpyHostToDevice, f->s_copy));
        MAGE_CUDA_TRY(cudaEventRecord(me.ev_in[k], f->s_copy));
        MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_compute, me.ev_in[k], 0));
        if (me.pending) {                                                                                           // call j-2 delivered from this result set
            const int ko = std::min(k, me.nch - 1);
            MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_compute, me.ev_out[ko], 0));
            if (k == nch - 1) for (int q = ko + 1; q < me.nch; q++) MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_compute, me.ev_out[q], 0));
        }
        if (k == 0 && me.pending) MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_compute, me.ev_all_matched, 0));           // matches of call j-2 read this set
        int rc = frontend_extract(f, b, stage + PT * H * c0, f->pitch, PT * H, c0, c1, f->s_compute);
        if (rc != MAGE_OK) return rc;
        MAGE_CUDA_TRY(cudaEventRecord(me.ev_done[k], f->s_compute));
        MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_match, me.ev_done[k], 0));
        rc = frontend_match(f, b, c0, c1, f->s_match);
        if (rc != MAGE_OK) return rc;
        MAGE_CUDA_TRY(cudaEventRecord(me.ev_match[k], f->s_match));
        MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_out, me.ev_done[k], 0));
        const int m = c1 - c0;
        const size_t r0 = (size_t)b * f->batch + c0;                      // first result slot (minus one) of the chunk
        MAGE_CUDA_TRY(cudaMemcpyAsync(kps + cap * c0, f->d_kps + cap * (r0 + 1), sizeof(mage_keypoint) * cap * m, cudaMemcpyDeviceToHost, f->s_out));
        MAGE_CUDA_TRY(cudaMemcpyAsync(desc + 32 * cap * c0, f->d_desc + 32 * cap * (r0 + 1), 32 * cap * m, cudaMemcpyDeviceToHost, f->s_out));
        MAGE_CUDA_TRY(cudaMemcpyAsync(counts + c0, f->d_counts + r0 + 1, sizeof(int) * m, cudaMemcpyDeviceToHost, f->s_out));
        MAGE_CUDA_TRY(cudaStreamWaitEvent(f->s_out, me.ev_match[k], 0));
        MAGE_CUDA_TRY(cudaMemcpyAsync(matches + cap * c0, f->d_matches + cap * r0, sizeof(mage_dmatch) * cap * m, cudaMemcpyDeviceToHost, f->s_out));
        MAGE_CUDA_TRY(cudaMemcpyAsync(match_counts + c0, f->d_match_counts + r0, sizeof(int) * m, cudaMemcpyDeviceToHost, f->s_out));
        MAGE_CUDA_TRY(cudaEventRecord(me.ev_out[k], f->s_out));
    }
    me.nch = nch; me.pending = true;
    f->submitted++;
    int rc = frontend_roll(f, b, n, f->s_match);
    if (rc != MAGE_OK) return rc;
    MAGE_CUDA_TRY(cudaEventRecord(me.ev_all_matched, f->s_match));
    return MAGE_OK;
}

extern "C" int mage_frontend_submit(mage_frontend_s* f, const uint8_t* images, int n, int stride, size_t frame_stride, mage_keypoint* kps,
                                    uint8_t* desc, int* counts, mage_dmatch* matches, int* match_counts)
{
    MAGE_REQUIRE(f && images && kps && desc && counts && matches && match_counts && n >= 1 && n <= f->batch, MAGE_ERR_INVALID,
                 "mage_frontend_submit: bad argument");
    MAGE_REQUIRE(stride >= f->width, MAGE_ERR_INVALID, "stride smaller than width");
    MAGE_REQUIRE(f->submitted - f->waited < 2, MAGE_ERR_INVALID, "two calls are already in flight: mage_frontend_wait first");
    return frontend_submit(f, images, n, stride, frame_stride, kps, desc, counts, matches, match_counts);
}

extern "C" int mage_frontend_wait(mage_frontend_s* f)
{
    MAGE_REQUIRE(f, MAGE_ERR_INVALID, "null handle");
    MAGE_REQUIRE(f->submitted > f->waited, MAGE_ERR_INVALID, "mage_frontend_wait: nothing in flight");
    mage_frontend_s::Flight& fl = f->flight[(int)(f->waited & 1)];
    MAGE_CUDA_TRY(cudaEventSynchronize(fl.ev_out[fl.nch - 1]));          // downloads of a call are enqueued in chunk order on one stream
    f->waited++;
    if (f->submitted == f->waited) MAGE_CUDA_TRY(cudaStreamSynchronize(f->s_match));        // nothing left in flight: the roll copy too
    return MAGE_OK;
}

extern "C" int mage_frontend_process(mage_frontend_s* f, const uint8_t* images, int n, int stride, size_t frame_stride, mage_keypoint* kps,
                                     uint8_t* desc, int* counts, mage_dmatch* matches, int* match_counts)
{
    MAGE_REQUIRE(f && images && kps && desc && counts && matches && match_counts && n >= 1 && n <= f->batch, MAGE_ERR_INVALID,
                 "mage_frontend_process: bad argument");
    MAGE_REQUIRE(stride >= f->width, MAGE_ERR_INVALID, "stride smaller than width");
    while (f->submitted > f->waited) { int rc = mage_frontend_wait(f); if (rc != MAGE_OK) return rc; }
    int rc = frontend_submit(f, images, n, stride, frame_stride, kps, desc, counts, matches, match_counts);
    if (rc != MAGE_OK) return rc;
    return mage_frontend_wait(f);
}

extern "C" int mage_frontend_capacity(mage_frontend_s* f) { return f ? f->cap : 0; }
