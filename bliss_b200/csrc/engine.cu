// engine.cu — the C-ABI of blx.h: device memory, streams, batching and kernel sequencing.
//
// Per chunk of songs the engine enqueues
//     memset(hist, stats) -> pass1 -> epilogue -> envelope -> logcomp -> tail
// on one stream. The host-buffer entry points split a batch into chunks that fit a device
// staging buffer and alternate between two slots, so the host->device copy of chunk i+1
// (copy stream) overlaps the kernels of chunk i (compute stream).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "blx_common.cuh"
#include "kernels.h"
#include "../../include/blx_resample.h"

using namespace blx;

// ---------------------------------------------------------------- error reporting
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(BLX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char *blx_last_error(void) { return g_err; }
// for the other translation units of the library (multi.cu)
extern "C" int blx_set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---------------------------------------------------------------- grow-only device buffer
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = (bytes + (1u << 20) - 1) & ~((size_t)(1u << 20) - 1);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct Slot {
    DevBuf pcm, songs, partials, hist, stats, norm, energy, xlog, q, results, freq;
    SongDesc *h_songs = nullptr; // pinned
    size_t h_songs_cap = 0;
    blx_result *h_results = nullptr; // pinned
    size_t h_results_cap = 0;
    cudaEvent_t copied = nullptr, done = nullptr, env_done = nullptr, p1_done = nullptr, ep_done = nullptr;
    bool busy = false;
    int n_songs = 0;
    int first_song = 0; // index in the caller's batch
};

struct ProfRec {
    int kid;
    cudaEvent_t a, b;
};

struct blx_engine {
    int device = 0;
    // compute: pass 1, epilogue, envelope (and everything else); tail: log compression + the sequential tail of a
    // (sub-)batch at high priority, so that its few CTAs slip in between the envelope CTAs of the next one
    cudaStream_t compute = nullptr, copy = nullptr, tail = nullptr;
    cudaEvent_t fence = nullptr, joined_work = nullptr, joined_tail = nullptr;
    int sub_batch = 512;  // songs per kernel sequence of the device-resident entry points
    int next_dev_slot = 0;
    float *d_hann = nullptr;
    float2 *d_tw1f = nullptr, *d_tw2f = nullptr;
    double2 *d_tw1d = nullptr, *d_tw2d = nullptr;
    static constexpr int kSlots = 4; // sub-batches in flight (the host-buffer path alternates between the first two)
    Slot slot[kSlots];
    int next_slot = 0;
    size_t chunk_bytes = (size_t)1 << 30;
    unsigned debug = 0;
    bool prof = false;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> ev_pool;
    float prof_ms[BLX_K_COUNT] = {0};
    int prof_n[BLX_K_COUNT] = {0};
    long long launches = 0;
    DevBuf scratch_a, scratch_b, scratch_c, scratch_near;
};

static const char *kKernelNames[BLX_K_COUNT] = {"pass1_kernel", "epilogue_kernel", "envelope_kernel", "tail_kernel",
                                                "distance_kernel"};
extern "C" const char *blx_kernel_name(int id) { return (id >= 0 && id < BLX_K_COUNT) ? kKernelNames[id] : "?"; }

// ---------------------------------------------------------------- profiling helpers
static cudaEvent_t ev_get(blx_engine *e) {
    if (!e->ev_pool.empty()) {
        cudaEvent_t ev = e->ev_pool.back();
        e->ev_pool.pop_back();
        return ev;
    }
    cudaEvent_t ev = nullptr;
    cudaEventCreate(&ev);
    return ev;
}
struct ProfScope {
    blx_engine *e;
    int kid;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(blx_engine *e_, int kid_, cudaStream_t st_) : e(e_), kid(kid_), st(st_) {
        e->launches++;
        if (e->prof) {
            a = ev_get(e);
            b = ev_get(e);
            cudaEventRecord(a, st);
        }
    }
    ~ProfScope() {
        if (e->prof) {
            cudaEventRecord(b, st);
            e->prof_pending.push_back({kid, a, b});
        }
    }
};

static int prof_drain(blx_engine *e) {
    for (auto &r : e->prof_pending) {
        CK(cudaEventSynchronize(r.b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, r.a, r.b));
        e->prof_ms[r.kid] += ms;
        e->prof_n[r.kid] += 1;
        e->ev_pool.push_back(r.a);
        e->ev_pool.push_back(r.b);
    }
    e->prof_pending.clear();
    return BLX_OK;
}

extern "C" int blx_profile_enable(blx_engine *e, int on) {
    if (!e) return fail(BLX_ERR_ARG, "null engine");
    e->prof = on != 0;
    return BLX_OK;
}
extern "C" int blx_profile_reset(blx_engine *e) {
    if (!e) return fail(BLX_ERR_ARG, "null engine");
    int rc = prof_drain(e);
    for (int i = 0; i < BLX_K_COUNT; ++i) { e->prof_ms[i] = 0; e->prof_n[i] = 0; }
    return rc;
}
extern "C" int blx_profile_read(blx_engine *e, float *ms, int *launches) {
    if (!e) return fail(BLX_ERR_ARG, "null engine");
    int rc = prof_drain(e);
    for (int i = 0; i < BLX_K_COUNT; ++i) {
        if (ms) ms[i] = e->prof_ms[i];
        if (launches) launches[i] = e->prof_n[i];
    }
    return rc;
}
extern "C" long long blx_launch_count(blx_engine *e) { return e ? e->launches : 0; }

// ---------------------------------------------------------------- lifecycle
extern "C" int blx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int blx_init(int device, blx_engine **out) {
    if (!out) return fail(BLX_ERR_ARG, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t ce = cudaGetDeviceCount(&n);
    if (ce != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(BLX_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU path",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    }
    if (device < 0 || device >= n) return fail(BLX_ERR_ARG, "device %d out of range [0, %d)", device, n);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(BLX_ERR_CUDA, "device %d is sm_%d%d; the kernels are built for sm_100a only", device, prop.major,
                    prop.minor);
    blx_engine *e = new blx_engine();
    e->device = device;
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithPriority(&e->compute, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithPriority(&e->copy, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithPriority(&e->tail, cudaStreamNonBlocking, prio_hi));
    CK(cudaEventCreateWithFlags(&e->fence, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->joined_work, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->joined_tail, cudaEventDisableTiming));
    for (int i = 0; i < blx_engine::kSlots; ++i) {
        CK(cudaEventCreateWithFlags(&e->slot[i].copied, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->slot[i].done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->slot[i].env_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->slot[i].p1_done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&e->slot[i].ep_done, cudaEventDisableTiming));
    }
    if (const char *sb = getenv("BLX_SUB_BATCH")) e->sub_batch = std::max(1, atoi(sb));
    // constant tables, computed once in double on the host
    std::vector<float> hann(kWin);
    for (int i = 0; i < kWin; ++i) // reference src/frequency_sort.c:40-42
        hann[i] = (float)(0.5 * (1.0 - cos(2 * M_PI * i / (kWin - 1))));
    std::vector<float2> tw1f(256), tw2f(128);
    std::vector<double2> tw1d(256), tw2d(128);
    for (int c = 0; c < 16; ++c)
        for (int b = 0; b < 16; ++b) {
            const double a = -2.0 * M_PI * (double)((b * c) % 256) / 256.0;
            tw1d[c * 16 + b] = make_double2(cos(a), sin(a));
            tw1f[c * 16 + b] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int k = 0; k < 128; ++k) {
        const double a = -2.0 * M_PI * (double)k / 512.0;
        tw2d[k] = make_double2(cos(a), sin(a));
        tw2f[k] = make_float2((float)cos(a), (float)sin(a));
    }
    CK(cudaMalloc(&e->d_hann, kWin * sizeof(float)));
    CK(cudaMalloc(&e->d_tw1f, 256 * sizeof(float2)));
    CK(cudaMalloc(&e->d_tw2f, 128 * sizeof(float2)));
    CK(cudaMalloc(&e->d_tw1d, 256 * sizeof(double2)));
    CK(cudaMalloc(&e->d_tw2d, 128 * sizeof(double2)));
    CK(cudaMemcpy(e->d_hann, hann.data(), kWin * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_tw1f, tw1f.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_tw2f, tw2f.data(), 128 * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_tw1d, tw1d.data(), 256 * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(e->d_tw2d, tw2d.data(), 128 * sizeof(double2), cudaMemcpyHostToDevice));
    *out = e;
    return BLX_OK;
}

extern "C" void blx_shutdown(blx_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < blx_engine::kSlots; ++i) {
        Slot &s = e->slot[i];
        s.pcm.release(); s.songs.release(); s.partials.release(); s.hist.release(); s.stats.release();
        s.norm.release(); s.energy.release(); s.xlog.release(); s.q.release(); s.results.release(); s.freq.release();
        if (s.h_songs) cudaFreeHost(s.h_songs);
        if (s.h_results) cudaFreeHost(s.h_results);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.done) cudaEventDestroy(s.done);
        if (s.env_done) cudaEventDestroy(s.env_done);
        if (s.p1_done) cudaEventDestroy(s.p1_done);
        if (s.ep_done) cudaEventDestroy(s.ep_done);
    }
    if (e->fence) cudaEventDestroy(e->fence);
    if (e->joined_work) cudaEventDestroy(e->joined_work);
    if (e->joined_tail) cudaEventDestroy(e->joined_tail);
    if (e->tail) cudaStreamDestroy(e->tail);
    e->scratch_a.release();
    e->scratch_b.release();
    e->scratch_c.release();
    e->scratch_near.release();
    for (auto &r : e->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto ev : e->ev_pool) cudaEventDestroy(ev);
    cudaFree(e->d_hann); cudaFree(e->d_tw1f); cudaFree(e->d_tw2f); cudaFree(e->d_tw1d); cudaFree(e->d_tw2d);
    if (e->compute) cudaStreamDestroy(e->compute);
    if (e->copy) cudaStreamDestroy(e->copy);
    delete e;
}

extern "C" int blx_configure(blx_engine *e, size_t chunk_bytes) {
    if (!e) return fail(BLX_ERR_ARG, "null engine");
    if (chunk_bytes < ((size_t)1 << 20)) return fail(BLX_ERR_ARG, "chunk_bytes must be at least 1 MiB");
    e->chunk_bytes = chunk_bytes;
    return BLX_OK;
}

extern "C" int blx_configure_sub_batch(blx_engine *e, int songs) {
    if (!e) return fail(BLX_ERR_ARG, "null engine");
    if (songs < 1) return fail(BLX_ERR_ARG, "sub-batch must be at least one song");
    e->sub_batch = songs;
    return BLX_OK;
}

extern "C" int blx_debug_flags(blx_engine *e, unsigned flags) {
    if (!e) return fail(BLX_ERR_ARG, "null engine");
    e->debug = flags;
    return BLX_OK;
}

// ---------------------------------------------------------------- descriptors
static inline long long round_up(long long v, long long m) { return (v + m - 1) / m * m; }

struct ChunkPlan {
    int kind = 0;
    int max_parts = 0;
    int parts_total = 0;
    int max_hops = 0;
    long long energy_total = 0;
    long long q_total = 0;
    long long pcm_rows = 0; // float32 input: rows of 64 floats the packed buffer is readable for
};

// Fills n descriptors. offsets/lengths are in input elements.
static void plan_songs(int fmt, const long long *offsets, const long long *lengths, int channels_kind,
                       const unsigned long long *durations, int n, SongDesc *sd, ChunkPlan *plan, bool spectral_only = false) {
    const int tile_m = pass1_tile_msamples();
    // A song is cut into parts of kTilesPerPart consecutive tiles, one CTA each. The cut depends on the
    // song alone, so its partial spectra (and their float summation order) are the same in any batch.
    // (The full pass pays a histogram clear + flush per CTA, hence larger parts; the spectral-only pass has
    // almost no per-CTA cost and takes small parts so that short songs still fill the 148 SMs evenly.)
    const int kTilesPerPart = spectral_only ? 16 : 64;
    plan->kind = (fmt == BLX_FMT_F32) ? kInF32 : channels_kind;
    long long env = 0, q = 0;
    int parts = 0;
    for (int i = 0; i < n; ++i) {
        SongDesc d;
        memset(&d, 0, sizeof(d));
        d.pcm_off = offsets[i];
        d.kind = plan->kind;
        d.n_elems = (int)lengths[i];
        if (fmt == BLX_FMT_F32) {
            d.n_msamples = (int)(lengths[i] / 2);
            d.n_samples = 2 * d.n_msamples;
            d.duration = (unsigned)(lengths[i] / BLX_FE_IN_RATE);
            d.q_off = q;
            q += round_up(std::max(d.n_msamples, 1), 64);
            plan->pcm_rows = std::max(plan->pcm_rows, (offsets[i] + round_up(std::max(lengths[i], 1ll), BLX_ALIGN_ELEMS)) / 64);
        } else {
            d.n_samples = (int)lengths[i];
            d.n_msamples = (plan->kind == kInS16Mono) ? d.n_samples : d.n_samples / 2;
            d.duration = durations ? (unsigned)durations[i] : 0u;
        }
        d.n_frames = d.n_msamples / kWin;
        // tiles cover every input element (the histogram and the statistics need the tail too)
        const long long rows_elems = (plan->kind == kInS16Mono) ? 32 : 64;
        const long long rows = (d.n_elems + rows_elems - 1) / rows_elems;
        d.n_tiles = (int)std::max(1ll, (rows * 32 + tile_m - 1) / tile_m);
        d.F = d.n_samples / kWin;
        d.n_hops = std::max(0, 2 * d.F - 2);
        d.env_off = env;
        env += round_up(std::max(2 * d.F, 2), 8);
        d.n_parts = (d.n_tiles + kTilesPerPart - 1) / kTilesPerPart;
        d.part_off = parts;
        parts += d.n_parts;
        plan->max_parts = std::max(plan->max_parts, d.n_parts);
        plan->max_hops = std::max(plan->max_hops, d.n_hops);
        sd[i] = d;
    }
    plan->parts_total = parts;
    plan->energy_total = env;
    plan->q_total = q;
}

// ---------------------------------------------------------------- the kernel sequence for one chunk
static int ensure_host_songs(Slot &s, int n) {
    if ((size_t)n > s.h_songs_cap) {
        if (s.h_songs) cudaFreeHost(s.h_songs);
        s.h_songs = nullptr;
        s.h_songs_cap = 0;
        const size_t cap = std::max<size_t>(1024, (size_t)n * 2);
        CK(cudaMallocHost(&s.h_songs, cap * sizeof(SongDesc)));
        s.h_songs_cap = cap;
    }
    if ((size_t)n > s.h_results_cap) {
        if (s.h_results) cudaFreeHost(s.h_results);
        s.h_results = nullptr;
        s.h_results_cap = 0;
        const size_t cap = std::max<size_t>(1024, (size_t)n * 2);
        CK(cudaMallocHost(&s.h_results, cap * sizeof(blx_result)));
        s.h_results_cap = cap;
    }
    return BLX_OK;
}

// Songs already described in s.h_songs[0..n). d_pcm is the packed input.
// The descriptors are uploaded here unless the caller already queued that copy (host-buffer path: behind the
// chunk's PCM on the copy stream - a small host->device copy on the compute stream would wait in the copy
// engine behind the NEXT chunk's PCM and stall this chunk's kernels for a whole chunk copy).
// A chunk runs in two halves so that a caller with several chunks can put the first half of chunk k + 1 in front of
// the second half of chunk k:
//   chunk_front: pass 1 on `st`, then the epilogue (one latency-bound CTA per song) on `st_side`;
//   chunk_back:  envelope on `st` (behind the epilogue's event), then log compression + tail on `st_side`.
// With st_side == st everything is one in-order sequence. The caller records the chunk's completion on st_side.
static int chunk_front(blx_engine *e, Slot &s, const ChunkPlan &plan, const void *d_pcm, int n, unsigned what, float *d_freq_only,
                       cudaStream_t st, cudaStream_t st_side, bool songs_uploaded) {
    const bool full = (d_freq_only == nullptr);
    CK(s.songs.reserve((size_t)n * sizeof(SongDesc)));
    CK(s.partials.reserve((size_t)std::max(plan.parts_total, 1) * 256 * sizeof(float)));
    CK(s.norm.reserve((size_t)n * sizeof(SongNorm)));
    if (full) {
        CK(s.hist.reserve((size_t)n * kHistStride * sizeof(unsigned)));
        CK(s.stats.reserve((size_t)n * sizeof(SongStats)));
        CK(s.energy.reserve((size_t)std::max(plan.energy_total, 8ll) * sizeof(double)));
        CK(s.xlog.reserve((size_t)(std::max(plan.energy_total, 8ll) + 16) * sizeof(double))); // + read-ahead slack of the tail
        if (plan.kind == kInF32) CK(s.q.reserve((size_t)std::max(plan.q_total, 64ll) * sizeof(short)));
    }
    if (!songs_uploaded) CK(cudaMemcpyAsync(s.songs.p, s.h_songs, (size_t)n * sizeof(SongDesc), cudaMemcpyHostToDevice, st));
    if (full) {
        CK(cudaMemsetAsync(s.hist.p, 0, (size_t)n * kHistStride * sizeof(unsigned), st));
        CK(cudaMemsetAsync(s.stats.p, 0, (size_t)n * sizeof(SongStats), st));
    }
    const SongDesc *d_songs = static_cast<const SongDesc *>(s.songs.p);
    {
        Pass1Params p;
        p.pcm = d_pcm;
        p.songs = d_songs;
        p.hann = e->d_hann;
        p.tw1 = e->d_tw1f;
        p.tw2 = e->d_tw2f;
        p.partials = static_cast<float *>(s.partials.p);
        p.hist = static_cast<unsigned *>(s.hist.p);
        p.stats = static_cast<SongStats *>(s.stats.p);
        p.qout = static_cast<short *>(s.q.p);
        memset(&p.map_a, 0, sizeof(p.map_a));
        memset(&p.map_b, 0, sizeof(p.map_b));
        if (plan.kind == kInF32) CK(make_pass1_maps(&p, d_pcm, plan.pcm_rows));
        ProfScope ps(e, BLX_K_PASS1, st);
        CK(launch_pass1(plan.kind, full, p, plan.max_parts, n, st));
    }
    if (st_side != st) {
        CK(cudaEventRecord(s.p1_done, st));
        CK(cudaStreamWaitEvent(st_side, s.p1_done, 0));
    }
    {
        EpilogueParams p;
        p.songs = d_songs;
        p.partials = static_cast<const float *>(s.partials.p);
        p.hist = static_cast<const unsigned *>(s.hist.p);
        p.stats = full ? static_cast<const SongStats *>(s.stats.p) : nullptr;
        p.norm = static_cast<SongNorm *>(s.norm.p);
        p.frequency = d_freq_only;
        p.energy = (full && (what & BLX_DO_ENVELOPE)) ? static_cast<double *>(s.energy.p) : nullptr;
        p.what = full ? what : BLX_DO_FREQUENCY;
        ProfScope ps(e, BLX_K_EPILOGUE, st_side);
        CK(launch_epilogue(p, n, st_side));
    }
    if (st_side != st) CK(cudaEventRecord(s.ep_done, st_side));
    return BLX_OK;
}

static int chunk_back(blx_engine *e, Slot &s, const ChunkPlan &plan, const void *d_pcm, int n, unsigned what, blx_result *d_out,
                      cudaStream_t st, cudaStream_t st_side) {
    const SongDesc *d_songs = static_cast<const SongDesc *>(s.songs.p);
    if (st_side != st) CK(cudaStreamWaitEvent(st, s.ep_done, 0));
    if (what & BLX_DO_ENVELOPE) {
        EnvelopeParams p;
        p.stream = (plan.kind == kInF32) ? static_cast<const short *>(s.q.p) : static_cast<const short *>(d_pcm);
        p.songs = d_songs;
        p.norm = static_cast<const SongNorm *>(s.norm.p);
        p.tw1 = e->d_tw1d;
        p.tw2 = e->d_tw2d;
        p.energy = static_cast<double *>(s.energy.p);
        p.dup = (plan.kind == kInF32) ? 1 : 0;
        p.slow_chain = (e->debug & BLX_DEBUG_SLOW_CHAIN) ? 1 : 0;
        ProfScope ps(e, BLX_K_ENVELOPE, st);
        CK(launch_envelope(p, plan.max_hops, n, st));
    }
    if (st_side != st) {
        CK(cudaEventRecord(s.env_done, st));
        CK(cudaStreamWaitEvent(st_side, s.env_done, 0));
    }
    if (what & BLX_DO_ENVELOPE) {
        ProfScope ps(e, BLX_K_TAIL, st_side);
        CK(launch_logcomp(static_cast<const double *>(s.energy.p), static_cast<double *>(s.xlog.p), plan.energy_total, st_side));
    }
    {
        TailParams p;
        p.songs = d_songs;
        p.norm = static_cast<const SongNorm *>(s.norm.p);
        p.xlog = static_cast<const double *>(s.xlog.p);
        p.out = d_out;
        p.what = what;
        ProfScope ps(e, BLX_K_TAIL, st_side);
        CK(launch_tail(p, n, st_side));
    }
    return BLX_OK;
}

static int run_chunk(blx_engine *e, Slot &s, const ChunkPlan &plan, const void *d_pcm, int n, unsigned what,
                     blx_result *d_out, float *d_freq_only, cudaStream_t st, cudaStream_t st_side, bool songs_uploaded = false) {
    int rc = chunk_front(e, s, plan, d_pcm, n, what, d_freq_only, st, st_side, songs_uploaded);
    if (rc || d_freq_only) return rc;
    return chunk_back(e, s, plan, d_pcm, n, what, d_out, st, st_side);
}

static int check_engine(blx_engine *e) {
    if (!e) return fail(BLX_ERR_ARG, "null engine (blx_init failed or was not called)");
    CK(cudaSetDevice(e->device));
    return BLX_OK;
}

// ---------------------------------------------------------------- device-resident entry points
// The songs are analysed in sub-batches of e->sub_batch songs, each with its own intermediates (slot): pass 1,
// epilogue and envelope of all sub-batches go to the engine's compute stream in order, the latency-bound tail of a
// sub-batch (4 ms whatever its size) to the high-priority tail stream, where it runs under the next sub-batch's
// kernels. `stream` is fenced in front (the engine's streams wait for it) and, unless `async`, behind (it waits for
// the engine's streams): the caller sees ordinary stream semantics.
static int join_stream(blx_engine *e, cudaStream_t st) {
    CK(cudaEventRecord(e->joined_work, e->compute));
    CK(cudaEventRecord(e->joined_tail, e->tail));
    if (st != e->compute) CK(cudaStreamWaitEvent(st, e->joined_work, 0));
    CK(cudaStreamWaitEvent(st, e->joined_tail, 0));
    return BLX_OK;
}

static int analyze_device_impl(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets, const int64_t *lengths,
                               const int *channels, const uint64_t *duration_s, int n_songs, unsigned what,
                               blx_result *d_out, float *d_freq_only, void *stream, bool async) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (n_songs <= 0) return BLX_OK;
    if (!d_pcm || !offsets || !lengths) return fail(BLX_ERR_ARG, "null input array");
    if (fmt != BLX_FMT_S16 && fmt != BLX_FMT_F32) return fail(BLX_ERR_ARG, "unknown format %d", fmt);
    if (!(what & BLX_DO_ALL)) return fail(BLX_ERR_ARG, "empty analyser mask");
    cudaStream_t user = stream ? static_cast<cudaStream_t>(stream) : e->compute;
    const bool spectral = d_freq_only != nullptr;
    // the spectral-only form (two kernels, no tail) runs on the caller's stream itself
    cudaStream_t work = spectral ? user : e->compute;
    if (user != work) {
        CK(cudaEventRecord(e->fence, user));
        CK(cudaStreamWaitEvent(work, e->fence, 0));
    }
    // the spectral-only form has no tail: one sequence per run of songs, all on the compute stream
    const int sub = spectral ? 65535 : std::min(e->sub_batch, 65535);
    // songs are processed in runs of equal channel count
    Slot *pend = nullptr; // sub-batch whose second half (envelope, tail) is still to be enqueued
    ChunkPlan pend_plan;
    int pend_n = 0;
    blx_result *pend_out = nullptr;
    int i0 = 0;
    while (i0 < n_songs) {
        const int ch0 = (fmt == BLX_FMT_S16 && channels) ? channels[i0] : 2;
        int i1 = i0;
        while (i1 < n_songs && i1 - i0 < sub && ((fmt == BLX_FMT_S16 && channels) ? channels[i1] : 2) == ch0) ++i1;
        if (fmt == BLX_FMT_S16 && ch0 != 1 && ch0 != 2) return fail(BLX_ERR_ARG, "song %d: %d channels unsupported", i0, ch0);
        const int n = i1 - i0;
        Slot &s = e->slot[e->next_dev_slot];
        e->next_dev_slot = (e->next_dev_slot + 1) % blx_engine::kSlots;
        if (s.busy) {
            CK(cudaEventSynchronize(s.done));
            s.busy = false;
        }
        rc = ensure_host_songs(s, n);
        if (rc) return rc;
        for (int i = i0; i < i1; ++i) {
            if (offsets[i] % BLX_ALIGN_ELEMS) return fail(BLX_ERR_ARG, "song %d: offset not a multiple of %d", i, BLX_ALIGN_ELEMS);
            if (lengths[i] < 0 || lengths[i] > 0x7fffffffll) return fail(BLX_ERR_ARG, "song %d: bad length", i);
        }
        ChunkPlan plan;
        plan_songs(fmt, reinterpret_cast<const long long *>(offsets + i0), reinterpret_cast<const long long *>(lengths + i0),
                   ch0 == 1 ? kInS16Mono : kInS16Stereo,
                   duration_s ? reinterpret_cast<const unsigned long long *>(duration_s + i0) : nullptr, n, s.h_songs, &plan,
                   spectral);
        if (spectral) {
            rc = run_chunk(e, s, plan, d_pcm, n, what, nullptr, d_freq_only + i0, work, work);
            if (rc) return rc;
            CK(cudaEventRecord(s.done, work));
            s.busy = true;
        } else {
            // software pipeline over the sub-batches: pass 1 + epilogue of this one go in front of the envelope kernel of
            // the previous one, so that the epilogue (side stream) runs under an envelope kernel, not between two kernels
            rc = chunk_front(e, s, plan, d_pcm, n, what, nullptr, work, e->tail, false);
            if (rc) return rc;
            s.busy = true;
            if (pend) {
                rc = chunk_back(e, *pend, pend_plan, d_pcm, pend_n, what, pend_out, work, e->tail);
                if (rc) return rc;
                CK(cudaEventRecord(pend->done, e->tail));
            }
            pend = &s; pend_plan = plan; pend_n = n; pend_out = d_out + i0;
        }
        i0 = i1;
    }
    if (pend) {
        rc = chunk_back(e, *pend, pend_plan, d_pcm, pend_n, what, pend_out, work, e->tail);
        if (rc) return rc;
        CK(cudaEventRecord(pend->done, e->tail));
    }
    if (!async && !spectral) return join_stream(e, user);
    return BLX_OK;
}

extern "C" int blx_analyze_device(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets, const int64_t *lengths,
                                  const int *channels, const uint64_t *duration_s, int n_songs, unsigned what,
                                  blx_result *d_out, void *stream) {
    if (!d_out) return fail(BLX_ERR_ARG, "null d_out");
    return analyze_device_impl(e, fmt, d_pcm, offsets, lengths, channels, duration_s, n_songs, what, d_out, nullptr, stream, false);
}

extern "C" int blx_analyze_device_async(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets,
                                        const int64_t *lengths, const int *channels, const uint64_t *duration_s, int n_songs,
                                        unsigned what, blx_result *d_out, void *stream) {
    if (!d_out) return fail(BLX_ERR_ARG, "null d_out");
    return analyze_device_impl(e, fmt, d_pcm, offsets, lengths, channels, duration_s, n_songs, what, d_out, nullptr, stream, true);
}

extern "C" int blx_join(blx_engine *e, void *stream) {
    int rc = check_engine(e);
    if (rc) return rc;
    return join_stream(e, stream ? static_cast<cudaStream_t>(stream) : e->compute);
}

extern "C" int blx_spectral_device(blx_engine *e, int fmt, const void *d_pcm, const int64_t *offsets, const int64_t *lengths,
                                   const int *channels, int n_songs, float *d_frequency, void *stream) {
    if (!d_frequency) return fail(BLX_ERR_ARG, "null d_frequency");
    return analyze_device_impl(e, fmt, d_pcm, offsets, lengths, channels, nullptr, n_songs, BLX_DO_FREQUENCY, nullptr,
                               d_frequency, stream, false);
}

// ---------------------------------------------------------------- host-buffer entry points
template <typename T>
static int analyze_host_impl(blx_engine *e, int fmt, const T *const *pcm, const long long *lengths, const int *channels,
                             const uint64_t *duration_s, int n_songs, unsigned what, blx_result *out) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (n_songs <= 0) return BLX_OK;
    if (!pcm || !lengths || !out) return fail(BLX_ERR_ARG, "null input array");
    if (!(what & BLX_DO_ALL)) return fail(BLX_ERR_ARG, "empty analyser mask");
    const size_t cap_elems = e->chunk_bytes / sizeof(T);
    std::vector<long long> offs;
    int i0 = 0;
    // BLX_TRACE=1: per-chunk timeline (copy start/end, compute start/end, ms since the call started) on stderr
    const bool trace = getenv("BLX_TRACE") != nullptr;
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) { if (trace) { cudaEvent_t ev; cudaEventCreate(&ev); cudaEventRecord(ev, st); tev.push_back(ev); } };
    mark(e->copy);
    int pending[2] = {-1, -1}; // first song of the chunk in flight in each slot
    int pending_n[2] = {0, 0};
    auto collect = [&](int si) -> int {
        Slot &s = e->slot[si];
        if (pending[si] >= 0) {
            CK(cudaEventSynchronize(s.done));
            memcpy(out + pending[si], s.h_results, (size_t)pending_n[si] * sizeof(blx_result));
            pending[si] = -1;
            s.busy = false;
        }
        return BLX_OK;
    };
    while (i0 < n_songs) {
        const int ch0 = (fmt == BLX_FMT_S16 && channels) ? channels[i0] : 2;
        if (fmt == BLX_FMT_S16 && ch0 != 1 && ch0 != 2) return fail(BLX_ERR_ARG, "song %d: %d channels unsupported", i0, ch0);
        // greedy chunk: equal channel count, fits the staging buffer (a single oversized song grows it)
        offs.clear();
        long long used = 0;
        int i1 = i0;
        while (i1 < n_songs && i1 - i0 < 65535) {
            const int ch = (fmt == BLX_FMT_S16 && channels) ? channels[i1] : 2;
            if (ch != ch0) break;
            if (lengths[i1] < 0 || lengths[i1] > 0x7fffffffll || !pcm[i1]) return fail(BLX_ERR_ARG, "song %d: bad buffer", i1);
            const long long need = round_up(std::max(lengths[i1], 1ll), BLX_ALIGN_ELEMS) + BLX_ALIGN_ELEMS;
            if (i1 > i0 && (size_t)(used + need) > cap_elems) break;
            offs.push_back(used);
            used += need;
            ++i1;
        }
        const int n = i1 - i0;
        const int si = e->next_slot;
        e->next_slot ^= 1;
        rc = collect(si);
        if (rc) return rc;
        Slot &s = e->slot[si];
        if (s.busy) { CK(cudaEventSynchronize(s.done)); s.busy = false; }
        CK(s.pcm.reserve((size_t)used * sizeof(T)));
        rc = ensure_host_songs(s, n);
        if (rc) return rc;
        CK(s.results.reserve((size_t)n * sizeof(blx_result)));
        mark(e->copy);
        for (int i = 0; i < n; ++i)
            CK(cudaMemcpyAsync(static_cast<T *>(s.pcm.p) + offs[i], pcm[i0 + i], (size_t)lengths[i0 + i] * sizeof(T),
                               cudaMemcpyHostToDevice, e->copy));
        mark(e->copy);
        ChunkPlan plan;
        plan_songs(fmt, offs.data(), lengths + i0, ch0 == 1 ? kInS16Mono : kInS16Stereo,
                   duration_s ? reinterpret_cast<const unsigned long long *>(duration_s + i0) : nullptr, n, s.h_songs, &plan);
        CK(s.songs.reserve((size_t)n * sizeof(SongDesc)));
        CK(cudaMemcpyAsync(s.songs.p, s.h_songs, (size_t)n * sizeof(SongDesc), cudaMemcpyHostToDevice, e->copy));
        CK(cudaEventRecord(s.copied, e->copy));
        CK(cudaStreamWaitEvent(e->compute, s.copied, 0));
        mark(e->compute);
        rc = run_chunk(e, s, plan, s.pcm.p, n, what, static_cast<blx_result *>(s.results.p), nullptr, e->compute, e->tail, true);
        if (rc) return rc;
        mark(e->tail);
        CK(cudaMemcpyAsync(s.h_results, s.results.p, (size_t)n * sizeof(blx_result), cudaMemcpyDeviceToHost, e->tail));
        CK(cudaEventRecord(s.done, e->tail));
        s.busy = true;
        pending[si] = i0;
        pending_n[si] = n;
        i0 = i1;
    }
    rc = collect(e->next_slot);
    if (rc) return rc;
    rc = collect(e->next_slot ^ 1);
    if (trace) {
        cudaDeviceSynchronize();
        for (size_t k = 1; k + 3 < tev.size() + 1; k += 4) {
            float t[4];
            for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], tev[0], tev[k + j]);
            fprintf(stderr, "[blx trace] chunk %zu: copy %.2f..%.2f ms, compute %.2f..%.2f ms\n", (k - 1) / 4, t[0], t[1], t[2], t[3]);
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
    }
    return rc;
}

extern "C" int blx_analyze_batch_s16(blx_engine *e, const int16_t *const *pcm, const int *n_samples, const int *channels,
                                     const uint64_t *duration_s, int n_songs, unsigned what, blx_result *out) {
    if (n_songs > 0 && !n_samples) return fail(BLX_ERR_ARG, "null n_samples");
    std::vector<long long> len(n_songs > 0 ? n_songs : 0);
    for (int i = 0; i < n_songs; ++i) len[i] = n_samples[i];
    return analyze_host_impl<int16_t>(e, BLX_FMT_S16, pcm, len.data(), channels, duration_s, n_songs, what, out);
}

extern "C" int blx_analyze_batch_f32(blx_engine *e, const float *const *pcm, const int64_t *n_in, int n_songs, unsigned what,
                                     blx_result *out) {
    if (n_songs > 0 && !n_in) return fail(BLX_ERR_ARG, "null n_in");
    return analyze_host_impl<float>(e, BLX_FMT_F32, pcm, reinterpret_cast<const long long *>(n_in), nullptr, nullptr, n_songs,
                                    what, out);
}

// ---------------------------------------------------------------- distances
extern "C" int blx_distance_rows_device(blx_engine *e, const float *d_vectors, int n, int row0, int n_rows, int mode,
                                        float *d_out, void *stream) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!d_vectors || !d_out || n < 0 || row0 < 0 || n_rows < 0 || row0 + n_rows > n || (mode != 0 && mode != 1))
        return fail(BLX_ERR_ARG, "bad distance arguments");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : e->compute;
    ProfScope ps(e, BLX_K_DISTANCE, st);
    CK(launch_distance_rows(d_vectors, n, row0, n_rows, mode, d_out, st));
    return BLX_OK;
}

extern "C" int blx_distance_nearest_device(blx_engine *e, const float *d_vectors, int n, int row0, int n_rows,
                                           int *d_nearest_index, float *d_nearest_dist, double *d_row_sum, void *stream) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!d_vectors || n < 0 || row0 < 0 || n_rows < 0 || row0 + n_rows > n) return fail(BLX_ERR_ARG, "bad distance arguments");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : e->compute;
    // few rows (a rank's slab in a multi-GPU run): the columns are split over gridDim.y and merged through a
    // per-engine scratch array, so one nearest-neighbour call per engine may be in flight at a time
    const int splits = distance_nearest_splits(n, n_rows, d_row_sum != nullptr);
    if (splits > 1) CK(e->scratch_near.reserve((size_t)n_rows * sizeof(unsigned long long)));
    ProfScope ps(e, BLX_K_DISTANCE, st);
    CK(launch_distance_nearest(d_vectors, n, row0, n_rows, d_nearest_index, d_nearest_dist, d_row_sum,
                               static_cast<unsigned long long *>(e->scratch_near.p), splits, st));
    return BLX_OK;
}

extern "C" int blx_cosine_nearest_device(blx_engine *e, const float *d_vectors, int n, int row0, int n_rows, int *d_index,
                                         float *d_similarity, void *stream) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!d_vectors || n < 0 || row0 < 0 || n_rows < 0 || row0 + n_rows > n) return fail(BLX_ERR_ARG, "bad distance arguments");
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : e->compute;
    ProfScope ps(e, BLX_K_DISTANCE, st);
    CK(launch_cosine_nearest(d_vectors, n, row0, n_rows, d_index, d_similarity, st));
    return BLX_OK;
}

static int matrix_host(blx_engine *e, const float *vectors, int n, float *out, int mode) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (n <= 0) return BLX_OK;
    if (!vectors || !out) return fail(BLX_ERR_ARG, "null array");
    CK(e->scratch_a.reserve((size_t)n * 16));
    CK(cudaMemcpyAsync(e->scratch_a.p, vectors, (size_t)n * 16, cudaMemcpyHostToDevice, e->compute));
    // slabs of at most 256 MiB of output
    const int rows_per = (int)std::max<long long>(1, std::min<long long>(n, ((long long)256 << 20) / ((long long)n * 4)));
    CK(e->scratch_b.reserve((size_t)rows_per * n * 4));
    for (int r0 = 0; r0 < n; r0 += rows_per) {
        const int nr = std::min(rows_per, n - r0);
        rc = blx_distance_rows_device(e, static_cast<const float *>(e->scratch_a.p), n, r0, nr, mode,
                                      static_cast<float *>(e->scratch_b.p), nullptr);
        if (rc) return rc;
        CK(cudaMemcpyAsync(out + (size_t)r0 * n, e->scratch_b.p, (size_t)nr * n * 4, cudaMemcpyDeviceToHost, e->compute));
        CK(cudaStreamSynchronize(e->compute));
    }
    return BLX_OK;
}
extern "C" int blx_distance_matrix(blx_engine *e, const float *vectors, int n, float *out) { return matrix_host(e, vectors, n, out, 0); }
extern "C" int blx_cosine_matrix(blx_engine *e, const float *vectors, int n, float *out) { return matrix_host(e, vectors, n, out, 1); }

// ---------------------------------------------------------------- small helpers
extern "C" int blx_rectangular_filter(blx_engine *e, double *out, const double *in, int n, int width) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!out || !in || width <= 0 || n < width) return fail(BLX_ERR_ARG, "bad filter arguments");
    CK(e->scratch_a.reserve((size_t)n * 8));
    CK(e->scratch_b.reserve((size_t)n * 8));
    CK(cudaMemcpyAsync(e->scratch_a.p, in, (size_t)n * 8, cudaMemcpyHostToDevice, e->compute));
    CK(cudaMemcpyAsync(e->scratch_b.p, out, (size_t)n * 8, cudaMemcpyHostToDevice, e->compute));
    e->launches++;
    CK(launch_rect_filter(static_cast<double *>(e->scratch_b.p), static_cast<const double *>(e->scratch_a.p), n, width, e->compute));
    CK(cudaMemcpyAsync(out, e->scratch_b.p, (size_t)n * 8, cudaMemcpyDeviceToHost, e->compute));
    CK(cudaStreamSynchronize(e->compute));
    return BLX_OK;
}

extern "C" int blx_frontend_f32(blx_engine *e, const float *pcm, int64_t n_in, int16_t *out) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!pcm || !out || n_in < 2) return fail(BLX_ERR_ARG, "bad front-end arguments");
    CK(e->scratch_a.reserve((size_t)n_in * 4));
    CK(e->scratch_b.reserve((size_t)(n_in / 2) * 4));
    CK(cudaMemcpyAsync(e->scratch_a.p, pcm, (size_t)n_in * 4, cudaMemcpyHostToDevice, e->compute));
    e->launches++;
    CK(launch_frontend(static_cast<const float *>(e->scratch_a.p), n_in, static_cast<short *>(e->scratch_b.p), e->compute));
    CK(cudaMemcpyAsync(out, e->scratch_b.p, (size_t)(n_in / 2) * 4, cudaMemcpyDeviceToHost, e->compute));
    CK(cudaStreamSynchronize(e->compute));
    return BLX_OK;
}

static int resample_impl(blx_engine *e, const void *samples, bool storage16, int kind, int bits, int channels, int64_t n_frames,
                         int in_rate, int16_t *out, int64_t out_capacity_frames, int64_t *n_out_frames) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!samples || !n_out_frames || n_frames <= 0 || (channels != 1 && channels != 2) || kind < 0 || kind > 3 || bits < 1 ||
        bits > 32 || (kind == BLX_RS_KIND_S16 && bits > 16) || (kind == BLX_RS_KIND_U8 && bits != 8))
        return fail(BLX_ERR_ARG, "bad resampler arguments");
    ResampleParams p;
    memset(&p, 0, sizeof(p));
    p.kind = kind; p.bits = bits; p.channels = channels; p.n_in = n_frames;
    std::vector<float> bank_f;
    std::vector<int16_t> bank_i;
    if (in_rate == BLX_RS_OUT_RATE) {
        p.n_out = n_frames;
    } else {
        blx_rs_plan plan;
        if (blx_rs_plan_make(in_rate, BLX_RS_OUT_RATE, &plan))
            return fail(BLX_ERR_ARG, "sample rate %d Hz needs more than %d filter phases", in_rate, BLX_RS_MAX_PHASES);
        p.L = plan.L; p.P = plan.P; p.q = plan.q; p.center = plan.center;
        p.mono_gain_last = BLX_RS_MONO_GAIN_LAST(in_rate) ? 1 : 0;
        p.n_out = blx_rs_out_frames(&plan, n_frames, nullptr);
        if (kind == BLX_RS_KIND_U8) { bank_i.resize((size_t)plan.P * plan.L); blx_rs_build_s16(&plan, bank_i.data()); }
        else { bank_f.resize((size_t)plan.P * plan.L); blx_rs_build_f32(&plan, bank_f.data()); }
    }
    *n_out_frames = p.n_out;
    if (!out) return BLX_OK; // size query
    if (out_capacity_frames < p.n_out) return fail(BLX_ERR_ARG, "output buffer too small");
    if (p.n_out == 0) return BLX_OK;
    const size_t in_bytes = (size_t)n_frames * channels * (storage16 ? 2 : 4), out_bytes = (size_t)p.n_out * 2 * 2;
    const size_t bank_bytes = bank_f.size() * 4 + bank_i.size() * 2;
    CK(e->scratch_a.reserve(in_bytes));
    CK(e->scratch_b.reserve(out_bytes + ((bank_bytes + 255) & ~(size_t)255) + 256));
    cudaStream_t st = e->compute;
    unsigned char *d_bank = static_cast<unsigned char *>(e->scratch_b.p) + ((out_bytes + 255) & ~(size_t)255);
    CK(cudaMemcpyAsync(e->scratch_a.p, samples, in_bytes, cudaMemcpyHostToDevice, st));
    if (bank_bytes)
        CK(cudaMemcpyAsync(d_bank, bank_f.empty() ? (const void *)bank_i.data() : (const void *)bank_f.data(), bank_bytes,
                           cudaMemcpyHostToDevice, st));
    p.in = storage16 ? nullptr : static_cast<const int *>(e->scratch_a.p);
    p.in16 = storage16 ? static_cast<const short *>(e->scratch_a.p) : nullptr;
    p.out = static_cast<short *>(e->scratch_b.p);
    p.bank_f32 = reinterpret_cast<const float *>(d_bank);
    p.bank_s16 = reinterpret_cast<const short *>(d_bank);
    e->launches++;
    CK(launch_resample(p, st));
    CK(cudaMemcpyAsync(out, e->scratch_b.p, out_bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BLX_OK;
}

extern "C" int blx_resample_to_s16(blx_engine *e, const int32_t *samples, int kind, int bits, int channels, int64_t n_frames,
                                   int in_rate, int16_t *out, int64_t out_capacity_frames, int64_t *n_out_frames) {
    return resample_impl(e, samples, false, kind, bits, channels, n_frames, in_rate, out, out_capacity_frames, n_out_frames);
}

extern "C" int blx_resample_s16_to_s16(blx_engine *e, const int16_t *samples, int channels, int64_t n_frames, int in_rate,
                                       int16_t *out, int64_t out_capacity_frames, int64_t *n_out_frames) {
    return resample_impl(e, samples, true, BLX_RS_KIND_S16, 16, channels, n_frames, in_rate, out, out_capacity_frames, n_out_frames);
}

// FLAC frames on the device (flacdec.cu). `hdr` is the host reader's chain of frames (flac_hdr[n_frames], 40 bytes each,
// host/flac_core.h), `first` the first sample of every frame; out receives samples * channels interleaved int16 (out16)
// or int32 values. BLX_ERR_ARG with "frame" in the message = a frame did not check out: decode on the host instead.
// in_rate > 0: a 16-bit mono / stereo stream that is not in the analysers' format - the decoded PCM stays on the device
// and goes straight through the decode-stage resampler; `out` then receives the int16 / 22 050 Hz / stereo result
// (*n_out_frames frames, capacity out_capacity_frames; out == NULL only sizes it).
static int flac_decode_impl(blx_engine *e, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first, int n_frames,
                            int channels, int out16, uint64_t samples, void *out, int in_rate, int64_t out_capacity_frames,
                            int64_t *n_out_frames) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!file || !hdr || !first || n_frames <= 0 || channels < 1 || channels > 8 || samples == 0 || samples > ((uint64_t)1 << 32))
        return fail(BLX_ERR_ARG, "bad FLAC decode arguments");
    ResampleParams rp;
    memset(&rp, 0, sizeof(rp));
    std::vector<float> bank;
    if (in_rate > 0) {
        if (!out16 || channels > 2 || !n_out_frames) return fail(BLX_ERR_ARG, "the fused resampler takes 16-bit mono / stereo streams");
        rp.kind = BLX_RS_KIND_S16; rp.bits = 16; rp.channels = channels; rp.n_in = (long long)samples;
        if (in_rate == BLX_RS_OUT_RATE) {
            rp.n_out = (long long)samples;
        } else {
            blx_rs_plan plan;
            if (blx_rs_plan_make(in_rate, BLX_RS_OUT_RATE, &plan))
                return fail(BLX_ERR_ARG, "sample rate %d Hz needs more than %d filter phases", in_rate, BLX_RS_MAX_PHASES);
            rp.L = plan.L; rp.P = plan.P; rp.q = plan.q; rp.center = plan.center;
            rp.mono_gain_last = BLX_RS_MONO_GAIN_LAST(in_rate) ? 1 : 0;
            rp.n_out = blx_rs_out_frames(&plan, (long long)samples, nullptr);
            bank.resize((size_t)plan.P * plan.L);
            blx_rs_build_f32(&plan, bank.data());
        }
        *n_out_frames = rp.n_out;
        if (!out) return BLX_OK; // size query
        if (out_capacity_frames < rp.n_out || rp.n_out <= 0) return fail(BLX_ERR_ARG, "output buffer too small or empty stream");
    } else if (!out) {
        return fail(BLX_ERR_ARG, "null output");
    }
    const size_t hdr_bytes = (size_t)n_frames * 40, first_bytes = (size_t)n_frames * 8;
    const size_t o_hdr = (n_bytes + 255) & ~(size_t)255, o_first = o_hdr + ((hdr_bytes + 255) & ~(size_t)255);
    const size_t o_fail = o_first + ((first_bytes + 255) & ~(size_t)255);
    const size_t scratch_bytes = (size_t)samples * channels * 4, out_bytes = (size_t)samples * channels * (out16 ? 2 : 4);
    CK(e->scratch_a.reserve(o_fail + 256));
    CK(e->scratch_b.reserve(((scratch_bytes + 255) & ~(size_t)255) + out_bytes));
    cudaStream_t st = e->compute;
    unsigned char *a = static_cast<unsigned char *>(e->scratch_a.p), *b = static_cast<unsigned char *>(e->scratch_b.p);
    CK(cudaMemcpyAsync(a, file, n_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(a + o_hdr, hdr, hdr_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(a + o_first, first, first_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(a + o_fail, 0, 4, st));
    FlacDecodeParams p;
    p.data = a; p.n_bytes = n_bytes; p.hdr = a + o_hdr;
    p.first = reinterpret_cast<const unsigned long long *>(a + o_first);
    p.n_frames = n_frames; p.channels = channels; p.out16 = out16;
    p.scratch = reinterpret_cast<int *>(b);
    p.out = b + ((scratch_bytes + 255) & ~(size_t)255);
    p.fail = reinterpret_cast<int *>(a + o_fail);
    e->launches++;
    CK(launch_flac_decode(p, st));
    int failed = 0;
    CK(cudaMemcpyAsync(&failed, p.fail, 4, cudaMemcpyDeviceToHost, st));
    if (in_rate > 0) {
        // decoded PCM (device) -> resampler -> host; the decoded stream itself never leaves the device
        const size_t rs_out_bytes = (size_t)rp.n_out * 2 * 2, bank_bytes = bank.size() * 4;
        CK(e->scratch_c.reserve(((rs_out_bytes + 255) & ~(size_t)255) + bank_bytes + 256));
        unsigned char *c = static_cast<unsigned char *>(e->scratch_c.p);
        unsigned char *d_bank = c + ((rs_out_bytes + 255) & ~(size_t)255);
        if (bank_bytes) CK(cudaMemcpyAsync(d_bank, bank.data(), bank_bytes, cudaMemcpyHostToDevice, st));
        rp.in = nullptr;
        rp.in16 = static_cast<const short *>(p.out);
        rp.out = reinterpret_cast<short *>(c);
        rp.bank_f32 = reinterpret_cast<const float *>(d_bank);
        rp.bank_s16 = reinterpret_cast<const short *>(d_bank);
        e->launches++;
        CK(launch_resample(rp, st));
        CK(cudaMemcpyAsync(out, c, rs_out_bytes, cudaMemcpyDeviceToHost, st));
    } else {
        CK(cudaMemcpyAsync(out, p.out, out_bytes, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    if (failed) return fail(BLX_ERR_ARG, "a FLAC frame did not decode on the device (damaged stream or a false frame boundary)");
    return BLX_OK;
}

extern "C" int blx_flac_decode_frames(blx_engine *e, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first,
                                      int n_frames, int channels, int out16, uint64_t samples, void *out) {
    return flac_decode_impl(e, file, n_bytes, hdr, first, n_frames, channels, out16, samples, out, 0, 0, nullptr);
}

extern "C" int blx_flac_decode_resample(blx_engine *e, const uint8_t *file, size_t n_bytes, const void *hdr, const uint64_t *first,
                                        int n_frames, int channels, uint64_t samples, int in_rate, int16_t *out,
                                        int64_t out_capacity_frames, int64_t *n_out_frames) {
    if (in_rate <= 0) return fail(BLX_ERR_ARG, "bad sample rate");
    return flac_decode_impl(e, file, n_bytes, hdr, first, n_frames, channels, 1, samples, out, in_rate, out_capacity_frames, n_out_frames);
}

// 44.1 kHz (or any-rate) mono / stereo float32 host buffers -> the decode-stage resampler on the device -> the native
// int16 pipeline: what bl_analyze computes for a float file, for callers that hold the decoded float PCM themselves.
extern "C" int blx_analyze_batch_f32_exact(blx_engine *e, const float *const *pcm, const int64_t *n_frames, int channels, int in_rate,
                                           int n_songs, unsigned what, blx_result *out) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (n_songs <= 0) return BLX_OK;
    if (!pcm || !n_frames || !out || (channels != 1 && channels != 2)) return fail(BLX_ERR_ARG, "bad arguments");
    if (!(what & BLX_DO_ALL)) return fail(BLX_ERR_ARG, "empty analyser mask");
    ResampleParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.kind = BLX_RS_KIND_F32; rp.bits = 32; rp.channels = channels;
    std::vector<float> bank;
    blx_rs_plan plan;
    const bool same_rate = in_rate == BLX_RS_OUT_RATE;
    if (!same_rate) {
        if (blx_rs_plan_make(in_rate, BLX_RS_OUT_RATE, &plan))
            return fail(BLX_ERR_ARG, "sample rate %d Hz needs more than %d filter phases", in_rate, BLX_RS_MAX_PHASES);
        rp.L = plan.L; rp.P = plan.P; rp.q = plan.q; rp.center = plan.center;
        rp.mono_gain_last = BLX_RS_MONO_GAIN_LAST(in_rate) ? 1 : 0;
        bank.resize((size_t)plan.P * plan.L);
        blx_rs_build_f32(&plan, bank.data());
    }
    DevBuf d_in, d_s16, d_bank, d_res;
    auto cleanup = [&]() { d_in.release(); d_s16.release(); d_bank.release(); d_res.release(); };
    cudaStream_t st = e->compute;
    if (!bank.empty()) {
        CK(d_bank.reserve(bank.size() * 4));
        CK(cudaMemcpyAsync(d_bank.p, bank.data(), bank.size() * 4, cudaMemcpyHostToDevice, st));
    }
    const size_t cap_in = e->chunk_bytes / 4; // floats of input per group of songs
    std::vector<int64_t> offs, lens;
    std::vector<uint64_t> durs;
    std::vector<int64_t> in_off;
    int i0 = 0;
    while (i0 < n_songs) {
        // group: consecutive songs whose input fits the staging size (a single larger song forms its own group)
        offs.clear(); lens.clear(); durs.clear(); in_off.clear();
        size_t used_in = 0, used_out = 0;
        int i1 = i0;
        while (i1 < n_songs && i1 - i0 < 65535) {
            if (!pcm[i1] || n_frames[i1] <= 0 || n_frames[i1] > 0x3fffffffll) { cleanup(); return fail(BLX_ERR_ARG, "song %d: bad buffer", i1); }
            const size_t need = (size_t)n_frames[i1] * channels;
            if (i1 > i0 && used_in + need > cap_in) break;
            const long long n_out = same_rate ? n_frames[i1] : blx_rs_out_frames(&plan, n_frames[i1], nullptr);
            in_off.push_back((int64_t)used_in);
            offs.push_back((int64_t)used_out);
            lens.push_back(2 * n_out);
            durs.push_back((uint64_t)(n_frames[i1] / in_rate)); // whole seconds of the source (reference src/decode.c:235)
            used_in += need;
            used_out += (size_t)round_up(std::max(2 * n_out, 1ll), BLX_ALIGN_ELEMS) + BLX_ALIGN_ELEMS;
            ++i1;
        }
        const int n = i1 - i0;
        rc = [&]() -> int {
            CK(d_in.reserve(used_in * 4));
            CK(d_s16.reserve(used_out * 2));
            CK(d_res.reserve((size_t)n * sizeof(blx_result)));
            for (int i = 0; i < n; ++i)
                CK(cudaMemcpyAsync(static_cast<float *>(d_in.p) + in_off[i], pcm[i0 + i], (size_t)n_frames[i0 + i] * channels * 4,
                                   cudaMemcpyHostToDevice, st));
            for (int i = 0; i < n; ++i) {
                rp.in = static_cast<const int *>(d_in.p) + in_off[i];
                rp.out = static_cast<short *>(d_s16.p) + offs[i];
                rp.bank_f32 = static_cast<const float *>(d_bank.p);
                rp.n_in = n_frames[i0 + i];
                rp.n_out = lens[i] / 2;
                e->launches++;
                CK(launch_resample(rp, st));
            }
            return BLX_OK;
        }();
        if (rc) { cleanup(); return rc; }
        rc = blx_analyze_device(e, BLX_FMT_S16, d_s16.p, offs.data(), lens.data(), nullptr, durs.data(), n, what,
                                static_cast<blx_result *>(d_res.p), nullptr);
        if (rc) { cleanup(); return rc; }
        cudaError_t ce = cudaMemcpyAsync(out + i0, d_res.p, (size_t)n * sizeof(blx_result), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { cleanup(); return fail(BLX_ERR_CUDA, "copy of the results failed: %s", cudaGetErrorString(ce)); }
        i0 = i1;
    }
    cleanup();
    return BLX_OK;
}

// Runs the full sequence on one S16 stereo song and leaves the intermediates in slot 0.
static int one_song_s16(blx_engine *e, const int16_t *pcm, int n_samples, uint64_t duration, unsigned what, blx_result *res) {
    const int16_t *ptrs[1] = {pcm};
    const int ns[1] = {n_samples};
    const uint64_t du[1] = {duration};
    e->next_slot = 0;
    return blx_analyze_batch_s16(e, ptrs, ns, nullptr, du, 1, what, res);
}

extern "C" int blx_mean_variance_s16(blx_engine *e, const int16_t *pcm, int n_samples, const int *mean_in, int *mean,
                                     int *variance) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!pcm || n_samples <= 0) return fail(BLX_ERR_ARG, "bad arguments");
    blx_result r;
    rc = one_song_s16(e, pcm, n_samples, 1, BLX_DO_ALL, &r);
    if (rc) return rc;
    SongNorm nm;
    SongStats st;
    CK(cudaMemcpy(&nm, e->slot[0].norm.p, sizeof(nm), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, e->slot[0].stats.p, sizeof(st), cudaMemcpyDeviceToHost));
    if (mean) *mean = nm.mean;
    if (variance) {
        if (!mean_in || *mean_in == nm.mean) {
            *variance = nm.variance;
        } else { // same integer identity as the epilogue kernel, around the caller's mean
            const long long m = *mean_in;
            const long long dev = (long long)st.sumsq - 2ll * m * st.sum + (long long)n_samples * m * m;
            *variance = (int)(dev / n_samples);
        }
    }
    return BLX_OK;
}

extern "C" int blx_envelope_energy_s16(blx_engine *e, const int16_t *pcm, int n_samples, double *energy) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!pcm || n_samples <= 0 || !energy) return fail(BLX_ERR_ARG, "bad arguments");
    blx_result r;
    rc = one_song_s16(e, pcm, n_samples, 1, BLX_DO_ALL, &r);
    if (rc) return rc;
    const int nb = 2 * (n_samples / kWin);
    if (nb > 0) CK(cudaMemcpy(energy, e->slot[0].energy.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost));
    return BLX_OK;
}

extern "C" int blx_envelope_energy_f32(blx_engine *e, const float *pcm, int64_t n_in, double *energy) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!pcm || n_in < 2 || !energy) return fail(BLX_ERR_ARG, "bad arguments");
    const float *ptrs[1] = {pcm};
    const int64_t ns[1] = {n_in};
    blx_result r;
    e->next_slot = 0;
    rc = blx_analyze_batch_f32(e, ptrs, ns, 1, BLX_DO_ALL, &r);
    if (rc) return rc;
    const int nb = 2 * (int)((2 * (n_in / 2)) / kWin);
    if (nb > 0) CK(cudaMemcpy(energy, e->slot[0].energy.p, (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost));
    return BLX_OK;
}

// ---------------------------------------------------------------- roofline denominators
extern "C" int blx_measure_fp64_peak(blx_engine *e, double *tflops, double *sm_mhz_nominal) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!tflops) return fail(BLX_ERR_ARG, "null output");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, e->device));
    const int threads = 1024, blocks = prop.multiProcessorCount * 2, iters = 1 << 15;
    CK(e->scratch_a.reserve((size_t)blocks * threads * sizeof(double)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) { // the first repetitions warm the clocks up
        CK(cudaEventRecord(a, e->compute));
        CK(launch_dfma_peak(static_cast<double *>(e->scratch_a.p), blocks, threads, iters, e->compute));
        CK(cudaEventRecord(b, e->compute));
        CK(cudaEventSynchronize(b));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a, b));
        const double tf = 2.0 * 8.0 * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device);
    if (sm_mhz_nominal) *sm_mhz_nominal = khz / 1000.0;
    return BLX_OK;
}

// ---------------------------------------------------------------- stage-level views (kernel parity tests)
extern "C" int blx_frequency_spectrum_s16(blx_engine *e, const int16_t *pcm, int n_samples, int channels, float *ps) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!pcm || n_samples <= 0 || !ps || (channels != 1 && channels != 2)) return fail(BLX_ERR_ARG, "bad arguments");
    const int16_t *ptrs[1] = {pcm};
    const int ns[1] = {n_samples}, ch[1] = {channels};
    const uint64_t du[1] = {1};
    blx_result r;
    e->next_slot = 0;
    rc = blx_analyze_batch_s16(e, ptrs, ns, ch, du, 1, BLX_DO_FREQUENCY | BLX_DO_AMPLITUDE, &r);
    if (rc) return rc;
    const int n_parts = e->slot[0].h_songs[0].n_parts;
    std::vector<float> part((size_t)n_parts * 256);
    CK(cudaMemcpy(part.data(), e->slot[0].partials.p, part.size() * sizeof(float), cudaMemcpyDeviceToHost));
    // the epilogue kernel's own order: parts added one after the other, in float
    for (int d = 0; d <= 256; ++d) ps[d] = 0.0f;
    for (int d = 1; d < 256; ++d) {
        float acc = 0.0f;
        for (int q = 0; q < n_parts; ++q) acc += part[(size_t)q * 256 + d];
        ps[d] = acc;
    }
    return BLX_OK;
}

extern "C" int blx_histogram_s16(blx_engine *e, const int16_t *pcm, int n_samples, unsigned *hist, int *first_nonzero,
                                 int *last_nonzero) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!pcm || n_samples <= 0 || !hist) return fail(BLX_ERR_ARG, "bad arguments");
    blx_result r;
    rc = one_song_s16(e, pcm, n_samples, 1, BLX_DO_AMPLITUDE, &r);
    if (rc) return rc;
    SongStats st;
    CK(cudaMemcpy(hist, e->slot[0].hist.p, (size_t)kHistBins * sizeof(unsigned), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, e->slot[0].stats.p, sizeof(st), cudaMemcpyDeviceToHost));
    if (first_nonzero) *first_nonzero = st.last_p1 ? (int)(0x7fffffffu - st.first_inv) : -1;
    if (last_nonzero) *last_nonzero = (int)st.last_p1 - 1;
    return BLX_OK;
}

extern "C" int blx_envelope_tail(blx_engine *e, const double *energy, int nb_frames, int n_samples, uint64_t duration_s,
                                 int *beat, float *tempo, float *attack) {
    int rc = check_engine(e);
    if (rc) return rc;
    if (!energy || nb_frames < 10 || (nb_frames & 1) || n_samples <= 0 || duration_s == 0)
        return fail(BLX_ERR_ARG, "bad arguments (nb_frames = 2 * (n_samples / 512) >= 10, duration >= 1)");
    Slot &s = e->slot[0];
    if (s.busy) { CK(cudaEventSynchronize(s.done)); s.busy = false; }
    rc = ensure_host_songs(s, 1);
    if (rc) return rc;
    SongDesc d;
    memset(&d, 0, sizeof(d));
    d.n_samples = n_samples;
    d.F = nb_frames / 2;
    d.n_hops = nb_frames - 2;
    d.duration = (unsigned)duration_s;
    s.h_songs[0] = d;
    SongNorm nm;
    memset(&nm, 0, sizeof(nm));
    const size_t row = (size_t)round_up(nb_frames, 8);
    CK(s.songs.reserve(sizeof(SongDesc)));
    CK(s.norm.reserve(sizeof(SongNorm)));
    CK(s.energy.reserve(row * sizeof(double)));
    CK(s.xlog.reserve((row + 16) * sizeof(double)));
    CK(s.results.reserve(sizeof(blx_result)));
    cudaStream_t st = e->compute;
    CK(cudaMemcpyAsync(s.songs.p, s.h_songs, sizeof(SongDesc), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s.norm.p, &nm, sizeof(nm), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(s.energy.p, 0, row * sizeof(double), st));
    CK(cudaMemcpyAsync(s.energy.p, energy, (size_t)nb_frames * sizeof(double), cudaMemcpyHostToDevice, st));
    e->launches += 2;
    CK(launch_logcomp(static_cast<const double *>(s.energy.p), static_cast<double *>(s.xlog.p), (long long)row, st));
    TailParams p;
    p.songs = static_cast<const SongDesc *>(s.songs.p);
    p.norm = static_cast<const SongNorm *>(s.norm.p);
    p.xlog = static_cast<const double *>(s.xlog.p);
    p.out = static_cast<blx_result *>(s.results.p);
    p.what = BLX_DO_ENVELOPE;
    CK(launch_tail(p, 1, st));
    blx_result r;
    CK(cudaMemcpyAsync(&r, s.results.p, sizeof(r), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (beat) *beat = r.beat;
    if (tempo) *tempo = r.tempo;
    if (attack) *attack = r.attack;
    return BLX_OK;
}
