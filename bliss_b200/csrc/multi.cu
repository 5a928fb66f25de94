// multi.cu — one process, several B200s, behind the C-ABI (blx.h "multi-device" section).
//
// The reference's callers are C programmes (reference examples/analyze.c:15-53, README.md:80); the split of
// BASELINE.json configs[2] / configs[4] over the GPUs of a box must therefore be reachable from C, without Python or
// torch.distributed: blx_multi owns one engine per device and one host thread per device per call.
//   - analysis: no data-path collective; the device threads take units of consecutive songs from a shared counter
//     (a device behind a slower host link takes fewer) and run the ordinary host-buffer path (blx_analyze_batch_*)
//     on each unit;
//   - all-pairs distances: every device keeps the force vectors of its block resident; they are all-gathered device
//     to device - ncclAllGather over NVLink (one communicator per device, ncclCommInitAll), or peer copies when
//     NCCL cannot start - and every device reduces the rows of its own block against all columns
//     (distance_nearest_kernel), 16 bytes per song on the wire.
#include <dlfcn.h>
#include <nccl.h> // types and enums only: the library is bound at run time (below)
#include <pthread.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "blx_common.cuh"
#include "kernels.h"

extern "C" int blx_set_error(int code, const char *fmt, ...);

// NCCL is bound with dlopen, not at link time: a process that also hosts PyTorch carries PyTorch's own libnccl.so.2
// (newer than the system's), and a link-time dependency of libbliss.so on the system copy would shadow it by soname
// and break `import torch`. A copy that is already loaded is reused; otherwise the system's is loaded privately.
namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
pthread_mutex_t g_nccl_lock = PTHREAD_MUTEX_INITIALIZER;
bool nccl_load() {
    pthread_mutex_lock(&g_nccl_lock);
    if (!g_nccl.lib) {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (h) {
            g_nccl.lib = h;
            g_nccl.CommInitAll = reinterpret_cast<decltype(g_nccl.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
            g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
            g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(dlsym(h, "ncclAllGather"));
            g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(dlsym(h, "ncclGroupStart"));
            g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(dlsym(h, "ncclGroupEnd"));
            g_nccl.GetVersion = reinterpret_cast<decltype(g_nccl.GetVersion)>(dlsym(h, "ncclGetVersion"));
            g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
            g_nccl.ok = g_nccl.CommInitAll && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.GroupStart && g_nccl.GroupEnd &&
                        g_nccl.GetVersion && g_nccl.GetErrorString;
        }
    }
    pthread_mutex_unlock(&g_nccl_lock);
    return g_nccl.ok;
}

struct Dev {
    int device = 0;
    blx_engine *eng = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr}; // start / gathered / reduced (blx_multi_nearest)
    float *d_local = nullptr; // this device's block of force vectors, padded to `pad` rows
    float *d_all = nullptr;   // the gathered table (G * pad rows) and its compacted form (n rows)
    float *d_table = nullptr;
    int *d_idx = nullptr;
    float *d_dist = nullptr;
    size_t cap_rows = 0, cap_all = 0;
    int lo = 0, hi = 0; // block of the resident force vectors of the last analysed / uploaded batch
    int last_taken = 0; // songs this device analysed in the last batch (dynamic units)
};
} // namespace

struct blx_multi {
    std::vector<Dev> dev;
    bool nccl = false;
    int last_n = 0; // songs of the batch whose vectors are resident
    int unit_songs = 128;
    char transport[64] = "single device";
};

#define MCK(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return blx_set_error(BLX_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static void shard(int n, int r, int G, int *lo, int *hi) {
    const int base = n / G, extra = n % G;
    *lo = r * base + std::min(r, extra);
    *hi = *lo + base + (r < extra ? 1 : 0);
}

extern "C" int blx_multi_init(const int *devices, int n_devices, blx_multi **out) {
    if (!out) return blx_set_error(BLX_ERR_ARG, "null out pointer");
    *out = nullptr;
    const int avail = blx_device_count();
    if (avail <= 0) return blx_set_error(BLX_ERR_CUDA, "no CUDA device available; this engine has no CPU path");
    std::vector<int> ids;
    if (devices && n_devices > 0) ids.assign(devices, devices + n_devices);
    else for (int i = 0; i < (n_devices > 0 ? std::min(n_devices, avail) : avail); ++i) ids.push_back(i);
    blx_multi *m = new blx_multi();
    m->dev.resize(ids.size());
    for (size_t r = 0; r < ids.size(); ++r) {
        m->dev[r].device = ids[r];
        int rc = blx_init(ids[r], &m->dev[r].eng);
        if (rc != BLX_OK) { blx_multi_shutdown(m); return rc; }
        cudaError_t ce = cudaSetDevice(ids[r]);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&m->dev[r].st, cudaStreamNonBlocking);
        for (int k = 0; k < 3 && ce == cudaSuccess; ++k) ce = cudaEventCreate(&m->dev[r].ev[k]);
        if (ce != cudaSuccess) {
            blx_multi_shutdown(m);
            return blx_set_error(BLX_ERR_CUDA, "device %d: %s", ids[r], cudaGetErrorString(ce));
        }
    }
    const int G = (int)ids.size();
    if (G > 1) {
        std::vector<ncclComm_t> comms(G);
        const bool have = nccl_load();
        const ncclResult_t nr = have ? g_nccl.CommInitAll(comms.data(), G, ids.data()) : ncclSystemError;
        if (nr == ncclSuccess) {
            for (int r = 0; r < G; ++r) m->dev[r].comm = comms[r];
            m->nccl = true;
            int ver = 0;
            g_nccl.GetVersion(&ver);
            snprintf(m->transport, sizeof(m->transport), "ncclAllGather (NCCL %d.%d.%d)", ver / 10000, (ver / 100) % 100, ver % 100);
        } else {
            // no NCCL: peer-to-peer copies (NVLink when peer access can be enabled, staged through the host otherwise)
            for (int r = 0; r < G; ++r) {
                cudaSetDevice(ids[r]);
                for (int s = 0; s < G; ++s)
                    if (s != r) {
                        int can = 0;
                        cudaDeviceCanAccessPeer(&can, ids[r], ids[s]);
                        if (can && cudaDeviceEnablePeerAccess(ids[s], 0) != cudaSuccess) cudaGetLastError();
                    }
            }
            snprintf(m->transport, sizeof(m->transport), "cudaMemcpyPeerAsync (NCCL: %s)", have ? g_nccl.GetErrorString(nr) : "libnccl.so.2 not found");
        }
    }
    *out = m;
    return BLX_OK;
}

extern "C" void blx_multi_shutdown(blx_multi *m) {
    if (!m) return;
    for (Dev &d : m->dev) {
        cudaSetDevice(d.device);
        if (d.comm) g_nccl.CommDestroy(d.comm);
        if (d.st) { cudaStreamSynchronize(d.st); cudaStreamDestroy(d.st); }
        for (cudaEvent_t e : d.ev) if (e) cudaEventDestroy(e);
        cudaFree(d.d_local); cudaFree(d.d_all); cudaFree(d.d_table); cudaFree(d.d_idx); cudaFree(d.d_dist);
        if (d.eng) blx_shutdown(d.eng);
    }
    delete m;
}

extern "C" int blx_multi_device_count(blx_multi *m) { return m ? (int)m->dev.size() : 0; }
extern "C" const char *blx_multi_transport(blx_multi *m) { return m ? m->transport : ""; }
extern "C" int blx_multi_songs_taken(blx_multi *m, int rank) {
    return (m && rank >= 0 && rank < (int)m->dev.size()) ? m->dev[rank].last_taken : -1;
}
extern "C" blx_engine *blx_multi_engine(blx_multi *m, int rank) {
    return (m && rank >= 0 && rank < (int)m->dev.size()) ? m->dev[rank].eng : nullptr;
}

// ---------------------------------------------------------------- sharded analysis
namespace {
struct Job {
    blx_multi *m;
    int rank;
    int n_songs;
    int unit;                 // songs per unit of work taken from the shared counter
    int *next;                // shared: first song nobody has taken yet (guarded by *lock)
    pthread_mutex_t *lock;
    int taken;                // songs this device ended up analysing
    int fmt; // BLX_FMT_*
    const void *const *pcm;
    const int *n_samples;     // s16
    const int64_t *n_in;      // f32
    const int *channels;
    const uint64_t *duration_s;
    unsigned what;
    blx_result *out;
    int rc;
    char err[256];
};

int upload_vectors(Dev &d, const blx_result *res, int n_local, int pad) {
    MCK(cudaSetDevice(d.device));
    if ((size_t)pad > d.cap_rows) {
        cudaFree(d.d_local); cudaFree(d.d_idx); cudaFree(d.d_dist);
        d.d_local = nullptr; d.d_idx = nullptr; d.d_dist = nullptr;
        MCK(cudaMalloc(&d.d_local, (size_t)pad * 16));
        MCK(cudaMalloc(&d.d_idx, (size_t)pad * 4));
        MCK(cudaMalloc(&d.d_dist, (size_t)pad * 4));
        d.cap_rows = pad;
    }
    std::vector<float> v((size_t)pad * 4, 0.0f);
    for (int i = 0; i < n_local; ++i) {
        v[4 * i] = res[i].tempo; v[4 * i + 1] = res[i].amplitude; v[4 * i + 2] = res[i].frequency; v[4 * i + 3] = res[i].attack;
    }
    MCK(cudaMemcpyAsync(d.d_local, v.data(), v.size() * 4, cudaMemcpyHostToDevice, d.st));
    MCK(cudaStreamSynchronize(d.st));
    return BLX_OK;
}

// Every device thread takes units of consecutive songs from a shared counter until none are left: a device behind a
// slower host link (on the 8-GPU boxes of this pool four GPUs copy at 23 GB/s and four at 35 GB/s when all are busy,
// profiles/r2_h2d_probe_n8.json) simply ends up with fewer songs. A record depends on its song alone, so the
// result array is the same whichever device analysed which unit.
void *analyze_thread(void *arg) {
    Job *j = static_cast<Job *>(arg);
    Dev &d = j->m->dev[j->rank];
    j->rc = BLX_OK;
    j->taken = 0;
    for (;;) {
        pthread_mutex_lock(j->lock);
        const int lo = *j->next;
        const int hi = std::min(j->n_songs, lo + j->unit);
        *j->next = hi;
        pthread_mutex_unlock(j->lock);
        if (lo >= hi) break;
        const int n = hi - lo;
        if (j->fmt == BLX_FMT_S16)
            j->rc = blx_analyze_batch_s16(d.eng, reinterpret_cast<const int16_t *const *>(j->pcm) + lo, j->n_samples + lo,
                                          j->channels ? j->channels + lo : nullptr, j->duration_s ? j->duration_s + lo : nullptr, n,
                                          j->what, j->out + lo);
        else
            j->rc = blx_analyze_batch_f32(d.eng, reinterpret_cast<const float *const *>(j->pcm) + lo, j->n_in + lo, n, j->what,
                                          j->out + lo);
        if (j->rc != BLX_OK) {
            snprintf(j->err, sizeof(j->err), "device %d: %s", d.device, blx_last_error());
            break;
        }
        j->taken += n;
    }
    return nullptr;
}

int analyze_sharded(blx_multi *m, Job proto, int n_songs) {
    if (!m) return blx_set_error(BLX_ERR_ARG, "null blx_multi");
    m->last_n = 0; // nothing resident until this batch has gone through
    if (n_songs <= 0) return BLX_OK;
    if (!proto.pcm || !proto.out) return blx_set_error(BLX_ERR_ARG, "null input array");
    const int G = (int)m->dev.size();
    std::vector<Job> jobs(G, proto);
    std::vector<pthread_t> th(G);
    // units of work: large enough for the engine to overlap the copy of one staging chunk with the kernels of the
    // previous one inside a unit, small enough that the devices finish together
    int next = 0;
    pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
    const int unit = std::max(1, std::min(m->unit_songs, (n_songs + 4 * G - 1) / (4 * G)));
    for (int r = 0; r < G; ++r) {
        shard(n_songs, r, G, &m->dev[r].lo, &m->dev[r].hi); // blocks of the resident force vectors (blx_multi_nearest)
        jobs[r].m = m;
        jobs[r].rank = r;
        jobs[r].n_songs = n_songs;
        jobs[r].unit = unit;
        jobs[r].next = &next;
        jobs[r].lock = &lock;
        jobs[r].err[0] = 0;
        if (pthread_create(&th[r], nullptr, analyze_thread, &jobs[r]) != 0) {
            for (int s = 0; s < r; ++s) pthread_join(th[s], nullptr);
            return blx_set_error(BLX_ERR_NOMEM, "cannot start a host thread for device %d", m->dev[r].device);
        }
    }
    for (int r = 0; r < G; ++r) pthread_join(th[r], nullptr);
    for (int r = 0; r < G; ++r) {
        m->dev[r].last_taken = jobs[r].taken;
        if (jobs[r].rc != BLX_OK) return blx_set_error(jobs[r].rc, "%s", jobs[r].err);
    }
    // keep every block's force vectors resident on its device for blx_multi_nearest
    const int pad = (n_songs + G - 1) / G;
    for (int r = 0; r < G; ++r) {
        Dev &d = m->dev[r];
        int rc = upload_vectors(d, proto.out + d.lo, d.hi - d.lo, pad);
        if (rc) return rc;
    }
    m->last_n = n_songs;
    return BLX_OK;
}
} // namespace

extern "C" int blx_multi_analyze_batch_s16(blx_multi *m, const int16_t *const *pcm, const int *n_samples, const int *channels,
                                           const uint64_t *duration_s, int n_songs, unsigned what, blx_result *out) {
    if (n_songs > 0 && !n_samples) return blx_set_error(BLX_ERR_ARG, "null n_samples");
    Job j;
    memset(&j, 0, sizeof(j));
    j.fmt = BLX_FMT_S16; j.pcm = reinterpret_cast<const void *const *>(pcm); j.n_samples = n_samples; j.channels = channels;
    j.duration_s = duration_s; j.what = what; j.out = out;
    return analyze_sharded(m, j, n_songs);
}

extern "C" int blx_multi_analyze_batch_f32(blx_multi *m, const float *const *pcm, const int64_t *n_in, int n_songs, unsigned what,
                                           blx_result *out) {
    if (n_songs > 0 && !n_in) return blx_set_error(BLX_ERR_ARG, "null n_in");
    Job j;
    memset(&j, 0, sizeof(j));
    j.fmt = BLX_FMT_F32; j.pcm = reinterpret_cast<const void *const *>(pcm); j.n_in = n_in; j.what = what; j.out = out;
    return analyze_sharded(m, j, n_songs);
}

// ---------------------------------------------------------------- all-gather + all-pairs nearest neighbours
extern "C" int blx_multi_set_vectors(blx_multi *m, const float *vectors, int n) {
    if (!m || !vectors || n <= 0) return blx_set_error(BLX_ERR_ARG, "bad arguments");
    const int G = (int)m->dev.size(), pad = (n + G - 1) / G;
    std::vector<blx_result> tmp;
    for (int r = 0; r < G; ++r) {
        Dev &d = m->dev[r];
        shard(n, r, G, &d.lo, &d.hi);
        tmp.assign((size_t)std::max(d.hi - d.lo, 1), blx_result());
        for (int i = d.lo; i < d.hi; ++i) {
            blx_result &q = tmp[i - d.lo];
            q.tempo = vectors[4 * i]; q.amplitude = vectors[4 * i + 1]; q.frequency = vectors[4 * i + 2]; q.attack = vectors[4 * i + 3];
        }
        int rc = upload_vectors(d, tmp.data(), d.hi - d.lo, pad);
        if (rc) return rc;
    }
    m->last_n = n;
    return BLX_OK;
}

extern "C" int blx_multi_nearest(blx_multi *m, int *nearest_index, float *nearest_dist, float *gather_ms, float *nearest_ms) {
    if (!m || m->last_n <= 0) return blx_set_error(BLX_ERR_ARG, "no resident force vectors (analyse a batch or call blx_multi_set_vectors first)");
    if (!nearest_index && !nearest_dist) return blx_set_error(BLX_ERR_ARG, "no output requested");
    const int G = (int)m->dev.size(), n = m->last_n, pad = (n + G - 1) / G;
    for (int r = 0; r < G; ++r) {
        Dev &d = m->dev[r];
        MCK(cudaSetDevice(d.device));
        if ((size_t)G * pad > d.cap_all) {
            cudaFree(d.d_all); cudaFree(d.d_table);
            d.d_all = nullptr; d.d_table = nullptr;
            MCK(cudaMalloc(&d.d_all, (size_t)G * pad * 16));
            MCK(cudaMalloc(&d.d_table, (size_t)G * pad * 16));
            d.cap_all = (size_t)G * pad;
        }
        MCK(cudaEventRecord(d.ev[0], d.st));
    }
    // (1) all-gather of the padded blocks, device to device
    if (G == 1) {
        MCK(cudaMemcpyAsync(m->dev[0].d_all, m->dev[0].d_local, (size_t)pad * 16, cudaMemcpyDeviceToDevice, m->dev[0].st));
    } else if (m->nccl) {
        g_nccl.GroupStart();
        for (int r = 0; r < G; ++r) {
            Dev &d = m->dev[r];
            const ncclResult_t nr = g_nccl.AllGather(d.d_local, d.d_all, (size_t)pad * 4, ncclFloat, d.comm, d.st);
            if (nr != ncclSuccess) { g_nccl.GroupEnd(); return blx_set_error(BLX_ERR_CUDA, "ncclAllGather: %s", g_nccl.GetErrorString(nr)); }
        }
        const ncclResult_t nr = g_nccl.GroupEnd();
        if (nr != ncclSuccess) return blx_set_error(BLX_ERR_CUDA, "ncclGroupEnd: %s", g_nccl.GetErrorString(nr));
    } else {
        for (int r = 0; r < G; ++r)
            for (int s = 0; s < G; ++s) {
                MCK(cudaSetDevice(m->dev[r].device));
                MCK(cudaMemcpyPeerAsync(m->dev[r].d_all + (size_t)s * pad * 4, m->dev[r].device, m->dev[s].d_local, m->dev[s].device,
                                        (size_t)pad * 16, m->dev[r].st));
            }
    }
    // (2) drop the padding (blocks are ragged when G does not divide n), then (3) this device's rows against all columns
    for (int r = 0; r < G; ++r) {
        Dev &d = m->dev[r];
        MCK(cudaSetDevice(d.device));
        for (int s = 0; s < G; ++s) {
            int lo, hi;
            shard(n, s, G, &lo, &hi);
            if (hi > lo)
                MCK(cudaMemcpyAsync(d.d_table + (size_t)lo * 4, d.d_all + (size_t)s * pad * 4, (size_t)(hi - lo) * 16,
                                    cudaMemcpyDeviceToDevice, d.st));
        }
        MCK(cudaEventRecord(d.ev[1], d.st));
        if (d.hi > d.lo) {
            int rc = blx_distance_nearest_device(d.eng, d.d_table, n, d.lo, d.hi - d.lo, d.d_idx, d.d_dist, nullptr, d.st);
            if (rc) return rc;
            if (nearest_index) MCK(cudaMemcpyAsync(nearest_index + d.lo, d.d_idx, (size_t)(d.hi - d.lo) * 4, cudaMemcpyDeviceToHost, d.st));
            if (nearest_dist) MCK(cudaMemcpyAsync(nearest_dist + d.lo, d.d_dist, (size_t)(d.hi - d.lo) * 4, cudaMemcpyDeviceToHost, d.st));
        }
        MCK(cudaEventRecord(d.ev[2], d.st));
    }
    float g_ms = 0, n_ms = 0;
    for (int r = 0; r < G; ++r) {
        MCK(cudaSetDevice(m->dev[r].device));
        MCK(cudaStreamSynchronize(m->dev[r].st));
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, m->dev[r].ev[0], m->dev[r].ev[1]);
        cudaEventElapsedTime(&b, m->dev[r].ev[1], m->dev[r].ev[2]);
        g_ms = std::max(g_ms, a);
        n_ms = std::max(n_ms, b);
    }
    if (gather_ms) *gather_ms = g_ms;
    if (nearest_ms) *nearest_ms = n_ms;
    return BLX_OK;
}
