/*
 * api.cpp -- C ABI of libhvqm4_b200.so (include/hvqm4.h): the SDK-compatible seven entry
 * points, the batched multi-stream runtime, and the .h4m container walker.
 *
 * Runtime design (B200-first):
 *   - one HVQM4Batch per GPU; streams are independent, so multi-GPU = one batch per
 *     device/process with no communication;
 *   - a step decodes one picture for each of n streams:  host threads run the serial
 *     stage (entropy.c) in two parallel phases (sizes, then emission straight into a
 *     pinned arena), ONE cudaMemcpyAsync uploads the whole arena on the copy stream, ONE
 *     kernel launch reconstructs all n pictures on the compute stream;
 *   - arenas form a ring of kArenas, so the host stage of the following steps overlaps upload and
 *     reconstruction of step k; frame surfaces never leave HBM unless asked for.
 *
 * The reference equivalents: the call protocol of main()/decode_video()
 * (/root/reference/h4m_audio_decode.c:2409-2419, 2078-2138) and the buffer rotation
 * rule (2087-2093, 2131-2137), which HVQM4Batch applies per stream.
 */
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/hvqm4.h"
#include "entropy.h"
#include "entropy_dev.h"
#include "recon.h"

#define H4_API extern "C" __attribute__((visibility("default")))

namespace {

std::atomic<int> g_last_cuda_error{0};
std::atomic<long long> g_launches{0};

bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    g_last_cuda_error = (int)e;
    fprintf(stderr, "hvqm4_b200: %s failed: %s\n", what, cudaGetErrorString(e));
    return false;
}

bool have_device()
{
    static int state = -1;
    if (state < 0)
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        state = (e == cudaSuccess && n > 0) ? 1 : 0;
        if (!state)
        {
            cudaGetLastError();
            fprintf(stderr, "hvqm4_b200: no CUDA device available -- this decoder has no CPU reconstruction path\n");
        }
    }
    return state == 1;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

/* ------------------------------------------------------------------ host thread pool */

class Pool
{
    /* one Job object per parallel_for: a worker that is late leaving the previous job can
       only ever touch that job's own counters */
    struct Job
    {
        const std::function<void(int)> *fn;
        int n;
        std::atomic<int> next{0};
        std::atomic<int> pending;
        Job(const std::function<void(int)> *f, int count) : fn(f), n(count), pending(count) {}
    };

public:
    explicit Pool(int n_threads)
    {
        for (int i = 0; i < n_threads; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~Pool()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    /* runs f(i) for i in [0,n); the calling thread takes part */
    void parallel_for(int n, const std::function<void(int)> &f)
    {
        if (n <= 0) return;
        if (workers_.empty() || n == 1)
        {
            for (int i = 0; i < n; ++i) f(i);
            return;
        }
        auto job = std::make_shared<Job>(&f, n);
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = job;
            ++epoch_;
        }
        cv_.notify_all();
        run(*job);
        std::unique_lock<std::mutex> l(m_);
        done_.wait(l, [&] { return job->pending.load() == 0; });
        job_.reset();
    }
    int size() const { return (int)workers_.size() + 1; }

private:
    void run(Job &j)
    {
        for (;;)
        {
            const int i = j.next.fetch_add(1);
            if (i >= j.n) break;
            (*j.fn)(i);
            if (j.pending.fetch_sub(1) == 1)
            {
                std::lock_guard<std::mutex> l(m_);
                done_.notify_all();
            }
        }
    }
    void loop()
    {
        unsigned long seen = 0;
        for (;;)
        {
            std::shared_ptr<Job> j;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [&] { return stop_ || epoch_ != seen; });
                if (stop_) return;
                seen = epoch_;
                j = job_;
            }
            if (j) run(*j);
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::shared_ptr<Job> job_;
    unsigned long epoch_ = 0;
    bool stop_ = false;
};

}  // namespace

/* Device -> host copy of one frame that also works when the destination STRADDLES page-locked and pageable memory: a heap
   block next to a range registered with HVQM4HostRegister shares the range's first or last page (registration is by
   page), and cudaMemcpyAsync refuses such a destination with cudaErrorInvalidValue.  Then the frame takes the detour over
   a page-locked bounce buffer of the library (synchronous: the case is rare and only costs time). */
struct Bounce
{
    uint8_t *h = nullptr;
    size_t cap = 0;
    bool reserve(size_t bytes)
    {
        if (bytes <= cap) return true;
        if (h) cudaFreeHost(h);
        h = nullptr;
        cap = 0;
        if (!cuda_ok(cudaHostAlloc((void **)&h, bytes, cudaHostAllocDefault), "cudaHostAlloc(bounce)")) return false;
        cap = bytes;
        return true;
    }
    void release() { if (h) cudaFreeHost(h); h = nullptr; cap = 0; }
};
static bool copy_frame_to_host(void *dst, const void *src, size_t bytes, cudaStream_t stream, Bounce *bounce, const char *what)
{
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) return true;
    if (e != cudaErrorInvalidValue) return cuda_ok(e, what);
    cudaGetLastError();
    if (!cuda_ok(cudaStreamSynchronize(stream), what) || !bounce->reserve(bytes)) return false;
    if (!cuda_ok(cudaMemcpyAsync(bounce->h, src, bytes, cudaMemcpyDeviceToHost, stream), what) ||
        !cuda_ok(cudaStreamSynchronize(stream), what))
        return false;
    memcpy(dst, bounce->h, bytes);
    return true;
}

/* the same for an application frame on its way to the device (reference frames of the SDK entry points, RGB conversion) */
static bool copy_frame_to_device(void *dst, const void *src, size_t bytes, cudaStream_t stream, Bounce *bounce, const char *what)
{
    const cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) return true;
    if (e != cudaErrorInvalidValue) return cuda_ok(e, what);
    cudaGetLastError();
    if (!cuda_ok(cudaStreamSynchronize(stream), what) || !bounce->reserve(bytes)) return false;      /* nothing reads the buffer any more */
    memcpy(bounce->h, src, bytes);
    return cuda_ok(cudaMemcpyAsync(dst, bounce->h, bytes, cudaMemcpyHostToDevice, stream), what) && cuda_ok(cudaStreamSynchronize(stream), what);
}

/* ====================================================================== batch runtime */

struct StreamState
{
    H4Seq *seq = nullptr;
    int past = 0, present = 1, future = 2;   /* surface indices, rotated like the reference's Player */
    int spare = 3;                           /* B pictures alternate between `present` and this one (see batch_after_picture) */
    int last = -1;                           /* surface holding the most recently decoded picture */
};

/* arenas (pinned host + device staging of one step) form a ring: the host may run this many steps
   ahead of the reconstruction, so that a step's gather / entropy work and upload never wait for the
   read-back chain of the step two before (measured: depth 2 left the copy stream idle ~1 ms per step) */
constexpr int kArenas = 4;

/* Frame surfaces per stream: the reference's three (past / present / future, h4m:2343-2349) plus a
   spare.  A B picture is never a reference (h4m:2063-2064), so the next picture may be written
   into a different surface than the B picture that is still being read back: `present` and the
   spare swap after every B picture, and a step only has to wait for the read-backs issued before
   the PREVIOUS step instead of for the previous step's own (which would put every reconstruction
   between two read-backs on the PCIe critical path). */
constexpr int kSurfaces = 4;

struct Arena
{
    uint8_t *h = nullptr;   /* pinned */
    uint8_t *d = nullptr;
    size_t cap = 0;
    cudaEvent_t consumed = nullptr;   /* recorded after the kernel that reads this arena */
    bool in_flight = false;
};

struct RecordedStep
{
    uint8_t *d = nullptr;   /* jobs + blobs */
    int n = 0;
    std::vector<uint32_t> rec_prefix;   /* n + 1 entries */
};

struct HVQM4Batch
{
    int device = 0, n_streams = 0, width = 0, height = 0, version15 = 1;
    int mcb_w = 0, mcb_h = 0;
    size_t frame_bytes = 0, surf_stride = 0;
    uint8_t *d_surfaces = nullptr;
    bool slab_ok = false;      /* the surface slab is registered with the row kernel (tensor maps exist) */
    int band_rows = 8;         /* macroblock rows per record band of this batch's streams (symbuf.h) */
    std::vector<StreamState> st;
    Arena arena[kArenas];
    int cur = 0;
    cudaStream_t s_copy = nullptr, s_comp = nullptr, s_d2h = nullptr;
    /* GPU entropy mode: consecutive steps upload + parse on these streams in turn (s_parse[0] = s_copy) */
    cudaStream_t s_parse[H4_PARSE_SLOTS] = {};
    cudaEvent_t ev_parse[H4_PARSE_SLOTS] = {};
    int parse_turn = 0, ipic_slot = 0, ipic_fence_left = 0;
    cudaEvent_t ev_h2d = nullptr, ev_kernel = nullptr, ev_d2h = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    std::vector<const uint8_t *> gather_src;        /* GPU entropy mode: device-visible source of every picture of the step */
    /* HVQM4_BATCH_TRACE=1: where the submitting thread spends a GPU-entropy step (printed by HVQM4BatchDestroy) */
    double t_trace[4] = {0, 0, 0, 0};               /* waiting for a staging arena, copying pictures, descriptors, enqueueing */
    unsigned n_trace = 0, steps_seen = 0;
    /* HVQM4_BATCH_TIMELINE=1: device-side timeline of the last GPU-entropy steps -- events on the streams around the upload,
       the parse kernel, the reconstruction and the read-back of every step (printed by HVQM4BatchDestroy) */
    struct StepMarks { cudaEvent_t e[7]; };         /* upload begin, upload end, parse end, recon begin, recon end, d2h begin, d2h end */
    std::vector<StepMarks> marks;
    long long mark_step = -1;                       /* step whose read-back comes next */
    bool d2h_pending = false;
    cudaEvent_t ev_d2h_mark[2] = {nullptr, nullptr};   /* read-backs issued before step n, n - 1 (batch_wait_readbacks) */
    unsigned step_no = 0;
    Pool *pool = nullptr;
    uint32_t errors = 0;
    bool recording = false;
    std::vector<RecordedStep> recorded;
    uint64_t stats[8] = {0};
    std::vector<size_t> sizes, offs;
    std::vector<uint32_t> rec_prefix;
    std::vector<uint8_t> seen;

    /* GPU entropy stage (HVQM4BatchSetEntropyMode): per-stream parser state, blob arena, counters */
    Bounce bounce;                 /* copy_frame_to_host */
    bool gpu_entropy = false;
    int host_share = 0;        /* GPU entropy mode: streams [0, host_share) are parsed by the host threads (HVQM4BatchSetHostShare) */
    uint8_t *d_estate = nullptr;
    size_t eslot = 0;
    uint8_t *d_blobs = nullptr;
    size_t blobs_cap = 0;
    unsigned long long *d_blob_used = nullptr;
    uint32_t *d_eerrors = nullptr;

    /* RGB read-back (HVQM4BatchReadFramesRGBAsync): staging frames + two rings of surface pointers */
    uint8_t *d_rgb = nullptr;
    const uint8_t **h_rgb_src[2] = {nullptr, nullptr};
    const uint8_t **d_rgb_src[2] = {nullptr, nullptr};
    cudaEvent_t ev_rgb_src[2] = {nullptr, nullptr}, ev_rgb = nullptr;
    int rgb_cur = 0;

    uint8_t *surface(int stream, int idx) const { return d_surfaces + ((size_t)stream * kSurfaces + idx) * surf_stride; }
};

static bool batch_create_parse_streams(HVQM4Batch *b)
{
    b->s_parse[0] = b->s_copy;
    for (int j = 0; j < H4_PARSE_SLOTS; ++j)
    {
        if (j && !cuda_ok(cudaStreamCreateWithFlags(&b->s_parse[j], cudaStreamNonBlocking), "cudaStreamCreate")) return false;
        if (!cuda_ok(cudaEventCreateWithFlags(&b->ev_parse[j], cudaEventDisableTiming), "cudaEventCreate")) return false;
    }
    return true;
}

/* ---- page-locked application memory (HVQM4HostRegister): the GPU entropy mode fetches pictures that lie in
   such a range itself (entropy_dev.cu: dev_gather_kernel) instead of having host threads copy them ---- */
struct HostRange
{
    uintptr_t begin, end;       /* as the application sees it */
    uintptr_t lo, hi;           /* the pages it touches */
    uint8_t *dev[16];           /* device-visible address of `begin` per device (filled on demand) */
};
/* Ranges of separate allocations may share pages, and a page can be registered only once: the
   pages are kept as disjoint registered segments, each counting the ranges that touch it. */
struct PageSegment
{
    uintptr_t lo, hi;
    int refs;
};
static std::mutex g_ranges_lock;
static std::vector<HostRange> g_ranges;
static std::vector<PageSegment> g_segments;

static void release_segments(uintptr_t lo, uintptr_t hi)
{
    for (size_t i = 0; i < g_segments.size();)
    {
        PageSegment &sg = g_segments[i];
        if (sg.lo < hi && lo < sg.hi && --sg.refs == 0)
        {
            cuda_ok(cudaHostUnregister((void *)sg.lo), "cudaHostUnregister");
            g_segments.erase(g_segments.begin() + (long)i);
        }
        else
            ++i;
    }
}

H4_API int HVQM4HostRegister(void *ptr, size_t bytes)
{
    if (!ptr || !bytes) return HVQM4_ERR_ARGUMENT;
    if (!have_device()) return HVQM4_ERR_NO_DEVICE;
    const uintptr_t page = 4096, lo = (uintptr_t)ptr & ~(page - 1), hi = ((uintptr_t)ptr + bytes + page - 1) & ~(page - 1);
    std::lock_guard<std::mutex> guard(g_ranges_lock);
    /* pages of [lo, hi) not yet registered: walk the segments that overlap in address order */
    std::vector<PageSegment> gaps;
    uintptr_t at = lo;
    while (at < hi)
    {
        const PageSegment *next = nullptr;
        for (const PageSegment &sg : g_segments)
            if (sg.hi > at && sg.lo < hi && (!next || sg.lo < next->lo)) next = &sg;
        const uintptr_t gap_end = next ? (next->lo > at ? next->lo : at) : hi;
        if (gap_end > at) gaps.push_back(PageSegment{at, gap_end, 0});
        at = next ? next->hi : hi;
    }
    size_t done = 0;
    for (; done < gaps.size(); ++done)
        if (!cuda_ok(cudaHostRegister((void *)gaps[done].lo, gaps[done].hi - gaps[done].lo, cudaHostRegisterPortable | cudaHostRegisterMapped),
                     "cudaHostRegister"))
            break;
    if (done < gaps.size())
    {
        for (size_t i = 0; i < done; ++i) cudaHostUnregister((void *)gaps[i].lo);
        return HVQM4_ERR_CUDA;
    }
    for (const PageSegment &g : gaps) g_segments.push_back(g);
    for (PageSegment &sg : g_segments)
        if (sg.lo < hi && lo < sg.hi) ++sg.refs;
    HostRange r{};
    r.begin = (uintptr_t)ptr;
    r.end = r.begin + bytes;
    r.lo = lo;
    r.hi = hi;
    g_ranges.push_back(r);
    return HVQM4_OK;
}

H4_API int HVQM4HostUnregister(void *ptr)
{
    /* gather kernels of steps that have been submitted may still be reading the range over PCIe: they are asynchronous to
       HVQM4BatchDecode.  Nothing of this process may touch the pages once they are unpinned. */
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> guard(g_ranges_lock);
    for (size_t i = 0; i < g_ranges.size(); ++i)
        if (g_ranges[i].begin == (uintptr_t)ptr)
        {
            release_segments(g_ranges[i].lo, g_ranges[i].hi);
            g_ranges.erase(g_ranges.begin() + (long)i);
            return HVQM4_OK;
        }
    return HVQM4_ERR_ARGUMENT;
}

/* device-visible address of [p, p + bytes) if it lies in a registered range, else NULL (caller holds the lock) */
static const uint8_t *registered_device_ptr(const uint8_t *p, size_t bytes, int device)
{
    const uintptr_t a = (uintptr_t)p;
    for (HostRange &r : g_ranges)
        if (a >= r.begin && a + bytes <= r.end)
        {
            if (device < 0 || device >= 16) return nullptr;
            if (!r.dev[device])
            {
                void *d = nullptr;
                if (cudaHostGetDevicePointer(&d, (void *)r.begin, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
                if (d != (void *)r.begin)
                {
                    /* no identity mapping: device addresses are contiguous only inside one registered segment */
                    bool one = false;
                    for (const PageSegment &sg : g_segments) one = one || (sg.lo <= r.lo && r.hi <= sg.hi);
                    if (!one) return nullptr;
                }
                r.dev[device] = static_cast<uint8_t *>(d);
            }
            return r.dev[device] + (a - r.begin);
        }
    return nullptr;
}

/* Called once per step before its reconstruction is enqueued: the surfaces the step writes were
   last read back two steps ago or earlier (kSurfaces), so the compute stream waits for the
   read-backs that had been issued when the previous step was submitted. */
static void batch_wait_readbacks(HVQM4Batch *b)
{
    const unsigned slot = b->step_no & 1;
    cudaEventRecord(b->ev_d2h_mark[slot], b->s_d2h);
    cudaStreamWaitEvent(b->s_comp, b->ev_d2h_mark[slot ^ 1], 0);
    ++b->step_no;
    b->d2h_pending = false;
}

constexpr int kTimelineSteps = 96;
static bool timeline_on()
{
    static const bool on = getenv("HVQM4_BATCH_TIMELINE") != nullptr;
    return on;
}
/* event k of the step in flight (nullptr: timeline off) */
static cudaEvent_t timeline_mark(HVQM4Batch *b, long long step, int k, cudaStream_t stream)
{
    if (!timeline_on() || step < 0) return nullptr;
    if (b->marks.empty())
    {
        b->marks.resize(kTimelineSteps);
        for (auto &m : b->marks)
            for (auto &e : m.e) cudaEventCreate(&e);
    }
    cudaEvent_t e = b->marks[(size_t)(step % kTimelineSteps)].e[k];
    cudaEventRecord(e, stream);
    return e;
}
static void timeline_print(HVQM4Batch *b)
{
    if (b->marks.empty() || b->steps_seen < 8) return;
    cudaDeviceSynchronize();
    const long long last = (long long)b->steps_seen - 1, first = last - 39 > 0 ? last - 39 : 0;
    if (last - first >= kTimelineSteps) return;
    cudaEvent_t t0 = b->marks[(size_t)(first % kTimelineSteps)].e[0];
    fprintf(stderr, "hvqm4_b200: device timeline of GPU-entropy steps %lld..%lld, ms since the first upload began\n"
                    "  step   upload          parse end   recon            read-back\n", first, last);
    for (long long s = first; s <= last; ++s)
    {
        float t[7];
        bool ok = true;
        for (int k = 0; k < 7; ++k) ok = ok && cudaEventElapsedTime(&t[k], t0, b->marks[(size_t)(s % kTimelineSteps)].e[k]) == cudaSuccess;
        if (!ok) { cudaGetLastError(); continue; }
        fprintf(stderr, "  %4lld   %7.2f-%7.2f   %7.2f   %7.2f-%7.2f   %7.2f-%7.2f\n", s, t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
    }
}

static bool arena_reserve(HVQM4Batch *b, Arena &a, size_t need)
{
    if (need <= a.cap) return true;
    size_t cap = align_up(need + need / 4 + (1 << 20), 1 << 20);
    if (a.h) cudaFreeHost(a.h);
    if (a.d) cudaFree(a.d);
    a.h = a.d = nullptr;
    a.cap = 0;
    if (!cuda_ok(cudaHostAlloc((void **)&a.h, cap, cudaHostAllocDefault), "cudaHostAlloc(arena)")) return false;
    if (!cuda_ok(cudaMalloc((void **)&a.d, cap), "cudaMalloc(arena)")) return false;
    a.cap = cap;
    (void)b;
    return true;
}

H4_API HVQM4Batch *HVQM4BatchCreate(int device, int n_streams, int width, int height, int version, int host_threads)
{
    if (n_streams <= 0 || (version != 13 && version != 15)) return nullptr;
    if (!have_device()) return nullptr;
    if (device >= 0 && !cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    if (device < 0) cudaGetDevice(&device);
    HVQM4Batch *b = new HVQM4Batch;
    b->device = device;
    b->n_streams = n_streams;
    b->width = width;
    b->height = height;
    b->version15 = version == 15;
    b->st.resize(n_streams);
    b->band_rows = hvqm4_recon_band_rows();      /* by the reconstruction schedule selected at this moment */
    h4e_set_band_rows(b->band_rows);
    for (int i = 0; i < n_streams; ++i)
    {
        b->st[i].seq = h4e_seq_create(width, height, 2, 2, b->version15);
        if (!b->st[i].seq)
        {
            HVQM4BatchDestroy(b);
            return nullptr;
        }
    }
    int dims[6];
    h4e_seq_dims(b->st[0].seq, dims);
    b->mcb_w = dims[2];
    b->mcb_h = dims[3];
    b->frame_bytes = h4e_frame_bytes(b->st[0].seq);
    /* 256-byte aligned surfaces with a tail so that the aligned 8-byte row reads of the
       half-sample filter never leave the allocation */
    b->surf_stride = align_up(b->frame_bytes + 64, 256);
    size_t total = b->surf_stride * kSurfaces * (size_t)n_streams + 256;
    if (!cuda_ok(cudaMalloc((void **)&b->d_surfaces, total), "cudaMalloc(surfaces)") ||
        !cuda_ok(cudaMemset(b->d_surfaces, 0, total), "cudaMemset(surfaces)") ||
        !cuda_ok(cudaStreamCreateWithFlags(&b->s_copy, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !batch_create_parse_streams(b) ||
        !cuda_ok(cudaStreamCreateWithFlags(&b->s_comp, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !cuda_ok(cudaStreamCreateWithFlags(&b->s_d2h, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !cuda_ok(cudaEventCreateWithFlags(&b->ev_h2d, cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(cudaEventCreateWithFlags(&b->ev_kernel, cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(cudaEventCreateWithFlags(&b->ev_d2h, cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(cudaEventCreateWithFlags(&b->ev_d2h_mark[0], cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(cudaEventCreateWithFlags(&b->ev_d2h_mark[1], cudaEventDisableTiming), "cudaEventCreate") ||
        !cuda_ok(cudaEventCreate(&b->ev_t0), "cudaEventCreate") || !cuda_ok(cudaEventCreate(&b->ev_t1), "cudaEventCreate"))
    {
        HVQM4BatchDestroy(b);
        return nullptr;
    }
    for (auto &a : b->arena)
        if (!cuda_ok(cudaEventCreateWithFlags(&a.consumed, cudaEventDisableTiming), "cudaEventCreate"))
        {
            HVQM4BatchDestroy(b);
            return nullptr;
        }
    /* tensor maps over the slab for the row kernel's reference patches (ragged widths: no maps, other kernels) */
    b->slab_ok = hvqm4_row_register_slab(b->d_surfaces, b->surf_stride, kSurfaces * n_streams, width, height) == 0;
    if (host_threads <= 0)
    {
        host_threads = (int)std::thread::hardware_concurrency();
        if (host_threads > 64) host_threads = 64;
        if (host_threads < 1) host_threads = 1;
    }
    if (host_threads > n_streams) host_threads = n_streams;
    b->pool = new Pool(host_threads - 1);
    b->sizes.resize(n_streams);
    b->offs.resize(n_streams);
    b->seen.assign(n_streams, 0);
    return b;
}

H4_API void HVQM4BatchDestroy(HVQM4Batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    b->bounce.release();
    timeline_print(b);
    for (auto &m : b->marks)
        for (auto &e : m.e) cudaEventDestroy(e);
    if (b->n_trace && getenv("HVQM4_BATCH_TRACE"))
        fprintf(stderr, "hvqm4_b200: %u GPU-entropy steps, submitting thread per step: arena wait %.2f ms, picture copies %.2f ms, descriptors %.2f ms, enqueue %.2f ms\n",
                b->n_trace, 1e3 * b->t_trace[0] / b->n_trace, 1e3 * b->t_trace[1] / b->n_trace, 1e3 * b->t_trace[2] / b->n_trace, 1e3 * b->t_trace[3] / b->n_trace);
    delete b->pool;
    for (auto &s : b->st) h4e_seq_destroy(s.seq);
    for (auto &r : b->recorded) cudaFree(r.d);
    for (auto &a : b->arena)
    {
        if (a.h) cudaFreeHost(a.h);
        if (a.d) cudaFree(a.d);
        if (a.consumed) cudaEventDestroy(a.consumed);
    }
    if (b->slab_ok) hvqm4_row_unregister_slab(b->d_surfaces);
    if (b->d_surfaces) cudaFree(b->d_surfaces);
    if (b->d_rgb) cudaFree(b->d_rgb);
    for (int i = 0; i < 2; ++i)
    {
        if (b->h_rgb_src[i]) cudaFreeHost((void *)b->h_rgb_src[i]);
        if (b->d_rgb_src[i]) cudaFree((void *)b->d_rgb_src[i]);
        if (b->ev_rgb_src[i]) cudaEventDestroy(b->ev_rgb_src[i]);
    }
    if (b->ev_rgb) cudaEventDestroy(b->ev_rgb);
    if (b->d_estate) cudaFree(b->d_estate);
    if (b->d_blobs) cudaFree(b->d_blobs);
    if (b->d_blob_used) cudaFree(b->d_blob_used);
    if (b->d_eerrors) cudaFree(b->d_eerrors);
    if (b->s_copy) cudaStreamDestroy(b->s_copy);
    for (int j = 1; j < H4_PARSE_SLOTS; ++j)
        if (b->s_parse[j]) cudaStreamDestroy(b->s_parse[j]);
    for (cudaEvent_t e : b->ev_parse)
        if (e) cudaEventDestroy(e);
    if (b->s_comp) cudaStreamDestroy(b->s_comp);
    if (b->s_d2h) cudaStreamDestroy(b->s_d2h);
    for (cudaEvent_t e : {b->ev_h2d, b->ev_kernel, b->ev_d2h, b->ev_t0, b->ev_t1, b->ev_d2h_mark[0], b->ev_d2h_mark[1]})
        if (e) cudaEventDestroy(e);
    delete b;
}

/*
 * Step with the bitstream stage on the GPU: the host only gathers the raw picture bytes into the
 * pinned arena (parallel memcpy) and fills in the surface pointers; one H2D copy, then the parse
 * kernel (one warp per picture, entropy_dev.cu) writes the symbol buffers into a device arena and
 * the fused band kernel reconstructs from them.  Symbol buffers never exist on the host.
 */
static int batch_decode_gpu_entropy(HVQM4Batch *b, int n, const int32_t *stream_ids, const int32_t *frame_types,
                                    const uint8_t *const *frames, const uint32_t *frame_bytes)
{
    if (b->recording) return HVQM4_ERR_ARGUMENT;
    auto t_host0 = std::chrono::steady_clock::now();
    const size_t pics_bytes = align_up((size_t)n * sizeof(H4DevPicture), 256);
    const size_t jobs_bytes = align_up((size_t)n * sizeof(ReconJob), 256);
    const size_t gather_bytes = align_up((size_t)n * sizeof(H4Gather), 256);
    /* pictures in registered (page-locked, mapped) memory are fetched by the GPU; all of a step or none */
    bool gather = true;
    {
        std::lock_guard<std::mutex> guard(g_ranges_lock);
        gather = !g_ranges.empty();
        b->gather_src.resize((size_t)n);
        for (int i = 0; i < n && gather; ++i)
        {
            b->gather_src[i] = registered_device_ptr(frames[i], frame_bytes[i], b->device);
            gather = b->gather_src[i] != nullptr;
        }
    }
    const size_t head_bytes = pics_bytes + jobs_bytes + (gather ? gather_bytes : 0);
    /* Pictures of the host's share of the streams (HVQM4BatchSetHostShare) go through the host stage of entropy.c on the
       thread pool: their symbol buffers are built in the pinned arena, right behind the tables, and the parse kernel skips
       them.  Phase A (sizes) comes first, like in the host-entropy step. */
    const int share = b->host_share;
    int n_host = 0;
    for (int i = 0; i < n; ++i) n_host += stream_ids[i] < share;
    if (n_host)
        b->pool->parallel_for(n, [&](int i) {
            if (stream_ids[i] < share) b->sizes[i] = h4e_parse_begin(b->st[stream_ids[i]].seq, frame_types[i], frames[i], frame_bytes[i]);
        });
    size_t total = align_up(head_bytes, 128);
    for (int i = 0; i < n; ++i)
    {
        if (stream_ids[i] >= share) continue;
        if (b->sizes[i] == 0) return HVQM4_ERR_GEOMETRY;
        b->offs[i] = total;
        total += align_up(b->sizes[i], 128);
    }
    const size_t upload_bytes = total;      /* what the host always uploads: tables + the host share's symbol buffers */
    for (int i = 0; i < n; ++i)
    {
        if (stream_ids[i] < share) continue;
        /* the device copy keeps the source's alignment modulo 16 when the GPU gathers it */
        b->offs[i] = total + (gather ? (size_t)((uintptr_t)b->gather_src[i] & 15u) : 0);
        total += align_up((size_t)frame_bytes[i] + 16, 16) + (gather ? 16 : 0);
    }
    /* everything that can refuse the step comes before anything of the batch changes (surface rotation, parser turn,
       I-picture fences): an error return leaves the batch where it was */
    if (total > 0xFFFFFFFFull) return HVQM4_ERR_OVERFLOW;
    Arena &a = b->arena[b->cur];
    const auto t_w0 = std::chrono::steady_clock::now();
    if (a.in_flight)
    {
        if (!cuda_ok(cudaEventSynchronize(a.consumed), "cudaEventSynchronize")) return HVQM4_ERR_CUDA;
        a.in_flight = false;
    }
    const auto t_w1 = std::chrono::steady_clock::now();
    if (!arena_reserve(b, a, total)) return HVQM4_ERR_NOMEM;
    H4DevPicture *pics = reinterpret_cast<H4DevPicture *>(a.h);
    ReconJob *jobs = reinterpret_cast<ReconJob *>(a.h + pics_bytes);
    H4Gather *gd = reinterpret_cast<H4Gather *>(a.h + pics_bytes + jobs_bytes);
    std::atomic<uint32_t> host_err{0};
    if (!gather || n_host)
        b->pool->parallel_for(n, [&](int i) {
            uint8_t *dst = a.h + b->offs[i];
            if (stream_ids[i] < share)
                host_err |= h4e_parse_finish(b->st[stream_ids[i]].seq, dst);      /* phase B: the symbol buffer itself */
            else if (!gather)
            {
                memcpy(dst, frames[i], frame_bytes[i]);
                memset(dst + frame_bytes[i], 0, 16);
            }
        });
    b->errors |= host_err.load();
    const auto t_w2 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i)
    {
        const bool on_host = stream_ids[i] < share;
        if (gather)
        {   /* a picture of the host's share: nothing to fetch */
            gd[i].src = b->gather_src[i];
            gd[i].dst_off = (uint32_t)b->offs[i];
            gd[i].bytes = on_host ? 0u : frame_bytes[i];
        }
        StreamState &s = b->st[stream_ids[i]];
        const int t = frame_types[i];
        pics[i].data = a.d + b->offs[i];
        pics[i].bytes = frame_bytes[i];
        pics[i].pic_type = on_host ? 0 : t;      /* 0: the parse kernel leaves the picture and its job alone */
        pics[i].stream = stream_ids[i];
        pics[i].pad = 0;
        if (t != SYM_PIC_B) std::swap(s.past, s.future);
        jobs[i].blob = on_host ? a.d + b->offs[i] : nullptr;
        jobs[i].present = b->surface(stream_ids[i], s.present);
        jobs[i].past = b->surface(stream_ids[i], s.past);
        jobs[i].future = b->surface(stream_ids[i], t == SYM_PIC_P ? s.present : s.future);
        jobs[i].rec_cta_begin = 0;
        jobs[i].n_chunks = on_host ? h4e_last_chunks(s.seq) : 0;
        jobs[i].pad[0] = jobs[i].pad[1] = 0;
        s.last = s.present;
        if (t != SYM_PIC_B) std::swap(s.present, s.future);
        else std::swap(s.present, s.spare);
    }
    auto t_host1 = std::chrono::steady_clock::now();
    b->stats[4] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_host1 - t_host0).count();
    const bool traced = ++b->steps_seen > 32;       /* the first steps size the arenas (allocations synchronise the device) */
    if (traced)
    {
        b->t_trace[0] += std::chrono::duration<double>(t_w1 - t_w0).count();
        b->t_trace[1] += std::chrono::duration<double>(t_w2 - t_w1).count();
        b->t_trace[2] += std::chrono::duration<double>(t_host1 - t_w2).count() + std::chrono::duration<double>(t_w0 - t_host0).count();
        ++b->n_trace;
    }

    /* Upload and parser run on one of H4_PARSE_SLOTS streams, taken in turn, so that the parse
       kernels of consecutive steps are in flight together (a warp per picture is latency bound: a
       thousand warps leave most of the GPU idle) while the steps before reconstruct and read
       back.  Parser slots and blob counters go with the stream; blob arenas follow the staging
       ring.  A step that contains an I picture runs alone: it rewrites the nest that every slot of
       its stream reads -- it waits for the parsers in flight, and the next steps wait for it. */
    const int par = b->parse_turn;
    b->parse_turn = (par + 1) % H4_PARSE_SLOTS;
    cudaStream_t sp = b->s_parse[par];
    bool has_ipic = false;
    for (int i = 0; i < n && !has_ipic; ++i) has_ipic = frame_types[i] == SYM_PIC_I;
    if (has_ipic)
    {
        for (int j = 0; j < H4_PARSE_SLOTS; ++j)
            if (j != par) cudaStreamWaitEvent(sp, b->ev_parse[j], 0);
        b->ipic_slot = par;
        b->ipic_fence_left = H4_PARSE_SLOTS - 1;
    }
    else if (b->ipic_fence_left > 0)
    {
        cudaStreamWaitEvent(sp, b->ev_parse[b->ipic_slot], 0);
        --b->ipic_fence_left;
    }
    /* from here on work is queued on the arena: it counts as in flight whatever happens next */
    a.in_flight = true;
    cudaEventRecord(a.consumed, sp);
    const long long tl_step = (long long)b->steps_seen - 1;
    timeline_mark(b, tl_step, 0, sp);
    if (!cuda_ok(cudaMemcpyAsync(a.d, a.h, gather ? upload_bytes : total, cudaMemcpyHostToDevice, sp), "cudaMemcpyAsync(H2D)")) return HVQM4_ERR_CUDA;
    timeline_mark(b, tl_step, 1, sp);
    if (gather)
    {
        const int grc = hvqm4_dev_gather(reinterpret_cast<const H4Gather *>(a.d + pics_bytes + jobs_bytes), n, a.d, sp);
        if (grc != 0)
        {
            cuda_ok((cudaError_t)grc, "gather kernel launch");
            return HVQM4_ERR_CUDA;
        }
        ++g_launches;
        b->stats[1] += 1;
    }
    cudaMemsetAsync(b->d_blob_used + par, 0, sizeof(unsigned long long), sp);
    ReconJob *d_jobs = reinterpret_cast<ReconJob *>(a.d + pics_bytes);
    int rc = hvqm4_dev_entropy_parse(b->d_estate, b->eslot, reinterpret_cast<const H4DevPicture *>(a.d), n, par,
                                     b->d_blobs + (size_t)b->cur * b->blobs_cap, b->d_blob_used + par, (unsigned long long)b->blobs_cap,
                                     d_jobs, b->d_eerrors, sp);
    cudaEventRecord(b->ev_parse[par], sp);
    timeline_mark(b, tl_step, 2, sp);
    cudaStreamWaitEvent(b->s_comp, b->ev_parse[par], 0);
    batch_wait_readbacks(b);
    timeline_mark(b, tl_step, 3, b->s_comp);
    b->mark_step = tl_step;
    if (rc == 0)
    {
        ++g_launches;
        b->stats[1] += 1;
        rc = hvqm4_recon_launch_band(d_jobs, n, b->mcb_w, b->mcb_h, b->slab_ok ? b->d_surfaces : nullptr, b->band_rows, b->s_comp);
        if (rc == 0)
        {
            ++g_launches;
            b->stats[1] += 1;
        }
    }
    if (rc != 0)
    {
        cuda_ok((cudaError_t)rc, "GPU entropy / recon kernel launch");
        return HVQM4_ERR_CUDA;
    }
    cudaEventRecord(a.consumed, b->s_comp);
    cudaEventRecord(b->ev_kernel, b->s_comp);
    timeline_mark(b, tl_step, 4, b->s_comp);
    a.in_flight = true;
    b->cur = (b->cur + 1) % kArenas;
    b->stats[0] += n;
    b->stats[2] += total;
    if (traced) b->t_trace[3] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host1).count();
    return HVQM4_OK;
}

H4_API void HVQM4DevEntropyProfile(uint64_t out[8])
{
    unsigned long long tmp[8];
    hvqm4_dev_entropy_profile(tmp);
    for (int i = 0; i < 8; ++i) out[i] = tmp[i];
}

/* the GPU entropy stage's buffers exist together or not at all */
static void batch_free_entropy_state(HVQM4Batch *b)
{
    cudaGetLastError();
    if (b->d_estate) cudaFree(b->d_estate);
    if (b->d_blobs) cudaFree(b->d_blobs);
    if (b->d_blob_used) cudaFree(b->d_blob_used);
    if (b->d_eerrors) cudaFree(b->d_eerrors);
    b->d_estate = nullptr; b->d_blobs = nullptr; b->d_blob_used = nullptr; b->d_eerrors = nullptr;
    b->eslot = 0;
    b->gpu_entropy = false;
}

H4_API int HVQM4BatchSetEntropyMode(HVQM4Batch *b, int gpu)
{
    if (!b) return HVQM4_ERR_ARGUMENT;
    /* the two stages keep separate per-stream state (maps, the last I picture's nest): a switch after the first
       picture would decode the following P/B pictures against state the other stage never saw */
    if (b->stats[0] > 0 && (gpu != 0) != b->gpu_entropy) return HVQM4_ERR_ARGUMENT;
    cudaSetDevice(b->device);
    if (!cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return HVQM4_ERR_CUDA;
    if (!gpu)
    {
        b->gpu_entropy = false;
        return HVQM4_OK;
    }
    if (!b->d_estate)
    {
        const uint32_t blocks = (uint32_t)(b->mcb_w * b->mcb_h * 6);
        /* 16 symbols per block of a section's plane: every type an encoder emits fits (a block has
           at most 15 bases); a pathological section beyond that is flagged HVQM4_ERR_TRUNCATED */
        const uint32_t sym_cap = 16, work_cap = blocks;
        hvqm4_dev_entropy_set_band_rows(b->band_rows);
        b->eslot = hvqm4_dev_entropy_slot_bytes(b->width, b->height, sym_cap, work_cap);
        if (!b->eslot) return HVQM4_ERR_GEOMETRY;
        /* symbol buffers of one step: twice the frame size covers the densest records; the constant covers the
           fixed parts (header, bordered maps, nest, tables) that dominate in tiny pictures */
        b->blobs_cap = (size_t)b->n_streams * align_up(2 * b->frame_bytes + 8192, 256);
        if (!cuda_ok(cudaMalloc((void **)&b->d_estate, b->eslot * (size_t)b->n_streams * H4_PARSE_SLOTS), "cudaMalloc(entropy state)") ||
            !cuda_ok(cudaMalloc((void **)&b->d_blobs, kArenas * b->blobs_cap), "cudaMalloc(blob arena)") ||
            !cuda_ok(cudaMalloc((void **)&b->d_blob_used, H4_PARSE_SLOTS * sizeof(unsigned long long)), "cudaMalloc") ||
            !cuda_ok(cudaMalloc((void **)&b->d_eerrors, sizeof(uint32_t)), "cudaMalloc") ||
            !cuda_ok(cudaMemset(b->d_eerrors, 0, sizeof(uint32_t)), "cudaMemset"))
        {
            batch_free_entropy_state(b);      /* all or nothing: a later call starts over */
            return HVQM4_ERR_NOMEM;
        }
        int rc = hvqm4_dev_entropy_init(b->d_estate, b->eslot, b->n_streams, b->width, b->height, b->version15, sym_cap, work_cap, b->s_comp);
        if (rc != 0 || !cuda_ok(cudaStreamSynchronize(b->s_comp), "GPU entropy state init"))
        {
            batch_free_entropy_state(b);
            return HVQM4_ERR_CUDA;
        }
    }
    b->gpu_entropy = true;
    return HVQM4_OK;
}

/* GPU entropy mode: the pictures of streams [0, n_streams) are parsed by the host threads instead (the host stage of
   entropy.c writes their symbol buffers into the step's pinned arena), next to the parse kernel that takes the rest */
H4_API int HVQM4BatchSetHostShare(HVQM4Batch *b, int n_streams)
{
    if (!b || n_streams < 0 || n_streams > b->n_streams) return HVQM4_ERR_ARGUMENT;
    /* a stream's maps and nest live with the stage that parsed its earlier pictures */
    if (b->stats[0] > 0 && n_streams != b->host_share) return HVQM4_ERR_ARGUMENT;
    b->host_share = n_streams;
    return HVQM4_OK;
}

H4_API int HVQM4BatchDecode(HVQM4Batch *b, int n, const int32_t *stream_ids, const int32_t *frame_types,
                            const uint8_t *const *frames, const uint32_t *frame_bytes)
{
    if (!b || n <= 0 || n > b->n_streams || !stream_ids || !frame_types || !frames || !frame_bytes) return HVQM4_ERR_ARGUMENT;
    cudaSetDevice(b->device);
    std::fill(b->seen.begin(), b->seen.end(), 0);
    for (int i = 0; i < n; ++i)
    {
        int s = stream_ids[i], t = frame_types[i];
        if (s < 0 || s >= b->n_streams || b->seen[s] || (t != SYM_PIC_I && t != SYM_PIC_P && t != SYM_PIC_B)) return HVQM4_ERR_ARGUMENT;
        b->seen[s] = 1;
    }
    if (b->gpu_entropy) return batch_decode_gpu_entropy(b, n, stream_ids, frame_types, frames, frame_bytes);
    auto t_host0 = std::chrono::steady_clock::now();
    /* phase A: headers, trees, maps, work-order offsets -> exact blob sizes */
    b->pool->parallel_for(n, [&](int i) {
        b->sizes[i] = h4e_parse_begin(b->st[stream_ids[i]].seq, frame_types[i], frames[i], frame_bytes[i]);
    });
    const size_t jobs_bytes = align_up((size_t)n * sizeof(ReconJob), 256);
    size_t total = jobs_bytes;
    for (int i = 0; i < n; ++i)
    {
        if (b->sizes[i] == 0) return HVQM4_ERR_GEOMETRY;
        b->offs[i] = total;
        total += align_up(b->sizes[i], 128);
    }
    Arena &a = b->arena[b->cur];
    if (a.in_flight)
    {
        if (!cuda_ok(cudaEventSynchronize(a.consumed), "cudaEventSynchronize")) return HVQM4_ERR_CUDA;
        a.in_flight = false;
    }
    if (!arena_reserve(b, a, total)) return HVQM4_ERR_NOMEM;
    uint8_t *d_base = a.d;
    if (b->recording)
    {
        RecordedStep r;
        if (!cuda_ok(cudaMalloc((void **)&r.d, total), "cudaMalloc(recorded step)")) return HVQM4_ERR_NOMEM;
        r.n = n;
        b->recorded.push_back(r);
        d_base = r.d;
    }
    /* record-kernel CTA ranges (chunk counts are known since phase A) */
    ReconJob *jobs = reinterpret_cast<ReconJob *>(a.h);
    b->rec_prefix.resize((size_t)n + 1);
    b->rec_prefix[0] = 0;
    for (int i = 0; i < n; ++i)
    {
        const uint32_t chunks = h4e_last_chunks(b->st[stream_ids[i]].seq);
        jobs[i].rec_cta_begin = b->rec_prefix[i];
        jobs[i].n_chunks = chunks;
        jobs[i].pad[0] = jobs[i].pad[1] = 0;
        b->rec_prefix[i + 1] = b->rec_prefix[i] + hvqm4_rec_ctas(chunks);
    }
    if (b->recording) b->recorded.back().rec_prefix = b->rec_prefix;
    /* phase B: side words, motion vectors, maps -> pinned arena */
    std::atomic<uint32_t> err{0};
    std::atomic<uint64_t> inter{0};
    b->pool->parallel_for(n, [&](int i) {
        H4Seq *seq = b->st[stream_ids[i]].seq;
        err |= h4e_parse_finish(seq, a.h + b->offs[i]);
        inter += h4e_last_inter_mcbs(seq);
    });
    /* rotation (h4m:2087-2093, 2131-2137) and job descriptors */
    for (int i = 0; i < n; ++i)
    {
        StreamState &s = b->st[stream_ids[i]];
        const int t = frame_types[i];
        if (t != SYM_PIC_B) std::swap(s.past, s.future);
        jobs[i].blob = d_base + b->offs[i];
        jobs[i].present = b->surface(stream_ids[i], s.present);
        jobs[i].past = b->surface(stream_ids[i], s.past);
        /* HVQM4DecodePpic passes `present` as the future frame (h4m:2060); the host stage
           has already rejected type-2 macroblocks in P pictures, so it is never read */
        jobs[i].future = b->surface(stream_ids[i], t == SYM_PIC_P ? s.present : s.future);
        s.last = s.present;
        if (t != SYM_PIC_B) std::swap(s.present, s.future);
        else std::swap(s.present, s.spare);
    }
    auto t_host1 = std::chrono::steady_clock::now();
    b->stats[4] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t_host1 - t_host0).count();

    if (!cuda_ok(cudaMemcpyAsync(d_base, a.h, total, cudaMemcpyHostToDevice, b->s_copy), "cudaMemcpyAsync(H2D)")) return HVQM4_ERR_CUDA;
    cudaEventRecord(b->ev_h2d, b->s_copy);
    cudaStreamWaitEvent(b->s_comp, b->ev_h2d, 0);
    batch_wait_readbacks(b);
    int launched = 0;
    int rc = hvqm4_recon_launch(reinterpret_cast<const ReconJob *>(d_base), n, b->mcb_w, b->mcb_h, b->rec_prefix.data(), b->slab_ok ? b->d_surfaces : nullptr, b->band_rows, b->s_comp, &launched);
    g_launches += launched;
    b->stats[1] += (uint64_t)launched;
    if (rc != 0)
    {
        cuda_ok((cudaError_t)rc, "recon kernel launch");
        return HVQM4_ERR_CUDA;
    }
    cudaEventRecord(a.consumed, b->s_comp);
    cudaEventRecord(b->ev_kernel, b->s_comp);
    a.in_flight = true;
    b->cur = (b->cur + 1) % kArenas;

    const uint64_t mcbs = (uint64_t)b->mcb_w * b->mcb_h * n;
    b->stats[0] += n;
    b->stats[2] += total;
    b->stats[3] += (uint64_t)n * b->frame_bytes + inter.load() * 96 + (total - jobs_bytes);
    b->stats[5] += inter.load();
    b->stats[6] += mcbs;
    b->errors |= err.load();
    return (int)err.load();
}

H4_API int HVQM4BatchSync(HVQM4Batch *b)
{
    if (!b) return HVQM4_ERR_ARGUMENT;
    cudaSetDevice(b->device);
    uint32_t e = b->errors;
    b->errors = 0;
    bool ok = cuda_ok(cudaStreamSynchronize(b->s_comp), "sync compute") &
              cuda_ok(cudaStreamSynchronize(b->s_d2h), "sync d2h");
    for (cudaStream_t sp : b->s_parse) ok = cuda_ok(cudaStreamSynchronize(sp), "sync copy") && ok;
    for (auto &a : b->arena) a.in_flight = false;
    b->d2h_pending = false;
    if (ok && b->d_eerrors)
    {   /* stream errors raised by the GPU entropy stage since the last sync */
        uint32_t de = 0;
        if (cuda_ok(cudaMemcpy(&de, b->d_eerrors, sizeof de, cudaMemcpyDeviceToHost), "cudaMemcpy(errors)") && de)
        {
            e |= de;
            cudaMemset(b->d_eerrors, 0, sizeof de);
        }
    }
    if (!ok) e |= HVQM4_ERR_CUDA;
    return (int)e;
}

H4_API void *HVQM4BatchFramePtr(HVQM4Batch *b, int stream_id)
{
    if (!b || stream_id < 0 || stream_id >= b->n_streams || b->st[stream_id].last < 0) return nullptr;
    return b->surface(stream_id, b->st[stream_id].last);
}

H4_API int HVQM4BatchReadFrameAsync(HVQM4Batch *b, int stream_id, void *host_dst)
{
    void *src = HVQM4BatchFramePtr(b, stream_id);
    if (!src || !host_dst) return HVQM4_ERR_ARGUMENT;
    cudaSetDevice(b->device);
    if (!b->d2h_pending) cudaStreamWaitEvent(b->s_d2h, b->ev_kernel, 0);
    if (!copy_frame_to_host(host_dst, src, b->frame_bytes, b->s_d2h, &b->bounce, "cudaMemcpyAsync(D2H)")) return HVQM4_ERR_CUDA;
    cudaEventRecord(b->ev_d2h, b->s_d2h);
    b->d2h_pending = true;
    return HVQM4_OK;
}

H4_API int HVQM4BatchReadFramesAsync(HVQM4Batch *b, int n, const int32_t *stream_ids, void *host_base, size_t host_stride)
{
    if (!b || n <= 0 || !stream_ids || !host_base || host_stride < b->frame_bytes) return HVQM4_ERR_ARGUMENT;
    /* streams decoded in lock step sit at a constant pitch in the surface slab: one 2-D copy */
    bool regular = true;
    for (int i = 0; i < n && regular; ++i)
    {
        const int s = stream_ids[i];
        if (s < 0 || s >= b->n_streams || b->st[s].last < 0) return HVQM4_ERR_ARGUMENT;
        regular = s == stream_ids[0] + i && b->st[s].last == b->st[stream_ids[0]].last;
    }
    if (regular && n > 1)
    {
        cudaSetDevice(b->device);
        if (!b->d2h_pending) cudaStreamWaitEvent(b->s_d2h, b->ev_kernel, 0);
        timeline_mark(b, b->mark_step, 5, b->s_d2h);
        const cudaError_t e2 = cudaMemcpy2DAsync(host_base, host_stride, b->surface(stream_ids[0], b->st[stream_ids[0]].last), kSurfaces * b->surf_stride,
                                                 b->frame_bytes, (size_t)n, cudaMemcpyDeviceToHost, b->s_d2h);
        if (e2 == cudaErrorInvalidValue)
        {   /* a destination that straddles page-locked and pageable memory (copy_frame_to_host): frame by frame */
            cudaGetLastError();
            for (int i = 0; i < n; ++i)
                if (!copy_frame_to_host(static_cast<uint8_t *>(host_base) + (size_t)i * host_stride, b->surface(stream_ids[i], b->st[stream_ids[i]].last),
                                        b->frame_bytes, b->s_d2h, &b->bounce, "cudaMemcpyAsync(D2H)"))
                    return HVQM4_ERR_CUDA;
        }
        else if (!cuda_ok(e2, "cudaMemcpy2DAsync(D2H)"))
            return HVQM4_ERR_CUDA;
        cudaEventRecord(b->ev_d2h, b->s_d2h);
        timeline_mark(b, b->mark_step, 6, b->s_d2h);
        b->d2h_pending = true;
        return HVQM4_OK;
    }
    for (int i = 0; i < n; ++i)
    {
        int rc = HVQM4BatchReadFrameAsync(b, stream_ids[i], static_cast<uint8_t *>(host_base) + (size_t)i * host_stride);
        if (rc) return rc;
    }
    return HVQM4_OK;
}

H4_API int HVQM4BatchReadFrame(HVQM4Batch *b, int stream_id, void *host_dst)
{
    int rc = HVQM4BatchReadFrameAsync(b, stream_id, host_dst);
    if (rc) return rc;
    return cuda_ok(cudaStreamSynchronize(b->s_d2h), "sync d2h") ? HVQM4_OK : HVQM4_ERR_CUDA;
}

/* dumpRGB (h4m:895-926) for the last decoded picture of n streams: conversion kernel on the
   compute stream (behind the reconstruction it reads), then one pitched copy on the read-back stream */
H4_API int HVQM4BatchReadFramesRGBAsync(HVQM4Batch *b, int n, const int32_t *stream_ids, void *host_base, size_t host_stride)
{
    if (!b) return HVQM4_ERR_ARGUMENT;
    const size_t rgb_bytes = (size_t)b->width * b->height * 3;
    if (n <= 0 || n > b->n_streams || !stream_ids || !host_base || host_stride < rgb_bytes) return HVQM4_ERR_ARGUMENT;
    for (int i = 0; i < n; ++i)
        if (stream_ids[i] < 0 || stream_ids[i] >= b->n_streams || b->st[stream_ids[i]].last < 0) return HVQM4_ERR_ARGUMENT;
    cudaSetDevice(b->device);
    if (!b->d_rgb)
    {
        bool ok = cuda_ok(cudaMalloc((void **)&b->d_rgb, rgb_bytes * b->n_streams), "cudaMalloc(rgb staging)") &&
                  cuda_ok(cudaEventCreateWithFlags(&b->ev_rgb, cudaEventDisableTiming), "cudaEventCreate");
        for (int i = 0; i < 2 && ok; ++i)
            ok = cuda_ok(cudaHostAlloc((void **)&b->h_rgb_src[i], sizeof(void *) * b->n_streams, cudaHostAllocDefault), "cudaHostAlloc") &&
                 cuda_ok(cudaMalloc((void **)&b->d_rgb_src[i], sizeof(void *) * b->n_streams), "cudaMalloc") &&
                 cuda_ok(cudaEventCreateWithFlags(&b->ev_rgb_src[i], cudaEventDisableTiming), "cudaEventCreate");
        if (!ok) return HVQM4_ERR_NOMEM;
    }
    const int r = b->rgb_cur;
    b->rgb_cur ^= 1;
    cudaEventSynchronize(b->ev_rgb_src[r]);               /* the copy that last used this pointer ring is done */
    for (int i = 0; i < n; ++i) b->h_rgb_src[r][i] = b->surface(stream_ids[i], b->st[stream_ids[i]].last);
    /* the staging frames of the previous call may still be on their way out */
    cudaStreamWaitEvent(b->s_comp, b->ev_d2h, 0);
    b->d2h_pending = false;
    if (!cuda_ok(cudaMemcpyAsync((void *)b->d_rgb_src[r], (const void *)b->h_rgb_src[r], sizeof(void *) * n, cudaMemcpyHostToDevice, b->s_comp),
                 "cudaMemcpyAsync(frame pointers)"))
        return HVQM4_ERR_CUDA;
    cudaEventRecord(b->ev_rgb_src[r], b->s_comp);
    const int rc = hvqm4_rgb_launch(b->d_rgb_src[r], n, b->d_rgb, rgb_bytes, b->width, b->height, b->s_comp);
    if (rc != 0)
    {
        cuda_ok((cudaError_t)rc, "yuv2rgb kernel launch");
        return HVQM4_ERR_CUDA;
    }
    ++g_launches;
    b->stats[1] += 1;
    cudaEventRecord(b->ev_rgb, b->s_comp);
    cudaEventRecord(b->ev_kernel, b->s_comp);
    cudaStreamWaitEvent(b->s_d2h, b->ev_rgb, 0);
    const cudaError_t e2 = cudaMemcpy2DAsync(host_base, host_stride, b->d_rgb, rgb_bytes, rgb_bytes, (size_t)n, cudaMemcpyDeviceToHost, b->s_d2h);
    if (e2 == cudaErrorInvalidValue)
    {   /* destination straddling page-locked and pageable memory (copy_frame_to_host) */
        cudaGetLastError();
        for (int i = 0; i < n; ++i)
            if (!copy_frame_to_host(static_cast<uint8_t *>(host_base) + (size_t)i * host_stride, b->d_rgb + (size_t)i * rgb_bytes, rgb_bytes, b->s_d2h,
                                    &b->bounce, "cudaMemcpyAsync(D2H rgb)"))
                return HVQM4_ERR_CUDA;
    }
    else if (!cuda_ok(e2, "cudaMemcpy2DAsync(D2H rgb)"))
        return HVQM4_ERR_CUDA;
    cudaEventRecord(b->ev_d2h, b->s_d2h);
    b->d2h_pending = true;
    return HVQM4_OK;
}

H4_API int HVQM4BatchRecord(HVQM4Batch *b, int enable)
{
    if (!b) return HVQM4_ERR_ARGUMENT;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    if (enable)
    {
        for (auto &r : b->recorded) cudaFree(r.d);
        b->recorded.clear();
    }
    b->recording = enable != 0;
    return HVQM4_OK;
}

H4_API float HVQM4BatchReplay(HVQM4Batch *b, int repeats)
{
    if (!b || b->recorded.empty() || repeats <= 0) return -1.f;
    cudaSetDevice(b->device);
    if (!cuda_ok(cudaDeviceSynchronize(), "cudaDeviceSynchronize")) return -1.f;
    cudaEventRecord(b->ev_t0, b->s_comp);
    for (int r = 0; r < repeats; ++r)
        for (auto &st : b->recorded)
        {
            int launched = 0;
            int rc = hvqm4_recon_launch(reinterpret_cast<const ReconJob *>(st.d), st.n, b->mcb_w, b->mcb_h, st.rec_prefix.data(), b->slab_ok ? b->d_surfaces : nullptr, b->band_rows, b->s_comp, &launched);
            g_launches += launched;
            b->stats[1] += (uint64_t)launched;
            if (rc != 0)
            {
                cuda_ok((cudaError_t)rc, "recon kernel launch");
                return -1.f;
            }
        }
    cudaEventRecord(b->ev_t1, b->s_comp);
    if (!cuda_ok(cudaEventSynchronize(b->ev_t1), "cudaEventSynchronize")) return -1.f;
    float ms = 0;
    cudaEventElapsedTime(&ms, b->ev_t0, b->ev_t1);
    return ms;
}

H4_API void HVQM4BatchStats(HVQM4Batch *b, uint64_t out[8])
{
    for (int i = 0; i < 8; ++i) out[i] = b ? b->stats[i] : 0;
    out[7] = (uint64_t)hvqm4_recon_band_launches();   /* process-wide */
}

H4_API void *HVQM4HostAlloc(size_t bytes)
{
    void *p = nullptr;
    if (!have_device()) return nullptr;
    if (!cuda_ok(cudaHostAlloc(&p, bytes, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
    return p;
}

H4_API void HVQM4HostFree(void *p)
{
    if (p) cudaFreeHost(p);
}

H4_API int HVQM4GetLastCudaError(void) { return g_last_cuda_error.load(); }

/* ====================================================================== SDK-compatible layer */

namespace {

constexpr uint32_t kCompatMagic = 0x48344232u;   /* "H4B2" */
constexpr int kTwins = 4;

struct Twin
{
    const void *host = nullptr;
    uint8_t *dev = nullptr;
    uint64_t stamp = 0;
    uint64_t print = 0;       /* fingerprint of the host frame when the twin last equalled it (twin_print) */
    bool printed = false;
};

/* 64 words spread over a host frame, folded: cheap enough for every decode call, and an application that writes into
   a frame buffer behind the library's back (clears it on a seek, draws into it) changes it with near certainty, so
   the stale device twin is uploaded again instead of being used as a reference.  HVQM4InvalidateFrame stays for the
   writes this cannot see. */
static uint64_t twin_print(const void *host, size_t bytes)
{
    const uint8_t *p = static_cast<const uint8_t *>(host);
    const size_t words = bytes / 8, step = words / 64 ? words / 64 : 1;
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < words; i += step)
    {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        h = (h ^ w) * 0x100000001B3ull + (h >> 29);
    }
    return h;
}

struct Compat
{
    H4Seq *seq = nullptr;
    int width = 0, height = 0;
    size_t frame_bytes = 0, surf_bytes = 0;
    int mcb_w = 0, mcb_h = 0;
    int band_rows = 8;                              /* macroblock rows per record band of the stream (symbuf.h) */
    uint8_t *h_blob = nullptr, *d_blob = nullptr;   /* job descriptor (256 B) + blob */
    Bounce bounce;                                  /* copy_frame_to_host / copy_frame_to_device */
    uint8_t *d_rgb = nullptr;                       /* HVQM4ConvertRGB staging */
    size_t blob_cap = 0;
    Twin twin[kTwins];
    uint64_t clock = 0;
    cudaStream_t stream = nullptr;
    uint32_t errors = 0;
    uint32_t next_frame_bytes = 0;
    /* HVQM4_SDK_TRACE=1: seconds spent per phase of the synchronous calls, printed by HVQM4ReleaseBuffer */
    bool trace = false;
    double t_phase[4] = {0, 0, 0, 0};   /* parse_begin, parse_finish, resolve + enqueue, wait */
    unsigned n_calls = 0;
};

/* what lives in the caller's work buffer */
struct WorkHeader
{
    uint32_t magic;
    uint32_t version15;
    Compat *impl;
};

Compat *compat_of(SeqObj *so)
{
    if (!so || !so->state) return nullptr;
    WorkHeader *w = static_cast<WorkHeader *>(so->state);
    return w->magic == kCompatMagic ? w->impl : nullptr;
}

bool is_device_ptr(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

/* device copy of a caller frame: the pointer itself if it is device memory, otherwise a
   cached twin keyed by the host address (uploaded on first sight when it is a reference) */
uint8_t *resolve(Compat *c, void *p, bool is_reference)
{
    if (is_device_ptr(p)) return static_cast<uint8_t *>(p);
    Twin *slot = nullptr;
    for (auto &t : c->twin)
        if (t.host == p) slot = &t;
    bool fresh = false;
    if (!slot)
    {
        slot = &c->twin[0];
        for (auto &t : c->twin)
            if (t.stamp < slot->stamp) slot = &t;
        if (!slot->dev && !cuda_ok(cudaMalloc((void **)&slot->dev, c->surf_bytes), "cudaMalloc(frame twin)")) return nullptr;
        slot->host = p;
        fresh = true;
    }
    slot->stamp = ++c->clock;
    if (is_reference)
    {
        const uint64_t print = twin_print(p, c->frame_bytes);
        if (fresh || !slot->printed || slot->print != print)
        {   /* unknown, or the application has written into the frame since the library last saw it */
            if (!copy_frame_to_device(slot->dev, p, c->frame_bytes, c->stream, &c->bounce, "upload reference frame")) return nullptr;
            slot->print = print;
            slot->printed = true;
        }
    }
    else
        slot->printed = false;      /* about to be overwritten by this picture: fingerprinted after the read-back */
    return slot->dev;
}

void compat_decode(SeqObj *so, int type, const uint8_t *frame, void *present, void *past, void *future)
{
    Compat *c = compat_of(so);
    if (!c)
    {
        fprintf(stderr, "hvqm4_b200: decode called on a SeqObj without HVQM4SetBuffer\n");
        return;
    }
    if (!have_device() || !c->stream)
    {
        c->errors |= HVQM4_ERR_NO_DEVICE;
        return;
    }
    /* The SDK protocol has no picture length (the reference trusts the record, h4m:1061-1071).  Without
       HVQM4SetFrameBytes the readable length is bounded by what a picture of this geometry can possibly need -- every
       block at the format's maximum of side data -- so that a damaged section table is reported as truncated instead
       of sending the parser gigabytes away; applications that feed untrusted input call HVQM4SetFrameBytes. */
    const size_t len = c->next_frame_bytes ? c->next_frame_bytes : (size_t)64 * c->frame_bytes + 65536;
    c->next_frame_bytes = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t0 = now();
    const size_t need = h4e_parse_begin(c->seq, type, frame, len);
    const auto t1 = now();
    if (!need)
    {
        c->errors |= HVQM4_ERR_GEOMETRY;
        return;
    }
    const size_t total = 256 + need;
    if (total > c->blob_cap)
    {
        cudaStreamSynchronize(c->stream);
        if (c->h_blob) cudaFreeHost(c->h_blob);
        if (c->d_blob) cudaFree(c->d_blob);
        c->h_blob = c->d_blob = nullptr;
        c->blob_cap = 0;
        const size_t cap = align_up(total + total / 2, 1 << 16);
        if (!cuda_ok(cudaHostAlloc((void **)&c->h_blob, cap, cudaHostAllocDefault), "cudaHostAlloc(blob)") ||
            !cuda_ok(cudaMalloc((void **)&c->d_blob, cap), "cudaMalloc(blob)"))
        {
            c->errors |= HVQM4_ERR_NOMEM;
            return;
        }
        c->blob_cap = cap;
    }
    const auto t1b = now();
    c->errors |= h4e_parse_finish(c->seq, c->h_blob + 256);
    const auto t2 = now();
    ReconJob *job = reinterpret_cast<ReconJob *>(c->h_blob);
    uint8_t *d_present = resolve(c, present, false);
    uint8_t *d_past = past ? resolve(c, past, true) : d_present;
    uint8_t *d_future = (future && future != present) ? resolve(c, future, true) : d_present;
    if (!d_present || !d_past || !d_future)
    {
        c->errors |= HVQM4_ERR_CUDA;
        return;
    }
    job->blob = c->d_blob + 256;
    job->present = d_present;
    job->past = d_past;
    job->future = d_future;
    job->rec_cta_begin = 0;
    job->n_chunks = h4e_last_chunks(c->seq);
    job->pad[0] = job->pad[1] = 0;
    bool ok = cuda_ok(cudaMemcpyAsync(c->d_blob, c->h_blob, total, cudaMemcpyHostToDevice, c->stream), "upload symbols");
    if (ok)
    {
        int launched = 0;
        const uint32_t prefix[2] = {0, hvqm4_rec_ctas(job->n_chunks)};
        int rc = hvqm4_recon_launch(reinterpret_cast<const ReconJob *>(c->d_blob), 1, c->mcb_w, c->mcb_h, prefix, nullptr, c->band_rows, c->stream, &launched);
        g_launches += launched;
        ok = rc == 0 || cuda_ok((cudaError_t)rc, "recon kernel launch");
    }
    if (ok && !is_device_ptr(present))
        ok = copy_frame_to_host(present, d_present, c->frame_bytes, c->stream, &c->bounce, "download frame");
    const auto t3 = now();
    ok = ok && cuda_ok(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
    if (!ok) c->errors |= HVQM4_ERR_CUDA;
    if (ok && !is_device_ptr(present))
        for (auto &t : c->twin)
            if (t.host == present)
            {   /* the host frame and its twin are equal now */
                t.print = twin_print(present, c->frame_bytes);
                t.printed = true;
            }
    if (c->trace)
    {
        c->t_phase[0] += secs(t0, t1);
        c->t_phase[1] += secs(t1b, t2);
        c->t_phase[2] += secs(t2, t3);
        c->t_phase[3] += secs(t3, now());
        ++c->n_calls;
    }
}

}  // namespace

H4_API void HVQM4InitDecoder(void)
{
    /* The reference fills divTable/mcdivTable here (h4m:265-278); the kernels build them in
       shared memory per CTA, so only the device probe remains. */
    have_device();
}

H4_API void HVQM4InitSeqObj(SeqObj *seqobj, VideoInfo *videoinfo)
{
    seqobj->width = videoinfo->hres;
    seqobj->height = videoinfo->vres;
    seqobj->h_samp = videoinfo->h_samp;
    seqobj->v_samp = videoinfo->v_samp;
}

H4_API uint32_t HVQM4BuffSize(SeqObj *) { return 256; }

H4_API void HVQM4SetBuffer(SeqObj *seqobj, void *workbuff)
{
    WorkHeader *w = static_cast<WorkHeader *>(workbuff);
    seqobj->state = workbuff;
    w->magic = 0;
    w->impl = nullptr;
    w->version15 = 1;
    Compat *c = new Compat;
    c->band_rows = hvqm4_recon_band_rows();
    h4e_set_band_rows(c->band_rows);
    c->seq = h4e_seq_create(seqobj->width, seqobj->height, seqobj->h_samp, seqobj->v_samp, 1);
    if (!c->seq)
    {
        fprintf(stderr, "hvqm4_b200: unsupported geometry %ux%u (sampling %u x %u)\n", seqobj->width, seqobj->height,
                seqobj->h_samp, seqobj->v_samp);
        delete c;
        return;
    }
    c->width = seqobj->width;
    c->height = seqobj->height;
    int dims[6];
    h4e_seq_dims(c->seq, dims);
    c->mcb_w = dims[2];
    c->mcb_h = dims[3];
    c->frame_bytes = h4e_frame_bytes(c->seq);
    c->surf_bytes = align_up(c->frame_bytes + 64, 256);
    if (have_device() && !cuda_ok(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking), "cudaStreamCreate"))
        c->stream = nullptr;
    c->trace = getenv("HVQM4_SDK_TRACE") != nullptr;
    w->impl = c;
    w->magic = kCompatMagic;
}

H4_API void HVQM4ReleaseBuffer(SeqObj *seqobj)
{
    Compat *c = compat_of(seqobj);
    if (!c) return;
    if (c->trace && c->n_calls)
        fprintf(stderr, "hvqm4_b200: %u SDK calls, per call: parse_begin %.1f us, parse_finish %.1f us, enqueue %.1f us, wait %.1f us\n", c->n_calls,
                1e6 * c->t_phase[0] / c->n_calls, 1e6 * c->t_phase[1] / c->n_calls, 1e6 * c->t_phase[2] / c->n_calls, 1e6 * c->t_phase[3] / c->n_calls);
    if (c->stream)
    {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    for (auto &t : c->twin)
        if (t.dev) cudaFree(t.dev);
    if (c->h_blob) cudaFreeHost(c->h_blob);
    c->bounce.release();
    if (c->d_blob) cudaFree(c->d_blob);
    if (c->d_rgb) cudaFree(c->d_rgb);
    h4e_seq_destroy(c->seq);
    delete c;
    static_cast<WorkHeader *>(seqobj->state)->magic = 0;
    static_cast<WorkHeader *>(seqobj->state)->impl = nullptr;
}

H4_API int HVQM4SetVersion(SeqObj *seqobj, int version)
{
    Compat *c = compat_of(seqobj);
    if (!c || (version != 13 && version != 15)) return HVQM4_ERR_ARGUMENT;
    h4e_seq_set_version(c->seq, version == 15);
    static_cast<WorkHeader *>(seqobj->state)->version15 = version == 15;
    return HVQM4_OK;
}

H4_API void HVQM4SetFrameBytes(SeqObj *seqobj, uint32_t bytes)
{
    Compat *c = compat_of(seqobj);
    if (c) c->next_frame_bytes = bytes;
}

H4_API uint32_t HVQM4GetLastError(SeqObj *seqobj)
{
    Compat *c = compat_of(seqobj);
    if (!c) return HVQM4_ERR_ARGUMENT;
    uint32_t e = c->errors;
    c->errors = 0;
    return e;
}

H4_API void HVQM4InvalidateFrame(SeqObj *seqobj, void *host_frame)
{
    Compat *c = compat_of(seqobj);
    if (!c) return;
    for (auto &t : c->twin)
        if (t.host == host_frame)
        {
            t.host = nullptr;
            t.stamp = 0;
        }
}

/* dumpRGB (h4m:895-926) of one frame: `frame` is a planar picture of this SeqObj's geometry (host
   pointer: uploaded; device pointer: used in place); `rgb` receives width * height * 3 bytes of
   interleaved R, G, B in host memory. */
H4_API int HVQM4ConvertRGB(SeqObj *seqobj, const void *frame, void *rgb)
{
    Compat *c = compat_of(seqobj);
    if (!c || !frame || !rgb) return HVQM4_ERR_ARGUMENT;
    if (!have_device() || !c->stream) return HVQM4_ERR_NO_DEVICE;
    const size_t rgb_bytes = (size_t)c->width * c->height * 3;
    /* staging: [RGB frame | surface pointer | planar input].  A host frame is always uploaded:
       the cached device twins are keyed by address only, and nothing says this buffer still
       holds what the decoder wrote there. */
    const size_t ptr_off = align_up(rgb_bytes, 256), in_off = ptr_off + 256;
    if (!c->d_rgb && !cuda_ok(cudaMalloc((void **)&c->d_rgb, in_off + c->surf_bytes), "cudaMalloc(rgb)")) return HVQM4_ERR_NOMEM;
    const uint8_t *src = static_cast<const uint8_t *>(frame);
    bool ok = true;
    if (!is_device_ptr(frame))
    {
        src = c->d_rgb + in_off;
        ok = copy_frame_to_device(c->d_rgb + in_off, frame, c->frame_bytes, c->stream, &c->bounce, "upload frame");
    }
    const uint8_t **d_ptr = reinterpret_cast<const uint8_t **>(c->d_rgb + ptr_off);
    ok = ok && cuda_ok(cudaMemcpyAsync((void *)d_ptr, &src, sizeof src, cudaMemcpyHostToDevice, c->stream), "cudaMemcpyAsync");
    ok = ok && cuda_ok(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");   /* &src is a stack address */
    if (ok)
    {
        const int rc = hvqm4_rgb_launch(d_ptr, 1, c->d_rgb, rgb_bytes, c->width, c->height, c->stream);
        ok = rc == 0 || cuda_ok((cudaError_t)rc, "yuv2rgb kernel launch");
        if (rc == 0) ++g_launches;
    }
    ok = ok && copy_frame_to_host(rgb, c->d_rgb, rgb_bytes, c->stream, &c->bounce, "download rgb");
    ok = ok && cuda_ok(cudaStreamSynchronize(c->stream), "cudaStreamSynchronize");
    return ok ? HVQM4_OK : HVQM4_ERR_CUDA;
}

H4_API void HVQM4DecodeIpic(SeqObj *seqobj, uint8_t const *frame, void *present)
{
    compat_decode(seqobj, SYM_PIC_I, frame, present, nullptr, nullptr);
}

H4_API void HVQM4DecodePpic(SeqObj *seqobj, uint8_t const *frame, void *present, void *past)
{
    compat_decode(seqobj, SYM_PIC_P, frame, present, past, present);
}

H4_API void HVQM4DecodeBpic(SeqObj *seqobj, uint8_t const *frame, void *present, void *past, void *future)
{
    compat_decode(seqobj, SYM_PIC_B, frame, present, past, future);
}

H4_API long long HVQM4KernelLaunches(void) { return g_launches.load(); }

H4_API void HVQM4SetReconMode(int mode) { hvqm4_recon_set_mode(mode); }

extern "C" int hvqm4_sweep_errors(void);
extern "C" long long hvqm4_recon_sweep_launches(void);
extern "C" long long hvqm4_recon_row_launches(void);
extern "C" int hvqm4_row_errors(void);
H4_API int HVQM4RowErrors(void) { return hvqm4_row_errors(); }
H4_API long long HVQM4RowLaunches(void) { return hvqm4_recon_row_launches(); }
H4_API int HVQM4SweepErrors(void) { return hvqm4_sweep_errors(); }
H4_API long long HVQM4SweepLaunches(void) { return hvqm4_recon_sweep_launches(); }

/* ====================================================================== container walker */

static inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
static inline uint32_t be16(const uint8_t *p) { return (uint32_t)p[0] << 8 | p[1]; }

H4_API int HVQM4ParseFile(const uint8_t *data, size_t len, HVQM4FileInfo *info, HVQM4FrameRef *frames, int max_frames)
{
    if (!data || len < 0x44 || !info) return -1;
    if (!memcmp(data, "HVQM4 1.3\0\0\0\0\0\0\0", 16)) info->version = 13;
    else if (!memcmp(data, "HVQM4 1.5\0\0\0\0\0\0\0", 16)) info->version = 15;
    else return -2;
    if (be32(data + 0x10) != 0x44) return -3;                       /* h4m:2213 */
    info->n_gops = (int32_t)be32(data + 0x18);
    info->n_video_frames = (int32_t)be32(data + 0x1C);
    info->usec_per_frame = (int32_t)be32(data + 0x24);
    info->width = (int32_t)be16(data + 0x34);
    info->height = (int32_t)be16(data + 0x36);
    info->h_samp = data[0x38];
    info->v_samp = data[0x39];
    info->n_audio_frames = (int32_t)be32(data + 0x20);
    info->audio_channels = data[0x3C];
    info->audio_bits = data[0x3D];
    info->audio_format = data[0x3E];
    info->audio_sample_rate = (int32_t)be32(data + 0x40);
    if (info->n_gops == 0) return -4;                               /* h4m:2215-2219 */
    size_t pos = 0x44;
    int count = 0;
    for (int g = 0; g < info->n_gops; ++g)
    {
        if (pos + 20 > len) return -5;
        uint32_t nv = be32(data + pos + 8), na = be32(data + pos + 12);
        if (be32(data + pos + 16) != 0x01000000) return -6;         /* h4m:2436 */
        pos += 20;
        while (nv || na)
        {
            if (pos + 8 > len) return -5;
            uint32_t id1 = be16(data + pos), id2 = be16(data + pos + 2), size = be32(data + pos + 4);
            pos += 8;
            if (pos + size > len) return -5;
            if (id1 == 1)
            {
                if (!nv || size < 4) return -7;
                if (id2 != 0x10 && id2 != 0x20 && id2 != 0x30) return -8;   /* h4m:2113-2115 */
                if (frames && count < max_frames)
                {
                    frames[count].offset = (uint32_t)(pos + 4);
                    frames[count].bytes = size - 4;
                    frames[count].frame_type = (uint16_t)id2;
                    frames[count].gop = (uint16_t)g;
                    frames[count].disp_id = be32(data + pos);
                }
                ++count;
                --nv;
            }
            else if (id1 == 0)
            {
                if (!na) return -7;
                --na;
            }
            else
                return -9;                                          /* h4m:2509-2512 */
            pos += size;
        }
    }
    return count;
}

H4_API int HVQM4ParseFileAudio(const uint8_t *data, size_t len, HVQM4AudioRef *refs, int max_refs)
{
    HVQM4FileInfo info;
    const int nv = HVQM4ParseFile(data, len, &info, nullptr, 0);     /* validates the whole walk */
    if (nv < 0) return nv;
    size_t pos = 0x44;
    int count = 0;
    for (int g = 0; g < info.n_gops; ++g)
    {
        uint32_t left = be32(data + pos + 8) + be32(data + pos + 12);
        pos += 20;
        bool first = true;
        while (left--)
        {
            const uint32_t id1 = be16(data + pos), size = be32(data + pos + 4);
            pos += 8;
            if (id1 == 0)
            {
                if (size < 4) return -7;
                if (refs && count < max_refs)
                {
                    refs[count].offset = (uint32_t)pos;
                    refs[count].bytes = size;
                    refs[count].gop = (uint16_t)g;
                    refs[count].first = first ? 1 : 0;
                    refs[count].samples = be32(data + pos);
                }
                first = false;
                ++count;
            }
            pos += size;
        }
    }
    return count;
}
