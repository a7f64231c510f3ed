/* The CUDA runtime simulator declared in cuda_runtime.h (this directory).  Test infrastructure. */
#include "cuda_runtime.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <random>

namespace {

enum { OP_WORK, OP_RECORD, OP_WAIT };

struct range { uintptr_t lo, hi; };

struct op {
    int kind;
    std::function<void()> fn;
    fake_event *ev;
    uint64_t gen;
    std::vector<range> touches;            /* memory the op reads or writes when it runs */
    const char *what;
};

} /* namespace */

struct fake_stream {
    std::deque<op> q;
    bool alive = true;
    int device = 0;
};

struct fake_event {
    uint64_t recorded = 0, completed = 0;
    bool alive = true;
};

namespace {

struct alloc { size_t bytes; int kind; };  /* kind 0 device, 1 pinned */

struct sim {
    std::recursive_mutex m;
    std::mt19937 rng{1};
    int laziness = 1, n_devices = 1;
    std::map<uintptr_t, alloc> allocs;                     /* by base address */
    std::map<uintptr_t, size_t> registered;                /* cudaHostRegister ranges */
    std::vector<std::unique_ptr<fake_stream>> streams;
    std::vector<std::unique_ptr<fake_event>> events;
    std::vector<std::string> errors;
    cudaError_t last = cudaSuccess;
    int fail_kind = -1, fail_in = -1, fail_count = 0;
    int call_fail_in = -1, call_fail_count = 0;
    int depth = 0;
};

sim &S()
{
    static sim s;
    /* programs that never call fake_cuda_reset are configured from the environment (once; thread-safe static init) */
    static const bool configured = [] {
        if (const char *e = getenv("FAKE_CUDA_LAZINESS")) s.laziness = atoi(e);
        if (const char *e = getenv("FAKE_CUDA_SEED")) s.rng.seed((unsigned)atoi(e));
        if (const char *e = getenv("FAKE_CUDA_DEVICES")) s.n_devices = atoi(e);
        return true;
    }();
    (void)configured;
    return s;
}

thread_local int t_device = 0;

cudaError_t fail(cudaError_t e) { S().last = e; return e; }

void complain(const std::string &s)
{
    S().errors.push_back(s);
}

bool call_should_fail()
{
    sim &s = S();
    if (s.call_fail_in < 0) return false;
    if (s.call_fail_in > 0) { s.call_fail_in--; return false; }
    if (s.call_fail_count > 0 && --s.call_fail_count == 0) s.call_fail_in = -1;
    return true;
}

bool should_fail(int kind)
{
    sim &s = S();
    if (s.fail_kind != kind) return false;
    if (s.fail_in > 0) { s.fail_in--; return false; }
    if (s.fail_count > 0) { if (--s.fail_count == 0) s.fail_kind = -1; return true; }
    s.fail_kind = -1;
    return false;
}

const alloc *find_alloc(const void *p, size_t n, uintptr_t *base = nullptr)
{
    sim &s = S();
    const uintptr_t a = (uintptr_t)p;
    auto it = s.allocs.upper_bound(a);
    if (it == s.allocs.begin()) return nullptr;
    --it;
    if (a >= it->first && a + n <= it->first + it->second.bytes) { if (base) *base = it->first; return &it->second; }
    return nullptr;
}

bool in_registered(const void *p, size_t n)
{
    sim &s = S();
    const uintptr_t a = (uintptr_t)p;
    auto it = s.registered.upper_bound(a);
    if (it == s.registered.begin()) return false;
    --it;
    return a >= it->first && a + n <= it->first + it->second;
}

/* 0 pageable, 1 pinned (allocated or registered), 2 device, -1 straddles */
int classify(const void *p, size_t n)
{
    if (n == 0) n = 1;
    const alloc *a = find_alloc(p, n);
    if (a) return a->kind == 0 ? 2 : 1;
    if (in_registered(p, n)) return 1;
    /* partly inside something we know? */
    const alloc *a0 = find_alloc(p, 1), *a1 = find_alloc((const uint8_t *)p + n - 1, 1);
    if (a0 || a1 || in_registered(p, 1) || in_registered((const uint8_t *)p + n - 1, 1)) return -1;
    return 0;
}

bool step(fake_stream *st);

fake_stream *stream_with_record(fake_event *e, uint64_t gen)
{
    for (auto &st : S().streams)
        for (auto &o : st->q)
            if (o.kind == OP_RECORD && o.ev == e && o.gen >= gen) return st.get();
    return nullptr;
}

void drive_event(fake_event *e, uint64_t gen);

void unblock_head(fake_stream *st)
{
    op &o = st->q.front();
    if (o.kind == OP_WAIT && o.ev->completed < o.gen) drive_event(o.ev, o.gen);
}

void drive_event(fake_event *e, uint64_t gen)
{
    sim &s = S();
    if (++s.depth > 64) { complain("deadlock: streams wait for each other in a cycle"); e->completed = gen; --s.depth; return; }
    while (e->completed < gen) {
        fake_stream *st = stream_with_record(e, gen);
        if (!st) { complain("an event is waited for whose record was never enqueued"); e->completed = gen; break; }
        if (!step(st)) unblock_head(st);
    }
    --s.depth;
}

void drive_stream(fake_stream *st)
{
    while (!st->q.empty())
        if (!step(st)) unblock_head(st);
}

void drive_all()
{
    for (auto &st : S().streams) drive_stream(st.get());
}

bool step(fake_stream *st)
{
    if (st->q.empty()) return false;
    op &o = st->q.front();
    if (o.kind == OP_WAIT) {
        if (o.ev->completed < o.gen) return false;
        st->q.pop_front();
        return true;
    }
    if (o.kind == OP_RECORD) {
        if (o.ev->completed < o.gen) o.ev->completed = o.gen;
        st->q.pop_front();
        return true;
    }
    std::function<void()> fn = std::move(o.fn);
    st->q.pop_front();
    fn();
    return true;
}

/* called at the top of every runtime call */
void tick()
{
    sim &s = S();
    if (s.laziness == 0) { drive_all(); return; }
    if (s.laziness == 1 && !s.streams.empty()) {
        const int n = (int)(s.rng() % 5);
        for (int i = 0; i < n; i++) step(s.streams[s.rng() % s.streams.size()].get());
    }
}

void enqueue(fake_stream *st, op o)
{
    st->q.push_back(std::move(o));
    if (S().laziness == 0) drive_all();
}

/* does pending work touch [lo, hi)? */
bool pending_touches(uintptr_t lo, uintptr_t hi, const char **what)
{
    for (auto &st : S().streams)
        for (auto &o : st->q)
            for (auto &r : o.touches)
                if (r.lo < hi && lo < r.hi) { *what = o.what; return true; }
    return false;
}

bool valid_stream(cudaStream_t st)
{
    for (auto &p : S().streams) if (p.get() == st) return st->alive;
    return false;
}
bool valid_event(cudaEvent_t e)
{
    for (auto &p : S().events) if (p.get() == e) return e->alive;
    return false;
}

void check_device_side(const void *p, size_t n, const char *what)
{
    if (n && classify(p, n) != 2) complain(std::string(what) + ": device side of a copy is not inside a live device allocation");
}

#define LOCK std::lock_guard<std::recursive_mutex> lock_(S().m)

} /* namespace */

/* ---- controls --------------------------------------------------------------------------------------------------- */
void fake_cuda_reset(unsigned seed, int laziness, int n_devices)
{
    LOCK;
    sim &s = S();
    drive_all();
    s.rng.seed(seed);
    s.laziness = laziness;
    s.n_devices = n_devices;
    s.last = cudaSuccess;
    s.fail_kind = s.fail_in = -1;
    s.fail_count = 0;
    s.call_fail_in = -1;
    s.call_fail_count = 0;
    /* dead streams / events are only ever forgotten here */
    s.streams.erase(std::remove_if(s.streams.begin(), s.streams.end(), [](const std::unique_ptr<fake_stream> &p) { return !p->alive; }), s.streams.end());
    s.events.erase(std::remove_if(s.events.begin(), s.events.end(), [](const std::unique_ptr<fake_event> &p) { return !p->alive; }), s.events.end());
    t_device = 0;
}

void fake_cuda_enqueue(cudaStream_t st, std::function<void()> fn, const char *what)
{
    LOCK;
    tick();
    if (!valid_stream(st)) { complain("launch on a dead stream"); return; }
    op o{ OP_WORK, std::move(fn), nullptr, 0, {}, what };
    enqueue(st, std::move(o));
}

std::vector<std::string> fake_cuda_take_errors()
{
    LOCK;
    std::vector<std::string> e;
    e.swap(S().errors);
    return e;
}

fake_cuda_counts fake_cuda_live()
{
    LOCK;
    fake_cuda_counts c{};
    for (auto &a : S().allocs) (a.second.kind == 0 ? c.device : c.pinned)++;
    c.registered = S().registered.size();
    for (auto &p : S().streams) { if (p->alive) c.streams++; c.pending_ops += p->q.size(); }
    for (auto &p : S().events) if (p->alive) c.events++;
    return c;
}

bool fake_cuda_is_device_range(const void *p, size_t n)
{
    LOCK;
    return classify(p, n) == 2;
}

void fake_cuda_fail_call(int k, int count)
{
    LOCK;
    S().call_fail_in = k;
    S().call_fail_count = count;
}

bool fake_cuda_launch_should_fail()
{
    LOCK;
    return call_should_fail();
}

void fake_cuda_complain(const char *msg)
{
    LOCK;
    complain(msg);
}

void fake_cuda_fail_alloc(int kind, int k, int count)
{
    LOCK;
    S().fail_kind = k < 0 ? -1 : kind;
    S().fail_in = k;
    S().fail_count = count;
}

/* ---- runtime ---------------------------------------------------------------------------------------------------- */
cudaError_t cudaGetLastError(void)
{
    LOCK;
    cudaError_t e = S().last;
    S().last = cudaSuccess;
    return e;
}

const char *cudaGetErrorString(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument";
    case cudaErrorMemoryAllocation: return "out of memory";
    case cudaErrorNotReady: return "device not ready";
    case cudaErrorNoDevice: return "no CUDA-capable device is detected";
    case cudaErrorInvalidDevice: return "invalid device ordinal";
    default: return "error";
    }
}

cudaError_t cudaGetDeviceCount(int *n)
{
    LOCK;
    if (S().n_devices <= 0) { *n = 0; return fail(cudaErrorNoDevice); }
    *n = S().n_devices;
    return cudaSuccess;
}

cudaError_t cudaGetDevice(int *d) { *d = t_device; return cudaSuccess; }

cudaError_t cudaSetDevice(int d)
{
    LOCK;
    if (d < 0 || d >= S().n_devices) return fail(cudaErrorInvalidDevice);
    t_device = d;
    return cudaSuccess;
}

cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr a, int device)
{
    LOCK;
    if (device < 0 || device >= S().n_devices) return fail(cudaErrorInvalidDevice);
    *v = a == cudaDevAttrMultiProcessorCount ? 148 : 0;
    return cudaSuccess;
}

static cudaError_t do_alloc(void **p, size_t bytes, int kind)
{
    LOCK;
    tick();
    if (should_fail(kind)) { *p = nullptr; return fail(cudaErrorMemoryAllocation); }
    void *q = nullptr;
    if (posix_memalign(&q, 256, bytes ? bytes : 1) != 0) { *p = nullptr; return fail(cudaErrorMemoryAllocation); }
    memset(q, kind == 0 ? 0xDD : 0xD1, bytes ? bytes : 1);             /* never zero: nobody may rely on fresh memory */
    S().allocs[(uintptr_t)q] = alloc{ bytes ? bytes : 1, kind };
    *p = q;
    return cudaSuccess;
}

cudaError_t cudaMalloc(void **p, size_t bytes) { return do_alloc(p, bytes, 0); }

cudaError_t cudaMallocPitch(void **p, size_t *pitch, size_t width, size_t height)
{
    *pitch = (width + 511) & ~(size_t)511;
    return do_alloc(p, *pitch * height, 0);
}

cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned) { return do_alloc(p, bytes, 1); }

static cudaError_t do_free(void *p, int kind, bool syncs)
{
    LOCK;
    tick();
    if (!p) return cudaSuccess;
    auto it = S().allocs.find((uintptr_t)p);
    if (it == S().allocs.end() || it->second.kind != kind) { complain(kind == 0 ? "cudaFree of something that is not a live device allocation" : "cudaFreeHost of something that is not a live pinned allocation"); return fail(cudaErrorInvalidValue); }
    if (syncs) drive_all();                                            /* cudaFree waits for the device */
    const char *what = "";
    if (pending_touches(it->first, it->first + it->second.bytes, &what))
        complain(std::string("memory freed while queued work still uses it: ") + what);
    memset(p, 0xEE, it->second.bytes);
    free(p);
    S().allocs.erase(it);
    return cudaSuccess;
}

cudaError_t cudaFree(void *p) { return do_free(p, 0, true); }
/* the documentation does not promise that cudaFreeHost waits for anything: the simulator does not */
cudaError_t cudaFreeHost(void *p) { return do_free(p, 1, false); }

cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned)
{
    LOCK;
    tick();
    const uintptr_t a = (uintptr_t)p;
    if (!p || !bytes) return fail(cudaErrorInvalidValue);
    for (auto &r : S().registered)
        if (r.first < a + bytes && a < r.first + r.second) return fail(cudaErrorHostMemoryAlreadyRegistered);
    if (find_alloc(p, 1)) return fail(cudaErrorHostMemoryAlreadyRegistered);
    S().registered[a] = bytes;
    return cudaSuccess;
}

cudaError_t cudaHostUnregister(void *p)
{
    LOCK;
    tick();
    auto it = S().registered.find((uintptr_t)p);
    if (it == S().registered.end()) return fail(cudaErrorHostMemoryNotRegistered);
    const char *what = "";
    if (pending_touches(it->first, it->first + it->second, &what))
        complain(std::string("host range unregistered while queued work still uses it: ") + what);
    S().registered.erase(it);
    return cudaSuccess;
}

cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes *a, const void *p)
{
    LOCK;
    memset(a, 0, sizeof(*a));
    const int k = classify(p, 1);
    a->type = k == 2 ? cudaMemoryTypeDevice : (k == 1 ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered);
    a->device = t_device;
    return cudaSuccess;
}

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned)
{
    LOCK;
    S().streams.emplace_back(new fake_stream());
    S().streams.back()->device = t_device;
    *st = S().streams.back().get();
    return cudaSuccess;
}

cudaError_t cudaStreamDestroy(cudaStream_t st)
{
    LOCK;
    tick();
    if (!valid_stream(st)) { complain("cudaStreamDestroy of a dead stream"); return fail(cudaErrorInvalidResourceHandle); }
    drive_stream(st);                                                  /* resources are released once the work has completed */
    st->alive = false;
    return cudaSuccess;
}

cudaError_t cudaStreamSynchronize(cudaStream_t st)
{
    LOCK;
    tick();
    if (!valid_stream(st)) { complain("cudaStreamSynchronize on a dead stream"); return fail(cudaErrorInvalidResourceHandle); }
    drive_stream(st);
    return cudaSuccess;
}

cudaError_t cudaStreamWaitEvent(cudaStream_t st, cudaEvent_t e, unsigned)
{
    LOCK;
    tick();
    if (!valid_stream(st) || !valid_event(e)) { complain("cudaStreamWaitEvent with a dead stream or event"); return fail(cudaErrorInvalidResourceHandle); }
    if (call_should_fail()) return fail(cudaErrorInvalidValue);
    if (e->recorded == 0 || e->completed >= e->recorded) return cudaSuccess;      /* nothing to wait for */
    op o{ OP_WAIT, nullptr, e, e->recorded, {}, "wait" };
    enqueue(st, std::move(o));
    return cudaSuccess;
}

cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned)
{
    LOCK;
    if (should_fail(2)) { *e = nullptr; return fail(cudaErrorMemoryAllocation); }
    S().events.emplace_back(new fake_event());
    *e = S().events.back().get();
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { return cudaEventCreateWithFlags(e, 0); }

cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    LOCK;
    if (!valid_event(e)) { complain("cudaEventDestroy of a dead event"); return fail(cudaErrorInvalidResourceHandle); }
    e->alive = false;                                                  /* queued records / waits keep working, as in CUDA */
    return cudaSuccess;
}

cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st)
{
    LOCK;
    tick();
    if (!valid_stream(st) || !valid_event(e)) { complain("cudaEventRecord with a dead stream or event"); return fail(cudaErrorInvalidResourceHandle); }
    if (call_should_fail()) return fail(cudaErrorInvalidValue);
    e->recorded++;
    op o{ OP_RECORD, nullptr, e, e->recorded, {}, "record" };
    enqueue(st, std::move(o));
    return cudaSuccess;
}

cudaError_t cudaEventQuery(cudaEvent_t e)
{
    LOCK;
    tick();
    if (!valid_event(e)) { complain("cudaEventQuery of a dead event"); return fail(cudaErrorInvalidResourceHandle); }
    if (e->completed >= e->recorded) return cudaSuccess;
    return fail(cudaErrorNotReady);
}

cudaError_t cudaEventSynchronize(cudaEvent_t e)
{
    LOCK;
    tick();
    if (!valid_event(e)) { complain("cudaEventSynchronize of a dead event"); return fail(cudaErrorInvalidResourceHandle); }
    drive_event(e, e->recorded);
    return cudaSuccess;
}

cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    LOCK;
    if (!valid_event(a) || !valid_event(b)) return fail(cudaErrorInvalidResourceHandle);
    if (a->completed < a->recorded || b->completed < b->recorded) return fail(cudaErrorNotReady);
    *ms = 1.0f;
    return cudaSuccess;
}

static cudaError_t copy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                          enum cudaMemcpyKind kind, cudaStream_t st, const char *what)
{
    LOCK;
    tick();
    if (!valid_stream(st)) { complain(std::string(what) + " on a dead stream"); return fail(cudaErrorInvalidResourceHandle); }
    if (width == 0 || height == 0) return cudaSuccess;
    if (call_should_fail()) return fail(cudaErrorInvalidValue);
    if (dpitch < width || spitch < width) return fail(cudaErrorInvalidValue);
    const size_t dspan = (height - 1) * dpitch + width, sspan = (height - 1) * spitch + width;
    const bool h2d = kind == cudaMemcpyHostToDevice, d2h = kind == cudaMemcpyDeviceToHost;
    if (h2d || kind == cudaMemcpyDeviceToDevice) check_device_side(dst, dspan, what);
    if (d2h || kind == cudaMemcpyDeviceToDevice) check_device_side(src, sspan, what);
    const void *hostp = h2d ? src : (d2h ? dst : nullptr);
    const size_t hspan = h2d ? sspan : dspan;
    int hk = hostp ? classify(hostp, hspan) : 1;
    if (hk == 2) { complain(std::string(what) + ": host side of a copy is device memory"); return fail(cudaErrorInvalidValue); }
    if (hk == -1) return fail(cudaErrorInvalidValue);                  /* partly registered: CUDA refuses the copy */
    op o{ OP_WORK, nullptr, nullptr, 0, {}, what };
    if (h2d && hk == 0) {
        /* pageable source: staged at the call */
        auto stage = std::make_shared<std::vector<uint8_t>>(width * height);
        for (size_t r = 0; r < height; r++) memcpy(stage->data() + r * width, (const uint8_t *)src + r * spitch, width);
        o.fn = [=] { for (size_t r = 0; r < height; r++) memcpy((uint8_t *)dst + r * dpitch, stage->data() + r * width, width); };
        o.touches.push_back({ (uintptr_t)dst, (uintptr_t)dst + dspan });
        enqueue(st, std::move(o));
        return cudaSuccess;
    }
    o.fn = [=] { for (size_t r = 0; r < height; r++) memcpy((uint8_t *)dst + r * dpitch, (const uint8_t *)src + r * spitch, width); };
    o.touches.push_back({ (uintptr_t)dst, (uintptr_t)dst + dspan });
    o.touches.push_back({ (uintptr_t)src, (uintptr_t)src + sspan });
    enqueue(st, std::move(o));
    if (d2h && hk == 0) drive_stream(st);                              /* pageable destination: written before the call returns */
    return cudaSuccess;
}

cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              enum cudaMemcpyKind kind, cudaStream_t st)
{
    return copy2d(dst, dpitch, src, spitch, width, height, kind, st, "cudaMemcpy2DAsync");
}

cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t st)
{
    return copy2d(dst, bytes, src, bytes, bytes, 1, kind, st, "cudaMemcpyAsync");
}

cudaError_t cudaMemsetAsync(void *p, int byte, size_t bytes, cudaStream_t st)
{
    LOCK;
    tick();
    if (!valid_stream(st)) return fail(cudaErrorInvalidResourceHandle);
    if (call_should_fail()) return fail(cudaErrorInvalidValue);
    check_device_side(p, bytes, "cudaMemsetAsync");
    op o{ OP_WORK, [=] { memset(p, byte, bytes); }, nullptr, 0, { { (uintptr_t)p, (uintptr_t)p + bytes } }, "cudaMemsetAsync" };
    enqueue(st, std::move(o));
    return cudaSuccess;
}

/* The legacy default stream does not order with the (non-blocking) streams the library creates: these two act at once. */
cudaError_t cudaMemset(void *p, int byte, size_t bytes)
{
    LOCK;
    tick();
    check_device_side(p, bytes, "cudaMemset");
    memset(p, byte, bytes);
    return cudaSuccess;
}

cudaError_t cudaMemcpy2D(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, enum cudaMemcpyKind kind)
{
    LOCK;
    tick();
    if (width == 0 || height == 0) return cudaSuccess;
    if (dpitch < width || spitch < width) return fail(cudaErrorInvalidValue);
    if (kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice) check_device_side(dst, (height - 1) * dpitch + width, "cudaMemcpy2D");
    if (kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice) check_device_side(src, (height - 1) * spitch + width, "cudaMemcpy2D");
    for (size_t r = 0; r < height; r++) memcpy((uint8_t *)dst + r * dpitch, (const uint8_t *)src + r * spitch, width);
    return cudaSuccess;
}

cudaError_t cudaDeviceSynchronize(void)
{
    LOCK;
    drive_all();
    return cudaSuccess;
}
