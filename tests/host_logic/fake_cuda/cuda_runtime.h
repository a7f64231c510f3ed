/*
 * A CUDA runtime SIMULATOR for CPU tests of the library's host layer (tests/host_logic): the subset of the runtime
 * API that jm_nv_dec.cu / jmnv_enc.cu / jmc_runtime.cu use, with stream semantics made as unhelpful as the real thing
 * is allowed to be:
 *   - work enqueued on a stream runs LATER -- when somebody synchronises with it, or at random moments (seeded) --
 *     never at enqueue time (laziness 2), so host code that reads a result it has not waited for sees stale bytes;
 *   - cudaStreamWaitEvent / cudaEventRecord order streams exactly as CUDA defines (an event waits for the work
 *     captured at its most recent record), and cudaEventQuery says cudaErrorNotReady until that work has run;
 *   - copies from / to pageable host memory behave as CUDA's do (source staged at the call, destination written
 *     before the call returns); pinned and registered memory is accessed when the copy executes;
 *   - "device" memory is host memory filled with garbage at allocation, tracked so that frees with work still
 *     pending on the range, double frees, leaks, and copies that run outside an allocation are reported.
 * Test infrastructure only: nothing under jmcodec_b200/ includes this file (the product builds against the real
 * <cuda_runtime.h>; the test harness puts this directory first on the include path).
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

typedef enum cudaError {
    cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorInsufficientDriver = 35,
    cudaErrorNoDevice = 100, cudaErrorInvalidDevice = 101, cudaErrorInvalidResourceHandle = 400, cudaErrorNotReady = 600,
    cudaErrorHostMemoryAlreadyRegistered = 712, cudaErrorHostMemoryNotRegistered = 713
} cudaError_t;

struct fake_stream;
struct fake_event;
typedef fake_stream *cudaStream_t;
typedef fake_event *cudaEvent_t;

enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { enum cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

#define cudaStreamNonBlocking 0x01u
#define cudaEventDisableTiming 0x02u
#define cudaHostAllocDefault 0x00u
#define cudaHostAllocWriteCombined 0x04u
#define cudaHostRegisterDefault 0x00u

cudaError_t cudaGetLastError(void);
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaDeviceGetAttribute(int *v, enum cudaDeviceAttr a, int device);

cudaError_t cudaMalloc(void **p, size_t bytes);
cudaError_t cudaMallocPitch(void **p, size_t *pitch, size_t width, size_t height);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t bytes, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaHostRegister(void *p, size_t bytes, unsigned flags);
cudaError_t cudaHostUnregister(void *p);
cudaError_t cudaPointerGetAttributes(struct cudaPointerAttributes *a, const void *p);

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);

cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaEventQuery(cudaEvent_t e);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);

cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                              enum cudaMemcpyKind kind, cudaStream_t s);
cudaError_t cudaMemsetAsync(void *p, int byte, size_t bytes, cudaStream_t s);
cudaError_t cudaMemset(void *p, int byte, size_t bytes);
cudaError_t cudaMemcpy2D(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, enum cudaMemcpyKind kind);
cudaError_t cudaDeviceSynchronize(void);

#ifdef __cplusplus
/* the typed overloads of the real header */
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)(void *)p, bytes); }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t bytes, unsigned flags) { return cudaHostAlloc((void **)(void *)p, bytes, flags); }
template <class T> static inline cudaError_t cudaMallocPitch(T **p, size_t *pitch, size_t w, size_t h) { return cudaMallocPitch((void **)(void *)p, pitch, w, h); }
#endif

/* ---- simulator controls (tests only) -------------------------------------------------------------------------- */
#ifdef __cplusplus
#include <functional>
#include <string>
#include <vector>

/* laziness 0: work runs when it is enqueued (a GPU that is always ahead of the host);
 *          1: every runtime call lets a random amount of queued work run first (seeded);
 *          2: work runs only when the host synchronises with it (a GPU that is always behind). */
void fake_cuda_reset(unsigned seed, int laziness, int n_devices);
/* enqueue host code as a "kernel": reads / writes are declared so that frees under pending work are caught */
void fake_cuda_enqueue(cudaStream_t s, std::function<void()> fn, const char *what);
/* problems the simulator saw (frees under pending work, copies outside allocations, ...); empties the list */
std::vector<std::string> fake_cuda_take_errors();
/* live allocations by kind: device, pinned, registered ranges, streams, events */
struct fake_cuda_counts { size_t device, pinned, registered, streams, events, pending_ops; };
fake_cuda_counts fake_cuda_live();
/* is [p, p+n) inside one live device allocation? (for "kernels" to check their accesses) */
bool fake_cuda_is_device_range(const void *p, size_t n);
/* fail `count` allocations of the given kind (0 device, 1 pinned, 2 event) starting with the k-th next one: error-path
 * tests; k < 0 disables */
void fake_cuda_fail_alloc(int kind, int k, int count = 1);
/* make `count` ENQUEUE-type calls (async copies / memsets, event records, stream waits, launches) fail with
 * cudaErrorInvalidValue, starting with the k-th next one; nothing is enqueued by a failing call.  k < 0 disables */
void fake_cuda_fail_call(int k, int count = 1);
/* for stand-in launch code: does the armed call failure hit this launch? */
bool fake_cuda_launch_should_fail();
/* record a problem from test code that runs as a "kernel" */
void fake_cuda_complain(const char *msg);
#endif
