// TEST INFRASTRUCTURE - a stand-in for <cuda_runtime.h> that lets g++ compile the plain-CUDA-C parts of dreamer4_b200/csrc for
// the host and run them under a thread-per-CUDA-thread simulator, so that kernel logic and the engine's orchestration written
// without a GPU at hand can be checked on the CPU (tests/test_kernels_cusim_cpu.py).  It models exactly what the simulated code
// uses: 2-D grids of 1-D blocks, static and dynamic shared memory, __syncthreads / __syncwarp, warp shuffles, float4, __ldg,
// the usual math intrinsics, and the handful of runtime calls the engine makes ("device" memory is host memory, streams are
// synchronous, stream capture records closures).  It is NOT a CUDA emulator (no memory model, no divergence rules beyond "every lane of a warp reaches each
// shuffle"), kernels written in PTX (tcgen05 / TMA / mma.sync) are outside it, and it is never linked into the product.
#pragma once
#define D4_CUSIM 1
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <barrier>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static            // one copy per kernel (instantiation): blocks run one after the other

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorNotSupported = 801 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaStreamCaptureModeThreadLocal = 1, cudaDevAttrMultiProcessorCount = 16 };

namespace cusim {
struct Warp { std::barrier<>* bar; uint64_t slot[32]; };
extern thread_local Warp* warp;
extern thread_local std::barrier<>* block_bar;
extern unsigned char dyn_smem[];
void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t, std::function<void()> body);
}
extern thread_local dim3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

inline void __syncthreads() { cusim::block_bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cusim::warp->bar->arrive_and_wait(); }
template <class T> inline T cusim_shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    const int lane = threadIdx.x & 31;
    memcpy(&cusim::warp->slot[lane], &v, sizeof(T));
    cusim::warp->bar->arrive_and_wait();
    T r; memcpy(&r, &cusim::warp->slot[src & 31], sizeof(T));
    cusim::warp->bar->arrive_and_wait();
    return r;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int o) { return cusim_shfl(v, (int)(threadIdx.x & 31) ^ o); }
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return cusim_shfl(v, src); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int o) { const int l = threadIdx.x & 31; return cusim_shfl(v, l + o < 32 ? l + o : l); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int o) { const int l = threadIdx.x & 31; return cusim_shfl(v, l - o >= 0 ? l - o : l); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline float __fdividef(float a, float b) { return a / b; }

// ---- the runtime calls the engine and the launch wrappers make.  Stream capture is modelled as "record the closures instead of
// running them": a graph is the list of kernel launches / copies / memsets enqueued between Begin and EndCapture (the launch
// rewrite captures kernel arguments BY VALUE, as a real launch does), replayed in order by cudaGraphLaunch.
namespace cusim {
struct Graph { std::vector<std::function<void()>> nodes; };
extern Graph* capturing;
inline void enqueue(std::function<void()> op) { if (capturing) capturing->nodes.push_back(std::move(op)); else op(); }
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "cusim"; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { cusim::enqueue([=] { memmove(d, s, n); }); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    cusim::enqueue([=] { for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w); });
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { cusim::enqueue([=] { memset(d, v, n); }); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { static int token; *s = &token; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { if (cusim::capturing) return 900; cusim::capturing = new cusim::Graph(); return cudaSuccess; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = cusim::capturing; cusim::capturing = nullptr; return *g ? cudaSuccess : 901; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) { *e = new cusim::Graph(*static_cast<cusim::Graph*>(g)); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete static_cast<cusim::Graph*>(g); return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { delete static_cast<cusim::Graph*>(e); return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) { for (auto& op : static_cast<cusim::Graph*>(e)->nodes) op(); return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }
