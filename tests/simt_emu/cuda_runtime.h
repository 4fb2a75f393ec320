/* cuda_runtime.h -- STAND-IN for the CUDA runtime header, used ONLY by the CPU SIMT emulator build of the test-suite
 * (tests/simt_emu/README.md).  TEST INFRASTRUCTURE: it is found instead of the real header because tests/simt_emu comes
 * first on the include path of that one build; the product (msc-futhark-ray-tracer_b200/Makefile, nvcc) never sees it.
 *
 * What it provides, so that the UNMODIFIED product sources (the .cu files under csrc) compile with g++ and run on the host:
 *   - the CUDA qualifiers as no-ops, vector types, the device intrinsics the sources use;
 *   - threadIdx / blockIdx / blockDim / gridDim of the lane that is running;
 *   - warp collectives (__ballot_sync, __any_sync, __shfl_sync, __shfl_xor_sync, __reduce_add_sync, __match_any_sync,
 *     __syncwarp) and __syncthreads with their real meaning: every GPU thread of a CTA is a fiber (emu_engine.cpp), a lane
 *     that reaches a collective parks until all live lanes of its warp (CTA) have reached one, then all get their results;
 *   - a host implementation of the few runtime calls the library makes (memory = malloc, streams = in-order immediate
 *     execution, events = wall-clock stamps, one device with a handful of "SMs").
 * Kernel launches `k<<<grid, block, smem, stream>>>(args)` are rewritten to emu::launch(...) by tests/simt_emu/transform.py.
 */
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <functional>

#define LYS_SIMT_EMU 1

/* ---- qualifiers ---------------------------------------------------------------------- */
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__             /* empty: libstdc++ spells __attribute__((__noinline__)) */
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static          /* one CTA runs at a time */

/* ---- vector types -------------------------------------------------------------------- */
struct alignas(8) float2 { float x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint3 { unsigned int x, y, z; };
struct dim3 {
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

/* ---- the SIMT engine (emu_engine.cpp) ------------------------------------------------ */
namespace emu {
enum Op { OP_BALLOT = 1, OP_SHFL, OP_SHFL_XOR, OP_REDUCE_ADD, OP_MATCH_ANY, OP_SYNCWARP, OP_REDUCE_MAX };
struct LaneIds { uint3 tid; };
extern uint3 g_block_idx, g_block_dim, g_grid_dim;
const uint3 &cur_tid();
uint32_t warp_collective(Op op, uint32_t a, uint32_t b);     /* parks the lane; returns its result */
void cta_barrier();
void *dyn_smem();
void launch(const char *name, dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
struct Stats { uint64_t launches, ctas, lanes, warp_collectives, cta_barriers; };
Stats stats();
}
#define threadIdx (emu::cur_tid())
#define blockIdx (emu::g_block_idx)
#define blockDim (emu::g_block_dim)
#define gridDim (emu::g_grid_dim)

/* ---- device intrinsics --------------------------------------------------------------- */
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned int __float_as_uint(float f) { unsigned int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned int x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned int)x); }
template <class T> static inline T __ldg(const T *p) { return *p; }
template <class T> static inline T __ldcg(const T *p) { return *(const volatile T *)p; }
template <> inline float4 __ldcg<float4>(const float4 *p) { return *p; }
static inline void __threadfence() {}
static inline void __syncthreads() { emu::cta_barrier(); }
/* block-wide vote barriers: lanes run one at a time, so a plain counter between barriers is exact */
namespace emu { inline int &sync_acc() { static int a; return a; } }
static inline int __syncthreads_count(int pred) {
    emu::cta_barrier();
    if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0) emu::sync_acc() = 0;
    emu::cta_barrier();
    if (pred) emu::sync_acc()++;
    emu::cta_barrier();
    const int r = emu::sync_acc();
    emu::cta_barrier();
    return r;
}
static inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)p; }
static inline void __syncwarp(unsigned int = 0xffffffffu) { (void)emu::warp_collective(emu::OP_SYNCWARP, 0, 0); }
static inline unsigned int __ballot_sync(unsigned int, int pred) { return emu::warp_collective(emu::OP_BALLOT, pred ? 1u : 0u, 0); }
static inline int __any_sync(unsigned int m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __shfl_sync(unsigned int, int v, int src) { return (int)emu::warp_collective(emu::OP_SHFL, (uint32_t)v, (uint32_t)src); }
static inline unsigned int __shfl_sync(unsigned int, unsigned int v, int src) { return emu::warp_collective(emu::OP_SHFL, v, (uint32_t)src); }
static inline float __shfl_sync(unsigned int, float v, int src) { return __uint_as_float(emu::warp_collective(emu::OP_SHFL, __float_as_uint(v), (uint32_t)src)); }
static inline int __shfl_xor_sync(unsigned int, int v, int m) { return (int)emu::warp_collective(emu::OP_SHFL_XOR, (uint32_t)v, (uint32_t)m); }
static inline unsigned int __shfl_xor_sync(unsigned int, unsigned int v, int m) { return emu::warp_collective(emu::OP_SHFL_XOR, v, (uint32_t)m); }
static inline float __shfl_xor_sync(unsigned int, float v, int m) { return __uint_as_float(emu::warp_collective(emu::OP_SHFL_XOR, __float_as_uint(v), (uint32_t)m)); }
static inline unsigned int __reduce_add_sync(unsigned int, unsigned int v) { return emu::warp_collective(emu::OP_REDUCE_ADD, v, 0); }
static inline int __reduce_max_sync(unsigned int, int v) { return (int)emu::warp_collective(emu::OP_REDUCE_MAX, (unsigned int)v, 0); }
static inline unsigned int __match_any_sync(unsigned int, unsigned int v) { return emu::warp_collective(emu::OP_MATCH_ANY, v, 0); }
/* atomics: lanes run one at a time, so plain read-modify-write is atomic */
template <class T, class U> static inline T atomicAdd(T *p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <class T, class U> static inline T atomicMin(T *p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <class T, class U> static inline T atomicMax(T *p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
/* CUDA's global min / max overloads */
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
static inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long long min(long long a, int b) { return a < b ? a : b; }
static inline long long max(long long a, int b) { return a > b ? a : b; }
static inline long long min(int a, long long b) { return a < b ? a : b; }
static inline long long max(int a, long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

/* ---- the runtime calls the library makes --------------------------------------------- */
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct emu_stream *cudaStream_t;
typedef struct emu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaDeviceLmemResizeToMax = 16 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; };

const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDevice(int *d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceFlags(unsigned int *f);
cudaError_t cudaSetDeviceFlags(unsigned int f);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d);
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr a, int d);
cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreate(cudaStream_t *s);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned int flags = 0);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *fr = (size_t)8 << 30; *tot = (size_t)16 << 30; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
