// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE, not product code.
//
// A host-only stand-in for the slice of the CUDA runtime + device intrinsics that
// probreg_b200/csrc uses, so that the library's kernels AND its host orchestration can be
// executed on a machine without a GPU (the build container has none).  tests/emu/build.py
// compiles  csrc/cpd_b200.cu  (with `kernel<<<...>>>(...)` rewritten to emu::launch) against
// this header into tests/emu/_build/libcpd_b200_emu.so; only tests load that file.
//
// Execution model: blocks run one after another; the threads of a block are ucontext fibers
// scheduled round-robin, which makes __syncthreads / __syncwarp / warp shuffles / votes real
// barriers (a fiber yields until its peers arrive).  `__shared__` becomes `static`.
// mbarrier + TMA bulk copies are modelled by emu_device.h (copy performed when the barrier is
// waited on; the destination is poisoned at issue time so that a read-before-wait or a
// write-after-read hazard shows up as NaNs).  Nothing here models timing.
#pragma once
#define CPD_HOST_EMU 1

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <functional>

// ---- qualifiers --------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---- vector types ------------------------------------------------------------------------
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }
static inline int4 make_int4(int a, int b, int c, int d) { int4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }

// ---- runtime API subset --------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801,
       cudaErrorLaunchFailure = 719 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };

cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t s = nullptr);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind k,
                              cudaStream_t s = nullptr);
cudaError_t cudaMemset(void* p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
typedef struct emu_graph_exec* cudaGraphExec_t;      // never instantiated: the EM graph is a device-build feature
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
cudaError_t emu_func_set_attribute(const void* f, int attr, int value);
cudaError_t emu_occupancy(int* out, const void* f, int threads, size_t smem);
template <class F> cudaError_t cudaFuncSetAttribute(F* f, cudaFuncAttribute a, int v) { return emu_func_set_attribute((const void*)f, (int)a, v); }
template <class F> cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* out, F* f, int threads, size_t smem) {
    return emu_occupancy(out, (const void*)f, threads, smem);
}

// ---- device-side model ---------------------------------------------------------------------
namespace emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;       // dynamic shared memory of the running block (zero-size launches: still valid)
void launch(const char* name, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const std::function<void()>& body);
void syncthreads();
void syncwarp();
void exchange(const void* mine, size_t bytes, int src_lane, void* out);   // warp-wide: out = value published by src_lane
unsigned vote_any(bool pred);
void yield();
long long launches();                   // kernels launched since load (tests)
}  // namespace emu
#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) { emu::yield(); }
long long clock64();
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    T r;
    emu::exchange(&v, sizeof(T), (int)((threadIdx.x & 31u) ^ (unsigned)lane_mask), &r);
    return r;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    T r;
    emu::exchange(&v, sizeof(T), src & 31, &r);
    return r;
}
static inline int __any_sync(unsigned, int pred) { return emu::vote_any(pred != 0) != 0; }
static inline unsigned __ballot_sync(unsigned, int pred) { return emu::vote_any(pred != 0); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
// individually rounded float32 operations (the host build of the emulation is compiled without FMA contraction: -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { volatile float r = sqrtf(a); return r; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
