// cuda_emu.h - TEST-ONLY single-threaded simulator for the subset of CUDA the cgnn kernels use.
//
// Purpose: this container has no GPU.  To catch indexing / synchronisation / logic bugs in
// the kernels before spending GPU minutes, the *same kernel sources* (csrc/*.cu) are compiled
// with g++ against this header (-DCGNN_EMU) into tests/emu/_build/libcgnn_emu.so and driven
// through the same C ABI with HOST pointers by tests/test_emu_*.py.
//
// It is NOT a product path: the package (connectome_gnn/_lib.py) never loads this library,
// nothing is timed through it, and it lives under tests/.  Every CUDA thread of a block is a
// ucontext fiber; __syncthreads / warp collectives are fiber barriers; blocks run one after
// another.  A barrier that not all live threads reach is reported as a deadlock and aborts.
#pragma once
#ifndef CGNN_EMU
#error "cuda_emu.h is only for -DCGNN_EMU builds"
#endif

#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static
#define __grid_constant__

struct uint3_ { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount, cudaDevAttrMaxSharedMemoryPerBlockOptin };

namespace cgnn_emu {

extern uint3_ g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void sync_threads();
// Warp exchange: every live lane of the calling warp deposits `v`; returns pointer to the 32 slots.
// Must be followed by warp_release() once the caller has read what it needs.
const uint64_t* warp_gather(uint64_t v);
void warp_release();
int lane_id();
// ---- warp-specialised kernels: polling waits (mbarrier / named barriers) and tensor memory -------------------
// A fiber that spins on shared state calls poll_yield() inside its loop; whoever changes state that a poller may be
// waiting for calls note_event().  A sweep in which only pollers ran and no event was noted is a deadlock.
void poll_yield();
void note_event();
float* tmem();                 // [128 lanes][512 columns] of the running block, zeroed at block start
void named_barrier(int id, int nthreads);   // bar.sync id, nthreads

}  // namespace cgnn_emu

#define threadIdx (cgnn_emu::g_threadIdx)
#define blockIdx (cgnn_emu::g_blockIdx)
#define blockDim (cgnn_emu::g_blockDim)
#define gridDim (cgnn_emu::g_gridDim)
#define warpSize 32

static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
  // few "SMs" so that persistent loops run several items per block even in tiny tests
  *v = (a == cudaDevAttrMultiProcessorCount) ? 3 : 227 * 1024;
  return cudaSuccess;
}

// ---- thread-level intrinsics ------------------------------------------------------------
static inline void __syncthreads() { cgnn_emu::sync_threads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { cgnn_emu::warp_gather(0); cgnn_emu::warp_release(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 8, "shfl type");
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  const uint64_t* slots = cgnn_emu::warp_gather(raw);
  int lane = cgnn_emu::lane_id();
  int base = (lane / width) * width;
  uint64_t got = slots[base + ((src % width) + width) % width];
  cgnn_emu::warp_release();
  T out; memcpy(&out, &got, sizeof(T)); return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int laneMask, int width = 32) {
  return __shfl_sync(m, v, (cgnn_emu::lane_id() % width) ^ laneMask, width);
}
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) {
  int l = cgnn_emu::lane_id() % width;
  return __shfl_sync(m, v, (l + (int)d < width) ? l + (int)d : l, width);
}
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) {
  int l = cgnn_emu::lane_id() % width;
  return __shfl_sync(m, v, (l - (int)d >= 0) ? l - (int)d : l, width);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  const uint64_t* s = cgnn_emu::warp_gather(pred ? 1 : 0);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (s[i] == 1) r |= (1u << i);
  cgnn_emu::warp_release();
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p);
static inline unsigned __activemask() { return 0xffffffffu; }
static inline unsigned __match_any_sync(unsigned, int v) {
  // dead lanes hold the sentinel 0xDEAD...; tag live values so they never collide with it
  const uint64_t* s = cgnn_emu::warp_gather(((uint64_t)1 << 40) | (uint32_t)v);
  uint64_t mine = ((uint64_t)1 << 40) | (uint32_t)v;
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) if (s[i] == mine) r |= (1u << i);
  cgnn_emu::warp_release();
  return r;
}
static inline int __all_sync(unsigned, int p) {
  const uint64_t* s = cgnn_emu::warp_gather(p ? 1 : 0);
  int ok = 1;
  for (int i = 0; i < 32; ++i) if (s[i] == 0) ok = 0;  // dead lanes hold a non-zero sentinel
  cgnn_emu::warp_release();
  return ok;
}

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }

template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline long min(long a, long b) { return a < b ? a : b; }
static inline long max(long a, long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }
