// host shim of lfx_common.cuh: lets g++ compile the device functions of lfx_draw.cu as plain C++ (one "thread" per block)
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include <string.h>
#include <algorithm>
using std::max; using std::min;
#define __device__
#define __global__
#define __host__
#define __constant__ static const
#define __shared__ static
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct Dim3 { unsigned x, y, z; };
static Dim3 threadIdx = {0,0,0}, blockIdx = {0,0,0}, blockDim = {1,1,1};
static inline void __syncthreads() {}
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline long long __double2ll_rn(double a) { return llrint(a); }
static inline int atomicMax(int* p, int v) { int o = *p; if (v > o) *p = v; return o; }
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 ld_stream16(const void* p) { uint4 r; memcpy(&r, p, 16); return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
static inline uint4 __ldg(const uint4* p) { return *p; }
typedef uintptr_t lfx_uintptr_shim;
#define __noinline__
enum { LFX_DRAW_NONE = 0, LFX_DRAW_LINE = 1, LFX_DRAW_LINE_AA = 2, LFX_DRAW_CIRCLE_FILLED = 3, LFX_DRAW_RECTANGLE = 4,
       LFX_DRAW_MARKER_CROSS = 5 };   // include/leafx.h
static inline uint8_t __ldg(const uint8_t* p) { return *p; }
static inline void __syncwarp() {}
