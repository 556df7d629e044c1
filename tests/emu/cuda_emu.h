// cuda_emu -- a small CPU emulation of the CUDA execution model, TEST INFRASTRUCTURE ONLY.
//
// tests/emu/build_emu.py compiles the engine's unmodified CUDA sources (nrsc5_b200/csrc/*.cu, *.cuh) with g++
// against this header into tests/_build/libnrsc5_b200_emu.so, which exports the same C ABI as the product library.
// The product never loads it: only tests/test_emu_engine.py does, to run parity tests of the real kernel source
// on a box without a GPU and to catch barrier divergence / deadlocks before they can hang a device.
//
//  * a thread block = one fibre per CUDA thread (own stack, private locals = registers), blocks run one after the
//    other, kernels run synchronously at launch, streams and events are no-ops;
//  * __shared__ = static storage (one block is resident at a time), dynamic shared memory = one static arena;
//  * __syncthreads / __syncwarp / bar.sync / __shfl_*_sync / __ballot_sync are real barriers between fibres:
//    a fibre that arrives switches to the next runnable one; if no fibre can make progress the run aborts with
//    the list of waiting threads (on the GPU this would be a hang);
//  * device memory = host heap, filled with 0xA5 at allocation so that reads of never-written memory show up.
//
// Floating point: the sources are compiled with -ffp-contract=off (like -fmad=false); sincosf, atan2f, the
// approximate division etc. are glibc's, so floats can differ from the GPU's in the last place.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#define NB_EMU 1
#ifndef __CUDA_ARCH__
#define __CUDA_ARCH__ 1000
#endif

// ---- keywords ----
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

// ---- vector types ----
struct alignas(2) char2 { signed char x, y; };
struct alignas(2) uchar2 { unsigned char x, y; };
struct alignas(4) char4 { signed char x, y, z, w; };
struct alignas(4) uchar4 { unsigned char x, y, z, w; };
struct alignas(4) short2 { short x, y; };
struct alignas(4) ushort2 { unsigned short x, y; };
struct alignas(8) short4 { short x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(16) longlong2 { long long x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
inline char2 make_char2(signed char x, signed char y) { return char2{ x, y }; }
inline uchar2 make_uchar2(unsigned char x, unsigned char y) { return uchar2{ x, y }; }
inline char4 make_char4(signed char x, signed char y, signed char z, signed char w) { return char4{ x, y, z, w }; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{ x, y, z, w }; }
inline short2 make_short2(short x, short y) { return short2{ x, y }; }
inline ushort2 make_ushort2(unsigned short x, unsigned short y) { return ushort2{ x, y }; }
inline short4 make_short4(short x, short y, short z, short w) { return short4{ x, y, z, w }; }
inline int2 make_int2(int x, int y) { return int2{ x, y }; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{ x, y, z, w }; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{ x, y, z, w }; }
inline float2 make_float2(float x, float y) { return float2{ x, y }; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{ x, y, z, w }; }
inline double2 make_double2(double x, double y) { return double2{ x, y }; }

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- the fibre runtime (tests/emu/cuda_emu.cpp) ----
namespace emu {
struct Idx {
    unsigned x, y, z;
};
extern Idx g_block_idx, g_block_dim, g_grid_dim;
Idx cur_tid();
int cur_linear_tid();
void sync_threads();
void sync_warp(unsigned mask);
void sync_named(int id, int nthreads);
uint64_t *warp_mail();                       // 32 slots of the calling thread's warp
unsigned char *dyn_smem();                   // the dynamic shared memory arena of the resident block
void run_grid(dim3 grid, dim3 block, size_t smem, void (*body)(void *), void *arg);

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F &&f)
{
    run_grid(grid, block, smem, [](void *p) { (*static_cast<F *>(p))(); }, &f);
}
}   // namespace emu

#define threadIdx (emu::cur_tid())
#define blockIdx (emu::g_block_idx)
#define blockDim (emu::g_block_dim)
#define gridDim (emu::g_grid_dim)
constexpr int warpSize = 32;
// kernel<<<grid, block, smem, stream>>>(args) is rewritten by build_emu.py into this
#define EMU_LAUNCH(grid, block, smem, kernel, ...) emu::launch(dim3 grid, dim3 block, (size_t)(smem), [&]() { kernel(__VA_ARGS__); })

// ---- synchronisation and warp primitives ----
inline void __syncthreads() { emu::sync_threads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::sync_warp(mask); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}

namespace emu {
template <class T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "shuffle of more than 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
template <class T, class SrcOf>
inline T shuffle(unsigned mask, T v, int width, SrcOf src_of)
{
    const int lane = cur_linear_tid() & 31;
    uint64_t *mail = warp_mail();
    mail[lane] = to_bits(v);
    sync_warp(mask);
    int src = src_of(lane);
    const int base = lane & ~(width - 1);
    if (src < base || src >= base + width) src = lane;          // out of the segment: own value
    const T r = ((mask >> src) & 1) ? from_bits<T>(mail[src]) : v;   // a lane outside the mask: undefined on the GPU
    sync_warp(mask);
    return r;
}
}   // namespace emu

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    return emu::shuffle(mask, v, width, [=](int lane) { return (lane & ~(width - 1)) + (src & (width - 1)); });
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    return emu::shuffle(mask, v, width, [=](int lane) { return lane ^ lanemask; });
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    return emu::shuffle(mask, v, width, [=](int lane) { return lane + (int)delta; });
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    return emu::shuffle(mask, v, width, [=](int lane) { return lane - (int)delta; });
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    const int lane = emu::cur_linear_tid() & 31;
    uint64_t *mail = emu::warp_mail();
    mail[lane] = pred ? 1 : 0;
    emu::sync_warp(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (((mask >> l) & 1) && mail[l]) r |= 1u << l;
    emu::sync_warp(mask);
    return r;
}
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline unsigned __activemask() { return 0xffffffffu; }

// ---- memory ----
template <class T>
inline T __ldg(const T *p) { return *p; }
template <class T>
inline T __ldcg(const T *p) { return *p; }
template <class T>
inline T __ldcs(const T *p) { return *p; }
template <class T>
inline void __stcg(T *p, T v) { *p = v; }
template <class T>
inline void __stcs(T *p, T v) { *p = v; }

template <class T>
inline T atomicAdd(T *p, T v) { const T o = *p; *p = (T)(o + v); return o; }
template <class T>
inline T atomicExch(T *p, T v) { const T o = *p; *p = v; return o; }
template <class T>
inline T atomicMax(T *p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <class T>
inline T atomicMin(T *p, T v) { const T o = *p; if (v < o) *p = v; return o; }
template <class T>
inline T atomicOr(T *p, T v) { const T o = *p; *p = o | v; return o; }
template <class T>
inline T atomicAnd(T *p, T v) { const T o = *p; *p = o & v; return o; }
template <class T>
inline T atomicCAS(T *p, T cmp, T v) { const T o = *p; if (o == cmp) *p = v; return o; }

// ---- integer intrinsics ----
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned __brev(unsigned v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(v);
}
inline unsigned emu_prmt(unsigned a, unsigned b, unsigned sel)       // prmt.b32, default mode
{
    const uint64_t pool = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(pool >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;             // replicate the sign
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) { return emu_prmt(a, b, sel & 0x7777u); }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (shift & 31));
}
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned shift)
{
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)((v << (shift & 31)) >> 32);
}
inline int __dp4a(int a, int b, int c)
{
    for (int i = 0; i < 4; i++) c += (int)(int8_t)(a >> (8 * i)) * (int)(int8_t)(b >> (8 * i));
    return c;
}
inline unsigned __dp4a(unsigned a, unsigned b, unsigned c)
{
    for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xff) * ((b >> (8 * i)) & 0xff);
    return c;
}
namespace emu {
inline int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
template <class F>
inline unsigned simd2(unsigned a, unsigned b, F f)
{
    const unsigned lo = (unsigned)f((int)(short)(a & 0xffff), (int)(short)(b & 0xffff)) & 0xffff;
    const unsigned hi = (unsigned)f((int)(short)(a >> 16), (int)(short)(b >> 16)) & 0xffff;
    return lo | (hi << 16);
}
template <class F>
inline unsigned simd2u(unsigned a, unsigned b, F f)
{
    const unsigned lo = (unsigned)f((int)(a & 0xffff), (int)(b & 0xffff)) & 0xffff;
    const unsigned hi = (unsigned)f((int)(a >> 16), (int)(b >> 16)) & 0xffff;
    return lo | (hi << 16);
}
}   // namespace emu
inline unsigned __vaddss2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return emu::sat16(x + y); }); }
inline unsigned __vsubss2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return emu::sat16(x - y); }); }
inline unsigned __vadd2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x + y; }); }
inline unsigned __vsub2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x - y; }); }
inline unsigned __vmaxs2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x > y ? x : y; }); }
inline unsigned __vmins2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x < y ? x : y; }); }
inline unsigned __vcmpgts2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x > y ? 0xffff : 0; }); }
// the PTX the sources issue through inline asm (guarded by NB_EMU there)
inline unsigned emu_add_s16x2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x + y; }); }
inline unsigned emu_max_s16x2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x > y ? x : y; }); }
inline unsigned emu_min_s16x2(unsigned a, unsigned b) { return emu::simd2(a, b, [](int x, int y) { return x < y ? x : y; }); }
inline unsigned emu_max_u16x2(unsigned a, unsigned b) { return emu::simd2u(a, b, [](int x, int y) { return x > y ? x : y; }); }
inline unsigned emu_min_u16x2(unsigned a, unsigned b) { return emu::simd2u(a, b, [](int x, int y) { return x < y ? x : y; }); }
inline void emu_bar_sync(int id, int nthreads) { emu::sync_named(id, nthreads); }

// ---- float intrinsics ----
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline int __float2int_rz(float a) { return (int)a; }
inline int __float2int_rn(float a) { return (int)lrintf(a); }
inline float __int2float_rn(int a) { return (float)a; }
inline float __int_as_float(int a) { return emu::from_bits<float>((uint64_t)(uint32_t)a); }
inline int __float_as_int(float a) { return (int)(uint32_t)emu::to_bits(a); }
inline unsigned __float_as_uint(float a) { return (uint32_t)emu::to_bits(a); }
inline float __uint_as_float(unsigned a) { return emu::from_bits<float>((uint64_t)a); }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
inline float __sinf(float a) { return sinf(a); }
inline float __cosf(float a) { return cosf(a); }
inline void __sincosf(float a, float *s, float *c) { sincosf(a, s, c); }
inline float __expf(float a) { return expf(a); }
inline float __logf(float a) { return logf(a); }
inline float __log10f(float a) { return log10f(a); }
inline float __powf(float a, float b) { return powf(a, b); }
inline float __saturatef(float a) { return a < 0 ? 0.0f : (a > 1 ? 1.0f : a); }
inline long long clock64() { return (long long)__builtin_ia32_rdtsc(); }

// CUDA's overloaded min / max
#define EMU_MINMAX(T)                                  \
    inline T min(T a, T b) { return b < a ? b : a; } \
    inline T max(T a, T b) { return a < b ? b : a; }
EMU_MINMAX(int)
EMU_MINMAX(unsigned)
EMU_MINMAX(long)
EMU_MINMAX(unsigned long)
EMU_MINMAX(long long)
EMU_MINMAX(unsigned long long)
EMU_MINMAX(float)
EMU_MINMAX(double)
#undef EMU_MINMAX
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
inline long long min(long long a, int b) { return min(a, (long long)b); }
inline long long min(int a, long long b) { return min((long long)a, b); }
inline long long max(long long a, int b) { return max(a, (long long)b); }
inline long long max(int a, long long b) { return max((long long)a, b); }
inline unsigned long min(unsigned long a, int b) { return min(a, (unsigned long)b); }
inline unsigned long min(unsigned long a, unsigned b) { return min(a, (unsigned long)b); }
inline unsigned long min(unsigned a, unsigned long b) { return min((unsigned long)a, b); }
inline unsigned long max(unsigned long a, unsigned b) { return max(a, (unsigned long)b); }

// ---- runtime API (synchronous, everything lives on the host heap) ----
enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 11, cudaErrorNotReady = 600 };
typedef cudaError_t cudaError;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
    void *devicePointer, *hostPointer;
};
struct emuStream {
    int dummy;
};
struct emuEvent {
    double t_ms;
};
typedef emuStream *cudaStream_t;
typedef emuEvent *cudaEvent_t;

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t n)
{
    void *q = malloc(n ? n : 1);
    if (!q) return cudaErrorMemoryAllocation;
    memset(q, 0xA5, n);                       // device memory is not zeroed: make reads of unwritten bytes visible
    *p = static_cast<T *>(q);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T>
inline cudaError_t cudaMallocHost(T **p, size_t n)
{
    void *q = malloc(n ? n : 1);
    if (!q) return cudaErrorMemoryAllocation;
    *p = static_cast<T *>(q);
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2D(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k); }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset2DAsync(void *d, size_t pitch, int v, size_t w, size_t h, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < h; r++) memset((char *)d + r * pitch, v, w);
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *s, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice)
{
    memcpy((char *)&sym + off, s, n);
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMemcpyFromSymbol(void *d, const T &sym, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyDeviceToHost)
{
    memcpy(d, (const char *)&sym + off, n);
    return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMemcpyFromSymbolAsync(void *d, const T &sym, size_t n, size_t off, cudaMemcpyKind, cudaStream_t = nullptr)
{
    memcpy(d, (const char *)&sym + off, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new emuStream{ 0 }; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
double emu_now_ms();
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent{ 0 }; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t_ms = emu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
    a->type = cudaMemoryTypeHost;
    a->device = 0;
    a->devicePointer = a->hostPointer = const_cast<void *>(p);
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
