// Register-resident small DFTs and the 2048-point shared-memory FFT used by the
// OFDM demodulator (replaces the reference's FFTW call, reference
// src/acquire.c:254, plan at :318).  Forward transform, exp(-2*pi*i*n*k/N).
#pragma once
#include "common.cuh"

namespace nb {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// complex multiply with explicit fused multiply-adds (this header is also used by translation units
// built with -fmad=false)
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(__fmaf_rn(a.x, b.x, -(a.y * b.y)), __fmaf_rn(a.x, b.y, a.y * b.x));
}
// barrier over one 128-thread half of a CTA (id 1 or 2), or the whole CTA (id 0)
__device__ __forceinline__ void bar_sync(int id)
{
    if (id == 0) __syncthreads();
#if defined(NB_EMU)                      // tests/emu: the CPU emulator has no PTX
    else emu_bar_sync(id, 128);
#else
    else asm volatile("bar.sync %0, 128;" :: "r"(id) : "memory");
#endif
}
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }   // * (-j)

__device__ __forceinline__ void fft4(float2 &a0, float2 &a1, float2 &a2, float2 &a3)
{
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mj(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

// in-place 8-point DFT, natural order in and out
__device__ __forceinline__ void fft8(float2 *v)
{
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    fft4(e0, e1, e2, e3);
    fft4(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    float2 w1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));      // o1 * (1-j)/sqrt2
    float2 w2 = mul_mj(o2);
    float2 w3 = make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));     // o3 * (-1-j)/sqrt2
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
    v[2] = cadd(e2, w2); v[6] = csub(e2, w2);
    v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
}

// outputs 1, 2, 5, 6 only of the 8-point DFT of v (same arithmetic as fft8 for those outputs)
__device__ __forceinline__ void fft8_kept(const float2 *v, float2 &x1, float2 &x2, float2 &x5, float2 &x6)
{
    // even/odd 4-point DFTs, outputs 1 and 2 each
    const float2 et0 = cadd(v[0], v[4]), et1 = csub(v[0], v[4]), et2 = cadd(v[2], v[6]), et3 = mul_mj(csub(v[2], v[6]));
    const float2 ot0 = cadd(v[1], v[5]), ot1 = csub(v[1], v[5]), ot2 = cadd(v[3], v[7]), ot3 = mul_mj(csub(v[3], v[7]));
    const float2 e1 = cadd(et1, et3), e2 = csub(et0, et2);
    const float2 o1 = cadd(ot1, ot3), o2 = csub(ot0, ot2);
    const float h = 0.70710678118654752440f;
    const float2 w1 = make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));      // o1 * (1-j)/sqrt2
    const float2 w2 = mul_mj(o2);
    x1 = cadd(e1, w1); x5 = csub(e1, w1);
    x2 = cadd(e2, w2); x6 = csub(e2, w2);
}

// in-place 16-point DFT, natural order in and out
__device__ __forceinline__ void fft16(float2 *v)
{
    float2 e[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    fft8(e);
    fft8(o);
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;   // cos, sin (pi/8)
    const float h = 0.70710678118654752440f;
    const float2 w[8] = { make_float2(1.f, 0.f), make_float2(c1, -s1), make_float2(h, -h), make_float2(s1, -c1),
                          make_float2(0.f, -1.f), make_float2(-s1, -c1), make_float2(-h, -h), make_float2(-c1, -s1) };
#pragma unroll
    for (int k = 0; k < 8; k++) {
        float2 t = (k == 0) ? o[0] : (k == 4 ? mul_mj(o[4]) : cmul(o[k], w[k]));
        v[k] = cadd(e[k], t);
        v[k + 8] = csub(e[k], t);
    }
}

constexpr int FFT_THREADS = 128;
constexpr int FFT_LD = 129;                        // padded leading dimension of the pass-1 layout
constexpr int FFT_SMEM_ELEMS = 16 * FFT_LD;        // 2064 float2

// twiddle tables of fft2048_block: tw1[k1*128 + t] = W^(t*k1) (16 x 128), tw2[k2*8 + n3] = W^(16*n3*k2)
// (16 x 8), W = exp(-2*pi*i/2048): laid out so that a warp's loads are consecutive (tw1) or a two-address
// broadcast (tw2), from shared or global memory alike
constexpr int FFT_TW1 = 16 * 128, FFT_TW2 = 16 * 8, FFT_TW = FFT_TW1 + FFT_TW2;

// 2048-point FFT by 128 threads.  On entry thread r holds x[n1*128 + r] in
// v[n1], n1 = 0..15.  On exit thread t holds X[q + 256*k3] in out[h][k3] for
// q = t + 128*h, h = 0..1, k3 = 0..7.  `buf` is FFT_SMEM_ELEMS float2 of
// shared memory; `tw` = the FFT_TW-entry twiddle table above.
// KEPT_ONLY: the last pass only produces out[h][1,2,5,6] (the bins the receiver keeps); the rest is garbage.
template <bool KEPT_ONLY = false>
__device__ __forceinline__ void fft2048_block(float2 *v, float2 (*out)[8], float2 *buf,
                                              const float2 *__restrict__ tw, int t, int bar = 0)
{
    const float2 *tw2 = tw + FFT_TW1;
    // pass 1: DFT-16 over n1, twiddle W^(r*k1), store A[k1][r]
    fft16(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; k1++) {
        float2 x = v[k1];
        if (k1) x = cmul(x, tw[k1 * 128 + t]);
        buf[k1 * FFT_LD + t] = x;
    }
    bar_sync(bar);
    // pass 2: thread (k1, n3): DFT-16 over n2, twiddle W^(16*n3*k2)
    {
        const int k1 = t & 15, n3 = t >> 4;
        float2 u[16];
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) u[n2] = buf[k1 * FFT_LD + n2 * 8 + n3];
        fft16(u);
        bar_sync(bar);
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            float2 x = u[k2];
            if (k2) x = cmul(x, tw2[k2 * 8 + n3]);
            buf[n3 * 256 + k2 * 16 + k1] = x;
        }
    }
    bar_sync(bar);
    // pass 3: DFT-8 over n3 for q = k1 + 16*k2
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int q = t + 128 * h;
#pragma unroll
        for (int n3 = 0; n3 < 8; n3++) out[h][n3] = buf[n3 * 256 + q];
        if (KEPT_ONLY) {
            // the receiver keeps bins 478..744 and 1304..1570 of the shifted spectrum (sync.c:785-789): natural
            // bins 280..546 and 1502..1768, i.e. only k3 = 1, 2, 5, 6 of X[q + 256*k3]
            float2 x1, x2, x5, x6;
            fft8_kept(out[h], x1, x2, x5, x6);
            out[h][1] = x1; out[h][2] = x2; out[h][5] = x5; out[h][6] = x6;
        } else {
            fft8(out[h]);
        }
    }
}

}  // namespace nb
