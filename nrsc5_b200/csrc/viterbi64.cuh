// Register-resident chunk-parallel Viterbi (K=7, rate 1/3, tail-biting): the fast path of the P1 decode.
//
// The reference decodes a frame with one sequential pass of len+64 add-compare-select steps (reference
// src/conv_dec.c:402-427, SSE kernel src/conv_sse.h:56-66,233-315).  Here
//
//   k_v64_fwd    one THREAD per chunk of `ch` steps: all 64 path metrics live in 32 registers as packed
//                u16x2, a step is 16 fully unrolled butterfly pairs with no cross-lane traffic at all;
//                every chunk first replays V64_WARM warm-up steps from all-zero metrics
//   k_v64_check  accepts the frame only if every chunk's warmed-up metric vector equals - up to one common
//                constant - the vector its predecessor ended with (equal vectors => equal decisions from
//                there on, so accepted decisions are exactly the sequential pass's); picks the end state
//   k_v64_emit   one warp per 1024-step window: the survivor state at the window's end is the one all 64
//                survivors of the following 128 steps merge into (checked, not assumed); 32 lanes then
//                emit 32 bits each from guessed segment end states that are verified against their
//                neighbours from the known window end downwards
//
// Anything that cannot be proven on this path (a chunk that did not converge, survivors that did not
// merge, metrics that could saturate in the reference's int16 arithmetic) flags the frame `retry`; it is
// then decoded by the exact half-warp kernels of viterbi_chunk.cuh.  Speculation only ever costs time.
//
// Arithmetic.  Add-compare-select decisions depend only on metric differences, so - as long as nothing
// saturates in the reference and nothing wraps here - they are unchanged by (a) adding the same bias to
// all branch metrics of a step and (b) subtracting a common constant from all path metrics.  With the
// bias V64_BIAS = 384 >= 3*127 every branch metric is positive, path metrics are unsigned, halves never
// carry into each other and plain 32-bit integer adds do two states at a time; the minimum is subtracted
// every V64_NORM = 32 steps, which keeps all metrics below 2^15 (spread <= 12*381, growth <= 765/step), the
// range in which  (Y + 0x8000 - X) >> 15  is the comparison  Y >= X  (ties: the odd predecessor wins, as
// in src/conv_gen.h:47,55).
//
// Decision word of a step (uint2 w): the bit of new state n is bit 8*(n>>4) + (n&7) of w.x (n&8 == 0) or
// w.y (n&8 != 0); set = the survivor comes from the odd predecessor 2*(n&31)+1.
#pragma once
#include "common.cuh"
#include "viterbi_chunk.cuh"

namespace nb {

constexpr int V64_WARM = 256;
constexpr int V64_NORM = 32;
constexpr int V64_BIAS = 384;
constexpr int V64_WIN = 1024;                           // traceback window (steps per emit warp)
constexpr int V64_HEAD = 128;                           // look-ahead in which all survivors must merge
constexpr int V64_GUESS = 64;                           // look-ahead of a segment's (verified) end-state guess

struct V64Args {
    const int8_t *vin;      // [frames][3*len]
    uint2 *dec;             // [frames][dec_stride] decision words, one per step
    uint32_t *vspec;        // [frames][nch][32] metrics at each chunk start (after warm-up)
    uint32_t *vend;         // [frames][nch][32] metrics at each chunk end
    int *endstate;          // [frames] survivor state after the last step
    uint32_t *bitsw;        // [frames][len/32] decoded bits
    const int *ready;       // ready[f*stride] != 0 selects the frames to decode
    int *retry;             // retry[f*stride] = 1: decode this frame with the exact fallback kernels
    int stride;             // in ints
    int len;                // frame length in bits (multiple of 32)
    int ch;                 // chunk length in steps (multiple of 32)
    int nch;                // chunks per frame
    size_t dec_stride;      // uint2 per frame
};

__device__ __forceinline__ unsigned v64_umax(unsigned a, unsigned b)
{
    unsigned r;
#if defined(NB_EMU)
    r = emu_max_u16x2(a, b);
#else
    asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
    return r;
}
__device__ __forceinline__ unsigned v64_umin(unsigned a, unsigned b)
{
    unsigned r;
#if defined(NB_EMU)
    r = emu_min_u16x2(a, b);
#else
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
    return r;
}
__device__ __forceinline__ unsigned v64_prmt(unsigned a, unsigned b, unsigned sel)
{
    unsigned r;
#if defined(NB_EMU)
    r = emu_prmt(a, b, sel);
#else
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
#endif
    return r;
}

// sign pattern of butterfly b: expected code bits of the branch (state 2b, input 0), src/conv_dec.c:139-154
__host__ __device__ constexpr int v64_parity(unsigned v)
{
    return (int)((v ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4) ^ (v >> 5) ^ (v >> 6)) & 1u);
}
// index into the 8 packed branch-metric constants M[] below for butterfly pair l (butterflies l, l+16)
__host__ __device__ constexpr int v64_mp_index(int l)
{
    const unsigned reg = (unsigned)l << 1;
    const int k0 = v64_parity(reg & 0133u), k1 = v64_parity(reg & 0171u), k2 = v64_parity(reg & 0165u);   // 1 = +
    // M index: bit 2 = sign of s0 (0: +), bits 1..0 = u in { s1+s2, -(s1+s2), s1-s2, -(s1-s2) }
    const int u = (k1 && k2) ? 0 : (!k1 && !k2) ? 1 : (k1 && !k2) ? 2 : 3;
    return (k0 ? 0 : 4) | u;
}
__host__ __device__ constexpr int v64_mm_index(int l)      // the negated metric: other sign of s0, -u
{
    const int i = v64_mp_index(l);
    return (i ^ 4) ^ 1;
}

// 64-state path metrics of one chunk, P[i] = (pm[i], pm[i+32]) as u16x2
struct V64State {
    unsigned P[32];

    // one step; s0,s1,s2 = soft values of the step; returns the decision word
    __device__ __forceinline__ uint2 step(int s0, int s1, int s2)
    {
        // the 8 packed, biased branch metrics: M[(s0 sign, u)] = (c + u, c - u), c = BIAS +- s0
        const int c1 = V64_BIAS + s0, c2 = V64_BIAS - s0, b1 = s1 + s2, b2 = s1 - s2;
        unsigned M[8];
        M[0] = (unsigned)(c1 + b1) | ((unsigned)(c1 - b1) << 16);
        M[2] = (unsigned)(c1 + b2) | ((unsigned)(c1 - b2) << 16);
        M[4] = (unsigned)(c2 + b1) | ((unsigned)(c2 - b1) << 16);
        M[6] = (unsigned)(c2 + b2) | ((unsigned)(c2 - b2) << 16);
        M[1] = v64_prmt(M[0], 0, 0x1032);
        M[3] = v64_prmt(M[2], 0, 0x1032);
        M[5] = v64_prmt(M[4], 0, 0x1032);
        M[7] = v64_prmt(M[6], 0, 0x1032);
        unsigned N[32];
        unsigned w0 = 0, w1 = 0;
#pragma unroll
        for (int l = 0; l < 16; l++) {
            const unsigned E = P[2 * l], O = P[2 * l + 1];
            const unsigned Mp = M[v64_mp_index(l)], Mm = M[v64_mm_index(l)];
            const unsigned X1 = E + Mp, Y1 = O + Mm, X2 = E + Mm, Y2 = O + Mp;
            const unsigned N0 = v64_umax(X1, Y1), N1 = v64_umax(X2, Y2);
            const unsigned T1 = Y1 + 0x80008000u - X1, T2 = Y2 + 0x80008000u - X2;     // bit 15/31: Y >= X
            N[l] = v64_prmt(N0, N1, 0x5410);
            N[l + 16] = v64_prmt(N0, N1, 0x7632);
            const unsigned R = v64_prmt(T1, T2, 0xFDB9);      // sign masks of new states l, l+16, l+32, l+48
            if (l < 8) w0 |= R & (0x01010101u << (l & 7));
            else w1 |= R & (0x01010101u << (l & 7));
        }
#pragma unroll
        for (int i = 0; i < 32; i++) P[i] = N[i];
        return make_uint2(w0, w1);
    }

    __device__ __forceinline__ void normalise()
    {
        unsigned m = P[0];
#pragma unroll
        for (int i = 1; i < 32; i++) m = v64_umin(m, P[i]);
        m = v64_umin(m, v64_prmt(m, 0, 0x1032));
#pragma unroll
        for (int i = 0; i < 32; i++) P[i] -= m;
    }
};

__device__ __forceinline__ int v64_prev(int state, uint2 w)
{
    const unsigned word = (state & 8) ? w.y : w.x;
    return ((state << 1) & 62) | (int)((word >> (8 * (state >> 4) + (state & 7))) & 1u);
}

// ---------------------------------------------------------------------------
// forward pass: one thread per chunk
// ---------------------------------------------------------------------------
constexpr int V64_FWD_THREADS = 128;                   // 4 warps: one per warp scheduler of an SM

__global__ void __launch_bounds__(V64_FWD_THREADS) k_v64_fwd(V64Args a)
{
    const int f = blockIdx.y;
    if (!a.ready[(size_t)f * a.stride]) return;
    const int c = blockIdx.x * V64_FWD_THREADS + threadIdx.x;
    if (c >= a.nch) return;
    const int total = a.len + 64;
    const int8_t *vin = a.vin + (size_t)f * 3 * a.len;
    uint2 *dec = a.dec + (size_t)f * a.dec_stride;
    const int s_begin = c * a.ch, s_end = min(total, s_begin + a.ch);
    // saturation guard (see viterbi_chunk.cuh): between two of the reference's normalisations (every 79
    // steps) the metrics grow by at most the sum of |s0|+|s1|+|s2|; with a spread of at most 12*381 after
    // a normalisation, sums <= 32767 - 12*381 cannot saturate.  Every 79-step window lies inside the
    // range (warm-up included) of at least one chunk.
    const int sat_limit = 32767 - 12 * 381;
    int wsum = 0;
    bool bad = false;

    V64State vs;
#pragma unroll
    for (int i = 0; i < 32; i++) vs.P[i] = 0;

    // 4 steps (12 soft bytes = three words) per iteration: the unrolled body must stay inside the 32 KB
    // instruction cache, one warp per scheduler cannot hide instruction fetches.  The next iteration's
    // words are loaded one iteration ahead.
    auto fetch = [&](int g, uint32_t &x0, uint32_t &x1, uint32_t &x2) {
        if (g >= 0 && g < s_end) {
            int j = g + a.len - 32;
            if (j >= a.len) j -= a.len;
            if (j >= a.len) j -= a.len;
            const uint32_t *q = reinterpret_cast<const uint32_t *>(vin + 3 * (size_t)j);
            x0 = q[0]; x1 = q[1]; x2 = q[2];
        } else {
            x0 = x1 = x2 = 0;
        }
    };
    const int g_first = s_begin - V64_WARM;
    int wc = ((g_first % 79) + 79) % 79;               // step index modulo 79
    uint32_t n0, n1, n2;
    fetch(g_first, n0, n1, n2);
#pragma unroll 1
    for (int g0 = g_first; g0 < s_end; g0 += 4) {
        const uint32_t sw[3] = { n0, n1, n2 };
        fetch(g0 + 4, n0, n1, n2);
        const bool store = g0 >= s_begin;
        uint2 held = make_uint2(0, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            auto sb = [&](int pos) -> int {               // sign-extended byte `pos` of the 12-byte group
                const unsigned sel = (unsigned)(pos & 3) * 0x1111u | 0x8880u;
                return (int)v64_prmt(sw[pos >> 2], 0, sel);
            };
            const int s0 = sb(3 * k), s1 = sb(3 * k + 1), s2 = sb(3 * k + 2);
            wsum += abs(s0) + abs(s1) + abs(s2);
            if (wc == 0) {                                 // the reference normalises here (src/conv_dec.c:419)
                bad |= wsum > sat_limit;
                wsum = 0;
            }
            wc = wc == 78 ? 0 : wc + 1;
            const uint2 w = vs.step(s0, s1, s2);
            if (k & 1) {
                if (store) *reinterpret_cast<uint4 *>(dec + g0 + k - 1) = make_uint4(held.x, held.y, w.x, w.y);
            } else {
                held = w;
            }
        }
        if (((g0 + 4) & (V64_NORM - 1)) == 0) vs.normalise();
        if (g0 + 4 == s_begin) {                           // warm-up done: publish the speculative start vector
            uint32_t *o = a.vspec + ((size_t)f * a.nch + c) * 32;
#pragma unroll
            for (int i = 0; i < 32; i++) o[i] = vs.P[i];
        }
    }
    {
        uint32_t *o = a.vend + ((size_t)f * a.nch + c) * 32;
#pragma unroll
        for (int i = 0; i < 32; i++) o[i] = vs.P[i];
    }
    if (bad) atomicExch(&a.retry[(size_t)f * a.stride], 1);
}

// ---------------------------------------------------------------------------
// check: one CTA per frame; warp w verifies chunks w, w+nwarps, ...; warp 0 also picks the end state
// ---------------------------------------------------------------------------
constexpr int V64_CHECK_THREADS = 256;

__global__ void __launch_bounds__(V64_CHECK_THREADS) k_v64_check(V64Args a)
{
    const int f = blockIdx.x;
    if (!a.ready[(size_t)f * a.stride]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = V64_CHECK_THREADS / 32;
    const uint32_t *vspec = a.vspec + (size_t)f * a.nch * 32, *vend = a.vend + (size_t)f * a.nch * 32;
    bool ok = true;
    for (int c = 1 + warp; c < a.nch; c += nwarps) {
        const unsigned x = vend[(size_t)(c - 1) * 32 + lane], y = vspec[(size_t)c * 32 + lane];
        // equal up to one constant added to all 64 metrics (reference point: state 0)
        const int xr = (int)(__shfl_sync(0xffffffffu, x, 0) & 0xffff), yr = (int)(__shfl_sync(0xffffffffu, y, 0) & 0xffff);
        ok &= ((int)(x & 0xffff) - xr == (int)(y & 0xffff) - yr) && ((int)(x >> 16) - xr == (int)(y >> 16) - yr);
    }
    if (!__all_sync(0xffffffffu, ok) && lane == 0) atomicExch(&a.retry[(size_t)f * a.stride], 1);
    if (warp == 0) {
        // first maximum in state order (reference src/conv_dec.c:310-317); lane i holds states i and i+32
        const unsigned x = vend[(size_t)(a.nch - 1) * 32 + lane];
        int v = (int)(x & 0xffff), idx = lane;
        const int v2 = (int)(x >> 16);
        if (v2 > v) { v = v2; idx = lane + 32; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ov = __shfl_xor_sync(0xffffffffu, v, o), oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (lane == 0) a.endstate[f] = idx;
    }
}

// ---------------------------------------------------------------------------
// emit: one warp per 1024-step window
// ---------------------------------------------------------------------------
constexpr int V64_EMIT_WARPS = 4;
constexpr int V64_EMIT_STEPS = V64_WIN + V64_HEAD;
// staged decisions are padded by one word per 32 steps: the lanes walk segments 32 steps apart at the same time
constexpr int V64_EMIT_LD = V64_EMIT_STEPS + V64_EMIT_STEPS / 32 + 1;
constexpr size_t V64_EMIT_SMEM = (size_t)V64_EMIT_WARPS * V64_EMIT_LD * sizeof(uint2);
__device__ __forceinline__ int v64_pad(int q) { return q + (q >> 5); }

__global__ void __launch_bounds__(V64_EMIT_WARPS * 32) k_v64_emit(V64Args a)
{
#if defined(NB_EMU)
    unsigned char *v64_smem = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char v64_smem[];
#endif
    const int f = blockIdx.y;
    if (!a.ready[(size_t)f * a.stride]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int total = a.len + 64;
    const int w = blockIdx.x * V64_EMIT_WARPS + warp;
    const int lo = w * V64_WIN;
    if (lo >= total) return;
    uint2 *sd = reinterpret_cast<uint2 *>(v64_smem) + (size_t)warp * V64_EMIT_LD;
    const uint2 *dec = a.dec + (size_t)f * a.dec_stride + lo;
    const int n = min(total - lo, V64_WIN);             // steps of this window
    const int nst = min(total - lo, V64_EMIT_STEPS);    // staged steps (incl. look-ahead)
    for (int i = lane; i < nst; i += 32) sd[v64_pad(i)] = dec[i];
    __syncwarp();
    auto walk = [&](int state, int from, int to) {      // state after local step `from` -> state after local step `to`
        for (int q = from; q > to; q--) state = v64_prev(state, sd[v64_pad(q)]);
        return state;
    };
    // the window's end state: the frame's end state for the last window, otherwise the state all survivors
    // of the look-ahead merge into
    int true_end;
    if (nst == n) {
        true_end = a.endstate[f];
    } else {
        // all 64 survivors (two per lane) of a look-ahead of 64 steps, else of the whole staged look-ahead
        bool merged = false;
        int ref = 0;
        for (int look = min(V64_HEAD / 2, nst - n); ; look = nst - n) {
            int e0 = lane, e1 = lane + 32;
            for (int q = n + look - 1; q > n - 1; q--) {
                const uint2 w = sd[v64_pad(q)];
                e0 = v64_prev(e0, w);
                e1 = v64_prev(e1, w);
            }
            ref = __shfl_sync(0xffffffffu, e0, 0);
            merged = __all_sync(0xffffffffu, e0 == ref && e1 == ref);
            if (merged || look == nst - n) break;
        }
        if (!merged) {
            if (lane == 0) atomicExch(&a.retry[(size_t)f * a.stride], 1);
            return;
        }
        true_end = ref;
    }
    const int seg_end = 32 * lane + 31;                 // local index of this lane's last step
    const bool have = 32 * lane < n;
    int g;                                              // state after local step seg_end
    if (!have) g = 0;
    else if (seg_end >= n - 1) g = true_end;
    else if (seg_end + V64_GUESS >= nst) g = walk(true_end, n - 1, seg_end);         // near the frame end
    else g = walk(0, seg_end + V64_GUESS, seg_end);                                 // guess
    unsigned word = 0;
    int b = 0;                                          // state before the segment's first step
    auto emit = [&](int gstate) {
        int state = gstate;
        unsigned ww = 0;
        for (int k = 31; k >= 0; k--) {
            ww |= (unsigned)((state >> 5) & 1) << k;
            state = v64_prev(state, sd[33 * lane + k]);
        }
        word = ww;
        b = state;
    };
    if (have) emit(g);
    // verification from the window end downwards
    const int nseg = (n + 31) / 32;
    for (;;) {
        const int bnext = __shfl_down_sync(0xffffffffu, b, 1);
        const bool wrong = have && lane < nseg - 1 && g != bnext;
        const unsigned m = __ballot_sync(0xffffffffu, wrong);
        if (!m) break;
        const int jj = 31 - __clz(m);                   // highest wrong segment: its right neighbour is already final
        if (lane == jj) { g = bnext; emit(g); }
    }
    const int widx = (lo >> 5) + lane - 1;              // frame bit 32*widx = step lo + 32*lane - 32
    if (have && widx >= 0 && widx < a.len / 32) a.bitsw[(size_t)f * (a.len / 32) + widx] = word;
}

}  // namespace nb
