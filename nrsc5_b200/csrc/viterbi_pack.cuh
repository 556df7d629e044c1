// Chunk-parallel, exact, tail-biting Viterbi (K=7, rate 1/3) for long frames.
//
// The reference decodes a frame with one sequential pass of len+64 trellis
// steps (reference src/conv_dec.c:402-427).  Here the steps are cut into chunks
// of CH_LEN; every chunk is decoded by half a warp that first replays CH_WARM
// warm-up steps from all-zero metrics.  A chunk's result is *accepted* only if
// its path-metric vector at the chunk start equals (up to a common constant) the
// true vector handed over by the previous chunk; otherwise that chunk is
// recomputed from the true vector (k_p1_post).  Equal metric vectors give equal
// add-compare-select decisions from there on, so the accepted decisions are
// exactly those of the sequential pass — the speculation only ever costs time.
//
// Arithmetic: int16 path metrics as in the reference's SSE kernel
// (src/conv_sse.h:56-66).  The fast path uses wrapping packed adds (VIADD.16x2 /
// VIMNMX.S16x2) and is only taken when the frame provably cannot saturate
// (viterbi_cannot_saturate); SAT=true reproduces saturating arithmetic.
//
// Lane layout (16 lanes per chunk, lane l): E = (pm[2l], pm[2l+32]),
// O = (pm[2l+1], pm[2l+33]) as packed s16x2; lane l owns butterflies l and l+16.
#pragma once
#include "common.cuh"

namespace nb {

constexpr int CH_LEN = 1024;
constexpr int CH_WARM = 256;
constexpr int VITC_NORM = 32767 / (3 * 127) - 7;                 // 79
constexpr int VITC_HEAD_STEPS = 128;

__device__ __forceinline__ unsigned vadd16(unsigned a, unsigned b)
{
    unsigned r;
#if defined(NB_EMU)
    r = emu_add_s16x2(a, b);
#else
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
    return r;
}
__device__ __forceinline__ unsigned vmax16(unsigned a, unsigned b)
{
    unsigned r;
#if defined(NB_EMU)
    r = emu_max_s16x2(a, b);
#else
    asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
    return r;
}
__device__ __forceinline__ unsigned vmin16(unsigned a, unsigned b)
{
    unsigned r;
#if defined(NB_EMU)
    r = emu_min_s16x2(a, b);
#else
    asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
#endif
    return r;
}
__device__ __forceinline__ unsigned vneg16(unsigned a) { return vadd16(~a, 0x00010001u); }

// decision-bit storage: one uint2 per (16-step group, lane): x = sign history of
// the "A" outputs (new states l | l+16 in the low | high half), y = "B" outputs
// (new states l+32 | l+48); bit k of a half = step (group*16 + k); bit set means
// the survivor comes from the ODD predecessor (2b+1).
__device__ __forceinline__ int vitc_prev(int state, const uint2 *__restrict__ dec, int step)
{
    const int l = state & 15, q = state >> 4;
    const uint2 w = dec[(size_t)(step >> 4) * 16 + l];
    const unsigned word = (q & 2) ? w.y : w.x;
    const unsigned bit = (word >> (((q & 1) << 4) + (step & 15))) & 1u;
    return ((state << 1) & 62) | (int)bit;
}

template <bool SAT>
struct VitHalf {
    unsigned E, O;          // packed metrics
    unsigned accA, accB;    // decision sign history
    int c0, c1;             // dp4a constants of butterflies l and l+16

    __device__ __forceinline__ void init(int l)
    {
        E = O = 0;
        accA = accB = 0;
        auto mk = [](int b) {
            const unsigned reg = (unsigned)b << 1;
            int k0 = (__popc(reg & 0133u) & 1) ? 1 : -1;
            int k1 = (__popc(reg & 0171u) & 1) ? 1 : -1;
            int k2 = (__popc(reg & 0165u) & 1) ? 1 : -1;
            return (int)((unsigned)(k0 & 0xff) | ((unsigned)(k1 & 0xff) << 8) | ((unsigned)(k2 & 0xff) << 16));
        };
        c0 = mk(l);
        c1 = mk(l + 16);
    }

    // one trellis step for both chunks of the warp; w = soft bytes (s0 | s1<<8 | s2<<16)
    // `norm` may differ between the two half-warps (they sit at different steps), so the
    // normalisation's reduction only synchronises the lanes of this half (hmask)
    __device__ __forceinline__ void step(int w, bool norm, bool active, int l, unsigned hmask)
    {
        const int m0 = __dp4a(w, c0, 0), m1 = __dp4a(w, c1, 0);
        const unsigned M = __byte_perm((unsigned)m0, (unsigned)m1, 0x5410);
        const unsigned Mn = vneg16(M);
        unsigned X1, Y1, X2, Y2, N0, N1, dA, dB;
        if (SAT) {
            X1 = __vaddss2(E, M); Y1 = __vaddss2(O, Mn); X2 = __vaddss2(E, Mn); Y2 = __vaddss2(O, M);
            N0 = vmax16(X1, Y1);
            N1 = vmax16(X2, Y2);
            dA = ~__vcmpgts2(X1, Y1);            // 0xffff where the odd predecessor wins (ties included)
            dB = ~__vcmpgts2(X2, Y2);
        } else {
            X1 = vadd16(E, M); Y1 = vadd16(O, Mn); X2 = vadd16(E, Mn); Y2 = vadd16(O, M);
            N0 = vmax16(X1, Y1);
            N1 = vmax16(X2, Y2);
            dA = vadd16(X1, ~Y1);                // X - Y - 1 < 0  <=>  X <= Y  <=> odd wins
            dB = vadd16(X2, ~Y2);
        }
        accA = ((accA >> 1) & 0x7fff7fffu) | (dA & 0x80008000u);
        accB = ((accB >> 1) & 0x7fff7fffu) | (dB & 0x80008000u);
        if (norm) {                              // subtract the minimum over the 64 states
            unsigned t = vmin16(N0, N1);
            t = vmin16(t, __byte_perm(t, t, 0x1032));
#pragma unroll
            for (int o = 8; o; o >>= 1) t = vmin16(t, __shfl_xor_sync(hmask, t, o, 16));
            if (SAT) { N0 = __vsubss2(N0, t); N1 = __vsubss2(N1, t); }
            else { const unsigned tn = vneg16(t); N0 = vadd16(N0, tn); N1 = vadd16(N1, tn); }
        }
        const int a = (2 * l) & 15;
        const unsigned n0a = __shfl_sync(0xffffffffu, N0, a, 16), n0b = __shfl_sync(0xffffffffu, N0, a + 1, 16);
        const unsigned n1a = __shfl_sync(0xffffffffu, N1, a, 16), n1b = __shfl_sync(0xffffffffu, N1, a + 1, 16);
        const unsigned sel = (l & 8) ? 0x7632u : 0x5410u;
        if (active) {                            // steps past the end of the frame leave the metrics alone
            E = __byte_perm(n0a, n1a, sel);
            O = __byte_perm(n0b, n1b, sel);
        }
    }
};

// Runs steps [s_from, s_to) of the frame for the two chunks held by this warp
// (one per half-warp; the halves may work on different step ranges as long as
// the ranges have the same length).  Steps outside [0, total) consume zero soft
// values.  Decisions are stored for steps >= s_store.  `vin` = 3*len soft
// values of this half's frame; dec = that frame's decision array.
template <bool SAT>
__device__ inline void vitc_run(VitHalf<SAT> &vh, const int8_t *__restrict__ vin, int len, int total,
                                int s_from, int nsteps, int s_store, uint2 *__restrict__ dec, bool store_ok, int l,
                                uint2 *head = nullptr)
{
    const unsigned hmask = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
    for (int base = 0; base < nsteps; base += 16) {
        // lane l of each half fetches the soft triple of step s_from + base + l
        int mine = 0;
        {
            const int s = s_from + base + l;
            if (s >= 0 && s < total) {
                int j = s + len - 32;
                while (j >= len) j -= len;
                const int8_t *q = vin + 3 * j;
                mine = (uint8_t)q[0] | ((uint8_t)q[1] << 8) | ((uint8_t)q[2] << 16);
            }
        }
        const int n = min(16, nsteps - base);
        for (int k = 0; k < n; k++) {
            const int w = __shfl_sync(0xffffffffu, mine, k, 16);
            const int s = s_from + base + k;
            vh.step(w, s >= 0 && (s % VITC_NORM) == 0, s < total, l, hmask);
        }
        const int g0 = s_from + base;                    // groups are 16-aligned by construction
        if (store_ok && n == 16 && g0 >= s_store && g0 < total)
            dec[(size_t)(g0 >> 4) * 16 + l] = make_uint2(vh.accA, vh.accB);
        if (head && g0 >= s_store && g0 < s_store + VITC_HEAD_STEPS)      // first steps of the chunk, kept on chip
            head[((g0 - s_store) >> 4) * 16 + l] = make_uint2(vh.accA, vh.accB);
    }
}

__device__ __forceinline__ int vitc_prev_head(int state, const uint2 *hd, int q)
{
    const int ll = state & 15, qq = state >> 4;
    const uint2 w = hd[(q >> 4) * 16 + ll];
    const unsigned word = (qq & 2) ? w.y : w.x;
    return ((state << 1) & 62) | (int)((word >> (((qq & 1) << 4) + (q & 15))) & 1u);
}


}  // namespace nb
