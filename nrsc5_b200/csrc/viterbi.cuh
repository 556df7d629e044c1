// Warp-synchronous tail-biting Viterbi decoder, K=7, rate 1/3 mother code
// g = (0133, 0171, 0165).  Replaces reference src/conv_dec.c:359-453 with the
// SSE path's arithmetic (src/conv_sse.h:56-66,233-315): int16 saturating
// add-compare-select, strict-greater tie-break, min-normalisation when
// step % 79 == 0, first-maximum end state, 32-step pre/post-roll.
//
// One warp decodes one frame.  Lane l owns butterfly l: it reads old states
// 2l and 2l+1 and produces new states l and l+32 (new bit enters at bit 5).
#pragma once
#include "common.cuh"

namespace nb {

constexpr int VIT_NORM = 32767 / (3 * 127) - 7;     // 79, reference src/conv_dec.c:370

__device__ __forceinline__ int sat16(int v) { return max(-32768, min(32767, v)); }

struct VitWarp {
    int lo, hi;          // path metrics of old states 2l, 2l+1
    int c0, c1, c2;      // expected outputs (+1/-1) of branch (state 2l, input 0)

    __device__ __forceinline__ void init(int lane)
    {
        lo = hi = 0;
        const unsigned reg = (unsigned)lane << 1;
        c0 = (__popc(reg & 0133u) & 1) ? 1 : -1;
        c1 = (__popc(reg & 0171u) & 1) ? 1 : -1;
        c2 = (__popc(reg & 0165u) & 1) ? 1 : -1;
    }

    // one trellis step; returns the survivor bits of new states 0..31 (x) and 32..63 (y)
    __device__ __forceinline__ uint2 step(int s0, int s1, int s2, bool norm, int lane)
    {
        const int m = s0 * c0 + s1 * c1 + s2 * c2;
        const int a0 = sat16(lo + m), a1 = sat16(hi - m);
        const int b0 = sat16(lo - m), b1 = sat16(hi + m);
        const bool da = !(a0 > a1), db = !(b0 > b1);      // 1 = survivor comes from the odd state
        int n0 = da ? a1 : a0, n1 = db ? b1 : b0;
        uint2 dec;
        dec.x = __ballot_sync(0xffffffffu, da);
        dec.y = __ballot_sync(0xffffffffu, db);
        if (norm) {
            int mn = min(n0, n1);
#pragma unroll
            for (int o = 16; o; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            n0 = sat16(n0 - mn);
            n1 = sat16(n1 - mn);
        }
        // lane j next needs new[2j], new[2j+1]: both live in lanes (2j)&31 and (2j+1)&31,
        // in n0 when 2j < 32 and in n1 otherwise
        const int packed = (n0 & 0xffff) | (n1 << 16);
        const int src = (2 * lane) & 31;
        const int p0 = __shfl_sync(0xffffffffu, packed, src);
        const int p1 = __shfl_sync(0xffffffffu, packed, src + 1);
        if (lane < 16) { lo = (short)(p0 & 0xffff); hi = (short)(p1 & 0xffff); }
        else           { lo = p0 >> 16;             hi = p1 >> 16; }
        return dec;
    }

    // first maximum over the 64 states as they stand after the last step
    __device__ __forceinline__ int best_state(int lane)
    {
        // after step(), lane j holds states 2j (lo) and 2j+1 (hi)
        int v = lo, idx = 2 * lane;
        if (hi > v) { v = hi; idx = 2 * lane + 1; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            int ov = __shfl_xor_sync(0xffffffffu, v, o);
            int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        return idx;
    }
};

// Forward pass over len+64 steps; `in` holds 3*len soft values (0 = punctured).
// Survivor words go to dec[step]. Returns the traceback start state (all lanes).
template <typename DecT>
__device__ inline int viterbi_forward(const int8_t *__restrict__ in, int len, DecT *dec, int lane)
{
    VitWarp vw;
    vw.init(lane);
    const int steps = len + 64;
    int j0 = len - 32;                                  // input index of step s is (s + len - 32) mod len
    for (int sbase = 0; sbase < steps; sbase += 32) {
        // each lane fetches the three soft values of step sbase+lane
        int mine = 0;
        {
            int s = sbase + lane;
            if (s < steps) {
                int j = j0 + lane;
                if (j >= len) j -= len;
                if (j >= len) j -= len;
                const int8_t *q = in + 3 * j;
                mine = (uint8_t)q[0] | ((uint8_t)q[1] << 8) | ((uint8_t)q[2] << 16);
            }
        }
        const int n = min(32, steps - sbase);
        for (int k = 0; k < n; k++) {
            const int w = __shfl_sync(0xffffffffu, mine, k);
            const int s = sbase + k;
            uint2 d = vw.step((int)(int8_t)(w & 0xff), (int)(int8_t)((w >> 8) & 0xff), (int)(int8_t)((w >> 16) & 0xff),
                              (s % VIT_NORM) == 0, lane);
            if (lane == 0) dec[s] = d;
        }
        j0 += 32;
        if (j0 >= len) j0 -= len;
    }
    return vw.best_state(lane);
}

// survivor-walk helpers: state after step s -> state after step s-1
__device__ __forceinline__ int vit_prev(int state, uint2 d)
{
    const unsigned bit = state < 32 ? (d.x >> state) & 1u : (d.y >> (state - 32)) & 1u;
    return ((state << 1) & 62) | (int)bit;
}

}  // namespace nb
