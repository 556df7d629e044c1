// Reed-Solomon (255,247) over GF(2^8)/0x11d with roots alpha^1..alpha^8 - the code the reference sets up at
// src/frame.c:747 (init_rs_char(8, 0x11d, 1, 1, 8, 0)) and applies to the 96-byte L2 audio-PDU header
// (src/frame.c:158-179 -> decode_rs_char, src/rs_decode.c:16) - decoded by ONE WARP per codeword:
//
//   syndromes   every lane folds its own bytes of the codeword into all eight syndromes (log-domain multiply),
//               one packed XOR reduction across the warp; all-zero syndromes (every clean header) end here
//   locator     inversion-free Berlekamp-Massey with one coefficient per lane: the discrepancy is a warp reduction,
//               the update one shuffle; the result is the textbook (Blahut) locator up to a non-zero factor, so its
//               degree, roots and the error values are those the reference's decoder arrives at - including what it
//               makes of words with more than four errors (a locator of degree 5..8 that happens to split is
//               "corrected" there as well; > 10^5 such words are compared in tests/test_oracle.py / test_gpu_stages.py)
//   roots       Chien search with eight positions per lane; the word is rejected unless the locator has exactly
//               deg(lambda) roots among the 255 positions
//   values      Forney (fcr = 1: value = omega(X^-1) / lambda'(X^-1)) by the lanes that found the roots
//
// Codeword position p (0..254) is the coefficient of x^(254-p); an error at p shows as the locator root alpha^(p+1).
#pragma once
#include "common.cuh"

namespace nb {

__constant__ uint8_t c_gf_exp[256];          // alpha^i (i < 255); uploaded by the host (engine.cu: upload_tables)
__constant__ uint8_t c_gf_log[256];

// log / antilog tables in shared memory (the lanes of a warp look up unrelated entries)
struct GfTab {
    uint8_t exp[512];                        // alpha^(i mod 255), i < 510: sums of two logs need no reduction
    uint8_t log[256];                        // log[0] is never used
};

// cooperative fill by `nt` threads; the caller synchronises before the first decode
__device__ inline void gf_tab_load(GfTab &g, int t, int nt)
{
    for (int i = t; i < 510; i += nt) g.exp[i] = c_gf_exp[i < 255 ? i : i - 255];
    for (int i = t; i < 256; i += nt) g.log[i] = c_gf_log[i];
    if (t == 0) { g.exp[510] = 0; g.exp[511] = 0; }
}

__device__ __forceinline__ unsigned gf_mul(const GfTab &g, unsigned a, unsigned b)
{
    return (a && b) ? g.exp[g.log[a] + g.log[b]] : 0u;
}

__device__ __forceinline__ unsigned warp_xor(unsigned v)
{
#if defined(NB_EMU)
    for (int o = 16; o; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
#else
    return __reduce_xor_sync(0xffffffffu, v);
#endif
}

struct RsFix {                               // what a lane found: up to eight of its positions are error locations
    int count;                               // corrected symbols (the locator's degree), 0 = clean, -1 = uncorrectable
    unsigned mask;                           // bit m: position lane + 32 m is an error location
    uint8_t val[8];                          // its error value (may be zero)
};

// One warp decodes one codeword.  rd(p) returns the byte at codeword position p; positions below `first` are known
// to be zero (the shortened header code: 159).  All 32 lanes call this together; every lane gets the same count.
template <typename Rd>
__device__ inline RsFix rs8_locate(const GfTab &g, Rd rd, int first, int lane)
{
    RsFix out;
    out.count = 0;
    out.mask = 0;
    // ---- syndromes S_i = sum_p c[p] * alpha^((i+1)(254-p)), i = 0..7, four per 32-bit word
    unsigned s_lo = 0, s_hi = 0;
    for (int p = first + lane; p < 255; p += 32) {
        const unsigned c = rd(p);
        if (!c) continue;
        const unsigned e = 254u - (unsigned)p;                        // < 255
        unsigned x = g.log[c];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x += e;
            if (x >= 255u) x -= 255u;
            const unsigned term = g.exp[x];
            if (i < 4) s_lo ^= term << (8 * i);
            else s_hi ^= term << (8 * (i - 4));
        }
    }
    s_lo = warp_xor(s_lo);
    s_hi = warp_xor(s_hi);
    if ((s_lo | s_hi) == 0) return out;
    const unsigned long long S = ((unsigned long long)s_hi << 32) | s_lo;      // byte i = S_i
    auto syn = [&](int i) -> unsigned { return (unsigned)(S >> (8 * i)) & 0xffu; };

    // ---- locator: lane j holds lambda_j and b_j (j = 0..8), inversion-free update
    //      lambda <- gamma * lambda + d * x * b;  on a length change b <- lambda (old), gamma <- d;  else b <- x * b
    unsigned lam = lane == 0, bb = lane == 0, gamma = 1;
    int L = 0;
#pragma unroll 1
    for (int r = 1; r <= 8; r++) {
        const unsigned sv = lane < r ? syn(r - 1 - lane) : 0u;
        const unsigned d = warp_xor(gf_mul(g, lam, sv));
        unsigned bsh = __shfl_up_sync(0xffffffffu, bb, 1);
        if (lane == 0 || lane > 8) bsh = 0;
        const unsigned nl = gf_mul(g, gamma, lam) ^ gf_mul(g, d, bsh);
        if (d != 0 && 2 * L <= r - 1) {
            bb = lam;
            L = r - L;
            gamma = d;
        } else {
            bb = bsh;
        }
        lam = lane <= 8 ? nl : 0u;
    }
    const unsigned nz = __ballot_sync(0xffffffffu, lam != 0);
    const int deg = 31 - __clz((int)nz);                              // lambda_0 = product of the gammas, never zero
    // every lane needs the whole locator (as logs; 255 = zero coefficient) and, later, omega
    unsigned ll[9];
#pragma unroll
    for (int j = 0; j <= 8; j++) {
        const unsigned v = __shfl_sync(0xffffffffu, lam, j);
        ll[j] = v ? g.log[v] : 255u;
    }
    // ---- roots: lambda(alpha^(p+1)) == 0 at position p
    unsigned mine = 0;
#pragma unroll 1
    for (int m = 0; m < 8; m++) {
        const int p = lane + 32 * m;
        if (p >= 255) break;
        const unsigned x = (unsigned)(p + 1) % 255u;
        unsigned acc = 0, step = 0;                                    // step = j * x mod 255
#pragma unroll
        for (int j = 0; j <= 8; j++) {
            if (ll[j] != 255u) acc ^= g.exp[ll[j] + step];
            step += x;
            if (step >= 255u) step -= 255u;
        }
        if (acc == 0) mine |= 1u << m;
    }
    int nroots = __popc(mine);
#pragma unroll
    for (int o = 16; o; o >>= 1) nroots += __shfl_xor_sync(0xffffffffu, nroots, o);
    if (nroots != deg) {
        out.count = -1;
        return out;
    }
    // ---- omega(x) = S(x) lambda(x) mod x^8: lane k < 8 computes omega_k, then every lane gets all of them
    unsigned om = 0;
    if (lane < 8) {
        for (int j = 0; j <= lane; j++) {
            const unsigned sj = syn(lane - j);
            if (sj && ll[j] != 255u) om ^= g.exp[g.log[sj] + ll[j]];
        }
    }
    unsigned lo[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned v = __shfl_sync(0xffffffffu, om, k);
        lo[k] = v ? g.log[v] : 255u;
    }
    // ---- error values at this lane's roots: omega(X^-1) / lambda'(X^-1) with X^-1 = alpha^(p+1)
#pragma unroll 1
    for (int m = 0; m < 8; m++) {
        out.val[m] = 0;
        if (!((mine >> m) & 1u)) continue;
        const unsigned x = (unsigned)(lane + 32 * m + 1) % 255u;
        unsigned num = 0, den = 0, step = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {                                  // step = k * x mod 255
            if (lo[k] != 255u) num ^= g.exp[lo[k] + step];
            // lambda'(x) = sum over odd j of lambda_j x^(j-1): term j = k + 1 for even k
            if (!(k & 1) && ll[k + 1] != 255u) den ^= g.exp[ll[k + 1] + step];
            step += x;
            if (step >= 255u) step -= 255u;
        }
        // (den != 0: deg distinct roots make every root a simple one)
        out.val[m] = (num && den) ? g.exp[g.log[num] + 255u - g.log[den]] : 0;
    }
    out.mask = mine;
    out.count = deg;
    return out;
}

// decode_rs_char(rs, data, NULL, 0) on data[255] in place (global or shared memory), by one warp: returns the number
// of corrected symbols or -1 (data untouched)
__device__ inline int rs8_decode_warp(const GfTab &g, uint8_t *data, int lane)
{
    const RsFix f = rs8_locate(g, [&](int p) -> unsigned { return data[p]; }, 0, lane);
    if (f.count > 0)
        for (int m = 0; m < 8; m++)
            if ((f.mask >> m) & 1u) data[lane + 32 * m] ^= f.val[m];
    __syncwarp();
    return f.count;
}

// fix_header (reference src/frame.c:158-179) by one warp: buf[96] is the header as it lies in the PDU, i.e. codeword
// position 254 - i holds buf[i] and positions 0..158 are zero.  Returns 1 and corrects buf in place when the word
// decodes and no correction falls into the zero padding; returns 0 and leaves buf alone otherwise.
__device__ inline int rs8_fix_header_warp(const GfTab &g, uint8_t *buf, int lane)
{
    const RsFix f = rs8_locate(g, [&](int p) -> unsigned { return buf[254 - p]; }, 159, lane);
    if (f.count < 0) return 0;
    if (f.count == 0) return 1;
    int bad = 0;
    for (int m = 0; m < 8; m++)
        if (((f.mask >> m) & 1u) && f.val[m] && lane + 32 * m < 159) bad = 1;      // a padding byte would become non-zero
    if (__any_sync(0xffffffffu, bad)) return 0;
    for (int m = 0; m < 8; m++)
        if (((f.mask >> m) & 1u) && lane + 32 * m >= 159) buf[254 - (lane + 32 * m)] ^= f.val[m];
    __syncwarp();
    return 1;
}

}  // namespace nb
