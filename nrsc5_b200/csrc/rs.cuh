// Reed-Solomon (255,247) decoder over GF(256)/0x11d, fcr = 1, prim = 1, as the
// reference configures it (reference src/frame.c:747 -> src/rs_init.c:31-133)
// and runs it on the 96-byte L2 audio-PDU header (src/frame.c:158-179,
// src/rs_decode.c:16-210).  Syndromes, Berlekamp-Massey, Chien search and
// Forney in index form; one thread decodes one block.
#pragma once
#include "common.cuh"

namespace nb {

__constant__ uint8_t c_gf_exp[256];
__constant__ uint8_t c_gf_log[256];

__device__ __forceinline__ unsigned mod255(unsigned x)
{
    while (x >= 255) { x -= 255; x = (x >> 8) + (x & 255); }
    return x;
}

// data[255] in place; returns number of corrected symbols or -1.  lead_zeros: the caller knows data[0 .. lead_zeros)
// to be zero (the shortened code of the 96-byte header: 159); the syndromes' Horner evaluation over that prefix
// leaves zeros, so it starts behind it.
__device__ inline int rs_decode_255_247(uint8_t *data, int lead_zeros = 0)
{
    constexpr int R = 8, NN = 255, A0 = 255;
    uint8_t s[R], lambda[R + 1], b[R + 1], t[R + 1], omega[R + 1], reg[R + 1], root[R], loc[R];
    for (int i = 0; i < R; i++) s[i] = data[lead_zeros];
    for (int j = lead_zeros + 1; j < NN; j++) {
        const uint8_t dj = data[j];
        for (int i = 0; i < R; i++)
            s[i] = s[i] == 0 ? dj : (uint8_t)(dj ^ c_gf_exp[mod255(c_gf_log[s[i]] + 1 + i)]);
    }
    unsigned any = 0;
    for (int i = 0; i < R; i++) { any |= s[i]; s[i] = c_gf_log[s[i]]; }
    if (!any) return 0;

    for (int i = 0; i <= R; i++) lambda[i] = 0;
    lambda[0] = 1;
    for (int i = 0; i <= R; i++) b[i] = c_gf_log[lambda[i]];
    int el = 0;
    for (int r = 1; r <= R; r++) {
        uint8_t disc = 0;
        for (int i = 0; i < r; i++)
            if (lambda[i] != 0 && s[r - i - 1] != A0)
                disc ^= c_gf_exp[mod255(c_gf_log[lambda[i]] + s[r - i - 1])];
        disc = c_gf_log[disc];
        if (disc == A0) {
            for (int i = R; i > 0; i--) b[i] = b[i - 1];
            b[0] = A0;
        } else {
            t[0] = lambda[0];
            for (int i = 0; i < R; i++)
                t[i + 1] = b[i] != A0 ? (uint8_t)(lambda[i + 1] ^ c_gf_exp[mod255(disc + b[i])]) : lambda[i + 1];
            if (2 * el <= r - 1) {
                el = r - el;
                for (int i = 0; i <= R; i++)
                    b[i] = lambda[i] == 0 ? (uint8_t)A0 : (uint8_t)mod255(c_gf_log[lambda[i]] - disc + NN);
            } else {
                for (int i = R; i > 0; i--) b[i] = b[i - 1];
                b[0] = A0;
            }
            for (int i = 0; i <= R; i++) lambda[i] = t[i];
        }
    }
    int deg = 0;
    for (int i = 0; i <= R; i++) {
        lambda[i] = c_gf_log[lambda[i]];
        if (lambda[i] != A0) deg = i;
    }
    for (int i = 1; i <= R; i++) reg[i] = lambda[i];
    int count = 0;
    unsigned k = 0;
    for (unsigned i = 1; i <= NN; i++, k = mod255(k + 1)) {
        uint8_t q = 1;
        for (int j = deg; j > 0; j--)
            if (reg[j] != A0) {
                reg[j] = (uint8_t)mod255(reg[j] + j);
                q ^= c_gf_exp[reg[j]];
            }
        if (q != 0) continue;
        root[count] = (uint8_t)i;
        loc[count] = (uint8_t)k;
        if (++count == deg) break;
    }
    if (deg != count) return -1;
    int dego = 0;
    for (int i = 0; i < R; i++) {
        uint8_t tmp = 0;
        for (int j = deg < i ? deg : i; j >= 0; j--)
            if (s[i - j] != A0 && lambda[j] != A0)
                tmp ^= c_gf_exp[mod255(s[i - j] + lambda[j])];
        if (tmp != 0) dego = i;
        omega[i] = c_gf_log[tmp];
    }
    omega[R] = A0;
    for (int j = count - 1; j >= 0; j--) {
        uint8_t num1 = 0, den = 0;
        for (int i = dego; i >= 0; i--)
            if (omega[i] != A0)
                num1 ^= c_gf_exp[mod255(omega[i] + i * root[j])];
        const uint8_t num2 = c_gf_exp[mod255(NN)];                   // root^(fcr-1) = 1
        for (int i = (deg < R - 1 ? deg : R - 1) & ~1; i >= 0; i -= 2)
            if (lambda[i + 1] != A0)
                den ^= c_gf_exp[mod255(lambda[i + 1] + i * root[j])];
        if (den == 0) return -1;
        if (num1 != 0)
            data[loc[j]] ^= c_gf_exp[mod255(c_gf_log[num1] + c_gf_log[num2] + NN - c_gf_log[den])];
    }
    return count;
}

// reference src/frame.c:158-179: returns 1 when the 96-byte header decodes
__device__ inline int fix_header_96(uint8_t *buf, uint8_t *blk /* 255 bytes scratch */)
{
    for (int i = 0; i < 159; i++) blk[i] = 0;
    for (int i = 0; i < 96; i++) blk[254 - i] = buf[i];
    if (rs_decode_255_247(blk, 159) == -1) return 0;
    for (int i = 0; i < 159; i++)
        if (blk[i] != 0) return 0;
    for (int i = 0; i < 96; i++) buf[i] = blk[254 - i];
    return 1;
}

}  // namespace nb
