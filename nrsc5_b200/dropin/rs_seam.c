/* Reed-Solomon (255,247) over GF(2^8)/0x11d, fcr = 1, prim = 1, 8 roots, behind the three entry points
 * the reference's L2 (frame.c:167,747,753) links against: init_rs_char / free_rs_char / decode_rs_char
 * (replaces reference src/rs_init.c, src/rs_decode.c for the one code the path uses).  Host-side twin of
 * csrc/rs.cuh: syndromes, Berlekamp-Massey, Chien search and Forney in index form. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NN 255
#define R 8
#define A0 255

struct rs255 {
    uint8_t exp[256], log[256];
};

static unsigned mod255(unsigned x)
{
    while (x >= 255) { x -= 255; x = (x >> 8) + (x & 255); }
    return x;
}

void *init_rs_char(unsigned int symsize, unsigned int gfpoly, unsigned int fcr, unsigned int prim, unsigned int nroots)
{
    if (symsize != 8 || gfpoly != 0x11d || fcr != 1 || prim != 1 || nroots != R) return NULL;
    struct rs255 *rs = calloc(1, sizeof(*rs));
    if (!rs) return NULL;
    unsigned v = 1;
    rs->log[0] = A0;
    rs->exp[255] = 0;
    for (int i = 0; i < 255; i++) {
        rs->exp[i] = (uint8_t)v;
        rs->log[v] = (uint8_t)i;
        v <<= 1;
        if (v & 0x100) v ^= 0x11d;
    }
    return rs;
}

void free_rs_char(void *p) { free(p); }

/* data[255] corrected in place; returns the number of corrected symbols, -1 if uncorrectable.
 * Erasures are not used by the caller (frame.c:167 passes NULL, 0). */
int decode_rs_char(void *p, unsigned char *data, int *eras_pos, int no_eras)
{
    const struct rs255 *rs = p;
    const uint8_t *EX = rs->exp, *LG = rs->log;
    uint8_t s[R], lambda[R + 1], b[R + 1], t[R + 1], omega[R + 1], reg[R + 1], root[R], loc[R];
    (void)eras_pos;
    if (no_eras != 0) return -1;
    for (int i = 0; i < R; i++) s[i] = data[0];
    for (int j = 1; j < NN; j++) {
        const uint8_t dj = data[j];
        for (int i = 0; i < R; i++)
            s[i] = s[i] == 0 ? dj : (uint8_t)(dj ^ EX[mod255(LG[s[i]] + 1 + i)]);
    }
    unsigned any = 0;
    for (int i = 0; i < R; i++) { any |= s[i]; s[i] = LG[s[i]]; }
    if (!any) return 0;

    memset(lambda, 0, sizeof(lambda));
    lambda[0] = 1;
    for (int i = 0; i <= R; i++) b[i] = LG[lambda[i]];
    int el = 0;
    for (int r = 1; r <= R; r++) {
        uint8_t disc = 0;
        for (int i = 0; i < r; i++)
            if (lambda[i] != 0 && s[r - i - 1] != A0)
                disc ^= EX[mod255(LG[lambda[i]] + s[r - i - 1])];
        disc = LG[disc];
        if (disc == A0) {
            for (int i = R; i > 0; i--) b[i] = b[i - 1];
            b[0] = A0;
        } else {
            t[0] = lambda[0];
            for (int i = 0; i < R; i++)
                t[i + 1] = b[i] != A0 ? (uint8_t)(lambda[i + 1] ^ EX[mod255(disc + b[i])]) : lambda[i + 1];
            if (2 * el <= r - 1) {
                el = r - el;
                for (int i = 0; i <= R; i++)
                    b[i] = lambda[i] == 0 ? (uint8_t)A0 : (uint8_t)mod255(LG[lambda[i]] - disc + NN);
            } else {
                for (int i = R; i > 0; i--) b[i] = b[i - 1];
                b[0] = A0;
            }
            memcpy(lambda, t, sizeof(lambda));
        }
    }
    int deg = 0;
    for (int i = 0; i <= R; i++) {
        lambda[i] = LG[lambda[i]];
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
                q ^= EX[reg[j]];
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
                tmp ^= EX[mod255(s[i - j] + lambda[j])];
        if (tmp != 0) dego = i;
        omega[i] = LG[tmp];
    }
    omega[R] = A0;
    for (int j = count - 1; j >= 0; j--) {
        uint8_t num1 = 0, den = 0;
        for (int i = dego; i >= 0; i--)
            if (omega[i] != A0)
                num1 ^= EX[mod255(omega[i] + i * root[j])];
        const uint8_t num2 = EX[mod255(NN)];             /* root^(fcr-1) = 1 */
        for (int i = (deg < R - 1 ? deg : R - 1) & ~1; i >= 0; i -= 2)
            if (lambda[i + 1] != A0)
                den ^= EX[mod255(lambda[i + 1] + i * root[j])];
        if (den == 0) return -1;
        if (num1 != 0)
            data[loc[j]] ^= EX[mod255(LG[num1] + LG[num2] + NN - LG[den])];
    }
    return count;
}
