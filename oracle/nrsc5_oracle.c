/* nrsc5_oracle.c — CPU restatement of the reference FM receive chain.
 * TEST INFRASTRUCTURE ONLY (see nrsc5_oracle.h).  Every stage cites the
 * reference file:line it restates.  Floating-point expressions keep the
 * reference's evaluation order and types (float vs double promotion) so that
 * the soft bits are bit-identical to the reference built with the same FFT
 * (oracle/shim/fftshim.c stands in for FFTW in both builds).
 *
 * Not restated here: the AM (MA1/MA3) chain; L2 and above (only the one L2
 * predicate that feeds back into sync, reference src/frame.c:527-541).
 */
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "nrsc5_oracle.h"
#include "shim/fftw3.h"

typedef float complex cf;

#define NFFT 2048
#define NCP 112
#define NSYM (NFFT + NCP)           /* 2160 samples per OFDM symbol            */
#define BLK 32                      /* symbols per L1 block                    */
#define NACQ (NSYM * (BLK + 1))     /* 71280: acquisition window               */
#define LB0 (NFFT / 2 - 546)        /* first lower-sideband bin, 478           */
#define UB1 (NFFT / 2 + 546)        /* last upper-sideband bin, 1570           */
#define PW 19                       /* bins per partition                      */
#define MAXPART 14
#define PM_BLOCK 23040
#define P1_LEN 146176
#define P1_ENC (P1_LEN * 5 / 2)
#define PIDS_LEN 80
#define P3_LEN_MAX 4608
#define ST_NONE 0
#define ST_COARSE 1
#define ST_FINE 2

/* ------------------------------------------------------------------------ */
/* log                                                                      */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint8_t *p;
    size_t len, cap;
} olog_t;

static void olog_put(olog_t *l, uint32_t type, const void *a, size_t alen, const void *b, size_t blen)
{
    size_t plen = alen + blen, need = 8 + ((plen + 3) & ~(size_t)3);
    if (l->len + need > l->cap) {
        size_t nc = l->cap ? l->cap * 2 : (1u << 20);
        while (nc < l->len + need) nc *= 2;
        l->p = (uint8_t *)realloc(l->p, nc);
        l->cap = nc;
    }
    uint32_t hdr[2] = { type, (uint32_t)plen };
    memcpy(l->p + l->len, hdr, 8);
    if (alen) memcpy(l->p + l->len + 8, a, alen);
    if (blen) memcpy(l->p + l->len + 8 + alen, b, blen);
    memset(l->p + l->len + 8 + plen, 0, need - 8 - plen);
    l->len += need;
}

/* ------------------------------------------------------------------------ */
/* stream object                                                            */
/* ------------------------------------------------------------------------ */
typedef struct {
    int8_t pair[2 * 144 * BLK];           /* two blocks of PX soft bits      */
    int8_t ring[P3_LEN_MAX * 32];         /* convolutional interleaver store */
    unsigned wr;                          /* write position in ring          */
    unsigned rd[4];                       /* per-partition read counters     */
    int primed, started;
} px_deint_t;

struct orc {
    olog_t log;
    int want_soft, want_blocks;

    /* front-end decimator: last 15 Q15 samples (reference src/firdecim_q15.c:58-67) */
    int16_t hb_r[15], hb_i[15];
    int16_t hb_tap[4];

    /* acquisition window and its coarse band-pass filter */
    int16_t win_r[NACQ], win_i[NACQ];
    unsigned fill;
    long long start_index;                /* decimated index of win[0]       */
    int16_t bp_r[32], bp_i[32];           /* last 32 samples fed to the FIR  */
    int16_t bp_tap[32];
    cf tbuf[NACQ];
    cf sums[NSYM];
    float shape[NSYM];
    float prev_angle;
    cf phase;
    int keep_extra, cfo;
    int state;
    fftwf_complex *fin, *fout;
    fftwf_plan plan;

    /* sync */
    cf bins[NFFT][BLK];
    float phs[NFFT][BLK];
    float cfreq[NFFT], cphase[NFFT];
    float alpha, beta;
    unsigned sym;
    int psmi, cfo_wait, samperr, mer_cnt;
    unsigned bc;
    float angle, err_lb, err_ub;

    /* decode */
    int8_t pm[PM_BLOCK * 16];
    int started_pm;
    px_deint_t px1, px2;
    int8_t vit_p1[P1_LEN * 3];
    uint8_t out_p1[P1_LEN];
};

static const int compat_mode[64] = {
    0, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5, 6, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5,
    6, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5, 6, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5
};

/* ------------------------------------------------------------------------ */
/* a1/a2: cu8 -> Q15 -> halfband /2   (reference src/input.c:52-94,          */
/*        src/firdecim_q15.c:137-165, src/defines.h:93)                      */
/* ------------------------------------------------------------------------ */
static void hb_taps(int16_t tap[4])
{
    /* reference src/input.c:34-39 coefficients, reversed and truncated to
     * int16 as in src/firdecim_q15.c:37-41; tap[i] pairs x[n-14+2i] with x[n-2i] */
    static const float t[4] = { 0.6062333583831787f, -0.13481467962265015f,
                                0.032919470220804214f, -0.00410953676328063f };
    for (int i = 0; i < 4; i++)
        tap[i] = (int16_t)(t[3 - i] * 32767.0f);
}

static inline int16_t hb_axis(const int16_t *w, const int16_t *tap)
{
    /* w[0..14], w[14] newest. int16 accumulator wraps like the reference's */
    int16_t acc = 0;
    for (int i = 0; i < 4; i++)
        acc = (int16_t)(acc + (((w[2 * i] + w[14 - 2 * i]) * tap[i]) >> 15));
    return (int16_t)(acc + w[7]);
}

static inline void hb_shift(int16_t *w, int16_t v)
{
    memmove(w, w + 1, 14 * sizeof(int16_t));
    w[14] = v;
}

static inline int16_t u8_q15(uint8_t v) { return (int16_t)(((int16_t)v - 127) * 64); }

void orc_halfband_fm(const uint8_t *cu8, size_t npairs, int16_t *out)
{
    int16_t wr[15] = { 0 }, wi[15] = { 0 }, tap[4];
    hb_taps(tap);
    for (size_t n = 0; n < npairs; n++) {
        hb_shift(wr, u8_q15(cu8[4 * n + 0]));
        hb_shift(wi, u8_q15(cu8[4 * n + 1]));
        out[2 * n + 0] = hb_axis(wr, tap);
        out[2 * n + 1] = hb_axis(wi, tap);
        hb_shift(wr, u8_q15(cu8[4 * n + 2]));
        hb_shift(wi, u8_q15(cu8[4 * n + 3]));
    }
}

/* ------------------------------------------------------------------------ */
/* a16: tail-biting Viterbi, n = 3 (reference src/conv_dec.c:359-453,        */
/*      src/conv_gen.h:32-123, src/conv_sse.h:56-66,233-315)                 */
/* ------------------------------------------------------------------------ */
static inline int16_t sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : (int16_t)v); }

void orc_viterbi(const int8_t *in, uint8_t *out, int k, int len, unsigned g0, unsigned g1, unsigned g2)
{
    const int ns = 1 << (k - 1), half = ns / 2;
    const int steps = len + 64;                      /* 32 pre-roll + len + 32 post-roll */
    const int interval = 32767 / (3 * 127) - k;      /* conv_dec.c:370 */
    const unsigned gens[3] = { g0, g1, g2 };
    const int words = ns / 64 ? ns / 64 : 1;
    uint64_t *dec = (uint64_t *)calloc((size_t)steps * words, sizeof(uint64_t));
    int16_t *pmv = (int16_t *)calloc(ns, sizeof(int16_t));
    int16_t *nxt = (int16_t *)calloc(ns, sizeof(int16_t));
    /* expected outputs (+1/-1) of the branch old state 2b, input 0 (conv_dec.c:139-154) */
    int8_t (*sgn)[3] = (int8_t (*)[3])calloc(half, 3);
    for (int b = 0; b < half; b++)
        for (int g = 0; g < 3; g++)
            sgn[b][g] = __builtin_parity(((unsigned)b << 1) & gens[g]) ? 1 : -1;

    int j = len - 32;
    for (int s = 0; s < steps; s++, j++) {
        if (j == len) j = 0;
        const int8_t *q = in + 3 * j;
        uint64_t *d = dec + (size_t)s * words;
        for (int b = 0; b < half; b++) {
            int m = q[0] * sgn[b][0] + q[1] * sgn[b][1] + q[2] * sgn[b][2];
            int a0 = sat16(pmv[2 * b] + m), a1 = sat16(pmv[2 * b + 1] - m);
            int c0 = sat16(pmv[2 * b] - m), c1 = sat16(pmv[2 * b + 1] + m);
            /* survivor bit = 1 when the predecessor is the odd state (tie -> odd) */
            if (a0 > a1) nxt[b] = (int16_t)a0;
            else { nxt[b] = (int16_t)a1; d[b >> 6] |= 1ull << (b & 63); }
            int hb = b + half;
            if (c0 > c1) nxt[hb] = (int16_t)c0;
            else { nxt[hb] = (int16_t)c1; d[hb >> 6] |= 1ull << (hb & 63); }
        }
        if (s % interval == 0) {
            int16_t mn = nxt[0];
            for (int i = 1; i < ns; i++) if (nxt[i] < mn) mn = nxt[i];
            for (int i = 0; i < ns; i++) nxt[i] = sat16(nxt[i] - mn);
        }
        memcpy(pmv, nxt, ns * sizeof(int16_t));
    }
    /* first maximum wins (conv_dec.c:310-317) */
    int best = -1;
    unsigned state = 0;
    for (int i = 0; i < ns; i++)
        if (pmv[i] > best) { best = pmv[i]; state = (unsigned)i; }
    const unsigned mask = (unsigned)ns - 2;
    for (int s = steps - 1; s >= 0; s--) {
        const uint64_t *d = dec + (size_t)s * words;
        unsigned bit = (unsigned)((d[state >> 6] >> (state & 63)) & 1);
        if (s >= 32 && s < 32 + len)
            out[s - 32] = (uint8_t)((state >> (k - 2)) & 1);
        state = ((state << 1) & mask) | bit;
    }
    free(dec); free(pmv); free(nxt); free(sgn);
}

/* ------------------------------------------------------------------------ */
/* a14: interleaver I / II + depuncture (reference src/decode.c:296-342)     */
/* ------------------------------------------------------------------------ */
static const int8_t PMV[20] = { 10, 2, 18, 6, 14, 8, 16, 0, 12, 4, 11, 3, 19, 7, 15, 9, 17, 1, 13, 5 };

void orc_deinterleave_p1(const int8_t *pm, int8_t *out)
{
    unsigned o = 0;
    for (unsigned i = 0; i < P1_ENC; i++) {
        unsigned part = (unsigned)PMV[i % 20];
        unsigned block = (i / 20 + part * 7) % 16;
        unsigned k = i / 320;
        unsigned row = (k * 11) % 32;
        unsigned col = (k * 11 + k / 288) % 36;
        out[o++] = pm[(block * 32 + row) * 720 + part * 36 + col];
        if (o % 6 == 5) out[o++] = 0;
    }
}

void orc_deinterleave_pids(const int8_t *pm, unsigned bc, int8_t *out)
{
    unsigned o = 0;
    for (unsigned i = bc * 200; i < (bc + 1) * 200; i++) {
        unsigned part = (unsigned)PMV[i % 20];
        unsigned block = i / 200;
        unsigned k = (i / 20) % 10 + P1_ENC / 320;
        unsigned row = (k * 11) % 32;
        unsigned col = (k * 11 + k / 288) % 36;
        out[o++] = pm[(block * 32 + row) * 720 + part * 36 + col];
        if (o % 6 == 5) out[o++] = 0;
    }
}

/* a15: interleaver IV (reference src/decode.c:344-376) */
static void px_reset(px_deint_t *x)
{
    x->wr = 0;
    memset(x->rd, 0, sizeof(x->rd));
    x->primed = 0;
    x->started = 0;
}

static void px_deinterleave(px_deint_t *x, int8_t *out, unsigned frame_len)
{
    const unsigned J = frame_len == 4608 ? 4 : 2, M = frame_len == 4608 ? 2 : 4;
    const unsigned N = frame_len == 4608 ? 147456 : 73728, C = 36, B = 32;
    const unsigned per_blk = 32 * C;
    if (x->wr == N) {
        x->wr = 0;
        memset(x->rd, 0, sizeof(x->rd));
        x->primed = 1;
    }
    unsigned o = 0;
    for (unsigned i = 0; i < 2 * frame_len; i++) {
        unsigned part = ((x->wr + 2 * (M / 4)) / M) % J;
        unsigned t = x->rd[part]++;
        unsigned block = (t + part * 7 - (per_blk - 1) * (t / per_blk)) % B;
        unsigned row = ((11 * t) % per_blk) / C;
        unsigned col = (t * 11) % C;
        out[o++] = x->ring[(block * 32 + row) * (J * C) + part * C + col];
        if (o % 6 == 1 || o % 6 == 4) out[o++] = 0;
        x->ring[x->wr++] = x->pair[i];
    }
}

/* a18: reference src/decode.c:279-294 */
void orc_descramble(uint8_t *bits, unsigned len)
{
    unsigned reg = 0x3ff;
    for (unsigned i = 0; i < len; i += 8)
        for (unsigned j = 0; j < 8; j++) {
            unsigned b = ((reg >> 9) ^ reg) & 1;
            reg |= b << 11;
            reg >>= 1;
            if (i + j < len) bits[i + j] ^= (uint8_t)b;
        }
}

/* a17: reference src/decode.c:234-265 */
int orc_bit_errors_fm(const int8_t *coded, const uint8_t *decoded, int len)
{
    static const unsigned g[3] = { 0133, 0171, 0165 };
    unsigned reg = 0;
    int errs = 0;
    for (int i = 0; i < 6; i++)
        reg = (reg >> 1) | ((unsigned)decoded[len - 6 + i] << 6);
    for (int i = 0, j = 0; i < len; i++, j += 3) {
        reg = (reg >> 1) | ((unsigned)decoded[i] << 6);
        for (int c = 0; c < 3; c++)
            if ((j + c) % 6 != 5 && ((coded[j + c] > 0) != __builtin_parity(reg & g[c])))
                errs++;
    }
    return errs;
}

/* ------------------------------------------------------------------------ */
/* a20: RS(255,247) over GF(256)/0x11d, fcr=1, prim=1 (reference             */
/*      src/rs_decode.c:16-210 as configured at src/frame.c:747)             */
/* ------------------------------------------------------------------------ */
static uint8_t gf_exp[256], gf_log[256];
static int gf_ready;

static void gf_init(void)
{
    if (gf_ready) return;
    unsigned v = 1;
    gf_log[0] = 255;
    gf_exp[255] = 0;
    for (int i = 0; i < 255; i++) {
        gf_exp[i] = (uint8_t)v;
        gf_log[v] = (uint8_t)i;
        v <<= 1;
        if (v & 0x100) v ^= 0x11d;
    }
    gf_ready = 1;
}

static inline unsigned mod255(unsigned x)
{
    while (x >= 255) { x -= 255; x = (x >> 8) + (x & 255); }
    return x;
}

int orc_rs_decode(uint8_t *data)
{
    enum { R = 8, NN = 255, A0 = 255 };
    gf_init();
    uint8_t s[R], lambda[R + 1], b[R + 1], t[R + 1], omega[R + 1], reg[R + 1], root[R], loc[R];
    /* syndromes by Horner at alpha^(1+i) */
    for (int i = 0; i < R; i++) s[i] = data[0];
    for (int j = 1; j < NN; j++)
        for (int i = 0; i < R; i++)
            s[i] = s[i] == 0 ? data[j] : (uint8_t)(data[j] ^ gf_exp[mod255(gf_log[s[i]] + (1 + i))]);
    unsigned any = 0;
    for (int i = 0; i < R; i++) { any |= s[i]; s[i] = gf_log[s[i]]; }
    if (!any) return 0;

    memset(lambda, 0, sizeof(lambda));
    lambda[0] = 1;
    for (int i = 0; i <= R; i++) b[i] = gf_log[lambda[i]];
    /* Berlekamp-Massey */
    int el = 0;
    for (int r = 1; r <= R; r++) {
        uint8_t disc = 0;
        for (int i = 0; i < r; i++)
            if (lambda[i] != 0 && s[r - i - 1] != A0)
                disc ^= gf_exp[mod255(gf_log[lambda[i]] + s[r - i - 1])];
        disc = gf_log[disc];
        if (disc == A0) {
            memmove(&b[1], b, R);
            b[0] = A0;
        } else {
            t[0] = lambda[0];
            for (int i = 0; i < R; i++)
                t[i + 1] = b[i] != A0 ? (uint8_t)(lambda[i + 1] ^ gf_exp[mod255(disc + b[i])]) : lambda[i + 1];
            if (2 * el <= r - 1) {
                el = r - el;
                for (int i = 0; i <= R; i++)
                    b[i] = lambda[i] == 0 ? A0 : (uint8_t)mod255(gf_log[lambda[i]] - disc + NN);
            } else {
                memmove(&b[1], b, R);
                b[0] = A0;
            }
            memcpy(lambda, t, R + 1);
        }
    }
    int deg = 0;
    for (int i = 0; i <= R; i++) {
        lambda[i] = gf_log[lambda[i]];
        if (lambda[i] != A0) deg = i;
    }
    /* Chien search */
    memcpy(&reg[1], &lambda[1], R);
    int count = 0;
    for (unsigned i = 1, k = 0; i <= NN; i++, k = mod255(k + 1)) {
        /* iprim = 1 so the location counter starts at iprim-1 = 0 */
        uint8_t q = 1;
        for (int j = deg; j > 0; j--)
            if (reg[j] != A0) {
                reg[j] = (uint8_t)mod255(reg[j] + j);
                q ^= gf_exp[reg[j]];
            }
        if (q != 0) continue;
        root[count] = (uint8_t)i;
        loc[count] = (uint8_t)k;
        if (++count == deg) break;
    }
    if (deg != count) return -1;
    /* omega = s * lambda mod x^R */
    int dego = 0;
    for (int i = 0; i < R; i++) {
        uint8_t tmp = 0;
        for (int j = deg < i ? deg : i; j >= 0; j--)
            if (s[i - j] != A0 && lambda[j] != A0)
                tmp ^= gf_exp[mod255(s[i - j] + lambda[j])];
        if (tmp != 0) dego = i;
        omega[i] = gf_log[tmp];
    }
    omega[R] = A0;
    /* Forney */
    for (int j = count - 1; j >= 0; j--) {
        uint8_t num1 = 0, den = 0;
        for (int i = dego; i >= 0; i--)
            if (omega[i] != A0)
                num1 ^= gf_exp[mod255(omega[i] + i * root[j])];
        uint8_t num2 = gf_exp[mod255(root[j] * 0 + NN)];      /* fcr - 1 = 0 */
        for (int i = (deg < R - 1 ? deg : R - 1) & ~1; i >= 0; i -= 2)
            if (lambda[i + 1] != A0)
                den ^= gf_exp[mod255(lambda[i + 1] + i * root[j])];
        if (den == 0) return -1;
        if (num1 != 0)
            data[loc[j]] ^= gf_exp[mod255(gf_log[num1] + gf_log[num2] + NN - gf_log[den])];
    }
    return count;
}

/* reference src/frame.c:158-179 */
int orc_fix_header(uint8_t *buf)
{
    uint8_t blk[255];
    memset(blk, 0, 159);
    for (int i = 0; i < 96; i++) blk[254 - i] = buf[i];
    if (orc_rs_decode(blk) == -1) return 0;
    for (int i = 0; i < 159; i++)
        if (blk[i] != 0) return 0;
    for (int i = 0; i < 96; i++) buf[i] = blk[254 - i];
    return 1;
}

/* reference src/frame.c:645-714 (PCI + first PDU bytes only) and :146-156,527-541.
 * Fixed-data frames whose audio region ends at exactly byte 96 are not
 * distinguished (the reference would skip the check there). */
int orc_p1_sync_lost(const uint8_t *bits, uint32_t *pci_out)
{
    uint32_t pci = 0;
    for (int h = 0; h < 24; h++) {
        unsigned i = 116176 + 1248 * (unsigned)h;
        pci |= (uint32_t)bits[(i & ~7u) + 7 - (i & 7)] << (23 - h);
    }
    if (pci_out) *pci_out = pci;
    if ((pci & 0xFFFFFC) == (0x3634CE & 0xFFFFFC)) return 0;       /* no audio */
    uint8_t hdr[96];
    for (int n = 0; n < 96; n++) {
        unsigned v = 0;
        for (int j = 0; j < 8; j++) {
            unsigned i = 8u * (unsigned)n + (unsigned)j;            /* PCI bits lie far beyond */
            v |= (unsigned)bits[(i & ~7u) + 7 - (i & 7)] << (7 - j);
        }
        hdr[n] = (uint8_t)v;
    }
    return !orc_fix_header(hdr);
}

/* ------------------------------------------------------------------------ */
/* state changes (reference src/input.c:172-188)                             */
/* ------------------------------------------------------------------------ */
static void set_state(orc_t *o, int ns)
{
    if (o->state == ns) return;
    if (o->state == ST_FINE)
        olog_put(&o->log, ORC_REC_LOST_SYNC, NULL, 0, NULL, 0);
    if (ns == ST_FINE) {
        float fo = (o->prev_angle - 2 * M_PI * o->cfo) * 744187.5 / (2 * M_PI * NFFT);
        struct { float f; int32_t v[5]; } p = { fo, { o->psmi, -1, -1, -1, -1 } };   /* FM leaves pli..rdbi at sync_reset's -1, sync.c:821-824 */
        olog_put(&o->log, ORC_REC_SYNC, &p, sizeof(p), NULL, 0);
    }
    o->state = ns;
}

/* CRC-12 of a PIDS frame, restated from reference src/pids.c:52-86 with the per-byte bit reversal of
 * pids_frame_push (src/pids.c:1036-1040); pk = 80 frame bits packed MSB-first */
int orc_pids_crc12_ok(const uint8_t *pk)
{
    uint8_t bits[80];
    for (int i = 0; i < 80; i++) {
        const int src = ((i >> 3) << 3) + 7 - (i & 7);              /* frame bit index read by pids_frame_push */
        bits[i] = (pk[src >> 3] >> (7 - (src & 7))) & 1;
    }
    uint16_t poly = 0xD010, reg = 0;
    for (int i = 67; i >= 0; i--) {
        int lowbit = reg & 1;
        reg >>= 1;
        reg ^= (uint16_t)(bits[i] << 15);
        if (lowbit) reg ^= poly;
    }
    for (int i = 0; i < 16; i++) {
        int lowbit = reg & 1;
        reg >>= 1;
        if (lowbit) reg ^= poly;
    }
    reg ^= 0x955;
    uint16_t expected = 0;
    for (int i = 68; i < 80; i++) expected = (uint16_t)((expected << 1) | bits[i]);
    return expected == (reg & 0xfff);
}

static void decode_reset(orc_t *o)
{
    o->started_pm = 0;
    px_reset(&o->px1);
    px_reset(&o->px2);
}

/* ------------------------------------------------------------------------ */
/* a13..a19: L1 assembly per block (reference src/decode.c:378-471)          */
/* ------------------------------------------------------------------------ */
static void emit_frame(orc_t *o, const uint8_t *bits, unsigned len, unsigned lc)
{
    size_t nb = (len + 7) / 8;
    uint8_t *pk = (uint8_t *)calloc(nb, 1);
    for (unsigned i = 0; i < len; i++) pk[i >> 3] |= (uint8_t)((bits[i] & 1) << (7 - (i & 7)));
    uint32_t hdr[2] = { lc, len };
    olog_put(&o->log, ORC_REC_FRAME, hdr, sizeof(hdr), pk, nb);
    free(pk);
}

static void push_pm(orc_t *o, const int8_t *soft, unsigned bc)
{
    if (o->want_soft) {
        uint32_t b = bc;
        olog_put(&o->log, ORC_REC_SOFT_PM, &b, 4, soft, PM_BLOCK);
    }
    memcpy(o->pm + PM_BLOCK * bc, soft, PM_BLOCK);
    /* PIDS every block */
    int8_t vp[PIDS_LEN * 3];
    uint8_t bp[PIDS_LEN];
    orc_deinterleave_pids(o->pm, bc, vp);
    orc_viterbi(vp, bp, 7, PIDS_LEN, 0133, 0171, 0165);
    orc_descramble(bp, PIDS_LEN);
    uint8_t pk[10] = { 0 };
    for (int i = 0; i < PIDS_LEN; i++) pk[i >> 3] |= (uint8_t)(bp[i] << (7 - (i & 7)));
    const uint8_t crc_ok = (uint8_t)orc_pids_crc12_ok(pk);
    olog_put(&o->log, ORC_REC_PIDS, pk, 10, &crc_ok, 1);

    if (bc == 0) o->started_pm = 1;
    if (o->started_pm && bc == 15) {
        orc_deinterleave_p1(o->pm, o->vit_p1);
        orc_viterbi(o->vit_p1, o->out_p1, 7, P1_LEN, 0133, 0171, 0165);
        float ber = (float)orc_bit_errors_fm(o->vit_p1, o->out_p1, P1_LEN) / P1_ENC;
        olog_put(&o->log, ORC_REC_BER, &ber, 4, NULL, 0);
        orc_descramble(o->out_p1, P1_LEN);
        emit_frame(o, o->out_p1, P1_LEN, 0);
        if (orc_p1_sync_lost(o->out_p1, NULL))
            set_state(o, ST_NONE);
    }
}

static void push_px(orc_t *o, px_deint_t *x, const int8_t *soft, unsigned len, unsigned bc, unsigned lc)
{
    if (bc % 2 == 0) x->started = 1;
    if (!x->started) return;
    memcpy(x->pair + len * (bc % 2), soft, len);
    if (bc % 2 == 1) {
        int8_t vit[P3_LEN_MAX * 3];
        uint8_t bits[P3_LEN_MAX];
        px_deinterleave(x, vit, len);
        if (x->primed) {
            orc_viterbi(vit, bits, 7, (int)len, 0133, 0171, 0165);
            orc_descramble(bits, len);
            emit_frame(o, bits, len, lc);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a8..a12: sync / equalise / demap (reference src/sync.c)                   */
/* ------------------------------------------------------------------------ */
static void costas_ref(orc_t *o, unsigned ref, int cfo)            /* sync.c:90-130 */
{
    static const signed char pat[BLK] = { -1, 1, -1, -1, -1, 1, 1, 0, 1, -1, 0, 0, 0, -1, -1, 0,
                                          0, 0, 0, 0, -1, 1, -1, 0, 0, 0, 0, 0, 0, 0, 0, -1 };
    float cfo_freq = 2 * M_PI * cfo * NCP / NFFT;
    for (unsigned n = 0; n < BLK; n++) {
        float error = cargf(o->bins[ref][n] * o->bins[ref][n] * cexpf(-I * 2 * o->cphase[ref])) * 0.5;
        o->phs[ref][n] = o->cphase[ref];
        o->bins[ref][n] *= cexpf(-I * o->cphase[ref]);
        o->cfreq[ref] += o->beta * error;
        if (o->cfreq[ref] > 0.5) o->cfreq[ref] = 0.5;
        if (o->cfreq[ref] < -0.5) o->cfreq[ref] = -0.5;
        o->cphase[ref] += o->cfreq[ref] + cfo_freq + (o->alpha * error);
        if (o->cphase[ref] > M_PI) o->cphase[ref] -= 2 * M_PI;
        if (o->cphase[ref] < -M_PI) o->cphase[ref] += 2 * M_PI;
    }
    float x = 0;
    for (unsigned n = 0; n < BLK; n++) x += crealf(o->bins[ref][n]) * pat[n];
    if (x < 0) {
        for (unsigned n = 0; n < BLK; n++) {
            o->phs[ref][n] += M_PI;
            o->bins[ref][n] *= -1;
        }
        o->cphase[ref] += M_PI;
    }
}

static void uncostas_ref(orc_t *o, unsigned ref)                    /* sync.c:132-136 */
{
    for (unsigned n = 0; n < BLK; n++) o->bins[ref][n] *= cexpf(I * o->phs[ref][n]);
}

static void ref_needle(signed char nd[BLK], unsigned rsid)
{
    static const signed char base[BLK] = { 0, 1, 0, 0, 0, 1, 1, -1, 1, 0, 0, 0, -1, 0, 0, -1,
                                           -1, -1, -1, -1, 0, 1, 0, -1, -1, -1, -1, -1, -1, -1, -1, 0 };
    memcpy(nd, base, BLK);
    nd[10] = (signed char)(rsid >> 1);
    nd[11] = (signed char)((rsid >> 1) ^ (rsid & 1));
}

static int ref_decode(orc_t *o, unsigned ref, unsigned rsid, unsigned *bc, unsigned *psmi)   /* sync.c:169-186 */
{
    signed char nd[BLK];
    unsigned char raw[BLK], d[BLK];
    ref_needle(nd, rsid);
    for (int n = 0; n < BLK; n++)
        if (nd[n] >= 0 && nd[n] != (crealf(o->bins[ref][n]) > 0)) return -1;
    unsigned char prev = 0;
    for (int n = 0; n < BLK; n++) {
        raw[n] = crealf(o->bins[ref][n]) <= 0 ? 0 : 1;
        d[n] = raw[n] ^ prev;
        prev = raw[n];
    }
    *bc = (unsigned)(d[16] << 3 | d[17] << 2 | d[18] << 1 | d[19]);
    *psmi = (unsigned)(d[25] << 5 | d[26] << 4 | d[27] << 3 | d[28] << 2 | d[29] << 1 | d[30]);
    return 0;
}

static int needle_search(const signed char *nd, const unsigned char *d)   /* sync.c:150-167 */
{
    for (int n = 0; n < BLK; n++) {
        int i;
        for (i = 0; i < BLK; i++) {
            if (nd[i] < 0) continue;
            if (nd[i] != d[(n + i) % BLK]) break;
        }
        if (i == BLK) return n;
    }
    return -1;
}

static int ref_find(orc_t *o, unsigned ref, unsigned rsid)          /* sync.c:188-207 */
{
    signed char nd[BLK];
    unsigned char d[BLK];
    ref_needle(nd, rsid);
    for (int n = 0; n < BLK; n++) d[n] = crealf(o->bins[ref][n]) <= 0 ? 0 : 1;
    int m = needle_search(nd, d);
    if (m >= 0) return m;
    for (int n = 0; n < BLK; n++) d[n] ^= 1;
    return needle_search(nd, d);
}

static void search_cfo(orc_t *o)                                    /* sync.c:292-337 */
{
    for (int cfo = -2 * PW; cfo < 2 * PW; cfo++) {
        unsigned votes[BLK] = { 0 };
        for (int i = 0; i <= 10; i++) {
            unsigned refs[2] = { (unsigned)(cfo + LB0 + i * PW), (unsigned)(cfo + UB1 - i * PW) };
            for (int s = 0; s < 2; s++) {
                costas_ref(o, refs[s], cfo);
                int off = ref_find(o, refs[s], (unsigned)(30 - i) & 3);
                uncostas_ref(o, refs[s]);
                if (off >= 0) votes[off]++;
            }
        }
        int best = -1;
        unsigned bestn = 0;
        for (int k = 0; k < BLK; k++)
            if (votes[k] > bestn) { best = k; bestn = votes[k]; }
        if (best >= 0 && bestn >= 3) {
            o->keep_extra = ((BLK - best) % BLK) * NSYM;
            o->cfo += cfo;
            o->cfo_wait = 8;
            break;
        }
    }
}

static float half_pi_wrap(float a, float b)                         /* sync.c:284-290 */
{
    float d = a - b;
    while (d > M_PI / 2) d -= M_PI;
    while (d < -M_PI / 2) d += M_PI;
    return d;
}

static void equalise_partition(orc_t *o, unsigned lo, unsigned hi)  /* sync.c:254-282 */
{
    float m0 = 0, m19 = 0;
    for (int n = 0; n < BLK; n++) m0 += fabsf(crealf(o->bins[lo][n]));
    m0 = m0 / BLK;
    for (int n = 0; n < BLK; n++) m19 += fabsf(crealf(o->bins[hi][n]));
    m19 = m19 / BLK;
    for (int n = 0; n < BLK; n++) {
        cf up = cexpf(o->phs[hi][n] * I);
        cf lp = cexpf(o->phs[lo][n] * I);
        for (int k = 1; k < PW; k++) {
            cf c = CMPLXF(PW, PW) / (k * m19 * up + (PW - k) * m0 * lp);
            o->bins[lo + k][n] *= c;
        }
    }
}

static inline int8_t soft_demap(float x, float mult)                /* sync.c:69-73 */
{
    float c = fmaxf(fminf(x, 1), -1);
    return (int8_t)lroundf(c * mult);
}

static unsigned demap_span(orc_t *o, int8_t *dst, unsigned n, unsigned first, unsigned parts, float mult)
{
    unsigned w = 0;
    for (unsigned p = 0; p < parts; p++)
        for (unsigned j = 1; j < PW; j++) {
            cf c = o->bins[first + p * PW + j][n];
            dst[w++] = soft_demap(crealf(c), mult);
            dst[w++] = soft_demap(cimagf(c), mult);
        }
    return w;
}

static void sync_block_fm(orc_t *o)                                 /* sync.c:339-610 */
{
    int ppb;
    /* partitions_per_band is fixed at entry (sync.c:343-357) even if the vote below changes psmi ... */
    switch (compat_mode[o->psmi]) {
    case 2: ppb = 11; break;
    case 3: ppb = 12; break;
    case 5: case 6: case 11: ppb = 14; break;
    default: ppb = 10;
    }
    for (int i = 0; i < ppb * PW + 1; i += PW) {
        costas_ref(o, (unsigned)(LB0 + i), 0);
        costas_ref(o, (unsigned)(UB1 - i), 0);
    }
    if (o->state == ST_COARSE) {
        unsigned good = 0, seen_bc[16] = { 0 }, seen_psmi[64] = { 0 };
        for (int i = 0; i <= ppb; i++) {
            unsigned bc, psmi;
            if (ref_decode(o, (unsigned)(LB0 + i * PW), (unsigned)(30 - i) & 3, &bc, &psmi) == 0) {
                good++; seen_bc[bc]++; seen_psmi[psmi]++;
            }
            if (ref_decode(o, (unsigned)(UB1 - i * PW), (unsigned)(30 - i) & 3, &bc, &psmi) == 0) {
                good++; seen_bc[bc]++; seen_psmi[psmi]++;
            }
        }
        if (good >= 4) {
            int mbc = -1, mps = -1;
            for (unsigned v = 0; v < 16; v++) if (seen_bc[v] > good / 2) mbc = (int)v;
            for (unsigned v = 0; v < 16; v++) if (seen_psmi[v] > good / 2) mps = (int)v;   /* 0..15 only */
            if (mbc >= 0 && mps >= 0) {
                o->bc = (unsigned)mbc;
                o->psmi = mps;
                set_state(o, ST_FINE);
                decode_reset(o);
            }
        } else if (o->cfo_wait == 0) {
            search_cfo(o);
        } else {
            o->cfo_wait--;
        }
    }
    if (o->state != ST_FINE) return;

    float samperr = 0, angle = 0, sum_xy = 0, sum_x2 = 0;
    for (int i = 0; i < ppb * PW; i += PW) {
        equalise_partition(o, (unsigned)(LB0 + i), (unsigned)(LB0 + i + PW));
        equalise_partition(o, (unsigned)(UB1 - i - PW), (unsigned)(UB1 - i));
        samperr += half_pi_wrap(o->phs[LB0 + i][0], o->phs[LB0 + i + PW][0]);
        samperr += half_pi_wrap(o->phs[UB1 - i - PW][0], o->phs[UB1 - i][0]);
    }
    samperr = samperr / (ppb * 2) * NFFT / PW / (2 * M_PI);
    for (int i = 0; i < ppb * PW + 1; i += PW) {
        float x, y;
        x = LB0 + i - (NFFT / 2);
        y = o->cfreq[LB0 + i];
        angle += y; sum_xy += x * y; sum_x2 += x * x;
        x = UB1 - i - (NFFT / 2);
        y = o->cfreq[UB1 - i];
        angle += y; sum_xy += x * y; sum_x2 += x * x;
    }
    samperr -= (sum_xy / sum_x2) * NFFT / (2 * M_PI) * BLK;
    o->samperr = (int)roundf(samperr);
    angle /= (ppb + 1) * 2;
    o->angle = angle;
    for (int i = 0; i < ppb * PW + 1; i += PW) {
        o->cfreq[LB0 + i] -= angle;
        o->cfreq[UB1 - i] -= angle;
    }

    float e_lb = 0, e_ub = 0;
    for (int n = 0; n < BLK; n++)
        for (int i = 0; i < ppb * PW; i += PW)
            for (int j = 1; j < PW; j++) {
                cf c = o->bins[LB0 + i + j][n];
                cf ideal = CMPLXF(crealf(c) >= 0 ? 1 : -1, cimagf(c) >= 0 ? 1 : -1);
                cf d = ideal - c;
                e_lb += crealf(d) * crealf(d) + cimagf(d) * cimagf(d);
                c = o->bins[UB1 - i - PW + j][n];
                ideal = CMPLXF(crealf(c) >= 0 ? 1 : -1, cimagf(c) >= 0 ? 1 : -1);
                d = ideal - c;
                e_ub += crealf(d) * crealf(d) + cimagf(d) * cimagf(d);
            }
    o->err_lb += e_lb;
    o->err_ub += e_ub;
    if (++o->mer_cnt == 16) {
        float signal = 2 * BLK * (ppb * 18) * o->mer_cnt;
        float db[2] = { 10 * log10f(signal / o->err_lb), 10 * log10f(signal / o->err_ub) };
        olog_put(&o->log, ORC_REC_MER, db, sizeof(db), NULL, 0);
        o->mer_cnt = 0;
        o->err_lb = 0;
        o->err_ub = 0;
    }
    const float mer_lb = 2.0f * BLK * (float)(ppb * 18) / e_lb;
    const float mer_ub = 2.0f * BLK * (float)(ppb * 18) / e_ub;
    const float mult_lb = fmaxf(fminf(mer_lb * 10, 127), 1);
    const float mult_ub = fmaxf(fminf(mer_ub * 10, 127), 1);

    /* ... while the PX1/PX2 demap looks psmi up again (sync.c:537,552,574): in the block that reaches FINE
     * sync the new mode's extra partitions are demapped although they were not equalised */
    const int cm = compat_mode[o->psmi];
    int8_t pm[PM_BLOCK], px1[P3_LEN_MAX], px2[P3_LEN_MAX];
    unsigned n_pm = 0, n_px1 = 0, n_px2 = 0;
    for (unsigned n = 0; n < BLK; n++) {
        n_pm += demap_span(o, pm + n_pm, n, LB0, 10, mult_lb);
        n_pm += demap_span(o, pm + n_pm, n, UB1 - 10 * PW, 10, mult_ub);
        if (cm == 2) {
            n_px1 += demap_span(o, px1 + n_px1, n, LB0 + 10 * PW, 1, mult_lb);
            n_px1 += demap_span(o, px1 + n_px1, n, UB1 - 11 * PW, 1, mult_ub);
        }
        if (cm == 3 || cm == 11) {
            n_px1 += demap_span(o, px1 + n_px1, n, LB0 + 10 * PW, 2, mult_lb);
            n_px1 += demap_span(o, px1 + n_px1, n, UB1 - 12 * PW, 2, mult_ub);
        }
        if (cm == 11) {
            n_px2 += demap_span(o, px2 + n_px2, n, LB0 + 12 * PW, 2, mult_lb);
            n_px2 += demap_span(o, px2 + n_px2, n, UB1 - 14 * PW, 2, mult_lb);  /* sic: sync.c:591-592 */
        }
    }
    push_pm(o, pm, o->bc);
    if (n_px1 > 0) push_px(o, &o->px1, px1, n_px1, o->bc, 1);
    if (n_px2 > 0) push_px(o, &o->px2, px2, n_px2, o->bc, 2);
    o->bc = (o->bc + 1) % 16;
}

/* ------------------------------------------------------------------------ */
/* a3..a7: acquisition / OFDM demodulation of one 33-symbol window           */
/*          (reference src/acquire.c:98-263)                                 */
/* ------------------------------------------------------------------------ */
static const float bp_coeff[32] = {
    -0.000685643230099231f, 0.005636964458972216f, 0.009015781804919243f, -0.015486305579543114f,
    -0.035108357667922974f, 0.017446253448724747f, 0.08155813068151474f, 0.007995186373591423f,
    -0.13311293721199036f, -0.0727422907948494f, 0.15914097428321838f, 0.16498781740665436f,
    -0.1324498951435089f, -0.2484012246131897f, 0.051773931831121445f, 0.2821577787399292f,
    0.051773931831121445f, -0.2484012246131897f, -0.1324498951435089f, 0.16498781740665436f,
    0.15914097428321838f, -0.0727422907948494f, -0.13311293721199036f, 0.007995186373591423f,
    0.08155813068151474f, 0.017446253448724747f, -0.035108357667922974f, -0.015486305579543114f,
    0.009015781804919243f, 0.005636964458972216f, -0.000685643230099231f, 0.0f
};

static inline int16_t bp_axis(const int16_t *w, const int16_t *tap)   /* firdecim_q15.c:95-109 */
{
    /* w[0..31], w[31] newest; tap[i] multiplies w[i] and w[32-i] */
    int16_t acc = 0;
    for (int i = 1; i < 16; i++)
        acc = (int16_t)(acc + (((w[i] + w[32 - i]) * tap[i]) >> 15));
    return (int16_t)(acc + ((w[16] * tap[16]) >> 15));
}

static void process_window(orc_t *o)
{
    int samperr = 0;
    float angle, angle_diff;
    const int state_in = o->state;

    if (o->state == ST_FINE) {
        samperr = NSYM / 2 + o->samperr;
        o->samperr = 0;
        angle_diff = -o->angle;
        o->angle = 0;
        angle = o->prev_angle + angle_diff;
        o->prev_angle = angle;
    } else {
        for (int i = 0; i < NACQ; i++) {
            memmove(o->bp_r, o->bp_r + 1, 31 * sizeof(int16_t));
            memmove(o->bp_i, o->bp_i + 1, 31 * sizeof(int16_t));
            o->bp_r[31] = o->win_r[i];
            o->bp_i[31] = o->win_i[i];
            int16_t yr = bp_axis(o->bp_r, o->bp_tap), yi = bp_axis(o->bp_i, o->bp_tap);
            o->tbuf[i] = CMPLXF((float)yr / 32767.0f, (float)yi / -32767.0f);
        }
        memset(o->sums, 0, sizeof(o->sums));
        for (int i = 0; i < NSYM; i++)
            for (int j = 0; j < BLK; j++)
                o->sums[i] += o->tbuf[i + j * NSYM] * conjf(o->tbuf[i + j * NSYM + NFFT]);
        float best = -1.0f;
        cf best_v = 0;
        for (int i = 0; i < NSYM; i++) {
            cf v = 0;
            for (int j = 0; j < NCP; j++)
                v += o->sums[(i + j) % NSYM] * o->shape[j] * o->shape[j + NFFT];
            float mag = crealf(v) * crealf(v) + cimagf(v) * cimagf(v);
            if (mag > best) {
                best = mag;
                best_v = v;
                samperr = (i + NSYM - 15) % NSYM;
            }
        }
        angle_diff = cargf(best_v * cexpf(I * -o->prev_angle));
        float factor = (o->prev_angle) ? 0.25 : 1.0;
        angle = o->prev_angle + (angle_diff * factor);
        o->prev_angle = angle;
        set_state(o, ST_COARSE);
    }

    for (int i = 0; i < NACQ; i++)
        o->tbuf[i] = CMPLXF((float)o->win_r[i] / 32767.0f, (float)o->win_i[i] / -32767.0f);

    /* sync_adjust (sync.c:769-777) */
    {
        int adj = NSYM / 2 - samperr;
        for (int i = 0; i < MAXPART * PW + 1; i++) {
            o->cphase[LB0 + i] -= adj * (LB0 + i - (NFFT / 2)) * 2 * M_PI / NFFT;
            o->cphase[UB1 - i] -= adj * (UB1 - i - (NFFT / 2)) * 2 * M_PI / NFFT;
        }
    }
    angle -= 2 * M_PI * o->cfo;
    o->phase *= cexpf(-(NSYM / 2 - samperr) * angle / NFFT * I);
    cf inc = cexpf(angle / NFFT * I);

    if (o->want_blocks) {
        struct { int32_t st, se; float ang, pr, pi; int32_t cfo; int64_t start; } r =
            { state_in, samperr, angle, crealf(o->phase), cimagf(o->phase), o->cfo, o->start_index };
        olog_put(&o->log, ORC_REC_BLOCK, &r, sizeof(r), NULL, 0);
    }

    for (int s = 0; s < BLK; s++) {
        for (int j = 0; j < NSYM; j++) {
            cf x = o->phase * o->tbuf[s * NSYM + j + samperr];
            if (j < NCP) o->fin[j] = o->shape[j] * x;
            else if (j < NFFT) o->fin[j] = x;
            else o->fin[j - NFFT] += o->shape[j] * x;
            o->phase *= inc;
        }
        o->phase /= cabsf(o->phase);
        fftwf_execute(o->plan);
        /* fftshift + sync_push (defines.h:123-138, sync.c:779-808): keep bins
         * LB0..LB0+266 and UB1-266..UB1 of the shifted spectrum */
        for (int i = 0; i < MAXPART * PW + 1; i++) {
            o->bins[LB0 + i][o->sym] = o->fout[(LB0 + i + NFFT / 2) % NFFT];
            o->bins[UB1 - i][o->sym] = o->fout[(UB1 - i + NFFT / 2) % NFFT];
        }
        if (++o->sym == BLK) {
            o->sym = 0;
            sync_block_fm(o);
        }
    }

    int keep = NSYM + (NSYM / 2 - samperr) + o->keep_extra;
    o->keep_extra = 0;
    memmove(o->win_r, o->win_r + (NACQ - keep), sizeof(int16_t) * (size_t)keep);
    memmove(o->win_i, o->win_i + (NACQ - keep), sizeof(int16_t) * (size_t)keep);
    o->start_index += NACQ - keep;
    o->fill = (unsigned)keep;
}

/* ------------------------------------------------------------------------ */
/* public                                                                    */
/* ------------------------------------------------------------------------ */
orc_t *orc_new(void)
{
    orc_t *o = (orc_t *)calloc(1, sizeof(*o));
    hb_taps(o->hb_tap);
    /* reversed + truncated like firdecim_q15.c:37-41; bp_tap[i] pairs w[i], w[32-i] */
    for (int i = 0; i < 32; i++) o->bp_tap[i] = (int16_t)(bp_coeff[31 - i] * 32767.0f);
    for (int i = 0; i < NSYM; i++) {
        if (i < NCP) o->shape[i] = sinf(M_PI / 2 * i / NCP);
        else if (i < NFFT) o->shape[i] = 1;
        else o->shape[i] = cosf(M_PI / 2 * (i - NFFT) / NCP);
    }
    o->fin = fftwf_alloc_complex(NFFT);
    o->fout = fftwf_alloc_complex(NFFT);
    o->plan = fftwf_plan_dft_1d(NFFT, o->fin, o->fout, FFTW_FORWARD, FFTW_ESTIMATE);
    float loop_bw = 0.05, damping = 0.70710678;
    float denom = 1 + (2 * damping * loop_bw) + (loop_bw * loop_bw);
    o->alpha = (4 * damping * loop_bw) / denom;
    o->beta = (4 * loop_bw * loop_bw) / denom;
    o->phase = 1;
    o->psmi = 1;
    o->state = ST_NONE;
    decode_reset(o);
    return o;
}

void orc_free(orc_t *o)
{
    if (!o) return;
    fftwf_destroy_plan(o->plan);
    fftwf_free(o->fin);
    fftwf_free(o->fout);
    free(o->log.p);
    free(o);
}

void orc_want_soft(orc_t *o, int on) { o->want_soft = on; }
void orc_want_blocks(orc_t *o, int on) { o->want_blocks = on; }
size_t orc_log_size(const orc_t *o) { return o->log.len; }
const uint8_t *orc_log_data(const orc_t *o) { return o->log.p; }
void orc_log_clear(orc_t *o) { o->log.len = 0; }

/* mirrors input_push_cs16 (reference src/input.c:119-124): FM samples already at 744 187.5 S/s go straight
 * to the acquisition window; nvalues counts int16 values */
void orc_push_cs16(orc_t *o, const int16_t *buf, size_t nvalues)
{
    for (size_t n = 0; n + 1 < nvalues; n += 2) {
        o->win_r[o->fill] = buf[n];
        o->win_i[o->fill] = buf[n + 1];
        if (++o->fill == NACQ)
            process_window(o);
    }
}

void orc_push_cu8(orc_t *o, const uint8_t *buf, size_t nbytes)
{
    for (size_t n = 0; n + 3 < nbytes; n += 4) {
        hb_shift(o->hb_r, u8_q15(buf[n + 0]));
        hb_shift(o->hb_i, u8_q15(buf[n + 1]));
        o->win_r[o->fill] = hb_axis(o->hb_r, o->hb_tap);
        o->win_i[o->fill] = hb_axis(o->hb_i, o->hb_tap);
        hb_shift(o->hb_r, u8_q15(buf[n + 2]));
        hb_shift(o->hb_i, u8_q15(buf[n + 3]));
        if (++o->fill == NACQ)
            process_window(o);
    }
}
