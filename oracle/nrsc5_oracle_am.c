/* nrsc5_oracle_am — plain-C CPU restatement of the NRSC-5 AM (hybrid MA1, all-digital MA3) physical-layer receive chain of
 * theori-io/nrsc5 (reference @ a5c0972), cs16 input at 46 511.72 S/s.
 *
 * TEST INFRASTRUCTURE ONLY (see nrsc5_oracle.h): the checker for the AM rows of the scope table (SURVEY §8
 * a21).  Pinned by tests/test_oracle_am.py against the UNMODIFIED reference in oracle/_ref/libnrsc5_ref.so
 * (PDUs, events and BER bit-identical on the synthetic MA1 and MA3 captures of nrsc5_b200/synth_am.py).
 *
 * Each stage cites the reference file:line it follows.  As in the reference every service mode other than
 * MA3 (psmi 2) is decoded as MA1.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <fftw3.h>

#include "nrsc5_oracle.h"

typedef float complex cf;

#define FFT_AM 256
#define CP_AM 14
#define SYM_AM (FFT_AM + CP_AM)                 /* 270 */
#define BLK 32
#define NACQ_AM (SYM_AM * (BLK + 1))            /* 8910, reference src/acquire.h:12 */
#define CENTER 128
#define REF_IDX 1
#define PIDS_INNER 27
#define PIDS_OUTER 53
#define INNER_START 2
#define MIDDLE_START 28
#define OUTER_START 57
#define MAX_IDX 81
#define PW_AM 25
#define P1_LEN_AM 3750
#define P3_LEN_MA1 24000
#define P3_LEN_MA3 30000
#define MODE_MA3 2                              /* SERVICE_MODE_MA3, reference src/defines.h:39 */
#define PIDS_LEN 80
#define DIVERSITY (18000 * 3)
#define ST_NONE 0
#define ST_COARSE 1
#define ST_FINE 2

typedef struct {
    uint8_t *p;
    size_t len, cap;
} alog_t;

static void alog_put(alog_t *l, uint32_t type, const void *a, size_t alen, const void *b, size_t blen)
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

struct orc_am {
    alog_t log;
    /* acquisition (reference src/acquire.h:7-33) */
    int16_t win_r[NACQ_AM], win_i[NACQ_AM];
    unsigned fill;
    int16_t bp_r[32], bp_i[32];                   /* coarse band-pass history */
    int16_t bp_tap[32];
    cf buf[NACQ_AM];
    cf sums[SYM_AM];
    float shape[SYM_AM];
    float prev_angle;
    cf phase;
    int keep_extra, cfo, state;
    fftwf_complex *fin, *fout;
    fftwf_plan plan;
    /* sync (reference src/sync.h:7-32) */
    cf bins[FFT_AM][BLK];
    unsigned sym;
    int psmi, pli, hppi, aabi, rdbi, cfo_wait, samperr;
    unsigned bc, offset_history;
    float angle;
    /* cu8 input: five cascaded halfband decimators (reference src/input.h:21-25, src/input.c:52-94) */
    int16_t hb_r[5][15], hb_i[5][15], hb_tap[4];
    int16_t stage_r[4][2], stage_i[4][2];
    unsigned offset;
    /* decode (reference src/decode.h:19-62) */
    uint8_t buffer_pl[PW_AM * BLK * 8], buffer_pu[PW_AM * BLK * 8], buffer_s[PW_AM * BLK * 8], buffer_t[PW_AM * BLK * 8];
    uint8_t bl[18000], bu[18000], ml[DIVERSITY + 18000], mu[DIVERSITY + 18000], el[12000], eu[24000];
    uint8_t ebl[18000], ebu[18000], eml[DIVERSITY + 18000], emu[DIVERSITY + 18000];      /* MA3, decode.h:49-52 */
    uint8_t p1_am[8 * 9000], p3_am[72000];
    int8_t vit_p1[8 * P1_LEN_AM * 3], vit_p3[P3_LEN_MA3 * 3], vit_pids[PIDS_LEN * 3];
    uint8_t out_p1[P1_LEN_AM + 8], out_p3[P3_LEN_MA3 + 8], out_pids[PIDS_LEN + 8];
    int am_errors, am_diversity_wait;
};

/* reference src/acquire.c:63-96, reversed and truncated to int16 as in src/firdecim_q15.c:37-41 */
static const float bp_coeff_am[32] = {
    -0.00038464731187559664f, -0.00021618751634377986f, 0.0026779419276863337f, -0.00029802651260979474f,
    -0.0012626448879018426f, -0.0013182522961869836f, -0.012252614833414555f, 0.015980124473571777f,
    0.037112727761268616f, -0.05451361835002899f, -0.05804193392395973f, 0.11320608854293823f,
    0.055298302322626114f, -0.16878043115139008f, -0.022917453199625015f, 0.19178225100040436f,
    -0.022917453199625015f, -0.16878043115139008f, 0.055298302322626114f, 0.11320608854293823f,
    -0.05804193392395973f, -0.05451361835002899f, 0.037112727761268616f, 0.015980124473571777f,
    -0.012252614833414555f, -0.0013182522961869836f, -0.0012626448879018426f, -0.00029802651260979474f,
    0.0026779419276863337f, -0.00021618751634377986f, -0.00038464731187559664f, 0.0f
};

static inline int16_t bp_axis(const int16_t *w, const int16_t *tap)       /* firdecim_q15.c:95-109 */
{
    int16_t acc = 0;
    for (int i = 1; i < 16; i++)
        acc = (int16_t)(acc + (((w[i] + w[32 - i]) * tap[i]) >> 15));
    return (int16_t)(acc + ((w[16] * tap[16]) >> 15));
}

static void emit_frame(orc_am_t *o, const uint8_t *bits, unsigned len, unsigned lc)
{
    size_t nb = (len + 7) / 8;
    uint8_t *pk = (uint8_t *)calloc(nb, 1);
    for (unsigned i = 0; i < len; i++) pk[i >> 3] |= (uint8_t)((bits[i] & 1) << (7 - (i & 7)));
    uint32_t hdr[2] = { lc, len };
    alog_put(&o->log, ORC_REC_FRAME, hdr, sizeof(hdr), pk, nb);
    free(pk);
}

static void decode_reset(orc_am_t *o)                 /* decode.c:556-565 */
{
    o->am_errors = 0;
    o->am_diversity_wait = 4;
}

static void set_state(orc_am_t *o, int ns)            /* input.c:172-188 */
{
    if (o->state == ns) return;
    if (o->state == ST_FINE)
        alog_put(&o->log, ORC_REC_LOST_SYNC, NULL, 0, NULL, 0);
    if (ns == ST_FINE) {
        float fo = (o->prev_angle - 2 * M_PI * o->cfo) * 46511.71875 / (2 * M_PI * FFT_AM);
        struct { float f; int32_t v[5]; } p = { fo, { o->psmi, o->pli, o->hppi, o->aabi, o->rdbi } };
        alog_put(&o->log, ORC_REC_SYNC, &p, sizeof(p), NULL, 0);
    }
    o->state = ns;
}

/* ------------------------------------------------------------------------ */
/* decode (reference src/decode.c:67-231, 234-277, 474-554)                  */
/* ------------------------------------------------------------------------ */
static int bit_map(const uint8_t *matrix, int b, int k, int p)            /* decode.c:67-72 */
{
    const int col = (9 * k) % 25;
    const int row = (11 * col + 16 * (k / 25) + 11 * (k / 50)) % 32;
    return (matrix[PW_AM * (b * BLK + row) + col] >> p) & 1;
}

static void interleaver_ma1(orc_am_t *o)                                  /* decode.c:74-231 */
{
    const int ma3 = o->psmi == MODE_MA3;
    static const int bl_delay[] = { 2, 1, 5 }, ml_delay[] = { 11, 6, 7 }, bu_delay[] = { 10, 8, 9 }, mu_delay[] = { 4, 3, 0 };
    static const int el_delay[] = { 0, 1 }, eu_delay[] = { 2, 3, 5, 4 };
    for (int n = 0; n < 18000; n++) {
        o->bl[n] = (uint8_t)bit_map(o->buffer_pl, n / 2250, (n + n / 750 + 1) % 750, n % 3);
        o->ml[DIVERSITY + n] = (uint8_t)bit_map(o->buffer_pl, (3 * n + 3) % 8, (n + n / 3000 + 3) % 750, 3 + (n % 3));
        o->bu[n] = (uint8_t)bit_map(o->buffer_pu, n / 2250, (n + n / 750) % 750, n % 3);
        o->mu[DIVERSITY + n] = (uint8_t)bit_map(o->buffer_pu, (3 * n) % 8, (n + n / 3000 + 2) % 750, 3 + (n % 3));
    }
    if (!ma3) {
        for (int n = 0; n < 12000; n++)
            o->el[n] = (uint8_t)bit_map(o->buffer_t, (3 * n + n / 3000) % 8, (n + (n / 6000)) % 750, n % 2);
        for (int n = 0; n < 24000; n++)
            o->eu[n] = (uint8_t)bit_map(o->buffer_s, (3 * n + n / 3000 + 2 * (n / 12000)) % 8, (n + (n / 6000)) % 750, n % 4);
    } else {
        for (int n = 0; n < 18000; n++) {                                  /* decode.c:119-140 */
            o->ebl[n] = (uint8_t)bit_map(o->buffer_t, (3 * n + 3) % 8, (n + n / 3000 + 3) % 750, n % 3);
            o->eml[DIVERSITY + n] = (uint8_t)bit_map(o->buffer_t, (3 * n + 3) % 8, (n + n / 3000 + 3) % 750, 3 + (n % 3));
            o->ebu[n] = (uint8_t)bit_map(o->buffer_s, (3 * n) % 8, (n + n / 3000 + 2) % 750, n % 3);
            o->emu[DIVERSITY + n] = (uint8_t)bit_map(o->buffer_s, (3 * n) % 8, (n + n / 3000 + 2) % 750, 3 + (n % 3));
        }
    }
    for (int i = 0; i < 6000; i++) {
        for (int j = 0; j < 3; j++) {
            o->p1_am[i * 12 + bl_delay[j]] = o->bl[i * 3 + j];
            o->p1_am[i * 12 + ml_delay[j]] = o->ml[i * 3 + j];
            o->p1_am[i * 12 + bu_delay[j]] = o->bu[i * 3 + j];
            o->p1_am[i * 12 + mu_delay[j]] = o->mu[i * 3 + j];
        }
        if (!ma3) {
            for (int j = 0; j < 2; j++) o->p3_am[i * 6 + el_delay[j]] = o->el[i * 2 + j];
            for (int j = 0; j < 4; j++) o->p3_am[i * 6 + eu_delay[j]] = o->eu[i * 4 + j];
        } else {
            for (int j = 0; j < 3; j++) {                                  /* decode.c:163-170 */
                o->p3_am[i * 12 + bl_delay[j]] = o->ebl[i * 3 + j];
                o->p3_am[i * 12 + ml_delay[j]] = o->eml[i * 3 + j];
                o->p3_am[i * 12 + bu_delay[j]] = o->ebu[i * 3 + j];
                o->p3_am[i * 12 + mu_delay[j]] = o->emu[i * 3 + j];
            }
        }
    }
    memmove(o->ml, o->ml + 18000, DIVERSITY);
    memmove(o->mu, o->mu + 18000, DIVERSITY);
    if (ma3) {
        memmove(o->eml, o->eml + 18000, DIVERSITY);
        memmove(o->emu, o->emu + 18000, DIVERSITY);
    }
    int off = 0;
    for (int i = 0; i < 8 * P1_LEN_AM * 3; i++) {
        const int r = i % 15;
        o->vit_p1[i] = (r == 1 || r == 4 || r == 7) ? 0 : (o->p1_am[off++] ? 1 : -1);
    }
    off = 0;
    if (!ma3) {
        for (int i = 0; i < P3_LEN_MA1 * 3; i++) {
            const int r = i % 6;
            o->vit_p3[i] = (r == 1 || r == 4 || r == 5) ? 0 : (o->p3_am[off++] ? 1 : -1);
        }
    } else {
        for (int i = 0; i < P3_LEN_MA3 * 3; i++) {                         /* decode.c:214-229 */
            const int r = i % 15;
            o->vit_p3[i] = (r == 1 || r == 4 || r == 7) ? 0 : (o->p3_am[off++] ? 1 : -1);
        }
    }
}

static int bit_errors(const int8_t *coded, const uint8_t *decoded, unsigned k, unsigned len, const unsigned gens[3],
                      const uint8_t *punct, int plen)                       /* decode.c:234-259 */
{
    uint16_t r = 0;
    unsigned errors = 0;
    for (unsigned i = 0; i < k - 1; i++)
        r = (uint16_t)((r >> 1) | (decoded[len - (k - 1) + i] << (k - 1)));
    for (unsigned i = 0, j = 0; i < len; i++, j += 3) {
        r = (uint16_t)((r >> 1) | (decoded[i] << (k - 1)));
        for (unsigned g = 0; g < 3; g++)
            if (punct[(j + g) % plen] && ((coded[j + g] > 0) != __builtin_parity(r & gens[g])))
                errors++;
    }
    return (int)errors;
}

static const unsigned GENS_E1[3] = { 0561, 0657, 0711 }, GENS_E2[3] = { 0561, 0753, 0711 };

static void process_pids(orc_am_t *o, const uint8_t *sbit)                 /* decode.c:474-505 */
{
    static const int il_delay[] = { 0, 1, 12, 13, 6, 5, 18, 17, 11, 7, 23, 19 };
    static const int iu_delay[] = { 2, 4, 14, 16, 3, 8, 15, 20, 9, 10, 21, 22 };
    uint8_t il[120], iu[120];
    for (int n = 0; n < 120; n++) {
        int p = n % 4, k, row;
        k = (n + (n / 60) + 11) % 30;
        row = (11 * (k + (k / 15)) + 3) % 32;
        il[n] = (sbit[row * 2] >> p) & 1;
        k = (n + (n / 60)) % 30;
        row = (11 * (k + (k / 15)) + 3) % 32;
        iu[n] = (sbit[row * 2 + 1] >> p) & 1;
    }
    const int pids1_disabled = (o->psmi == 1) && o->rdbi;
    for (int i = 0; i < 10; i++)
        for (int j = 0; j < 12; j++) {
            o->vit_pids[i * 24 + il_delay[j]] = pids1_disabled ? 0 : (il[i * 12 + j] ? 1 : -1);
            o->vit_pids[i * 24 + iu_delay[j]] = iu[i * 12 + j] ? 1 : -1;
        }
    orc_viterbi(o->vit_pids, o->out_pids, 9, PIDS_LEN, GENS_E2[0], GENS_E2[1], GENS_E2[2]);
    orc_descramble(o->out_pids, PIDS_LEN);
    uint8_t pk[10] = { 0 };
    for (int i = 0; i < PIDS_LEN; i++) pk[i >> 3] |= (uint8_t)(o->out_pids[i] << (7 - (i & 7)));
    const uint8_t crc_ok = (uint8_t)orc_pids_crc12_ok(pk);
    alog_put(&o->log, ORC_REC_PIDS, pk, 10, &crc_ok, 1);
}

/* frame.c:645-714 (PCI), :146-156, :527-541: an AM P1 PDU that announces audio but whose first header fails RS
 * sends the receiver back to acquisition */
static int p1_am_sync_lost(const uint8_t *bits)
{
    uint8_t pdu[96];
    unsigned h = 0, j = 0, nb = 0, val = 0;
    uint32_t pci = 0;
    memset(pdu, 0, sizeof(pdu));
    for (unsigned i = 0; i < P1_LEN_AM; i++) {
        unsigned byte_start = (i >> 3) << 3;
        unsigned byte_len = (P1_LEN_AM - byte_start < 8) ? P1_LEN_AM - byte_start : 8;
        uint8_t bit = bits[byte_start + byte_len - 1 - (i & 7)];
        if (i >= 120 && ((i - 120) % 160) == 0 && h < 22) {
            pci |= (uint32_t)bit << (23 - h);
            ++h;
        } else {
            val |= (unsigned)bit << (7 - j);
            if (++j == 8) {
                if (nb < 96) pdu[nb] = (uint8_t)val;
                nb++;
                val = 0;
                j = 0;
            }
        }
    }
    if ((pci & 0xFFFFFC) == (0x3634CE & 0xFFFFFC)) return 0;              /* fixed data only: no audio */
    return !orc_fix_header(pdu);
}

static void process_p1_p3(orc_am_t *o, unsigned bc)                        /* decode.c:507-554 */
{
    static const uint8_t punct_e1[] = { 1, 0, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1 }, punct_e2[] = { 1, 0, 1, 1, 0, 0 };
    if (bc == 0) o->am_errors = 0;
    if (o->am_diversity_wait == 0) {
        const int8_t *v = o->vit_p1 + bc * P1_LEN_AM * 3;
        orc_viterbi(v, o->out_p1, 9, P1_LEN_AM, GENS_E1[0], GENS_E1[1], GENS_E1[2]);
        o->am_errors += bit_errors(v, o->out_p1, 9, P1_LEN_AM, GENS_E1, punct_e1, 15);
        orc_descramble(o->out_p1, P1_LEN_AM);
        emit_frame(o, o->out_p1, P1_LEN_AM, 0);
        if (p1_am_sync_lost(o->out_p1)) set_state(o, ST_NONE);           /* inside frame_push, frame.c:538 */
        if (bc == 7) {
            unsigned total = 8 * 9000;
            if (!o->rdbi) {
                if (o->psmi != MODE_MA3) {
                    total += 36000;
                    orc_viterbi(o->vit_p3, o->out_p3, 9, P3_LEN_MA1, GENS_E2[0], GENS_E2[1], GENS_E2[2]);
                    o->am_errors += bit_errors(o->vit_p3, o->out_p3, 9, P3_LEN_MA1, GENS_E2, punct_e2, 6);
                    orc_descramble(o->out_p3, P3_LEN_MA1);
                    emit_frame(o, o->out_p3, P3_LEN_MA1, 1);
                } else {                                                   /* decode.c:533-539 */
                    total += 72000;
                    orc_viterbi(o->vit_p3, o->out_p3, 9, P3_LEN_MA3, GENS_E1[0], GENS_E1[1], GENS_E1[2]);
                    o->am_errors += bit_errors(o->vit_p3, o->out_p3, 9, P3_LEN_MA3, GENS_E1, punct_e1, 15);
                    orc_descramble(o->out_p3, P3_LEN_MA3);
                    emit_frame(o, o->out_p3, P3_LEN_MA3, 1);
                }
            }
            float cber = (float)o->am_errors / (float)total;
            alog_put(&o->log, ORC_REC_BER, &cber, sizeof(cber), NULL, 0);
        }
    }
    if (bc == 7) {
        interleaver_ma1(o);
        if (o->am_diversity_wait > 0) o->am_diversity_wait--;
    }
}

/* ------------------------------------------------------------------------ */
/* sync (reference src/sync.c:37-88, 208-252, 284-290, 612-767)              */
/* ------------------------------------------------------------------------ */
static uint8_t gray4(float f) { return f < -1 ? 0 : f < 0 ? 2 : f < 1 ? 3 : 1; }
static uint8_t gray8(float f)
{
    return f < -3 ? 0 : f < -2 ? 4 : f < -1 ? 6 : f < 0 ? 2 : f < 1 ? 3 : f < 2 ? 7 : f < 3 ? 5 : 1;
}
static uint8_t qpsk(cf c) { return (uint8_t)((crealf(c) < 0 ? 0 : 1) | (cimagf(c) < 0 ? 0 : 2)); }
static uint8_t qam16(cf c) { return (uint8_t)(gray4(crealf(c)) | (gray4(cimagf(c)) << 2)); }
static uint8_t qam64(cf c) { return (uint8_t)(gray8(crealf(c)) | (gray8(cimagf(c)) << 3)); }

static float phase_diff(float a, float b)
{
    float diff = a - b;
    while (diff > M_PI / 2) diff -= M_PI;
    while (diff < -M_PI / 2) diff += M_PI;
    return diff;
}

static int fuzzy_match(const signed char *needle, unsigned nn, const unsigned char *data, int size)   /* sync.c:150-167 */
{
    for (int n = 0; n < size; n++) {
        unsigned i;
        for (i = 0; i < nn; i++) {
            if (needle[i] < 0) continue;
            if (needle[i] != data[(n + i) % size]) break;
        }
        if (i == nn) return n;
    }
    return -1;
}

static const signed char needle_am[32] = {
    0, 1, 1, 0, 0, 1, 0, -1, -1, 1, -1, -1, -1, -1, 0, -1, -1, -1, -1, -1, -1, 1, 1, -1, -1, -1, -1, -1, -1, -1, -1, -1
};

static int find_block_am(orc_am_t *o, unsigned ref)                        /* sync.c:208-237 */
{
    unsigned char data[BLK];
    for (int n = 0; n < BLK; n++) {
        data[n] = cimagf(o->bins[ref][n]) <= 0 ? 0 : 1;
        if ((needle_am[n] >= 0) && (data[n] != needle_am[n])) return -1;
    }
    if (data[7] ^ data[8]) return -1;
    if (data[10] ^ data[11] ^ data[12] ^ data[13]) return -1;
    if (data[15] ^ data[16] ^ data[17] ^ data[18] ^ data[19] ^ data[20]) return -1;
    if (data[23] ^ data[24] ^ data[25] ^ data[26] ^ data[27] ^ data[28] ^ data[29] ^ data[30] ^ data[31]) return -1;
    int bc = (data[17] << 2) | (data[18] << 1) | data[19];
    if (bc == 0) {
        o->psmi = (data[26] << 4) | (data[27] << 3) | (data[28] << 2) | (data[29] << 1) | data[30];
        o->pli = data[7];
        o->hppi = data[11];
        o->aabi = data[12];
        o->rdbi = data[15];
    }
    return bc;
}

static int find_ref_am(orc_am_t *o, unsigned ref)                          /* sync.c:239-252 */
{
    unsigned char data[BLK];
    for (int n = 0; n < BLK; n++) data[n] = cimagf(o->bins[ref][n]) <= 0 ? 0 : 1;
    return fuzzy_match(needle_am, 23, data, BLK);
}

static void sync_block_am(orc_am_t *o)                                     /* sync.c:612-767 */
{
    for (int i = REF_IDX; i <= MAX_IDX; i++)
        for (int n = 0; n < BLK; n++)
            o->bins[CENTER - i][n] = -conjf(o->bins[CENTER - i][n]);
    if (o->psmi != MODE_MA3)                                               /* the mode known when the block starts */
        for (int i = REF_IDX; i <= PIDS_OUTER; i++)
            for (int n = 0; n < BLK; n++)
                o->bins[CENTER + i][n] += o->bins[CENTER - i][n];

    if (o->state == ST_COARSE && o->cfo_wait == 0) {
        int offset = find_ref_am(o, CENTER + REF_IDX);
        if (offset > 0) {
            o->keep_extra = ((BLK - offset) % BLK) * SYM_AM;
            o->cfo_wait = 8;
        }
    } else {
        o->cfo_wait--;
    }

    if (o->state == ST_COARSE) {
        int bc = find_block_am(o, CENTER + REF_IDX);
        if (bc == -1) o->offset_history = 0;
        else o->offset_history = (o->offset_history << 4) | (unsigned)bc;
        if ((o->offset_history & 0xffff) == 0x5670) {
            o->bc = 0;
            set_state(o, ST_FINE);
            decode_reset(o);
            o->offset_history = 0;
        }
    }

    if (o->state != ST_FINE) return;

    const int ma3 = o->psmi == MODE_MA3;
    const int pids1_index = !ma3 ? PIDS_INNER : -PIDS_INNER, pids2_index = !ma3 ? PIDS_OUTER : PIDS_INNER;   /* sync.c:670-671 */
    const cf pids1_mult = 2 * CMPLXF(1.5, -0.5) / (o->bins[CENTER + pids1_index][8] + o->bins[CENTER + pids1_index][24]);
    const cf pids2_mult = 2 * CMPLXF(1.5, -0.5) / (o->bins[CENTER + pids2_index][8] + o->bins[CENTER + pids2_index][24]);
    uint8_t pids[2 * BLK];
    int pids_out = 0;
    for (int n = 0; n < BLK; n++) {
        o->bins[CENTER + pids1_index][n] *= pids1_mult;
        pids[pids_out++] = qam16(o->bins[CENTER + pids1_index][n]);
        o->bins[CENTER + pids2_index][n] *= pids2_mult;
        pids[pids_out++] = qam16(o->bins[CENTER + pids2_index][n]);
    }
    process_pids(o, pids);

    /* sync.c:694-696: carrier index of column 0 and direction of each partition */
    const int primary = !ma3 ? OUTER_START : INNER_START, secondary = MIDDLE_START;
    const int tertiary = !ma3 ? INNER_START : MIDDLE_START, tdir = !ma3 ? 1 : -1;
    cf pl_mult[PW_AM], pu_mult[PW_AM], s_mult[PW_AM], t_mult[PW_AM];
    float samperr = 0;
    for (int col = 0; col < PW_AM; col++) {
        int train1 = (5 + 11 * col) % 32, train2 = (21 + 11 * col) % 32;
        const int ipl = CENTER - primary - col, ipu = CENTER + primary + col, is = CENTER + secondary + col,
                  it = CENTER + tdir * (tertiary + col);
        pl_mult[col] = 2 * CMPLXF(2.5, -2.5) / (o->bins[ipl][train1] + o->bins[ipl][train2]);
        pu_mult[col] = 2 * CMPLXF(2.5, -2.5) / (o->bins[ipu][train1] + o->bins[ipu][train2]);
        if (!ma3) {
            s_mult[col] = 2 * CMPLXF(1.5, -0.5) / (o->bins[is][train1] + o->bins[is][train2]);
            t_mult[col] = 2 * CMPLXF(-0.5, 0.5) / (o->bins[it][train1] + o->bins[it][train2]);
        } else {
            s_mult[col] = 2 * CMPLXF(2.5, -2.5) / (o->bins[is][train1] + o->bins[is][train2]);
            t_mult[col] = 2 * CMPLXF(2.5, -2.5) / (o->bins[it][train1] + o->bins[it][train2]);
        }
        if (col > 0) {
            samperr += phase_diff(cargf(pl_mult[col]), cargf(pl_mult[col - 1]));
            samperr += phase_diff(cargf(pu_mult[col]), cargf(pu_mult[col - 1]));
        }
    }
    samperr = samperr / (2 * (PW_AM - 1)) * FFT_AM / (2 * M_PI);
    o->samperr = roundf(samperr);

    uint8_t pl[BLK * PW_AM], pu[BLK * PW_AM], s[BLK * PW_AM], t[BLK * PW_AM];
    for (int n = 0; n < BLK; n++)
        for (int col = 0; col < PW_AM; col++) {
            const int ipl = CENTER - primary - col, ipu = CENTER + primary + col, is = CENTER + secondary + col,
                      it = CENTER + tdir * (tertiary + col);
            o->bins[ipl][n] *= pl_mult[col];
            o->bins[ipu][n] *= pu_mult[col];
            o->bins[is][n] *= s_mult[col];
            o->bins[it][n] *= t_mult[col];
            pl[n * PW_AM + col] = qam64(o->bins[ipl][n]);
            pu[n * PW_AM + col] = qam64(o->bins[ipu][n]);
            s[n * PW_AM + col] = !ma3 ? qam16(o->bins[is][n]) : qam64(o->bins[is][n]);
            t[n * PW_AM + col] = !ma3 ? qpsk(o->bins[it][n]) : qam64(o->bins[it][n]);
        }
    /* decode_push_pl_pu_s_t (decode.c:439-449) */
    memcpy(o->buffer_pl + o->bc * BLK * PW_AM, pl, BLK * PW_AM);
    memcpy(o->buffer_pu + o->bc * BLK * PW_AM, pu, BLK * PW_AM);
    memcpy(o->buffer_s + o->bc * BLK * PW_AM, s, BLK * PW_AM);
    memcpy(o->buffer_t + o->bc * BLK * PW_AM, t, BLK * PW_AM);
    process_p1_p3(o, o->bc);
    o->bc = (o->bc + 1) % 8;
}

/* ------------------------------------------------------------------------ */
/* acquisition + demodulation (reference src/acquire.c:98-263)               */
/* ------------------------------------------------------------------------ */
static void symbol_fft(orc_am_t *o, int sym, int samperr, cf *phase, cf inc)   /* acquire.c:178-195, 237-256 */
{
    const int offset = (FFT_AM - CP_AM) / 2;
    for (int j = 0; j < SYM_AM; ++j) {
        cf sample = *phase * o->buf[sym * SYM_AM + j + samperr];
        if (j < CP_AM) o->fin[(j + offset) % FFT_AM] = o->shape[j] * sample;
        else if (j < FFT_AM) o->fin[(j + offset) % FFT_AM] = sample;
        else o->fin[(j + offset) % FFT_AM] += o->shape[j] * sample;
        *phase *= inc;
    }
    *phase /= cabsf(*phase);
    fftwf_execute(o->plan);
}

static inline cf shifted(const orc_am_t *o, int bin)                       /* fftshift, defines.h:123-138 */
{
    return o->fout[(bin + FFT_AM / 2) % FFT_AM];
}

static void process_window(orc_am_t *o)
{
    int samperr = 0;
    float angle, angle_diff;
    cf max_v = 0;
    float max_mag = -1.0f;

    if (o->state == ST_FINE) {
        samperr = SYM_AM / 2 + o->samperr;
        o->samperr = 0;
        angle_diff = -o->angle;
        o->angle = 0;
        angle = o->prev_angle + angle_diff;
        o->prev_angle = angle;
    } else {
        for (int i = 0; i < NACQ_AM; i++) {
            memmove(o->bp_r, o->bp_r + 1, 31 * sizeof(int16_t));
            memmove(o->bp_i, o->bp_i + 1, 31 * sizeof(int16_t));
            o->bp_r[31] = o->win_r[i];
            o->bp_i[31] = o->win_i[i];
            int16_t yr = bp_axis(o->bp_r, o->bp_tap), yi = bp_axis(o->bp_i, o->bp_tap);
            o->buf[i] = CMPLXF((float)yr / 32767.0f, (float)yi / 32767.0f);
        }
        memset(o->sums, 0, sizeof(o->sums));
        for (int i = 0; i < SYM_AM; ++i)
            for (int j = 0; j < BLK; ++j)
                o->sums[i] += o->buf[i + j * SYM_AM] * conjf(o->buf[i + j * SYM_AM + FFT_AM]);
        for (int i = 0; i < SYM_AM; ++i) {
            cf v = 0;
            for (int j = 0; j < CP_AM; ++j)
                v += o->sums[(i + j) % SYM_AM] * o->shape[j] * o->shape[j + FFT_AM];
            float mag = crealf(v) * crealf(v) + cimagf(v) * cimagf(v);
            if (mag > max_mag) {
                max_mag = mag;
                max_v = v;
                samperr = (i + SYM_AM - 15) % SYM_AM;
            }
        }
        angle_diff = cargf(max_v * cexpf(I * -o->prev_angle));
        float factor = (o->prev_angle) ? 0.25 : 1.0;
        angle = o->prev_angle + (angle_diff * factor);
        o->prev_angle = angle;
        set_state(o, ST_COARSE);
    }

    for (int i = 0; i < NACQ_AM; i++)
        o->buf[i] = CMPLXF((float)o->win_r[i] / 32767.0f, (float)o->win_i[i] / 32767.0f);

    angle -= 2 * M_PI * o->cfo;
    o->phase *= cexpf(-(SYM_AM / 2 - samperr) * angle / FFT_AM * I);
    cf phase_increment = cexpf(angle / FFT_AM * I);

    /* AM only: carrier phase slope over the block and, while acquiring, the strongest bin (acquire.c:170-235) */
    {
        float y = 0, sum_y = 0, sum_xy = 0, sum_x2 = 0;
        cf last_carrier = 0;
        cf temp_phase = o->phase;
        float mag_sums[FFT_AM] = { 0 };
        for (int i = 0; i < BLK; ++i) {
            symbol_fft(o, i, samperr, &temp_phase, phase_increment);
            float x = SYM_AM * (i - (float)(BLK - 1) / 2);
            if (i == 0) y = cargf(shifted(o, CENTER));
            else y += cargf(shifted(o, CENTER) / last_carrier);
            last_carrier = shifted(o, CENTER);
            sum_y += y;
            sum_xy += x * y;
            sum_x2 += x * x;
            if (o->state != ST_FINE)
                for (int j = CENTER - PIDS_OUTER; j <= CENTER + PIDS_OUTER; j++)
                    mag_sums[j] += cabsf(shifted(o, j));
        }
        if (o->state != ST_FINE) {
            float mm = -1.0f;
            int max_index = -1;
            for (int j = CENTER - PIDS_OUTER; j <= CENTER + PIDS_OUTER; j++)
                if (mag_sums[j] > mm) {
                    mm = mag_sums[j];
                    max_index = j;
                }
            o->cfo += max_index - CENTER;
        }
        phase_increment *= cexpf(-sum_xy / sum_x2 * I);
        o->phase *= cexpf((-sum_y / BLK + (sum_xy / sum_x2) * (BLK) * SYM_AM / 2 - 0.06) * I);
    }

    for (int i = 0; i < BLK; ++i) {
        symbol_fft(o, i, samperr, &o->phase, phase_increment);
        for (int b = CENTER - MAX_IDX; b <= CENTER + MAX_IDX; b++)         /* sync_push, sync.c:793-797 */
            o->bins[b][o->sym] = shifted(o, b);
        if (++o->sym == BLK) {
            o->sym = 0;
            sync_block_am(o);
        }
    }

    int keep = SYM_AM + (SYM_AM / 2 - samperr) + o->keep_extra;
    o->keep_extra = 0;
    memmove(o->win_r, o->win_r + (NACQ_AM - keep), sizeof(int16_t) * (size_t)keep);
    memmove(o->win_i, o->win_i + (NACQ_AM - keep), sizeof(int16_t) * (size_t)keep);
    o->fill = (unsigned)keep;
}

/* ------------------------------------------------------------------------ */
/* public                                                                    */
/* ------------------------------------------------------------------------ */
orc_am_t *orc_am_new(void)
{
    orc_am_t *o = (orc_am_t *)calloc(1, sizeof(*o));
    for (int i = 0; i < 32; i++) o->bp_tap[i] = (int16_t)(bp_coeff_am[31 - i] * 32767.0f);
    for (int i = 0; i < SYM_AM; i++) {
        if (i < CP_AM) o->shape[i] = sinf(M_PI / 2 * i / CP_AM);
        else if (i < FFT_AM) o->shape[i] = 1;
        else o->shape[i] = cosf(M_PI / 2 * (i - FFT_AM) / CP_AM);
    }
    o->fin = fftwf_alloc_complex(FFT_AM);
    o->fout = fftwf_alloc_complex(FFT_AM);
    o->plan = fftwf_plan_dft_1d(FFT_AM, o->fin, o->fout, FFTW_FORWARD, FFTW_ESTIMATE);
    {
        /* reference src/input.c:33-38, reversed and truncated to int16 as in src/firdecim_q15.c:37-41 */
        static const float t[4] = { 0.6062333583831787f, -0.13481467962265015f, 0.032919470220804214f, -0.00410953676328063f };
        for (int i = 0; i < 4; i++) o->hb_tap[i] = (int16_t)(t[3 - i] * 32767.0f);
    }
    o->phase = 1;
    o->psmi = 1;
    o->pli = o->hppi = o->aabi = o->rdbi = -1;
    o->state = ST_NONE;
    decode_reset(o);
    return o;
}

void orc_am_free(orc_am_t *o)
{
    if (!o) return;
    fftwf_destroy_plan(o->plan);
    fftwf_free(o->fin);
    fftwf_free(o->fout);
    free(o->log.p);
    free(o);
}

/* mirrors input_push_cs16 (reference src/input.c:119-124) in AM mode; nvalues counts int16 values */
void orc_am_push_cs16(orc_am_t *o, const int16_t *buf, size_t nvalues)
{
    for (size_t n = 0; n + 1 < nvalues; n += 2) {
        o->win_r[o->fill] = buf[n];
        o->win_i[o->fill] = buf[n + 1];
        if (++o->fill == NACQ_AM)
            process_window(o);
    }
}

/* halfband decimator by 2 (reference src/firdecim_q15.c:137-165): w[0..14], w[14] newest; the output is taken
 * after the even sample of a pair went in, the odd one follows */
static inline int16_t hb_axis(const int16_t *w, const int16_t *tap)
{
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

static void hb_execute(orc_am_t *o, int stage, const int16_t xr[2], const int16_t xi[2], int16_t *yr, int16_t *yi)
{
    hb_shift(o->hb_r[stage], xr[0]);
    hb_shift(o->hb_i[stage], xi[0]);
    *yr = hb_axis(o->hb_r[stage], o->hb_tap);
    *yi = hb_axis(o->hb_i[stage], o->hb_tap);
    hb_shift(o->hb_r[stage], xr[1]);
    hb_shift(o->hb_i[stage], xi[1]);
}

/* mirrors input_push_cu8 in AM mode (reference src/input.c:52-117): (u8 - 127) * 64 >> 4, then /32 through five
 * halfband stages with ping-pong pair buffers; nbytes counts uint8 values and is a multiple of 4 */
void orc_am_push_cu8(orc_am_t *o, const uint8_t *buf, size_t nbytes)
{
    for (size_t n = 0; n + 3 < nbytes; n += 4) {
        int16_t xr[2], xi[2], yr, yi;
        xr[0] = (int16_t)((((int16_t)buf[n + 0] - 127) * 64) >> 4);
        xi[0] = (int16_t)((((int16_t)buf[n + 1] - 127) * 64) >> 4);
        xr[1] = (int16_t)((((int16_t)buf[n + 2] - 127) * 64) >> 4);
        xi[1] = (int16_t)((((int16_t)buf[n + 3] - 127) * 64) >> 4);
        const unsigned off = o->offset++;
        hb_execute(o, 0, xr, xi, &o->stage_r[0][off & 1], &o->stage_i[0][off & 1]);
        int produced = 0;
        for (int st = 1; st < 5; st++) {
            const unsigned mask = (1u << st) - 1;
            if ((off & mask) != mask) break;
            if (st < 4)
                hb_execute(o, st, o->stage_r[st - 1], o->stage_i[st - 1], &o->stage_r[st][(off >> st) & 1], &o->stage_i[st][(off >> st) & 1]);
            else {
                hb_execute(o, 4, o->stage_r[3], o->stage_i[3], &yr, &yi);
                produced = 1;
            }
        }
        if (produced) {
            o->win_r[o->fill] = yr;
            o->win_i[o->fill] = yi;
            if (++o->fill == NACQ_AM)
                process_window(o);
        }
    }
}

/* the decimator alone (kernel-level tests): npairs*... cu8 complex samples -> nbytes/64 cs16 samples */
size_t orc_am_decimate(const uint8_t *cu8, size_t nbytes, int16_t *out)
{
    orc_am_t *o = (orc_am_t *)calloc(1, sizeof(*o));
    static const float t[4] = { 0.6062333583831787f, -0.13481467962265015f, 0.032919470220804214f, -0.00410953676328063f };
    for (int i = 0; i < 4; i++) o->hb_tap[i] = (int16_t)(t[3 - i] * 32767.0f);
    size_t nout = 0;
    for (size_t n = 0; n + 3 < nbytes; n += 4) {
        int16_t xr[2], xi[2], yr = 0, yi = 0;
        xr[0] = (int16_t)((((int16_t)cu8[n + 0] - 127) * 64) >> 4);
        xi[0] = (int16_t)((((int16_t)cu8[n + 1] - 127) * 64) >> 4);
        xr[1] = (int16_t)((((int16_t)cu8[n + 2] - 127) * 64) >> 4);
        xi[1] = (int16_t)((((int16_t)cu8[n + 3] - 127) * 64) >> 4);
        const unsigned off = o->offset++;
        hb_execute(o, 0, xr, xi, &o->stage_r[0][off & 1], &o->stage_i[0][off & 1]);
        for (int st = 1; st < 5; st++) {
            const unsigned mask = (1u << st) - 1;
            if ((off & mask) != mask) break;
            if (st < 4)
                hb_execute(o, st, o->stage_r[st - 1], o->stage_i[st - 1], &o->stage_r[st][(off >> st) & 1], &o->stage_i[st][(off >> st) & 1]);
            else {
                hb_execute(o, 4, o->stage_r[3], o->stage_i[3], &yr, &yi);
                out[2 * nout] = yr;
                out[2 * nout + 1] = yi;
                nout++;
            }
        }
    }
    free(o);
    return nout;
}

size_t orc_am_log_size(const orc_am_t *o) { return o->log.len; }
const uint8_t *orc_am_log_data(const orc_am_t *o) { return o->log.p; }
