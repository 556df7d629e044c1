/* reftap — test-infrastructure harness around the UNMODIFIED reference
 * library compiled from /root/reference/src into oracle/_ref/ (see Makefile).
 *
 * It drives the reference through its public pipe API exactly as
 * reference src/main.c:1097-1119 does (nrsc5_open_pipe -> nrsc5_set_mode ->
 * nrsc5_set_callback -> nrsc5_pipe_samples_cu8 in 32768-byte pushes ->
 * nrsc5_close) and records, in call order, a binary log of
 *   - every L1 PDU handed to L2: frame_push() (reference src/frame.c:645) and
 *     pids_frame_push() (reference src/pids.c:1032), intercepted with
 *     -Wl,--wrap so the reference objects themselves are untouched;
 *   - optionally every block of soft bits entering decode_push_pm()
 *     (reference src/decode.c:378);
 *   - the public events SYNC / LOST_SYNC / MER / BER / HDC.
 * Nothing here is part of the product; only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may load the resulting library.
 *
 * Log record: u32 type, u32 payload_len, payload (padded to 4 bytes).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <nrsc5.h>

enum {
    REC_FRAME = 1,     /* payload: u32 lc, u32 nbits, bits packed MSB-first       */
    REC_PIDS = 2,      /* payload: 10 bytes (80 bits packed MSB-first)            */
    REC_SYNC = 3,      /* payload: f32 freq_offset, i32 psmi, pli, hppi, aabi, rdbi */
    REC_LOST_SYNC = 4, /* payload: none                                           */
    REC_MER = 5,       /* payload: f32 lower, f32 upper                           */
    REC_BER = 6,       /* payload: f32 cber                                       */
    REC_HDC = 7,       /* payload: u32 program, u32 count, bytes                  */
    REC_SOFT_PM = 8,   /* payload: u32 bc, 23040 int8 soft bits                   */
};

static __thread int tls_logging = 0; /* only the thread running reftap_decode logs */
static uint8_t *g_log;
static size_t g_len, g_cap;
static int g_want_soft;

static void log_put(uint32_t type, const void *a, size_t alen, const void *b, size_t blen)
{
    if (!tls_logging)
        return;
    size_t plen = alen + blen;
    size_t need = 8 + ((plen + 3) & ~(size_t)3);
    if (g_len + need > g_cap) {
        size_t ncap = g_cap ? g_cap * 2 : (1u << 20);
        while (ncap < g_len + need)
            ncap *= 2;
        g_log = (uint8_t *)realloc(g_log, ncap);
        g_cap = ncap;
    }
    uint32_t hdr[2] = { type, (uint32_t)plen };
    memcpy(g_log + g_len, hdr, 8);
    if (alen) memcpy(g_log + g_len + 8, a, alen);
    if (blen) memcpy(g_log + g_len + 8 + alen, b, blen);
    memset(g_log + g_len + 8 + plen, 0, need - 8 - plen);
    g_len += need;
}

/* for reftap_l2.c (the L2 -> L3 taps share this log) */
void reftap_log_put(uint32_t type, const void *a, size_t alen, const void *b, size_t blen) { log_put(type, a, alen, b, blen); }
void reftap_set_logging(int on) { tls_logging = on; }

/* ---- link-time taps (-Wl,--wrap=...) ---- */
struct frame_t;
struct pids_t;
struct decode_t;
void __real_frame_push(struct frame_t *st, uint8_t *bits, size_t length, int lc);
void __real_pids_frame_push(struct pids_t *st, const uint8_t *bits);
void __real_decode_push_pm(struct decode_t *st, const int8_t *sbit, unsigned int bc);

static size_t pack_bits(const uint8_t *bits, size_t n, uint8_t *out)
{
    size_t nb = (n + 7) / 8;
    memset(out, 0, nb);
    for (size_t i = 0; i < n; i++)
        out[i >> 3] |= (uint8_t)((bits[i] & 1) << (7 - (i & 7)));
    return nb;
}

void __wrap_frame_push(struct frame_t *st, uint8_t *bits, size_t length, int lc)
{
    if (tls_logging) {
        uint8_t *packed = (uint8_t *)malloc((length + 7) / 8);
        size_t nb = pack_bits(bits, length, packed);
        uint32_t hdr[2] = { (uint32_t)lc, (uint32_t)length };
        log_put(REC_FRAME, hdr, sizeof(hdr), packed, nb);
        free(packed);
    }
    __real_frame_push(st, bits, length, lc);
}

void __wrap_pids_frame_push(struct pids_t *st, const uint8_t *bits)
{
    if (tls_logging) {
        uint8_t packed[10];
        pack_bits(bits, 80, packed);
        log_put(REC_PIDS, packed, 10, NULL, 0);
    }
    __real_pids_frame_push(st, bits);
}

void __wrap_decode_push_pm(struct decode_t *st, const int8_t *sbit, unsigned int bc)
{
    if (tls_logging && g_want_soft) {
        uint32_t b = bc;
        log_put(REC_SOFT_PM, &b, 4, sbit, 23040);
    }
    __real_decode_push_pm(st, sbit, bc);
}

/* ---- public-callback recorder ---- */
static void on_event(const nrsc5_event_t *evt, void *opaque)
{
    (void)opaque;
    switch (evt->event) {
    case NRSC5_EVENT_SYNC: {
        struct { float f; int32_t v[5]; } p = { evt->sync.freq_offset,
            { evt->sync.psmi, evt->sync.pli, evt->sync.hppi, evt->sync.aabi, evt->sync.rdbi } };
        log_put(REC_SYNC, &p, sizeof(p), NULL, 0);
        break;
    }
    case NRSC5_EVENT_LOST_SYNC:
        log_put(REC_LOST_SYNC, NULL, 0, NULL, 0);
        break;
    case NRSC5_EVENT_MER: {
        float p[2] = { evt->mer.lower, evt->mer.upper };
        log_put(REC_MER, p, sizeof(p), NULL, 0);
        break;
    }
    case NRSC5_EVENT_BER: {
        float p = evt->ber.cber;
        log_put(REC_BER, &p, sizeof(p), NULL, 0);
        break;
    }
    case NRSC5_EVENT_HDC: {
        uint32_t hdr[2] = { evt->hdc.program, (uint32_t)evt->hdc.count };
        log_put(REC_HDC, hdr, sizeof(hdr), evt->hdc.data, evt->hdc.count);
        break;
    }
    default:
        break;
    }
}

/* ---- C API for ctypes ---- */
void reftap_reset(void) { g_len = 0; }
void reftap_want_soft(int on) { g_want_soft = on; }
size_t reftap_log_size(void) { return g_len; }
const uint8_t *reftap_log_data(void) { return g_log; }

static int run_one(const void *buf, size_t nvals, int mode, int is_cs16, size_t chunk, int with_cb)
{
    nrsc5_t *st = NULL;
    if (nrsc5_open_pipe(&st) != 0)
        return -1;
    nrsc5_set_mode(st, mode);
    if (with_cb)
        nrsc5_set_callback(st, on_event, NULL);
    if (is_cs16) {
        const int16_t *p = (const int16_t *)buf;
        for (size_t off = 0; off < nvals; off += chunk) {
            size_t n = nvals - off < chunk ? nvals - off : chunk;
            nrsc5_pipe_samples_cs16(st, p + off, (unsigned int)n);
        }
    } else {
        const uint8_t *p = (const uint8_t *)buf;
        for (size_t off = 0; off < nvals; off += chunk) {
            size_t n = nvals - off < chunk ? nvals - off : chunk;
            nrsc5_pipe_samples_cu8(st, p + off, (unsigned int)n);
        }
    }
    nrsc5_close(st);
    return 0;
}

/* Decode one capture held in RAM, appending to the log. nvals counts uint8
 * values (cu8) or int16 values (cs16), as the reference API does. */
int reftap_decode(const void *buf, size_t nvals, int mode, int is_cs16, size_t chunk)
{
    tls_logging = 1;
    int rc = run_one(buf, nvals, mode, is_cs16, chunk ? chunk : 32768, 1);
    tls_logging = 0;
    return rc;
}

/* ---- timing of the reference's own CPU path (bench.py reference arm) ---- */
struct bench_arg {
    const void *buf;
    size_t nvals;
    int mode, is_cs16, reps;
};

static void *bench_thread(void *p)
{
    struct bench_arg *a = (struct bench_arg *)p;
    for (int r = 0; r < a->reps; r++)
        run_one(a->buf, a->nvals, a->mode, a->is_cs16, a->is_cs16 ? 16384 : 32768, 0);
    return NULL;
}

/* nstreams independent sessions (one thread each, own nrsc5_t), each decoding
 * bufs[i] reps times. Returns wall seconds. */
double reftap_bench(const void *const *bufs, const size_t *nvals, int nstreams, int mode, int is_cs16, int reps)
{
    pthread_t *th = (pthread_t *)calloc(nstreams, sizeof(*th));
    struct bench_arg *args = (struct bench_arg *)calloc(nstreams, sizeof(*args));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int i = 0; i < nstreams; i++) {
        args[i] = (struct bench_arg){ bufs[i], nvals[i], mode, is_cs16, reps };
        pthread_create(&th[i], NULL, bench_thread, &args[i]);
    }
    for (int i = 0; i < nstreams; i++)
        pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th);
    free(args);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
