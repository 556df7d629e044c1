// AM (hybrid MA1, all-digital MA3) receive chain.  One CTA of AM_THREADS threads per stream (k_am, engine.cu) runs the
// whole chain block after block: coarse acquisition, carrier-phase tracking, 256-point OFDM demodulation, reference-
// carrier search, training-symbol equalisation, QAM/QPSK slicing, PIDS, interleaver MA1 with the diversity delay,
// K=9 tail-biting Viterbi (E1/E2/E3), descramble, L1 PDUs.
//
//   reference: src/acquire.c:98-263 (AM branches), src/sync.c:37-88,208-252,612-767, src/decode.c:67-231,
//              234-277,474-554, src/conv_dec.c (K=9), src/frame.c:645-714,527-541 (sync-loss predicate)
//
// AM decides on HARD symbols (QAM-64 / 16 / QPSK slicing, sync.c:37-88), so a float that differs in its last bits can
// flip a sliced bit and with it the channel-BER figure the reference reports: the float arithmetic here keeps the
// reference's order of operations throughout - including the per-sample NCO recurrence and the radix-2 butterfly
// order of the test oracle's FFT.  What is re-designed is where the work runs and what it waits for:
//   * the working set of a block lives in shared memory (AmSmem): the symbol being demodulated and its FFT, the K=9
//     path metrics, a tile of survivor decisions - nothing on a dependent path goes to global memory;
//   * the NCO phase chain (8 640 dependent complex multiplications per pass - sequential by definition) is run by ONE
//     warp, a symbol ahead of the three warps that window, fold and transform the previous symbol (demod_pass);
//   * the K=9 recursion is ONE WARP with the 256 path metrics in its registers, three trellis steps between two metric
//     exchanges and two butterflies per instruction (radix 8 on packed 16-bit pairs, viterbi_k9_warp) - no CTA barrier
//     in it; P1 and P3 of a frame's last block run side by side, a warp each.  Traceback is cut into 128 segments
//     walked concurrently from warmed-up start states and repaired until it equals the sequential one
//     (viterbi_k9_traceback);
//   * channel-BER re-encoding and the bit packing of the PDUs are spread over the CTA.
// Scalar receiver state (AmState) is held redundantly by every thread - each follows the same control flow on the
// same values - and written back by thread 0.
//
// The generic (host) branches of the functions below are the same algorithms in plain loops: tests/am_host.cu runs
// them on the CPU against the oracle; the device branches run under the CUDA emulator (tests/test_emu_engine.py)
// and on the GPU against the same oracle.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cuda_runtime.h>

#include "pids_crc.cuh"

#if defined(__CUDA_ARCH__)
#define AM_SYNC() __syncthreads()         // the lanes of a stream are the threads of its CTA
#elif defined(AM_HOST_SYNC)
#define AM_SYNC() AM_HOST_SYNC()          // tests/am_host.cu: lanes emulated by host threads meet at a barrier
#else
#define AM_SYNC() ((void)0)
#endif
// barrier of the cu8 front end's thread block (k_am_decim); the host harness uses the same fibre barrier
#if defined(__CUDA_ARCH__)
#define AM_BLOCK_SYNC() __syncthreads()
#elif defined(AM_HOST_SYNC)
#define AM_BLOCK_SYNC() AM_HOST_SYNC()
#else
#define AM_BLOCK_SYNC() ((void)0)
#endif
#define AM_HD __host__ __device__
// input samples may land (asynchronous pushes) while a kernel runs: on the device they are read through L2 only
#if defined(__CUDA_ARCH__)
#define AM_LDIN(ptr) __ldcg(ptr)
#else
#define AM_LDIN(ptr) (*(ptr))
#endif

namespace nbam {

constexpr int FFT = 256, CP = 14, SYM = FFT + CP, BLK = 32, NACQ = SYM * (BLK + 1);
constexpr int CENTER = 128, REF_IDX = 1, PIDS_INNER = 27, PIDS_OUTER = 53, INNER_START = 2, MIDDLE_START = 28,
              OUTER_START = 57, MAX_IDX = 81, PW = 25;
constexpr int P1_LEN = 3750, P3_LEN = 24000, P3_LEN_MA3 = 30000, PIDS_LEN = 80, DIVERSITY = 18000 * 3;
constexpr int MODE_MA3 = 2;                     // SERVICE_MODE_MA3, defines.h:39; every other psmi is decoded as MA1
constexpr int ST_NONE = 0, ST_COARSE = 1, ST_FINE = 2;
constexpr int VIT_MAX_STEPS = P3_LEN_MA3 + 64;
constexpr uint32_t REC_FRAME = 1, REC_PIDS = 2, REC_SYNC = 3, REC_LOST_SYNC = 4, REC_BER = 6;
constexpr double PI = 3.14159265358979323846;

constexpr int AM_THREADS = 128;                 // k_am: threads per stream (one K=9 butterfly each)

struct Lanes {
    int lane, n;
    void *smem;                                 // device: the CTA's AmSmem
    int dbg;                                    // experiment switches (nrsc5b_debug_set), 0 in production
};

// ---- complex helpers in the reference's (gcc, no FMA) evaluation order ----
AM_HD inline float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
AM_HD inline float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
AM_HD inline float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
AM_HD inline float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
AM_HD inline float2 cexpj(float a)
{
    float s, c;
    sincosf(a, &s, &c);
    return make_float2(c, s);
}
AM_HD inline float cabs2(float2 a) { return hypotf(a.x, a.y); }
AM_HD inline float carg(float2 a) { return atan2f(a.y, a.x); }
// complex division the way libgcc's __divsc3 does it (Smith's algorithm)
AM_HD inline float2 cdiv(float2 a, float2 b)
{
    float ratio, denom;
    if (fabsf(b.x) < fabsf(b.y)) {
        ratio = b.x / b.y;
        denom = (b.x * ratio) + b.y;
        return make_float2(((a.x * ratio) + a.y) / denom, ((a.y * ratio) - a.x) / denom);
    }
    ratio = b.y / b.x;
    denom = (b.y * ratio) + b.x;
    return make_float2(((a.y * ratio) + a.x) / denom, (a.y - (a.x * ratio)) / denom);
}

constexpr int AM_L2_QUEUE = 40;            // = nb::L2_QUEUE: a launch with L2 on takes at most 16 blocks (1 P1 + 1 P3 each)
constexpr int AM_L2_BLOCKS = 16;

// Per-stream scalars (reference src/acquire.h, src/sync.h, src/decode.h).
struct AmState {
    long long in_avail;        // cs16 complex samples available from absolute index 0 (advanced by the host)
    long long start;           // sample index of the acquisition window's first sample
    float prev_angle;
    float2 phase;
    int keep_extra, cfo, state;
    int psmi, pli, hppi, aabi, rdbi, cfo_wait, samperr;
    unsigned bc, offset_history;
    float angle;
    int am_errors, am_diversity_wait;
    unsigned log_len, log_overflow;
    unsigned long long blocks_done;
    int l2_on, l2_n;           // L2 on the device (l2.cuh): entries in AmWork::l2_* this launch handed to the L2 kernel
};

// Per-stream arrays.
struct AmWork {
    short bp_hist[31][2];      // the coarse band-pass filter's last 31 inputs
    // L2 on the device: the frames (log offset of their packed bits) and frame_resets (nbits == 0) this launch handed
    // to the L2 kernel that follows it, in the reference's call order
    unsigned l2_off[AM_L2_QUEUE], l2_lc[AM_L2_QUEUE], l2_nbits[AM_L2_QUEUE];
    float2 buf[NACQ];
    float2 sums[SYM];
    float2 fft[FFT];
    float2 spec[FFT];                          // one demodulated symbol, fftshift-ed
    float2 bins[FFT][BLK];
    uint8_t buffer_pl[PW * BLK * 8], buffer_pu[PW * BLK * 8], buffer_s[PW * BLK * 8], buffer_t[PW * BLK * 8];
    uint8_t bl[18000], bu[18000], el[12000], eu[24000];
    alignas(16) uint8_t ml[DIVERSITY + 18000];     // (16-byte aligned: the diversity delay moves them 16 bytes at a time)
    alignas(16) uint8_t mu[DIVERSITY + 18000];
    uint8_t ebl[18000], ebu[18000];                // MA3 (decode.h:49-52)
    alignas(16) uint8_t eml[DIVERSITY + 18000];
    alignas(16) uint8_t emu[DIVERSITY + 18000];
    uint8_t p1_am[8 * 9000], p3_am[72000];
    int8_t vit_p1[8 * P1_LEN * 3], vit_p3[P3_LEN_MA3 * 3], vit_pids[PIDS_LEN * 3];
    uint8_t out[P3_LEN_MA3 + 8];
    uint8_t out_p1[P1_LEN + 10];               // P1 bits when P1 and P3 are decoded side by side (a frame's last block)
    alignas(16) uint8_t dec_p1[(size_t)((P1_LEN + 64 + 2) / 3) * 128];   // ... and its survivor bits
    short pm[2][256];
    unsigned long long ph_cyc[16];             // SM cycles per phase (thread 0): window + acquisition, first pass, second pass, sync +
                                               // slicing, PIDS, P1/P3 group, of it P3's post-processing, interleaver; K=9 recursion,
                                               // traceback (all decodes), window load of a FINE block, spare
    alignas(16) uint8_t dec[(size_t)VIT_MAX_STEPS * 48];   // survivor bits: host 32 bytes per trellis step, device 128 per three steps (viterbi_k9)
    float2 mult[4][PW];
    uint8_t sym_pl[BLK * PW], sym_pu[BLK * PW], sym_s[BLK * PW], sym_t[BLK * PW], sym_pids[2 * BLK];
};

struct AmTables {
    float shape[SYM];
    float2 tw[FFT / 2];        // exp(-2*pi*i*k/256)
    short bp_tap[32];          // coarse band-pass taps, reversed and truncated like src/firdecim_q15.c:37-41
    uint8_t brev[FFT];         // bit reversal of 8 bits
    uint8_t pn[P3_LEN_MA3 + 8];    // descrambler sequence
};

// every lane calls this with identical arguments; lane 0 writes the entry
AM_HD inline void l2_enqueue(AmState &st, AmWork &w, Lanes L, unsigned off, unsigned lc, unsigned nbits)
{
    if (!st.l2_on) return;
    if (st.l2_n >= AM_L2_QUEUE) {               // cannot happen with AM_L2_BLOCKS blocks per launch
        st.log_overflow = 1;
        return;
    }
    if (L.lane == 0) {
        w.l2_off[st.l2_n] = off;
        w.l2_lc[st.l2_n] = lc;
        w.l2_nbits[st.l2_n] = nbits;
    }
    st.l2_n++;
}

struct AmIo {
    const int16_t *iq;         // this stream's cs16 samples, I/Q interleaved
    uint8_t *log;
    unsigned log_cap;
};

#if defined(__CUDA_ARCH__)
#define AM_LAP(w, L, slot, t0)                                   \
    do {                                                         \
        if ((L).lane == 0) {                                     \
            const long long t1_ = clock64();                     \
            (w).ph_cyc[slot] += (unsigned long long)(t1_ - (t0)); \
            (t0) = t1_;                                          \
        }                                                        \
    } while (0)
#define AM_T0() clock64()
#else
#define AM_LAP(w, L, slot, t0) ((void)0)
#define AM_T0() 0
#endif

// ---- records ----
// Every lane calls this with identical arguments (each keeps its own copy of the cursor); lane 0 writes.
AM_HD inline uint8_t *log_reserve(AmState &st, const AmIo &io, Lanes L, uint32_t type, uint32_t plen)
{
    const uint32_t need = 8 + ((plen + 3) & ~3u);
    if ((size_t)st.log_len + need > io.log_cap) {
        st.log_overflow = 1;
        return nullptr;
    }
    uint8_t *w = io.log + st.log_len;
    if (L.lane == 0) {
        uint32_t hdr[2] = { type, plen };
        memcpy(w, hdr, 8);
        memset(w + 8, 0, need - 8);
    }
    st.log_len += need;
    return w + 8;
}

// All lanes call this with identical arguments.  `st` is every lane's private copy.
AM_HD inline void emit_frame(AmState &st, AmWork &w, const AmIo &io, Lanes L, const uint8_t *bits, unsigned len, unsigned lc)
{
    uint8_t *rec = log_reserve(st, io, L, REC_FRAME, 8 + (len + 7) / 8);
    AM_SYNC();                                                      // (lane 0 cleared the payload)
    if (rec) {
        if (L.lane == 0) {
            uint32_t hdr[2] = { lc, len };
            memcpy(rec, hdr, 8);
        }
        for (unsigned by = L.lane; by < (len + 7) / 8; by += L.n) {     // MSB first, a byte per lane
            unsigned v = 0;
            for (unsigned k = 0; k < 8 && 8 * by + k < len; k++) v |= (unsigned)(bits[8 * by + k] & 1) << (7 - k);
            rec[8 + by] = (uint8_t)v;
        }
    }
    l2_enqueue(st, w, L, rec ? (unsigned)(rec + 8 - io.log) : 0xffffffffu, lc, len);     // frame_push -> frame_process
}

AM_HD inline void set_state(AmState &st, const AmIo &io, Lanes L, int ns)       // input.c:172-188
{
    if (st.state == ns) return;
    if (st.state == ST_FINE) log_reserve(st, io, L, REC_LOST_SYNC, 0);
    if (ns == ST_FINE) {
        float fo = (float)(((double)st.prev_angle - 2 * PI * st.cfo) * 46511.71875 / (2 * PI * FFT));
        uint8_t *w = log_reserve(st, io, L, REC_SYNC, 24);
        if (w && L.lane == 0) {
            memcpy(w, &fo, 4);
            const int32_t v[5] = { st.psmi, st.pli, st.hppi, st.aabi, st.rdbi };    // nrsc5_report_sync, input.c:184
            memcpy(w + 4, v, 20);
        }
    }
    st.state = ns;
}

// ---- K=9 tail-biting Viterbi, rate 1/3 (reference src/conv_dec.c:359-453 with src/conv_gen.h) ----
// int16 metrics, minimum subtracted when step % 77 == 0, survivor = odd predecessor unless the even one is
// strictly better, first maximum at the end, 32 steps of pre- and post-roll.  Lane l owns butterflies
// 4l..4l+3; byte dec[step*32 + l] holds the survivor bits of new states 4l..4l+3 (low nibble) and
// 128+4l..128+4l+3 (high nibble).
AM_HD inline int parity9(unsigned v)
{
    v ^= v >> 8;
    v ^= v >> 4;
    v ^= v >> 2;
    v ^= v >> 1;
    return (int)(v & 1u);
}
AM_HD inline int sat16(int v) { return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }

// ---- shared memory of a stream's CTA (device) ----
constexpr int VT = 192;                          // trellis steps per tile of staged inputs (a multiple of three)
constexpr int VIT_WARMUP = 96;                   // traceback walkers start this many steps late, from any state
constexpr int VIT_CHUNK_WARMUP = 192;            // the recursion's chunks start this many steps early (a multiple of three)
constexpr int VIT_TEST_SLOTS_BYTES = 4 * 5 * 1024;               // >= (AM_THREADS / 32) * sizeof(AmVitSlot), a multiple of 16 (stage test kernel)
constexpr int VIT_TB_WORDS = 2 * (AM_THREADS / 32) * 8 * 8 * 32;  // sizeof(AmVitRows) / 4 (declared after the decoder)
struct AmVitSlot {                               // one K=9 decoder: the recursion by ONE warp (viterbi_k9_warp)
    alignas(16) unsigned short pm[256];          // path metrics between two groups of three steps (slot order, see below)
    alignas(16) unsigned short chk[256];         // a chunk's metrics after its warm-up (viterbi_k9_forward)
    alignas(16) uint32_t qd[VT][4];              // the tile's soft inputs, (q + 1) in both halves of a word
    unsigned short ends[AM_THREADS + 1];         // traceback: the state each walker arrived at
    unsigned state;                              // state after the last step (first maximum)
    int changed;
};
struct AmSmem {
    union {
        struct {
            AmVitSlot vit[AM_THREADS / 32];      // the recursion's chunks, one per warp (viterbi_k9_forward)
            uint32_t tb[VIT_TB_WORDS];           // traceback: rows of decision words per walker (AmVitRows)
        };
        struct {                                 // demod_pass
            float2 ph[2][3][SYM];                // NCO phase per sample of three symbols, double-buffered (producer warp runs ahead)
            float2 phase_end[2];                 // the phase after the block, renormalised
            float2 fft[3][FFT];                  // a symbol per consumer warp: windowed, folded, shifted; transformed in place
            float2 carrier[BLK];                 // first pass: the carrier bin of every symbol
            float magv[2][3][2 * PIDS_OUTER + 1];    // first pass while acquiring: a symbol's magnitudes around the carrier ...
            float mag[2 * PIDS_OUTER + 1];       // ... and their sums over the symbols, added in symbol order
        } dem;
    };
    int red[AM_THREADS / 32];                    // CTA-wide sums (bit_errors)
};

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ int am_warp_min(int v)
{
#if defined(NB_EMU)
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
#else
    return __reduce_min_sync(0xffffffffu, v);
#endif
}
__device__ __forceinline__ unsigned am_umax2(unsigned a, unsigned b)       // max.u16x2
{
#if defined(NB_EMU)
    const unsigned lo = max(a & 0xffffu, b & 0xffffu), hi = max(a >> 16, b >> 16);
    return lo | (hi << 16);
#else
    unsigned r;
    asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
}
__device__ __forceinline__ unsigned am_umin2(unsigned a, unsigned b)       // min.u16x2
{
#if defined(NB_EMU)
    const unsigned lo = min(a & 0xffffu, b & 0xffffu), hi = min(a >> 16, b >> 16);
    return lo | (hi << 16);
#else
    unsigned r;
    asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
}
__device__ __forceinline__ unsigned am_lo2(unsigned a, unsigned b) { return (a & 0xffffu) | (b << 16); }          // (a.lo, b.lo)
__device__ __forceinline__ unsigned am_hi2(unsigned a, unsigned b) { return (a >> 16) | (b & 0xffff0000u); }      // (a.hi, b.hi)
#endif

#if defined(__CUDA_ARCH__)
// The K=9 recursion in ONE warp, three trellis steps between two metric exchanges (radix 8), two butterflies per
// instruction (16-bit metrics packed in pairs).  Lane g takes the eight old states 8g..8g+7 into four registers and
// runs three add-compare-select stages on them without talking to anybody; the states it holds after stage 1, 2, 3 are
// {4g+i, 4g+128+i}, {2g+(i&1)+64(i>>1) (+128)}, {g+32j}.  Only then do the metrics go through shared memory (eight
// 2-byte stores, one 16-byte load, two __syncwarp) - no CTA barrier anywhere in the recursion.
//   * In every stage butterfly i of a lane reads (unpacked) values 2i, 2i+1 and produces i (input bit 0) and 4+i (input
//     bit 1).  A packed operation takes E = the even inputs and O = the odd inputs of two butterflies and yields their
//     two bit-0 outputs in one register and their two bit-1 outputs in another: pairing the butterflies (0,2),(1,3) in
//     stage 1 makes stage 1's four output registers exactly the E/O operands of stage 2 paired (0,1),(2,3); stage 3
//     needs four byte permutes.  The exchange buffer is laid out so that the 16-byte load delivers stage 1's operands
//     as they are: 16-bit slot 8 (s >> 3) + 2 (s & 3) + ((s >> 2) & 1) holds state s.
//   * Metrics are unsigned with a common drift: every branch adds 3 + m or 3 - m (m = the branch metric, |m| <= 3
//     because AM slices hard: the inputs are -1, 0, +1 - a precondition of this function), so nothing ever borrows
//     between the halves of a word; differences between states - all that decisions, the minimum subtraction every 77
//     steps (conv_dec.c:417-421) and the final "first maximum" look at - are those of the reference's int16 metrics.
//     3 + m is built per polynomial by a bitwise select between (q + 1) and (1 - q) under the lane's sign masks.
//   * The decision "odd predecessor unless the even one is strictly better" is bit 15 / 31 of Y + 0x80008000 - X; the 24
//     decisions of a lane per group go to global memory as one word: bit 4 stage + 2 (input bit) + op + 16 half.
// Semantics as conv_dec.c / conv_gen.h: survivor = odd predecessor unless the even one is strictly better, first
// maximum at the end, 32 steps of pre- and post-roll.
__device__ __forceinline__ int vit_butterfly(int stage, int g, int i)
{
    return stage == 0 ? 4 * g + i : stage == 1 ? 2 * g + (i & 1) + 64 * (i >> 1) : g + 32 * i;
}
__device__ __forceinline__ int vit_slot(int s) { return 8 * (s >> 3) + 2 * (s & 3) + ((s >> 2) & 1); }

// The groups [g_begin, g_end) of the frame (three steps each; g_end past the last group = to the end).  start_pm: the
// path metrics to start from, in exchange-buffer order (a previous chunk's sm.pm), or nullptr = all equal (the frame's
// start, or a chunk's warm-up).  Decisions are stored from group g_store on; the metrics as they stand before that group
// are copied to sm.chk (chunk verification, viterbi_k9_forward).
__device__ __noinline__ void viterbi_k9_warp(AmVitSlot &sm, uint32_t *decw, int g, const int8_t *in, int len, unsigned g0, unsigned g1, unsigned g2,
                                             int g_begin, int g_store, int g_end, const unsigned short *start_pm)
{
    const int steps = min(len + 64, 3 * g_end), interval = 32767 / (3 * 127) - 9;
    // sign masks: mk[stage][op][poly], 0xffff in a half whose butterfly has code bit 1 on its (even state, input 0) branch
    unsigned mk[3][2][3];
#pragma unroll
    for (int sgi = 0; sgi < 3; sgi++)
#pragma unroll
        for (int op = 0; op < 2; op++) {
            const int iA = sgi == 0 ? op : 2 * op, iB = sgi == 0 ? op + 2 : 2 * op + 1;     // pairing (0,2),(1,3) | (0,1),(2,3)
            const unsigned rA = (unsigned)vit_butterfly(sgi, g, iA) << 1, rB = (unsigned)vit_butterfly(sgi, g, iB) << 1;
            mk[sgi][op][0] = (parity9(rA & g0) ? 0xffffu : 0u) | (parity9(rB & g0) ? 0xffff0000u : 0u);
            mk[sgi][op][1] = (parity9(rA & g1) ? 0xffffu : 0u) | (parity9(rB & g1) ? 0xffff0000u : 0u);
            mk[sgi][op][2] = (parity9(rA & g2) ? 0xffffu : 0u) | (parity9(rB & g2) ? 0xffff0000u : 0u);
        }
    const int my_slot = 8 * (g >> 3) + 2 * (g & 3) + ((g >> 2) & 1);        // vit_slot(g + 32 j) = my_slot + 32 j
    unsigned z0 = 0, z1 = 0, z2 = 0, z3 = 0;                                // (x0|x4), (x1|x5), (x2|x6), (x3|x7), x_i = state 8g+i
    if (start_pm) {
        const uint4 v = *reinterpret_cast<const uint4 *>(&start_pm[8 * g]);
        z0 = v.x; z1 = v.y; z2 = v.z; z3 = v.w;
    }
    int next_norm = ((3 * g_begin + interval - 1) / interval) * interval;   // the minimum goes when step % interval == 0
    for (int base = 3 * g_begin; base < steps; base += VT) {
        const int nst = min(VT, steps - base);
        for (int i = g; i < nst; i += 32) {
            int j = len - 32 + base + i;                                    // the input index wraps (tail biting)
            while (j >= len) j -= len;
            const unsigned a = (unsigned)(in[3 * j] + 1), b = (unsigned)(in[3 * j + 1] + 1), c = (unsigned)(in[3 * j + 2] + 1);
            *reinterpret_cast<uint4 *>(sm.qd[i]) = make_uint4(a * 0x00010001u, b * 0x00010001u, c * 0x00010001u, 0u);
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < nst; k += 3) {
            const int ns = min(3, nst - k);                                 // (only the very last group can be short)
            if (base + k == 3 * g_store && g_store > g_begin) {             // (sm.pm holds what the last exchange left there)
                *reinterpret_cast<uint4 *>(&sm.chk[8 * g]) = *reinterpret_cast<const uint4 *>(&sm.pm[8 * g]);
                __syncwarp();                                               // (before any lane's next exchange store)
            }
            unsigned dword = 0;
#pragma unroll
            for (int sgi = 0; sgi < 3; sgi++) {
                if (sgi < ns) {
                    const uint4 q = *reinterpret_cast<const uint4 *>(sm.qd[k + sgi]);
                    const unsigned n0 = 0x00020002u - q.x, n1 = 0x00020002u - q.y, n2 = 0x00020002u - q.z;
                    unsigned e0, o0, e1, o1;                                // operands of the two packed operations
                    if (sgi == 2) {
                        e0 = am_lo2(z0, z1); o0 = am_hi2(z0, z1);           // stage 2 left (v0|v1), (v2|v3), (v4|v5), (v6|v7):
                        e1 = am_lo2(z2, z3); o1 = am_hi2(z2, z3);           // evens (v0|v2), odds (v1|v3); (v4|v6), (v5|v7)
                    } else {
                        e0 = z0; o0 = z1; e1 = z2; o1 = z3;
                    }
                    unsigned r0, r1, r2, r3, d = 0;
                    {
                        const unsigned *m = mk[sgi][0];
                        const unsigned mp = ((q.x & m[0]) | (n0 & ~m[0])) + ((q.y & m[1]) | (n1 & ~m[1])) + ((q.z & m[2]) | (n2 & ~m[2]));
                        const unsigned mm = 0x00060006u - mp;
                        const unsigned x1 = e0 + mp, y1 = o0 + mm, x2 = e0 + mm, y2 = o0 + mp;
                        r0 = am_umax2(x1, y1);
                        r1 = am_umax2(x2, y2);
                        const unsigned t1 = y1 + 0x80008000u - x1, t2 = y2 + 0x80008000u - x2;
                        d |= ((t1 >> 15) & 0x00010001u) | ((t2 >> 13) & 0x00040004u);
                    }
                    {
                        const unsigned *m = mk[sgi][1];
                        const unsigned mp = ((q.x & m[0]) | (n0 & ~m[0])) + ((q.y & m[1]) | (n1 & ~m[1])) + ((q.z & m[2]) | (n2 & ~m[2]));
                        const unsigned mm = 0x00060006u - mp;
                        const unsigned x1 = e1 + mp, y1 = o1 + mm, x2 = e1 + mm, y2 = o1 + mp;
                        r2 = am_umax2(x1, y1);
                        r3 = am_umax2(x2, y2);
                        const unsigned t1 = y1 + 0x80008000u - x1, t2 = y2 + 0x80008000u - x2;
                        d |= ((t1 >> 14) & 0x00020002u) | ((t2 >> 12) & 0x00080008u);
                    }
                    // stage 1 -> 2: E/O of ops (0,1),(2,3) are (r0, r2) and (r1, r3); stage 2 -> 3 and stage 3 -> exchange
                    // keep the same order: op 0's bit-0 outputs, op 1's bit-0 outputs, op 0's bit-1 outputs, op 1's
                    z0 = r0; z1 = r2; z2 = r1; z3 = r3;
                    if (base + k + sgi == next_norm) {
                        unsigned mn = am_umin2(am_umin2(z0, z1), am_umin2(z2, z3));
                        mn = min(mn & 0xffffu, mn >> 16);
                        mn = (unsigned)am_warp_min((int)mn) * 0x00010001u;
                        z0 -= mn; z1 -= mn; z2 -= mn; z3 -= mn;
                        next_norm += interval;
                    }
                    dword |= d << (4 * sgi);
                }
            }
            if (base + k >= 3 * g_store) decw[(size_t)((base + k) / 3) * 32 + g] = dword;
            // back to the exchange buffer.  After a full group: z0 = (u0|u1), z1 = (u2|u3), z2 = (u4|u5), z3 = (u6|u7), u_j = state g + 32 j
            if (ns == 3) {
                sm.pm[my_slot] = (unsigned short)z0;        sm.pm[my_slot + 32] = (unsigned short)(z0 >> 16);
                sm.pm[my_slot + 64] = (unsigned short)z1;   sm.pm[my_slot + 96] = (unsigned short)(z1 >> 16);
                sm.pm[my_slot + 128] = (unsigned short)z2;  sm.pm[my_slot + 160] = (unsigned short)(z2 >> 16);
                sm.pm[my_slot + 192] = (unsigned short)z3;  sm.pm[my_slot + 224] = (unsigned short)(z3 >> 16);
            } else {
                // a short last group: after stage 1 the registers hold (y0|y2), (y1|y3), (y4|y6), (y5|y7); after stage 2
                // (z0|z1), (z2|z3), (z4|z5), (z6|z7) - value i of stage sgi is state vit_butterfly(sgi, g, i & 3) + 128 (i >> 2)
                const unsigned zz[4] = { z0, z1, z2, z3 };
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int ia = ns == 1 ? (r & 1) + 4 * (r >> 1) : 2 * r, ib = ns == 1 ? ia + 2 : ia + 1;
                    sm.pm[vit_slot(vit_butterfly(ns - 1, g, ia & 3) + 128 * (ia >> 2))] = (unsigned short)zz[r];
                    sm.pm[vit_slot(vit_butterfly(ns - 1, g, ib & 3) + 128 * (ib >> 2))] = (unsigned short)(zz[r] >> 16);
                }
            }
            __syncwarp();
            {
                const uint4 v = *reinterpret_cast<const uint4 *>(&sm.pm[8 * g]);
                z0 = v.x; z1 = v.y; z2 = v.z; z3 = v.w;
            }
            __syncwarp();
        }
    }
    // first maximum in state order (conv_dec.c:310-317): (z0..z3) = (x0|x4), (x1|x5), (x2|x6), (x3|x7)
    {
        const unsigned xs[8] = { z0 & 0xffffu, z1 & 0xffffu, z2 & 0xffffu, z3 & 0xffffu, z0 >> 16, z1 >> 16, z2 >> 16, z3 >> 16 };
        int v = (int)xs[0], idx = 8 * g;
#pragma unroll
        for (int i = 1; i < 8; i++)
            if ((int)xs[i] > v) { v = (int)xs[i]; idx = 8 * g + i; }
        for (int o = 16; o; o >>= 1) {
            const int ov = __shfl_xor_sync(0xffffffffu, v, o), oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (g == 0) sm.state = (unsigned)idx;
    }
}

// The recursion of a whole frame by the CTA.  A short frame (PIDS) is one warp's job.  A long one is cut into a chunk per
// warp: chunk c > 0 starts VIT_CHUNK_WARMUP steps early from all-equal metrics, keeps no decisions until its own part
// begins, and notes its metrics at that point.  If they equal - up to one common constant - the metrics the chunk before
// ended with, every later decision and metric difference of the chunk is what the sequential recursion computes (the
// recursion only ever looks at differences, and the minimum is subtracted at the same absolute steps), and so by
// induction from chunk 0 for the whole frame.  The first chunk that fails the comparison, and all after it, are run
// again one after the other from the true metrics - rare: survivors merge within a few constraint lengths.
// Leaves the end state (first maximum) in vit[0].state.
__device__ inline int viterbi_k9_forward(AmVitSlot *vit, uint32_t *decw, int t, const int8_t *in, int len, unsigned g0, unsigned g1, unsigned g2,
                                         int warmup = VIT_CHUNK_WARMUP)
{
    constexpr int NW = AM_THREADS / 32;
    const int steps = len + 64, G = (steps + 2) / 3, wp = t >> 5, lane = t & 31;
    if (steps < 1024) {
        if (wp == 0) viterbi_k9_warp(vit[0], decw, lane, in, len, g0, g1, g2, 0, 0, G, nullptr);
        __syncthreads();
        return 0;
    }
    auto gb = [&](int c) { return (int)((long long)c * G / NW); };
    {
        const int b = gb(wp), e = wp == NW - 1 ? G : gb(wp + 1);
        viterbi_k9_warp(vit[wp], decw, lane, in, len, g0, g1, g2, wp == 0 ? 0 : max(0, b - warmup / 3), b, e, nullptr);
    }
    __syncthreads();
    int bad = NW;
    for (int c = 1; c < NW && bad == NW; c++) {
        bool same = true;
        if (gb(c) - warmup / 3 > 0) {                                        // (a warm-up that reaches the frame's start is the true recursion)
            const int d0 = (int)vit[c].chk[0] - (int)vit[c - 1].pm[0];
            for (int i = t; i < 256; i += AM_THREADS) same = same && ((int)vit[c].chk[i] - (int)vit[c - 1].pm[i] == d0);
        }
        if (t == 0) vit[0].changed = 0;
        __syncthreads();
        if (!same) vit[0].changed = 1;
        __syncthreads();
        if (vit[0].changed) bad = c;
        __syncthreads();
    }
    for (int c = bad; c < NW; c++) {
        if (wp == c) viterbi_k9_warp(vit[c], decw, lane, in, len, g0, g1, g2, gb(c), gb(c), c == NW - 1 ? G : gb(c + 1), vit[c - 1].pm);
        __syncthreads();
    }
    if (t == 0) vit[0].state = vit[NW - 1].state;
    __syncthreads();
    return NW - bad;                                                         // chunks that had to be run again
}

// Where viterbi_k9_warp keeps the survivor bit of new state s after a step of stage sgi: lane and bit of the group's word
__device__ __forceinline__ void vit_where(int sgi, unsigned s, unsigned &ln, unsigned &bit)
{
    const unsigned b = s & 127u, hi = s >> 7;
    ln = sgi == 0 ? b >> 2 : sgi == 1 ? (b & 63u) >> 1 : b & 31u;
    const unsigned i = sgi == 0 ? b & 3u : sgi == 1 ? (b & 1u) + 2u * (b >> 6) : b >> 5;
    const unsigned op = sgi == 0 ? i & 1u : i >> 1, half = sgi == 0 ? i >> 1 : i & 1u;
    bit = 4u * (unsigned)sgi + 2u * hi + op + 16u * half;
}

// Traceback by the whole CTA: the steps are cut into VIT_WALKERS segments and eight lanes of every warp walk one each,
// backwards.  A walk is a dependent chain - the survivor bit of the state decides which bit is needed
// next - but WHICH ROWS of decision words it needs does not depend on the state: the warp brings the next VIT_G groups'
// rows of each of its walkers into shared memory with coalesced 128-byte asynchronous copies (cp.async, all 32 lanes),
// one batch of 24 steps ahead of the batch being walked - longer than the memory latency - so the chains read shared
// memory only and a warp's chains advance in lock-step for the price of one.
// The state a walk has to start from is only known once the walk of the following segment has arrived - so a walker
// starts VIT_WARMUP steps further on, from state 0: survivor paths merge going backwards, and by the time it crosses
// into its own segment it is (almost always) on the decoder's path.  Every walker then checks the state it started its
// segment from against the state the following segment's walker really arrived at, and walks again from that one if they
// differ, until nobody had to: the last segment starts from the true end state, so when the round without changes comes
// every segment was walked from the state the sequential traceback passes through - the output is that of the
// sequential traceback, whatever the warm-up did.  Returns the number of repair rounds.
constexpr int VIT_G = 8, VIT_LW = 8, VIT_WALKERS = VIT_LW * (AM_THREADS / 32);
struct AmVitRows {
    uint32_t w[2][AM_THREADS / 32][VIT_LW][VIT_G][32];   // [buffer][warp][walker = lane < VIT_LW][group, newest first][word]
};
static_assert((AM_THREADS / 32) * sizeof(AmVitSlot) <= VIT_TEST_SLOTS_BYTES, "stage test kernel's slots");
static_assert(sizeof(AmVitRows) == 4 * VIT_TB_WORDS, "AmSmem reserves the rows as plain words");

__device__ __forceinline__ void am_copy_async4(uint32_t *smem_dst, const uint32_t *gmem_src)
{
#if defined(NB_EMU)
    *smem_dst = *gmem_src;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void am_copy_async_commit()
{
#if !defined(NB_EMU)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void am_copy_async_wait_but_one()                 // all but the most recent group have landed
{
#if !defined(NB_EMU)
    asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}

// Every lane walks its own steps [from, to) going down, starting in `state` (updated); the decoded bits of the steps
// below emit_below are written (the others are warm-up).  The 32 lanes of a warp call it together.
__device__ inline void vit_walk(uint32_t (*rows)[AM_THREADS / 32][VIT_LW][VIT_G][32], int wp, const uint32_t *decw, int lane, int from, int to, int emit_below,
                                unsigned &state_io, uint8_t *out, int len)
{
    const int g_lo = from / 3;
    int t = to - 1;
    unsigned state = state_io;
    // the rows of every lane's groups g_top .. g_top - n + 1 into buffer `buf`
    auto issue = [&](int buf, int g_top, int n) {
#pragma unroll 2
        for (int j = 0; j < VIT_LW; j++) {
            const int gj = __shfl_sync(0xffffffffu, g_top, j), nj = __shfl_sync(0xffffffffu, n, j);
#pragma unroll
            for (int r = 0; r < VIT_G; r++)
                if (r < nj) am_copy_async4(&rows[buf][wp][j][r][lane], decw + (size_t)(gj - r) * 32 + lane);
        }
        am_copy_async_commit();
    };
    int g_cur = t >= from ? t / 3 : 0, n_cur = t >= from ? min(VIT_G, g_cur - g_lo + 1) : 0;
    int buf = 0;
    issue(0, g_cur, n_cur);
    while (__any_sync(0xffffffffu, n_cur > 0)) {
        const int g_next = g_cur - n_cur, n_next = (n_cur > 0 && g_next >= g_lo) ? min(VIT_G, g_next - g_lo + 1) : 0;
        issue(buf ^ 1, g_next, n_next);
        am_copy_async_wait_but_one();
        __syncwarp();
#pragma unroll
        for (int r = 0; r < VIT_G; r++) {
#pragma unroll
            for (int sgi = 2; sgi >= 0; sgi--) {
                const int tt = 3 * (g_cur - r) + sgi;
                const bool on = r < n_cur && tt <= t && tt >= from;
                unsigned ln, bit;
                vit_where(sgi, state, ln, bit);
                const unsigned v = rows[buf][wp][lane & (VIT_LW - 1)][r][ln];
                const unsigned nxt = ((state << 1) & 254u) | ((v >> bit) & 1u);
                if (on && tt < emit_below && tt >= 32 && tt < 32 + len) out[tt - 32] = (uint8_t)(state >> 7);
                state = on ? nxt : state;
            }
        }
        if (n_cur > 0) t = 3 * (g_cur - n_cur + 1) - 1;
        g_cur = g_next;
        n_cur = n_next;
        buf ^= 1;
        __syncwarp();
    }
    state_io = state;
}

// (not inlined: four call sites in k_am, and the kernel's code should stay within reach of the instruction cache)
__device__ __noinline__ int viterbi_k9_traceback(AmVitSlot &sm, AmVitRows &rows, const uint32_t *decw, int w, uint8_t *out, int len,
                                           int warmup = VIT_WARMUP)
{
    const int steps = len + 64, lane = w & 31, wp = w >> 5;
    // walkers: lanes 0 .. VIT_LW-1 of every warp; a short frame (PIDS) is walked by one of them, without any warm-up
    const int nwalk = steps >= 1024 ? VIT_WALKERS : 1;
    const int seg = (steps + nwalk - 1) / nwalk;
    const int id = lane < VIT_LW ? VIT_LW * wp + lane : VIT_WALKERS;         // (other lanes only help loading)
    const int lo = min(steps, id * seg), hi = min(steps, lo + seg);          // this walker's steps [lo, hi)
    const bool mine = lo < hi;
    // warm-up: from `warmup` steps beyond the segment (or from the true end state) down to its upper end
    const int top = mine ? min(steps, hi + warmup) : hi;
    unsigned start = top == steps ? sm.state : 0u;
    vit_walk(rows.w, wp, decw, lane, hi, top, 0, start, out, len);
    bool walk = mine;
    int round = 0;                                                           // repair rounds (0: the warm-up was right)
    for (; round <= VIT_WALKERS; round++) {
        unsigned st2 = start;
        vit_walk(rows.w, wp, decw, lane, lo, walk ? hi : lo, steps, st2, out, len);
        if (walk) sm.ends[id] = (unsigned short)st2;                         // state after step lo - 1
        if (w == 0) sm.changed = 0;
        __syncthreads();
        walk = false;
        if (mine) {
            const unsigned want = hi == steps ? sm.state : (unsigned)sm.ends[id + 1];
            if (want != start) { start = want; walk = true; }
        }
        if (walk) sm.changed = 1;
        __syncthreads();
        const int any = sm.changed;
        __syncthreads();
        if (!any) break;
    }
    return round;
}
#endif

// a decode job
struct VitJob {
    const int8_t *in;
    uint8_t *out;
    int len;
    unsigned g0, g1, g2;
};

AM_HD inline void viterbi_k9(AmWork &w, Lanes L, const int8_t *in, uint8_t *out, int len, unsigned g0, unsigned g1, unsigned g2)
{
    const int steps = len + 64, interval = 32767 / (3 * 127) - 9;
#if defined(__CUDA_ARCH__)
    (void)steps;
    (void)interval;
    AmSmem &sm = *static_cast<AmSmem *>(L.smem);
    uint32_t *decw = reinterpret_cast<uint32_t *>(w.dec);
    __syncthreads();
    long long tv = AM_T0();
    const int redone = viterbi_k9_forward(sm.vit, decw, L.lane, in, len, g0, g1, g2);
    if (L.lane == 0) w.ph_cyc[15] += (unsigned long long)redone;        // chunks of the recursion that had to be run again
    AM_LAP(w, L, 8, tv);
    const int rr = viterbi_k9_traceback(sm.vit[0], *reinterpret_cast<AmVitRows *>(sm.tb), decw, L.lane, out, len);
    if (L.lane == 0) w.ph_cyc[11] += (unsigned long long)rr;             // repair rounds of the segmented traceback, all decodes
    __syncthreads();
    AM_LAP(w, L, 9, tv);
#else
    for (int i = L.lane; i < 256; i += L.n) w.pm[0][i] = 0;
    AM_SYNC();
    int cur = 0;
    int j = len - 32;
    for (int s = 0; s < steps; s++, j++) {
        if (j == len) j = 0;
        const int q0 = in[3 * j], q1 = in[3 * j + 1], q2 = in[3 * j + 2];
        const short *pmv = w.pm[cur];
        short *nxt = w.pm[cur ^ 1];
        for (int l = L.lane; l < 32; l += L.n) {
            unsigned d = 0;
            for (int q = 0; q < 4; q++) {
                const int b = 4 * l + q;
                const unsigned reg = (unsigned)b << 1;
                const int m = q0 * (parity9(reg & g0) ? 1 : -1) + q1 * (parity9(reg & g1) ? 1 : -1) + q2 * (parity9(reg & g2) ? 1 : -1);
                const int a0 = sat16(pmv[2 * b] + m), a1 = sat16(pmv[2 * b + 1] - m);
                const int c0 = sat16(pmv[2 * b] - m), c1 = sat16(pmv[2 * b + 1] + m);
                if (a0 > a1) nxt[b] = (short)a0;
                else { nxt[b] = (short)a1; d |= 1u << q; }
                if (c0 > c1) nxt[b + 128] = (short)c0;
                else { nxt[b + 128] = (short)c1; d |= 16u << q; }
            }
            w.dec[(size_t)s * 32 + l] = (uint8_t)d;
        }
        AM_SYNC();
        if (s % interval == 0) {
            int mn = nxt[0];
            for (int i = 1; i < 256; i++) mn = nxt[i] < mn ? nxt[i] : mn;     // every lane scans: same value
            AM_SYNC();
            for (int i = L.lane; i < 256; i += L.n) nxt[i] = (short)sat16(nxt[i] - mn);
            AM_SYNC();
        }
        cur ^= 1;
    }
    // first maximum wins (conv_dec.c:310-317); every lane walks back, lane-strided writes of the output
    const short *pmv = w.pm[cur];
    int best = -1;
    unsigned state = 0;
    for (int i = 0; i < 256; i++)
        if (pmv[i] > best) { best = pmv[i]; state = (unsigned)i; }
    for (int s = steps - 1; s >= 0; s--) {
        const unsigned byte = w.dec[(size_t)s * 32 + ((state & 127) >> 2)];
        const unsigned bit = (byte >> ((state & 3) + (state >= 128 ? 4 : 0))) & 1u;
        if (s >= 32 && s < 32 + len && ((s - 32) % L.n) == L.lane) out[s - 32] = (uint8_t)((state >> 7) & 1);
        state = ((state << 1) & 254u) | bit;
    }
    AM_SYNC();
#endif
}

// P1 and P3 of a frame's last block into separate buffers
AM_HD inline void viterbi_k9_pair(AmWork &w, Lanes L, const VitJob &a, const VitJob &b)
{
#if defined(__CUDA_ARCH__)
    viterbi_k9(w, L, a.in, a.out, a.len, a.g0, a.g1, a.g2);              // (each is the whole CTA's job: chunks of the recursion, segments of the traceback)
    viterbi_k9(w, L, b.in, b.out, b.len, b.g0, b.g1, b.g2);
#else
    viterbi_k9(w, L, a.in, a.out, a.len, a.g0, a.g1, a.g2);
    viterbi_k9(w, L, b.in, b.out, b.len, b.g0, b.g1, b.g2);
#endif
}

AM_HD inline void descramble(const AmTables &tb, Lanes L, uint8_t *bits, int len)     // decode.c:279-294
{
    for (int i = L.lane; i < len; i += L.n) bits[i] ^= tb.pn[i];
    AM_SYNC();
}

// Every lane gets the total.  Device: the positions are spread over the CTA - the encoder register at bit i is just
// the nine decoded bits i-8 .. i (tail-biting), no running state - and the counts summed (integers: any order).
AM_HD inline int bit_errors(Lanes L, const int8_t *coded, const uint8_t *decoded, unsigned len, unsigned g0, unsigned g1, unsigned g2,
                            const uint8_t *punct, int plen)                             // decode.c:234-259
{
    const unsigned k = 9, gens[3] = { g0, g1, g2 };
#if defined(__CUDA_ARCH__)
    AmSmem &sm = *static_cast<AmSmem *>(L.smem);
    int errors = 0;
    for (unsigned i = L.lane; i < len; i += L.n) {
        unsigned r = 0;
        for (unsigned q = 0; q < k; q++) {
            const unsigned src = i >= q ? i - q : i + len - q;
            r |= (unsigned)decoded[src] << (k - 1 - q);
        }
        const unsigned j = 3 * i;
        for (unsigned g = 0; g < 3; g++)
            if (punct[(j + g) % plen] && ((coded[j + g] > 0) != parity9(r & gens[g]))) errors++;
    }
    for (int o = 16; o; o >>= 1) errors += __shfl_xor_sync(0xffffffffu, errors, o);
    __syncthreads();
    if ((L.lane & 31) == 0) sm.red[L.lane >> 5] = errors;
    __syncthreads();
    int total = 0;
    for (int q = 0; q < AM_THREADS / 32; q++) total += sm.red[q];
    return total;
#else
    (void)L;
    unsigned r = 0, errors = 0;
    for (unsigned i = 0; i < k - 1; i++) r = ((r >> 1) | ((unsigned)decoded[len - (k - 1) + i] << (k - 1))) & 0xffffu;
    for (unsigned i = 0, j = 0; i < len; i++, j += 3) {
        r = ((r >> 1) | ((unsigned)decoded[i] << (k - 1))) & 0xffffu;
        for (unsigned g = 0; g < 3; g++)
            if (punct[(j + g) % plen] && ((coded[j + g] > 0) != parity9(r & gens[g]))) errors++;
    }
    return (int)errors;
#endif
}

// ---- decode (reference src/decode.c) ----
AM_HD inline int bit_map(const uint8_t *matrix, int b, int k, int p)                    // decode.c:67-72
{
    const int col = (9 * k) % 25;
    const int row = (11 * col + 16 * (k / 25) + 11 * (k / 50)) % 32;
    return (matrix[PW * (b * BLK + row) + col] >> p) & 1;
}

// for (i = lane; i < N; i += n_lanes) st(i, ld(i)), B iterations at a time with all their loads ahead of their stores:
// the loops below move bytes between arrays of one struct, which the compiler will not reorder on its own, and every
// iteration would otherwise wait for its own load (the arrays live in global memory)
template <int B, typename Load, typename Store>
AM_HD inline void batched(Lanes L, int N, Load ld, Store st)
{
    for (int i0 = L.lane; i0 < N; i0 += B * L.n) {
        decltype(ld(0)) v[B];
#pragma unroll
        for (int k = 0; k < B; k++)
            if (i0 + k * L.n < N) v[k] = ld(i0 + k * L.n);
#pragma unroll
        for (int k = 0; k < B; k++)
            if (i0 + k * L.n < N) st(i0 + k * L.n, v[k]);
    }
}
struct AmBytes4 { uint8_t a, b, c, d; };
struct AmBytes12 { uint8_t v[12]; };
struct AmWords4 { uint32_t x, y, z, w; };

AM_HD inline void interleaver_ma1(AmWork &w, Lanes L, bool ma3)                        // decode.c:74-231
{
    const int bl_delay[3] = { 2, 1, 5 }, ml_delay[3] = { 11, 6, 7 }, bu_delay[3] = { 10, 8, 9 }, mu_delay[3] = { 4, 3, 0 };
    const int el_delay[2] = { 0, 1 }, eu_delay[4] = { 2, 3, 5, 4 };
    batched<2>(L, 18000,
               [&](int n) {
                   AmBytes4 r;
                   r.a = (uint8_t)bit_map(w.buffer_pl, n / 2250, (n + n / 750 + 1) % 750, n % 3);
                   r.b = (uint8_t)bit_map(w.buffer_pl, (3 * n + 3) % 8, (n + n / 3000 + 3) % 750, 3 + (n % 3));
                   r.c = (uint8_t)bit_map(w.buffer_pu, n / 2250, (n + n / 750) % 750, n % 3);
                   r.d = (uint8_t)bit_map(w.buffer_pu, (3 * n) % 8, (n + n / 3000 + 2) % 750, 3 + (n % 3));
                   return r;
               },
               [&](int n, const AmBytes4 &r) {
                   w.bl[n] = r.a;
                   w.ml[DIVERSITY + n] = r.b;
                   w.bu[n] = r.c;
                   w.mu[DIVERSITY + n] = r.d;
               });
    if (!ma3) {
        batched<2>(L, 12000, [&](int n) { return (uint8_t)bit_map(w.buffer_t, (3 * n + n / 3000) % 8, (n + (n / 6000)) % 750, n % 2); },
                   [&](int n, uint8_t v) { w.el[n] = v; });
        batched<2>(L, 24000,
                   [&](int n) { return (uint8_t)bit_map(w.buffer_s, (3 * n + n / 3000 + 2 * (n / 12000)) % 8, (n + (n / 6000)) % 750, n % 4); },
                   [&](int n, uint8_t v) { w.eu[n] = v; });
    } else {
        batched<2>(L, 18000,                                                             // decode.c:119-140
                   [&](int n) {
                       AmBytes4 r;
                       r.a = (uint8_t)bit_map(w.buffer_t, (3 * n + 3) % 8, (n + n / 3000 + 3) % 750, n % 3);
                       r.b = (uint8_t)bit_map(w.buffer_t, (3 * n + 3) % 8, (n + n / 3000 + 3) % 750, 3 + (n % 3));
                       r.c = (uint8_t)bit_map(w.buffer_s, (3 * n) % 8, (n + n / 3000 + 2) % 750, n % 3);
                       r.d = (uint8_t)bit_map(w.buffer_s, (3 * n) % 8, (n + n / 3000 + 2) % 750, 3 + (n % 3));
                       return r;
                   },
                   [&](int n, const AmBytes4 &r) {
                       w.ebl[n] = r.a;
                       w.eml[DIVERSITY + n] = r.b;
                       w.ebu[n] = r.c;
                       w.emu[DIVERSITY + n] = r.d;
                   });
    }
    AM_SYNC();
    batched<2>(L, 6000,
               [&](int i) {
                   AmBytes12 r;
#pragma unroll
                   for (int j = 0; j < 3; j++) {
                       r.v[j] = w.bl[i * 3 + j];
                       r.v[3 + j] = w.ml[i * 3 + j];
                       r.v[6 + j] = w.bu[i * 3 + j];
                       r.v[9 + j] = w.mu[i * 3 + j];
                   }
                   return r;
               },
               [&](int i, const AmBytes12 &r) {
#pragma unroll
                   for (int j = 0; j < 3; j++) {
                       w.p1_am[i * 12 + bl_delay[j]] = r.v[j];
                       w.p1_am[i * 12 + ml_delay[j]] = r.v[3 + j];
                       w.p1_am[i * 12 + bu_delay[j]] = r.v[6 + j];
                       w.p1_am[i * 12 + mu_delay[j]] = r.v[9 + j];
                   }
               });
    batched<2>(L, 6000,
               [&](int i) {
                   AmBytes12 r;
                   if (!ma3) {
#pragma unroll
                       for (int j = 0; j < 2; j++) r.v[j] = w.el[i * 2 + j];
#pragma unroll
                       for (int j = 0; j < 4; j++) r.v[2 + j] = w.eu[i * 4 + j];
#pragma unroll
                       for (int j = 6; j < 12; j++) r.v[j] = 0;
                   } else {
#pragma unroll
                       for (int j = 0; j < 3; j++) {                                     // decode.c:163-170
                           r.v[j] = w.ebl[i * 3 + j];
                           r.v[3 + j] = w.eml[i * 3 + j];
                           r.v[6 + j] = w.ebu[i * 3 + j];
                           r.v[9 + j] = w.emu[i * 3 + j];
                       }
                   }
                   return r;
               },
               [&](int i, const AmBytes12 &r) {
                   if (!ma3) {
#pragma unroll
                       for (int j = 0; j < 2; j++) w.p3_am[i * 6 + el_delay[j]] = r.v[j];
#pragma unroll
                       for (int j = 0; j < 4; j++) w.p3_am[i * 6 + eu_delay[j]] = r.v[2 + j];
                   } else {
#pragma unroll
                       for (int j = 0; j < 3; j++) {
                           w.p3_am[i * 12 + bl_delay[j]] = r.v[j];
                           w.p3_am[i * 12 + ml_delay[j]] = r.v[3 + j];
                           w.p3_am[i * 12 + bu_delay[j]] = r.v[6 + j];
                           w.p3_am[i * 12 + mu_delay[j]] = r.v[9 + j];
                       }
                   }
               });
    AM_SYNC();
    // the main bits move three frames towards the front (memmove of 54000 entries, in three non-overlapping steps),
    // sixteen bytes at a time
    for (int step = 0; step < 3; step++) {
        auto move16 = [&](uint8_t *arr) {
            const AmWords4 *src = reinterpret_cast<const AmWords4 *>(arr + (step + 1) * 18000);
            AmWords4 *dst = reinterpret_cast<AmWords4 *>(arr + step * 18000);
            batched<2>(L, 18000 / 16, [&](int i) { return src[i]; }, [&](int i, const AmWords4 &v) { dst[i] = v; });
        };
        move16(w.ml);
        move16(w.mu);
        if (ma3) {                                                                        // decode.c:176-180
            move16(w.eml);
            move16(w.emu);
        }
        AM_SYNC();
    }
    // depuncture: kept positions of every 15 (P1) / 6 (P3) code bits (decode.c:186-212)
    batched<2>(L, 8 * P1_LEN * 3,
               [&](int i) {
                   const int r = i % 15, base = (i / 15) * 12;
                   const int before = r - (r > 1) - (r > 4) - (r > 7);
                   return (int8_t)((r == 1 || r == 4 || r == 7) ? 0 : (w.p1_am[base + before] ? 1 : -1));
               },
               [&](int i, int8_t v) { w.vit_p1[i] = v; });
    if (!ma3) {
        batched<2>(L, P3_LEN * 3,
                   [&](int i) {
                       const int r = i % 6, base = (i / 6) * 3;
                       const int before = r - (r > 1);
                       return (int8_t)((r == 1 || r == 4 || r == 5) ? 0 : (w.p3_am[base + before] ? 1 : -1));
                   },
                   [&](int i, int8_t v) { w.vit_p3[i] = v; });
    } else {
        batched<2>(L, P3_LEN_MA3 * 3,                                                    // decode.c:214-229
                   [&](int i) {
                       const int r = i % 15, base = (i / 15) * 12;
                       const int before = r - (r > 1) - (r > 4) - (r > 7);
                       return (int8_t)((r == 1 || r == 4 || r == 7) ? 0 : (w.p3_am[base + before] ? 1 : -1));
                   },
                   [&](int i, int8_t v) { w.vit_p3[i] = v; });
    }
    AM_SYNC();
}

AM_HD inline void process_pids(AmState &st, AmWork &w, const AmTables &tb, const AmIo &io, Lanes L)   // decode.c:474-505
{
    const int il_delay[12] = { 0, 1, 12, 13, 6, 5, 18, 17, 11, 7, 23, 19 };
    const int iu_delay[12] = { 2, 4, 14, 16, 3, 8, 15, 20, 9, 10, 21, 22 };
    const uint8_t *sbit = w.sym_pids;
    const int pids1_disabled = (st.psmi == 1) && st.rdbi;
    for (int n = L.lane; n < 120; n += L.n) {
        const int p = n % 4, i = n / 12, j = n % 12;
        int k = (n + (n / 60) + 11) % 30;
        int row = (11 * (k + (k / 15)) + 3) % 32;
        const int il = (sbit[row * 2] >> p) & 1;
        k = (n + (n / 60)) % 30;
        row = (11 * (k + (k / 15)) + 3) % 32;
        const int iu = (sbit[row * 2 + 1] >> p) & 1;
        w.vit_pids[i * 24 + il_delay[j]] = pids1_disabled ? 0 : (il ? 1 : -1);
        w.vit_pids[i * 24 + iu_delay[j]] = iu ? 1 : -1;
    }
    AM_SYNC();
    viterbi_k9(w, L, w.vit_pids, w.out, PIDS_LEN, 0561, 0753, 0711);
    descramble(tb, L, w.out, PIDS_LEN);
    uint8_t *rec = log_reserve(st, io, L, REC_PIDS, 11);                                  // 80 bits + CRC verdict
    if (rec && L.lane == 0) {
        for (int i = 0; i < PIDS_LEN; i++) rec[i >> 3] |= (uint8_t)(w.out[i] << (7 - (i & 7)));
        rec[10] = (uint8_t)pids_crc12_ok(rec);                                            // pids.c:1042
    }
    AM_SYNC();
}

// frame.c:645-714 (PCI), :146-156, :527-541 + rs: an AM P1 PDU that announces audio but whose first header
// fails RS(255,247) sends the receiver back to acquisition.  `fix_header` is supplied by the caller
// (csrc/rs.cuh on the device, the same algorithm on the host).
template <typename FixHeader>
AM_HD inline int p1_sync_lost(const uint8_t *bits, FixHeader fix_header)
{
    uint8_t pdu[96];
    unsigned h = 0, j = 0, nb = 0, val = 0;
    uint32_t pci = 0;
    for (int i = 0; i < 96; i++) pdu[i] = 0;
    for (unsigned i = 0; i < (unsigned)P1_LEN && nb < 96; i++) {
        const unsigned byte_start = (i >> 3) << 3;
        const unsigned byte_len = (P1_LEN - byte_start < 8) ? P1_LEN - byte_start : 8;
        const unsigned bit = bits[byte_start + byte_len - 1 - (i & 7)];
        if (i >= 120 && ((i - 120) % 160) == 0 && h < 22) {
            pci |= bit << (23 - h);
            ++h;
        } else {
            val |= bit << (7 - j);
            if (++j == 8) {
                pdu[nb++] = (uint8_t)val;
                val = 0;
                j = 0;
            }
        }
    }
    // the PCI bits all lie behind the first 96 PDU bytes?  No: they start at bit 120 - collect the rest of them
    for (unsigned hh = h; hh < 22; hh++) {
        const unsigned i = 120 + 160 * hh;
        const unsigned byte_start = (i >> 3) << 3;
        const unsigned byte_len = (P1_LEN - byte_start < 8) ? P1_LEN - byte_start : 8;
        pci |= (uint32_t)bits[byte_start + byte_len - 1 - (i & 7)] << (23 - hh);
    }
    if ((pci & 0xFFFFFC) == (0x3634CE & 0xFFFFFC)) return 0;              // fixed data only: no audio
    return !fix_header(pdu);
}

template <typename FixHeader>
AM_HD inline void process_p1_p3(AmState &st, AmWork &w, const AmTables &tb, const AmIo &io, Lanes L, unsigned bc,
                                FixHeader fix_header)                                   // decode.c:507-554
{
    const uint8_t punct_e1[15] = { 1, 0, 1, 1, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 1 }, punct_e2[6] = { 1, 0, 1, 1, 0, 0 };
    if (bc == 0) st.am_errors = 0;
    if (st.am_diversity_wait == 0) {
        const int8_t *v = w.vit_p1 + bc * P1_LEN * 3;
        // the frame's last block also decodes P3 (unless the station sends none): both at once, P1 into its own buffer
        const bool with_p3 = bc == 7 && !st.rdbi;
        const bool ma3 = st.psmi == MODE_MA3;
        uint8_t *p1_out = with_p3 ? w.out_p1 : w.out;
        if (with_p3) {
            const VitJob ja = { v, w.out_p1, P1_LEN, 0561, 0657, 0711 };
            const VitJob jb = { w.vit_p3, w.out, ma3 ? P3_LEN_MA3 : P3_LEN, 0561, ma3 ? 0657u : 0753u, 0711 };
            viterbi_k9_pair(w, L, ja, jb);
        } else {
            viterbi_k9(w, L, v, w.out, P1_LEN, 0561, 0657, 0711);
        }
        st.am_errors += bit_errors(L, v, p1_out, P1_LEN, 0561, 0657, 0711, punct_e1, 15);
        AM_SYNC();
        descramble(tb, L, p1_out, P1_LEN);
        emit_frame(st, w, io, L, p1_out, P1_LEN, 0);
        if (p1_sync_lost(p1_out, fix_header)) set_state(st, io, L, ST_NONE);             // inside frame_push, frame.c:538
        AM_SYNC();
        long long tl3 = AM_T0();
        (void)tl3;
        if (bc == 7) {
            unsigned total = 8 * 9000;
            if (!st.rdbi) {
                if (!ma3) {
                    total += 36000;
                    st.am_errors += bit_errors(L, w.vit_p3, w.out, P3_LEN, 0561, 0753, 0711, punct_e2, 6);
                    AM_SYNC();
                    descramble(tb, L, w.out, P3_LEN);
                    emit_frame(st, w, io, L, w.out, P3_LEN, 1);
                } else {                                                                  // decode.c:533-539
                    total += 72000;
                    st.am_errors += bit_errors(L, w.vit_p3, w.out, P3_LEN_MA3, 0561, 0657, 0711, punct_e1, 15);
                    AM_SYNC();
                    descramble(tb, L, w.out, P3_LEN_MA3);
                    emit_frame(st, w, io, L, w.out, P3_LEN_MA3, 1);
                }
                AM_SYNC();
            }
            AM_LAP(w, L, 6, tl3);
            const float cber = (float)st.am_errors / (float)total;
            uint8_t *rec = log_reserve(st, io, L, REC_BER, 4);
            if (rec && L.lane == 0) memcpy(rec, &cber, 4);
        }
    }
    if (bc == 7) {
        long long tl7 = AM_T0();
        (void)tl7;
        interleaver_ma1(w, L, st.psmi == MODE_MA3);
        AM_LAP(w, L, 7, tl7);
        if (st.am_diversity_wait > 0) st.am_diversity_wait--;
    }
}

// ---- sync (reference src/sync.c) ----
AM_HD inline uint8_t gray4(float f) { return f < -1 ? 0 : f < 0 ? 2 : f < 1 ? 3 : 1; }
AM_HD inline uint8_t gray8(float f)
{
    return f < -3 ? 0 : f < -2 ? 4 : f < -1 ? 6 : f < 0 ? 2 : f < 1 ? 3 : f < 2 ? 7 : f < 3 ? 5 : 1;
}
AM_HD inline uint8_t qpsk(float2 c) { return (uint8_t)((c.x < 0 ? 0 : 1) | (c.y < 0 ? 0 : 2)); }
AM_HD inline uint8_t qam16(float2 c) { return (uint8_t)(gray4(c.x) | (gray4(c.y) << 2)); }
AM_HD inline uint8_t qam64(float2 c) { return (uint8_t)(gray8(c.x) | (gray8(c.y) << 3)); }

AM_HD inline float phase_diff(float a, float b)                                         // sync.c:284-290
{
    float diff = a - b;
    while ((double)diff > PI / 2) diff = (float)((double)diff - PI);
    while ((double)diff < -PI / 2) diff = (float)((double)diff + PI);
    return diff;
}

AM_HD inline int needle_am(int n)                                                       // sync.c:210-212
{
    const signed char nd[32] = { 0, 1, 1, 0, 0, 1, 0, -1, -1, 1, -1, -1, -1, -1, 0, -1, -1, -1, -1, -1, -1, 1, 1,
                                 -1, -1, -1, -1, -1, -1, -1, -1, -1 };
    return nd[n];
}

AM_HD inline int find_block_am(AmState &st, const AmWork &w, unsigned ref)              // sync.c:208-237
{
    unsigned char data[BLK];
    for (int n = 0; n < BLK; n++) {
        data[n] = w.bins[ref][n].y <= 0 ? 0 : 1;
        if ((needle_am(n) >= 0) && (data[n] != needle_am(n))) return -1;
    }
    if (data[7] ^ data[8]) return -1;
    if (data[10] ^ data[11] ^ data[12] ^ data[13]) return -1;
    if (data[15] ^ data[16] ^ data[17] ^ data[18] ^ data[19] ^ data[20]) return -1;
    if (data[23] ^ data[24] ^ data[25] ^ data[26] ^ data[27] ^ data[28] ^ data[29] ^ data[30] ^ data[31]) return -1;
    const int bc = (data[17] << 2) | (data[18] << 1) | data[19];
    if (bc == 0) {
        st.psmi = (data[26] << 4) | (data[27] << 3) | (data[28] << 2) | (data[29] << 1) | data[30];
        st.pli = data[7];
        st.hppi = data[11];
        st.aabi = data[12];
        st.rdbi = data[15];
    }
    return bc;
}

AM_HD inline int find_ref_am(const AmWork &w, unsigned ref)                             // sync.c:239-252, :150-167
{
    unsigned char data[BLK];
    for (int n = 0; n < BLK; n++) data[n] = w.bins[ref][n].y <= 0 ? 0 : 1;
    for (int n = 0; n < BLK; n++) {
        int i;
        for (i = 0; i < 23; i++) {
            if (needle_am(i) < 0) continue;
            if (needle_am(i) != data[(n + i) % BLK]) break;
        }
        if (i == 23) return n;
    }
    return -1;
}

template <typename FixHeader>
AM_HD inline void sync_block(AmState &st, AmWork &w, const AmTables &tb, const AmIo &io, Lanes L, FixHeader fix_header)
{
    // sync.c:616-633: mirror the lower sideband, fold it onto the upper one up to the outer PIDS carrier
    for (int i = REF_IDX + L.lane; i <= MAX_IDX; i += L.n)
        for (int n = 0; n < BLK; n++) {
            const float2 v = w.bins[CENTER - i][n];
            w.bins[CENTER - i][n] = make_float2(-v.x, v.y);                              // -conj
        }
    AM_SYNC();
    if (st.psmi != MODE_MA3)                                                             // the mode known when the block starts
        for (int i = REF_IDX + L.lane; i <= PIDS_OUTER; i += L.n)
            for (int n = 0; n < BLK; n++) w.bins[CENTER + i][n] = cadd(w.bins[CENTER + i][n], w.bins[CENTER - i][n]);
    AM_SYNC();

    if (st.state == ST_COARSE && st.cfo_wait == 0) {                                     // sync.c:635-647
        const int offset = find_ref_am(w, CENTER + REF_IDX);
        if (offset > 0) {
            st.keep_extra = ((BLK - offset) % BLK) * SYM;
            st.cfo_wait = 8;
        }
    } else {
        st.cfo_wait--;
    }
    if (st.state == ST_COARSE) {                                                         // sync.c:649-666
        const int bc = find_block_am(st, w, CENTER + REF_IDX);
        if (bc == -1) st.offset_history = 0;
        else st.offset_history = (st.offset_history << 4) | (unsigned)bc;
        if ((st.offset_history & 0xffff) == 0x5670) {
            st.bc = 0;
            set_state(st, io, L, ST_FINE);
            l2_enqueue(st, w, L, 0, 0, 0);                                               // frame_reset, sync.c:662-663
            st.am_errors = 0;                                                            // decode_reset, decode.c:556-565
            st.am_diversity_wait = 4;
            st.offset_history = 0;
        }
    }
    if (st.state != ST_FINE) return;

    const bool ma3 = st.psmi == MODE_MA3;
    // PIDS carriers (sync.c:668-685): equalise with the training symbols of rows 8 and 24, slice
    {
        const int b1 = CENTER + (!ma3 ? PIDS_INNER : -PIDS_INNER), b2 = CENTER + (!ma3 ? PIDS_OUTER : PIDS_INNER);
        const float2 tr = make_float2(2 * 1.5f, 2 * -0.5f);
        const float2 m1 = cdiv(tr, cadd(w.bins[b1][8], w.bins[b1][24]));
        const float2 m2 = cdiv(tr, cadd(w.bins[b2][8], w.bins[b2][24]));
        AM_SYNC();
        for (int n = L.lane; n < BLK; n += L.n) {
            w.bins[b1][n] = cmul(w.bins[b1][n], m1);
            w.sym_pids[2 * n] = qam16(w.bins[b1][n]);
            w.bins[b2][n] = cmul(w.bins[b2][n], m2);
            w.sym_pids[2 * n + 1] = qam16(w.bins[b2][n]);
        }
        AM_SYNC();
    }
    long long tlap = AM_T0();
    (void)tlap;
    process_pids(st, w, tb, io, L);
    AM_LAP(w, L, 4, tlap);

    // partitions (sync.c:687-717): per column the two training rows give the equaliser tap.  Column 0 of the
    // primary partitions sits at -/+ `primary`, of the secondary at +28, of the tertiary at +2 (MA1) / -28 (MA3)
    const int primary = !ma3 ? OUTER_START : INNER_START, tertiary = !ma3 ? INNER_START : MIDDLE_START, tdir = !ma3 ? 1 : -1;
    const float2 tr_p = make_float2(2 * 2.5f, 2 * -2.5f);
    const float2 tr_s = !ma3 ? make_float2(2 * 1.5f, 2 * -0.5f) : tr_p, tr_t = !ma3 ? make_float2(2 * -0.5f, 2 * 0.5f) : tr_p;
    for (int col = L.lane; col < PW; col += L.n) {
        const int t1 = (5 + 11 * col) % 32, t2 = (21 + 11 * col) % 32;
        const int ipl = CENTER - primary - col, ipu = CENTER + primary + col, is = CENTER + MIDDLE_START + col,
                  it = CENTER + tdir * (tertiary + col);
        w.mult[0][col] = cdiv(tr_p, cadd(w.bins[ipl][t1], w.bins[ipl][t2]));
        w.mult[1][col] = cdiv(tr_p, cadd(w.bins[ipu][t1], w.bins[ipu][t2]));
        w.mult[2][col] = cdiv(tr_s, cadd(w.bins[is][t1], w.bins[is][t2]));
        w.mult[3][col] = cdiv(tr_t, cadd(w.bins[it][t1], w.bins[it][t2]));
    }
    AM_SYNC();
    {
        float samperr = 0;
        for (int col = 1; col < PW; col++) {
            samperr += phase_diff(carg(w.mult[0][col]), carg(w.mult[0][col - 1]));
            samperr += phase_diff(carg(w.mult[1][col]), carg(w.mult[1][col - 1]));
        }
        samperr = (float)((double)(samperr / (2 * (PW - 1)) * FFT) / (2 * PI));
        st.samperr = (int)roundf(samperr);
    }
    for (int idx = L.lane; idx < BLK * PW; idx += L.n) {                                 // sync.c:725-754
        const int n = idx / PW, col = idx - n * PW;
        const int ipl = CENTER - primary - col, ipu = CENTER + primary + col, is = CENTER + MIDDLE_START + col,
                  it = CENTER + tdir * (tertiary + col);
        float2 v;
        v = cmul(w.bins[ipl][n], w.mult[0][col]);
        w.bins[ipl][n] = v;
        w.sym_pl[idx] = qam64(v);
        v = cmul(w.bins[ipu][n], w.mult[1][col]);
        w.bins[ipu][n] = v;
        w.sym_pu[idx] = qam64(v);
        v = cmul(w.bins[is][n], w.mult[2][col]);
        w.bins[is][n] = v;
        w.sym_s[idx] = !ma3 ? qam16(v) : qam64(v);
        v = cmul(w.bins[it][n], w.mult[3][col]);
        w.bins[it][n] = v;
        w.sym_t[idx] = !ma3 ? qpsk(v) : qam64(v);
    }
    AM_SYNC();
    for (int i = L.lane; i < BLK * PW; i += L.n) {                                        // decode_push_pl_pu_s_t, decode.c:439-449
        w.buffer_pl[st.bc * BLK * PW + i] = w.sym_pl[i];
        w.buffer_pu[st.bc * BLK * PW + i] = w.sym_pu[i];
        w.buffer_s[st.bc * BLK * PW + i] = w.sym_s[i];
        w.buffer_t[st.bc * BLK * PW + i] = w.sym_t[i];
    }
    AM_SYNC();
    AM_LAP(w, L, 3, tlap);
    process_p1_p3(st, w, tb, io, L, st.bc, fix_header);
    AM_LAP(w, L, 5, tlap);                     // (P1 + at the frame's last block P3 and the interleaver; split below)
    st.bc = (st.bc + 1) % 8;
}

// ---- acquisition and demodulation (reference src/acquire.c) ----
AM_HD inline float2 input_at(const AmIo &io, long long n)            // cq15_to_cf, defines.h:106-109 (no conjugate in AM)
{
    return make_float2((float)AM_LDIN(io.iq + 2 * n) / 32767.0f, (float)AM_LDIN(io.iq + 2 * n + 1) / 32767.0f);
}

// 256-point forward FFT of w.fft in place, result in natural order (radix-2, lanes share the butterflies)
AM_HD inline void fft256(AmWork &w, const AmTables &tb, Lanes L)
{
    for (int i = L.lane; i < FFT; i += L.n) {
        const int r = tb.brev[i];
        if (r > i) {
            const float2 t = w.fft[i];
            w.fft[i] = w.fft[r];
            w.fft[r] = t;
        }
    }
    AM_SYNC();
    for (int half = 1; half < FFT; half <<= 1) {
        const int tstep = FFT / (2 * half);
        for (int b = L.lane; b < FFT / 2; b += L.n) {
            const int grp = b / half, k = b - grp * half;
            const int i0 = grp * 2 * half + k, i1 = i0 + half;
            const float2 t = cmul(w.fft[i1], tb.tw[k * tstep]);
            const float2 u = w.fft[i0];
            w.fft[i0] = cadd(u, t);
            w.fft[i1] = make_float2(u.x - t.x, u.y - t.y);
        }
        AM_SYNC();
    }
}

// one OFDM symbol: rotate by the running phase, window, fold, shift by 121, FFT, fftshift (acquire.c:178-195,237-256)
AM_HD inline void symbol_fft(AmWork &w, const AmTables &tb, Lanes L, int sym, int samperr, float2 &phase, float2 inc)
{
    const int offset = (FFT - CP) / 2;
    // the phase recurrence is sequential: every lane runs it, each lane writes its share of the FFT input;
    // the folded tail (j >= 256) adds onto entries written by the same lane (same residue of j mod 256 mod n?)
    // - not in general, so the tail is added in a second sweep after a barrier
    float2 ph = phase;
    for (int j = 0; j < FFT; ++j) {
        if ((j % L.n) == L.lane) {
            const float2 sample = cmul(ph, w.buf[sym * SYM + j + samperr]);
            w.fft[(j + offset) % FFT] = j < CP ? cscale(sample, tb.shape[j]) : sample;
        }
        ph = cmul(ph, inc);
    }
    AM_SYNC();
    for (int j = FFT; j < SYM; ++j) {
        if (((j - FFT) % L.n) == L.lane) {
            const float2 sample = cmul(ph, w.buf[sym * SYM + j + samperr]);
            const int idx = (j + offset) % FFT;
            w.fft[idx] = cadd(w.fft[idx], cscale(sample, tb.shape[j]));
        }
        ph = cmul(ph, inc);
    }
    {
        const float a = cabs2(ph);
        phase = make_float2(ph.x / a, ph.y / a);
    }
    AM_SYNC();
    fft256(w, tb, L);
    for (int i = L.lane; i < FFT; i += L.n) w.spec[i] = w.fft[(i + FFT / 2) % FFT];       // fftshift, defines.h:123-138
    AM_SYNC();
}

#if defined(__CUDA_ARCH__)
// One pass over the 32 symbols of a block (acquire.c:178-195 first pass, :237-256 second pass), as a two-stage pipeline
// in batches of three symbols: warp 0 runs the NCO phase chain of symbols 3i .. 3i+2 - 270 dependent complex
// multiplications and the renormalisation each, exactly the reference's recurrence, the one thing in this pass that
// cannot be spread out - while warps 1, 2, 3 take ONE symbol each of batch i-1: window, fold, shift and the radix-2
// butterflies of symbol_fft / fft256 above in the warp's own shared-memory buffer, with __syncwarp between the stages
// and no CTA barrier (one per batch hands the phases over), then hand the bins on:
//   first pass  (bins == nullptr): the carrier bin of every symbol -> sm.dem.carrier, and - while acquiring - the
//               magnitudes of the 107 bins around it, summed over the symbols in symbol order -> sm.dem.mag
//   second pass: bins CENTER-81 .. CENTER+81 -> bins[b][symbol]                               (sync_push)
// `phase` is every thread's copy of the running NCO phase; all of them get the value after the last symbol.
__device__ inline void demod_pass(AmWork &w, const AmTables &tb, Lanes L, int samperr, float2 &phase, float2 inc,
                                  float2 (*bins)[BLK], bool want_mag)
{
    AmSmem &sm = *static_cast<AmSmem *>(L.smem);
    constexpr int NMAG = 2 * PIDS_OUTER + 1, NBATCH = (BLK + 2) / 3;
    const int t = L.lane, lane = t & 31, cw = (t >> 5) - 1;    // consumer warp 0..2 (warp 0 of the CTA produces)
    const int offset = (FFT - CP) / 2;
    float2 ph = phase;
    if (!bins && want_mag)
        for (int b = t; b < NMAG; b += AM_THREADS) sm.dem.mag[b] = 0.0f;
    for (int i = 0; i <= NBATCH; i++) {
        if (t < 32) {
            for (int q = 0; q < 3; q++) {
                if (3 * i + q >= BLK) break;
                float2 *out = sm.dem.ph[i & 1][q];
#pragma unroll 10
                for (int j = 0; j < SYM; ++j) {               // (unrolled: the stores and the loop leave the dependent chain alone)
                    if (t == 0) out[j] = ph;
                    ph = cmul(ph, inc);
                }
                const float a = cabs2(ph);
                ph = make_float2(ph.x / a, ph.y / a);
            }
        } else if (i > 0) {
            // the magnitudes of batch i-2, in symbol order (its three warps finished before the barrier that started this
            // iteration): consumer warp 0 adds them before it starts on its own symbol
            if (!bins && want_mag && cw == 0 && i >= 2) {
                const int nb = min(3, BLK - 3 * (i - 2));
                for (int b = lane; b < NMAG; b += 32) {
                    float m = sm.dem.mag[b];
                    for (int q = 0; q < nb; q++) m += sm.dem.magv[i & 1][q][b];
                    sm.dem.mag[b] = m;
                }
            }
            const int sym = 3 * (i - 1) + cw;
            if (sym < BLK && !(L.dbg & 16)) {               // (bit 4: timing experiment, the NCO chain alone - results are wrong)
                const float2 *pv = sm.dem.ph[(i - 1) & 1][cw];
                float2 *f = sm.dem.fft[cw];
                for (int j = lane; j < FFT; j += 32) {
                    const float2 sample = cmul(pv[j], w.buf[sym * SYM + j + samperr]);
                    f[(j + offset) % FFT] = j < CP ? cscale(sample, tb.shape[j]) : sample;
                }
                __syncwarp();
                for (int j = FFT + lane; j < SYM; j += 32) {
                    const float2 sample = cmul(pv[j], w.buf[sym * SYM + j + samperr]);
                    const int idx = (j + offset) % FFT;
                    f[idx] = cadd(f[idx], cscale(sample, tb.shape[j]));
                }
                __syncwarp();
                for (int k = lane; k < FFT; k += 32) {            // bit reversal
                    const int r = tb.brev[k];
                    if (r > k) {
                        const float2 tmp = f[k];
                        f[k] = f[r];
                        f[r] = tmp;
                    }
                }
                __syncwarp();
                for (int half = 1; half < FFT; half <<= 1) {
                    const int tstep = FFT / (2 * half);
#pragma unroll
                    for (int bf = lane; bf < FFT / 2; bf += 32) {
                        const int grp = bf / half, k = bf - grp * half;
                        const int i0 = grp * 2 * half + k, i1 = i0 + half;
                        const float2 tt = cmul(f[i1], tb.tw[k * tstep]);
                        const float2 u = f[i0];
                        f[i0] = cadd(u, tt);
                        f[i1] = make_float2(u.x - tt.x, u.y - tt.y);
                    }
                    __syncwarp();
                }
                // spec[k] = fft[(k + 128) % 256] (fftshift, defines.h:123-138)
                if (bins) {
                    for (int b = CENTER - MAX_IDX + lane; b <= CENTER + MAX_IDX; b += 32) bins[b][sym] = f[(b + FFT / 2) % FFT];
                } else {
                    if (lane == 0) sm.dem.carrier[sym] = f[(CENTER + FFT / 2) % FFT];
                    if (want_mag)
                        for (int b = lane; b < NMAG; b += 32) sm.dem.magv[(i - 1) & 1][cw][b] = cabs2(f[(CENTER - PIDS_OUTER + b + FFT / 2) % FFT]);
                }
            }
        }
        __syncthreads();
    }
    if (!bins && want_mag && cw == 0) {                        // the last batch's magnitudes
        const int nb = BLK - 3 * (NBATCH - 1);
        for (int b = lane; b < NMAG; b += 32) {
            float m = sm.dem.mag[b];
            for (int q = 0; q < nb; q++) m += sm.dem.magv[(NBATCH - 1) & 1][q][b];
            sm.dem.mag[b] = m;
        }
    }
    if (t == 0) sm.dem.phase_end[0] = ph;
    __syncthreads();
    phase = sm.dem.phase_end[0];
    __syncthreads();
}
#endif

template <typename FixHeader>
AM_HD inline void process_window(AmState &st, AmWork &w, const AmTables &tb, const AmIo &io, Lanes L, FixHeader fix_header)
{
    int samperr = 0;
    float angle, angle_diff;
    const long long start = st.start;
    long long tlap = AM_T0();
    (void)tlap;

    if (st.state == ST_FINE) {                                                            // acquire.c:110-119
        samperr = SYM / 2 + st.samperr;
        st.samperr = 0;
        angle_diff = -st.angle;
        st.angle = 0;
        angle = st.prev_angle + angle_diff;
        st.prev_angle = angle;
    } else {                                                                              // acquire.c:120-158
        for (int i = L.lane; i < NACQ; i += L.n) {
            int accr = 0, acci = 0;
            auto at = [&](int pos, int c) -> int {
                if (pos >= 0) return AM_LDIN(io.iq + 2 * (start + pos) + c);
                return w.bp_hist[31 + pos][c];
            };
            for (int k = 1; k < 16; k++) {
                accr = (short)(accr + (((at(i - 31 + k, 0) + at(i - 31 + 32 - k, 0)) * tb.bp_tap[k]) >> 15));
                acci = (short)(acci + (((at(i - 31 + k, 1) + at(i - 31 + 32 - k, 1)) * tb.bp_tap[k]) >> 15));
            }
            accr = (short)(accr + ((at(i - 31 + 16, 0) * tb.bp_tap[16]) >> 15));
            acci = (short)(acci + ((at(i - 31 + 16, 1) * tb.bp_tap[16]) >> 15));
            w.buf[i] = make_float2((float)accr / 32767.0f, (float)acci / 32767.0f);
        }
        AM_SYNC();
        for (int t = L.lane; t < 31; t += L.n) {
            w.bp_hist[t][0] = AM_LDIN(io.iq + 2 * (start + NACQ - 31 + t));
            w.bp_hist[t][1] = AM_LDIN(io.iq + 2 * (start + NACQ - 31 + t) + 1);
        }
        for (int i = L.lane; i < SYM; i += L.n) {
            float2 acc = make_float2(0.f, 0.f);
            for (int j = 0; j < BLK; ++j) acc = cadd(acc, cmul(w.buf[i + j * SYM], cconj(w.buf[i + j * SYM + FFT])));
            w.sums[i] = acc;
        }
        AM_SYNC();
        float max_mag = -1.0f;
        float2 max_v = make_float2(0.f, 0.f);
        for (int i = 0; i < SYM; ++i) {
            float2 v = make_float2(0.f, 0.f);
            for (int j = 0; j < CP; ++j) {
                const float2 sv = w.sums[(i + j) % SYM];
                v.x += (sv.x * tb.shape[j]) * tb.shape[j + FFT];
                v.y += (sv.y * tb.shape[j]) * tb.shape[j + FFT];
            }
            const float mag = v.x * v.x + v.y * v.y;
            if (mag > max_mag) {
                max_mag = mag;
                max_v = v;
                samperr = (i + SYM - 15) % SYM;
            }
        }
        angle_diff = carg(cmul(max_v, cexpj(-st.prev_angle)));
        const float factor = (st.prev_angle != 0.0f) ? 0.25f : 1.0f;
        angle = st.prev_angle + (angle_diff * factor);
        st.prev_angle = angle;
        set_state(st, io, L, ST_COARSE);
    }
    AM_SYNC();
#if defined(__CUDA_ARCH__)
    {   // acquire.c:160-161; one 4-byte load per sample (I and Q together), four in flight per thread
        const int *iqw = reinterpret_cast<const int *>(io.iq) + start;
        int i = L.lane;
        for (; i + 3 * L.n < NACQ; i += 4 * L.n) {
            const int v0 = __ldcg(iqw + i), v1 = __ldcg(iqw + i + L.n), v2 = __ldcg(iqw + i + 2 * L.n), v3 = __ldcg(iqw + i + 3 * L.n);
            w.buf[i] = make_float2((float)(short)(v0 & 0xffff) / 32767.0f, (float)(short)(v0 >> 16) / 32767.0f);
            w.buf[i + L.n] = make_float2((float)(short)(v1 & 0xffff) / 32767.0f, (float)(short)(v1 >> 16) / 32767.0f);
            w.buf[i + 2 * L.n] = make_float2((float)(short)(v2 & 0xffff) / 32767.0f, (float)(short)(v2 >> 16) / 32767.0f);
            w.buf[i + 3 * L.n] = make_float2((float)(short)(v3 & 0xffff) / 32767.0f, (float)(short)(v3 >> 16) / 32767.0f);
        }
        for (; i < NACQ; i += L.n) w.buf[i] = input_at(io, start + i);
    }
#else
    for (int i = L.lane; i < NACQ; i += L.n) w.buf[i] = input_at(io, start + i);          // acquire.c:160-161
#endif
    AM_SYNC();

#if defined(__CUDA_ARCH__)
    if (st.state == ST_FINE && L.lane == 0) w.ph_cyc[10] += (unsigned long long)(clock64() - tlap);
#endif
    AM_LAP(w, L, 0, tlap);
    angle = (float)((double)angle - 2 * PI * st.cfo);                                     // acquire.c:164-168
    st.phase = cmul(st.phase, cexpj((float)(-(SYM / 2 - samperr)) * angle / (float)FFT));
    float2 phase_increment = cexpj(angle / (float)FFT);

    {   // AM only (acquire.c:170-235): carrier phase slope over the block, strongest bin while acquiring
        float y = 0, sum_y = 0, sum_xy = 0, sum_x2 = 0;
        float2 last_carrier = make_float2(0.f, 0.f);
        float2 temp_phase = st.phase;
        float mag_sums[2 * PIDS_OUTER + 1];
        for (int j = 0; j < 2 * PIDS_OUTER + 1; j++) mag_sums[j] = 0;
#if defined(__CUDA_ARCH__)
        AmSmem &sm = *static_cast<AmSmem *>(L.smem);
        demod_pass(w, tb, L, samperr, temp_phase, phase_increment, nullptr, st.state != ST_FINE);
        for (int i = 0; i < BLK; ++i) {                        // the regression, on the carriers the pass collected
            const float2 carrier = sm.dem.carrier[i];
            const float x = SYM * (i - (float)(BLK - 1) / 2);
            if (i == 0) y = carg(carrier);
            else y += carg(cdiv(carrier, last_carrier));
            last_carrier = carrier;
            sum_y += y;
            sum_xy += x * y;
            sum_x2 += x * x;
        }
        if (st.state != ST_FINE)
            for (int j = 0; j < 2 * PIDS_OUTER + 1; j++) mag_sums[j] = sm.dem.mag[j];
        AM_SYNC();
#else
        for (int i = 0; i < BLK; ++i) {
            symbol_fft(w, tb, L, i, samperr, temp_phase, phase_increment);
            const float x = SYM * (i - (float)(BLK - 1) / 2);
            if (i == 0) y = carg(w.spec[CENTER]);
            else y += carg(cdiv(w.spec[CENTER], last_carrier));
            last_carrier = w.spec[CENTER];
            sum_y += y;
            sum_xy += x * y;
            sum_x2 += x * x;
            if (st.state != ST_FINE)
                for (int j = 0; j < 2 * PIDS_OUTER + 1; j++) mag_sums[j] += cabs2(w.spec[CENTER - PIDS_OUTER + j]);
            AM_SYNC();
        }
#endif
        if (st.state != ST_FINE) {
            float mm = -1.0f;
            int max_index = -1;
            for (int j = 0; j < 2 * PIDS_OUTER + 1; j++)
                if (mag_sums[j] > mm) {
                    mm = mag_sums[j];
                    max_index = CENTER - PIDS_OUTER + j;
                }
            st.cfo += max_index - CENTER;
        }
        phase_increment = cmul(phase_increment, cexpj(-sum_xy / sum_x2));
        st.phase = cmul(st.phase, cexpj((float)((double)(-sum_y / BLK + (sum_xy / sum_x2) * (BLK) * SYM / 2) - 0.06)));
    }

    AM_LAP(w, L, 1, tlap);
#if defined(__CUDA_ARCH__)
    demod_pass(w, tb, L, samperr, st.phase, phase_increment, w.bins, false);                // acquire.c:237-257, sync_push
#else
    for (int i = 0; i < BLK; ++i) {                                                       // acquire.c:237-257
        symbol_fft(w, tb, L, i, samperr, st.phase, phase_increment);
        for (int b = CENTER - MAX_IDX + L.lane; b <= CENTER + MAX_IDX; b += L.n) w.bins[b][i] = w.spec[b];   // sync_push
        AM_SYNC();
    }
#endif
    AM_LAP(w, L, 2, tlap);
    sync_block(st, w, tb, io, L, fix_header);

    const int keep = SYM + (SYM / 2 - samperr) + st.keep_extra;                           // acquire.c:259-262
    st.keep_extra = 0;
    st.start += NACQ - keep;
    st.blocks_done++;
}

// ---- cu8 input (reference src/input.c:52-117 in AM mode, src/firdecim_q15.c:137-165) ----
// (u8 - 127) * 64 >> 4, then five cascaded halfband decimators: 1 488 375 S/s cu8 -> 46 511.72 S/s cs16.  Every
// stage is y[m] = x[2m-7] + sum_i ((x[2m-14+2i] + x[2m-2i]) * tap[i]) >> 15 in wrapping int16 arithmetic, the
// output taken when the even sample of a pair has gone in.  The cascade is feed-forward, so output k is a pure
// function of the raw samples 32k-434 .. 32k (zeros before the stream's first sample: the reference's filter
// windows start cleared): a tile of DEC_T outputs recomputes its 434 samples of history instead of carrying
// filter state, and tiles are independent.
//
//   algorithmic bytes per cs16 output: 64 read (32 cu8 samples) + 4 written
//
// The raw samples live in a per-stream ring of `ring_bytes` (a power of two); `raw_avail` counts the raw complex
// samples written so far (absolute index).
constexpr int DEC_T = 64;                      // cs16 outputs per tile
constexpr int DEC_NX = 32 * (DEC_T - 1) + 435, DEC_N0 = 16 * (DEC_T - 1) + 211, DEC_N1 = 8 * (DEC_T - 1) + 99,
              DEC_N2 = 4 * (DEC_T - 1) + 43, DEC_N3 = 2 * (DEC_T - 1) + 15;
constexpr int DEC_RAW_HISTORY = 434;           // raw samples before 32k that output k depends on

struct DecimScratch {
    short2 x[DEC_NX], s0[DEC_N0], s1[DEC_N1], s2[DEC_N2], s3[DEC_N3];
};

AM_HD inline short2 halfband_q15(const short2 *w, int c)      // c: index of the newest (even) sample, c >= 14
{
    // taps of src/input.c:33-38 reversed and truncated to int16 (src/firdecim_q15.c:37-41)
    const int t0 = -134, t1 = 1078, t2 = -4417, t3 = 19864;
    int re = w[c - 7].x, im = w[c - 7].y;
    re += ((w[c - 14].x + w[c].x) * t0) >> 15;
    im += ((w[c - 14].y + w[c].y) * t0) >> 15;
    re += ((w[c - 12].x + w[c - 2].x) * t1) >> 15;
    im += ((w[c - 12].y + w[c - 2].y) * t1) >> 15;
    re += ((w[c - 10].x + w[c - 4].x) * t2) >> 15;
    im += ((w[c - 10].y + w[c - 4].y) * t2) >> 15;
    re += ((w[c - 8].x + w[c - 6].x) * t3) >> 15;
    im += ((w[c - 8].y + w[c - 6].y) * t3) >> 15;
    short2 y;
    y.x = (short)re;                               // the reference accumulates in int16: wraps modulo 2^16
    y.y = (short)im;
    return y;
}

// Outputs k0 .. k0+nout-1 (nout <= DEC_T) of one stream into out[0..nout-1].  Every raw sample they need
// (up to 32*(k0+nout-1)) must have been written: raw_avail >= 32*(k0+nout).
AM_HD inline void decim_tile(const uint8_t *ring, unsigned ring_bytes, long long raw_avail, long long k0, int nout,
                             short2 *out, DecimScratch &sc, Lanes L)
{
    const long long base = 32 * k0 - DEC_RAW_HISTORY;
    const unsigned mask = ring_bytes - 1;
    for (int i = L.lane; i < DEC_NX; i += L.n) {
        const long long a = base + i;
        short2 v;
        v.x = 0;
        v.y = 0;
        if (a >= 0 && a < raw_avail) {
            const unsigned o = (unsigned)((2 * a) & mask);      // rings hold whole samples: 2a and 2a+1 are adjacent
            v.x = (short)((((int)AM_LDIN(ring + o) - 127) * 64) >> 4);
            v.y = (short)((((int)AM_LDIN(ring + o + 1) - 127) * 64) >> 4);
        }
        sc.x[i] = v;
    }
    AM_BLOCK_SYNC();
    for (int m = L.lane; m < DEC_N0; m += L.n) sc.s0[m] = halfband_q15(sc.x, 14 + 2 * m);
    AM_BLOCK_SYNC();
    for (int m = L.lane; m < DEC_N1; m += L.n) sc.s1[m] = halfband_q15(sc.s0, 14 + 2 * m);
    AM_BLOCK_SYNC();
    for (int m = L.lane; m < DEC_N2; m += L.n) sc.s2[m] = halfband_q15(sc.s1, 14 + 2 * m);
    AM_BLOCK_SYNC();
    for (int m = L.lane; m < DEC_N3; m += L.n) sc.s3[m] = halfband_q15(sc.s2, 14 + 2 * m);
    AM_BLOCK_SYNC();
    for (int m = L.lane; m < nout; m += L.n) out[m] = halfband_q15(sc.s3, 14 + 2 * m);
    AM_BLOCK_SYNC();
}

}  // namespace nbam
