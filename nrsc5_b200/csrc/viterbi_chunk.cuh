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
    asm("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ unsigned vmax16(unsigned a, unsigned b)
{
    unsigned r;
    asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ unsigned vmin16(unsigned a, unsigned b)
{
    unsigned r;
    asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
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

// ===========================================================================
// Kernels.  A "frame" f has 3*len soft values at vin + f*3*len (len % 32 == 0),
// total = len+64 trellis steps, nch = ceil(total / CH_LEN) chunks, and per-frame
// work areas:
//   dec    [nch*CH_LEN + VITC_HEAD] uint2  decision history (layout above)
//   vspec  [nch][16]      uint2  metrics (E,O per lane) at each chunk start, speculative
//   vend   [nch][16]      uint2  metrics at each chunk end
//   hstate [nch]          int    state at the chunk's lower boundary found by the merge trace (-1: no merge)
//   tbend  [nch]          int    survivor state after the last step of each chunk
//   bitsw  [len/32]       u32    decoded bits, bit k of word w = frame bit 32w+k
// ready[f*ready_stride] != 0 selects the frames to decode; slow[f*ready_stride]
// is set when a frame must use saturating arithmetic.
// ===========================================================================
struct VitcArgs {
    const int8_t *vin;
    uint2 *dec;
    uint2 *vspec;
    uint2 *vend;
    int *hstate;
    int *tbend;
    uint32_t *bitsw;
    const int *ready;
    int *slow;
    int ready_stride;       // in ints
    int len;
    int nch;
    size_t dec_stride;      // uint2 per frame
};

constexpr int VITC_FWD_WARPS = 4;                      // 8 chunks per CTA
constexpr int VITC_HEAD = VITC_HEAD_STEPS;               // look-back that makes all 64 survivors merge (checked, not assumed)

__device__ __forceinline__ int vitc_prev_head(int state, const uint2 *hd, int q)
{
    const int ll = state & 15, qq = state >> 4;
    const uint2 w = hd[(q >> 4) * 16 + ll];
    const unsigned word = (qq & 2) ? w.y : w.x;
    return ((state << 1) & 62) | (int)((word >> (((qq & 1) << 4) + (q & 15))) & 1u);
}

__global__ void __launch_bounds__(VITC_FWD_WARPS * 32) k_vitc_fwd(VitcArgs a)
{
    const int f = blockIdx.y;
    if (!a.ready[(size_t)f * a.ready_stride]) return;
    __shared__ int sh_flag;
    __shared__ uint2 sh_head[2 * VITC_FWD_WARPS][VITC_HEAD];      // first VITC_HEAD steps of each chunk
    const int t = threadIdx.x, lane = t & 31, l = lane & 15, half = lane >> 4, warp = t >> 5;
    const int total = a.len + 64;
    const int8_t *vin = a.vin + (size_t)f * 3 * a.len;
    // Saturation pre-check over this CTA's step range (incl. warm-up): between two
    // normalisations the largest metric grows by at most the sum of |s0|+|s1|+|s2|
    // over the 79 steps in between, on top of a spread of at most 12*381 right after
    // a normalisation.  If that stays below 32767 wrapping == saturating arithmetic.
    {
        const int c_first = blockIdx.x * (2 * VITC_FWD_WARPS);
        const int s_lo = max(0, c_first * CH_LEN - CH_WARM), s_hi = min(total, (c_first + 2 * VITC_FWD_WARPS) * CH_LEN);
        if (t == 0) sh_flag = 1;
        __syncthreads();
        const int w_lo = s_lo / VITC_NORM, w_hi = (s_hi + VITC_NORM - 1) / VITC_NORM;
        for (int wdw = w_lo + t; wdw < w_hi; wdw += blockDim.x) {
            int sum = 0;
            for (int k = 1; k <= VITC_NORM; k++) {
                int s = wdw * VITC_NORM + k;
                if (s >= total) break;
                int j = s + a.len - 32;
                while (j >= a.len) j -= a.len;
                const int8_t *q = vin + 3 * j;
                sum += abs((int)q[0]) + abs((int)q[1]) + abs((int)q[2]);
            }
            if (sum + 12 * 381 > 32767) sh_flag = 0;
        }
        __syncthreads();
        if (!sh_flag) {
            if (t == 0) atomicExch(&a.slow[(size_t)f * a.ready_stride], 1);
            return;                                    // the whole frame is redone sequentially in k_vitc_ends
        }
    }
    const int c = blockIdx.x * (2 * VITC_FWD_WARPS) + warp * 2 + half;
    const bool valid = c < a.nch;
    const int cc = valid ? c : a.nch - 1;
    uint2 *dec = a.dec + (size_t)f * a.dec_stride;
    VitHalf<false> vh;
    vh.init(l);
    // warm-up from zero metrics (chunk 0 replays zero soft values: the true initial condition)
    vitc_run<false>(vh, vin, a.len, total, cc * CH_LEN - CH_WARM, CH_WARM, 0, dec, false, l);
    if (valid) a.vspec[((size_t)f * a.nch + cc) * 16 + l] = make_uint2(vh.E, vh.O);
    uint2 *head = sh_head[warp * 2 + half];
    vitc_run<false>(vh, vin, a.len, total, cc * CH_LEN, CH_LEN, cc * CH_LEN, dec, valid, l, head);
    if (valid) a.vend[((size_t)f * a.nch + cc) * 16 + l] = make_uint2(vh.E, vh.O);
    __syncwarp();
    // merge trace over the chunk head: every survivor at step lo+VITC_HEAD-1 walked back to the chunk's
    // lower boundary; if all 64 agree, that state lies on the final path whatever happens later
    {
        const int n = min(VITC_HEAD, total - cc * CH_LEN);
        int first = -1;
        bool same = true;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int state = l + 16 * r;
            for (int q = n - 1; q >= 0; q--) state = vitc_prev_head(state, head, q);
            if (r == 0) first = state;
            same &= state == first;
        }
        const int ref = __shfl_sync(0xffffffffu, first, 0, 16);
        same &= first == ref;
        const unsigned hm = half ? 0xffff0000u : 0x0000ffffu;
        const bool all = (__ballot_sync(0xffffffffu, same) & hm) == hm;
        if (valid && l == 0) a.hstate[(size_t)f * a.nch + cc] = all ? ref : -1;
    }
}

__device__ __forceinline__ bool vitc_same_shape(uint2 x, uint2 y)
{
    // equal up to one constant added to all 64 metrics (reference: the first metric of lane 0)
    const int xr = (short)(__shfl_sync(0xffffffffu, x.x, 0, 16) & 0xffff);
    const int yr = (short)(__shfl_sync(0xffffffffu, y.x, 0, 16) & 0xffff);
    auto lo = [](unsigned v) { return (int)(short)(v & 0xffff); };
    auto hi = [](unsigned v) { return (int)(short)(v >> 16); };
    bool ok = (lo(x.x) - xr == lo(y.x) - yr) && (hi(x.x) - xr == hi(y.x) - yr) &&
              (lo(x.y) - xr == lo(y.y) - yr) && (hi(x.y) - xr == hi(y.y) - yr);
    return __all_sync(0xffffffffu, ok);
}

// One warp per frame: accept / repair the speculative chunks, pick the end state,
// and fix the survivor state at every chunk boundary.
__global__ void __launch_bounds__(32) k_vitc_ends(VitcArgs a)
{
    const int f = blockIdx.x;
    if (!a.ready[(size_t)f * a.ready_stride]) return;
    const int t = threadIdx.x, l = t & 15;
    const int total = a.len + 64, nch = a.nch;
    const int8_t *vin = a.vin + (size_t)f * 3 * a.len;
    uint2 *dec = a.dec + (size_t)f * a.dec_stride;
    const uint2 *vspec = a.vspec + (size_t)f * nch * 16, *vend = a.vend + (size_t)f * nch * 16;
    int *tbend = a.tbend + (size_t)f * nch, *hstate = a.hstate + (size_t)f * nch;

    uint2 cur;
    if (a.slow[(size_t)f * a.ready_stride]) {
        // exact saturating arithmetic, sequential over the whole frame
        VitHalf<true> vs;
        vs.init(l);
        vitc_run<true>(vs, vin, a.len, total, 0, total, 0, dec, true, l);
        cur = make_uint2(vs.E, vs.O);
        for (int c = t; c < nch; c += 32) hstate[c] = -1;
    } else {
        cur = vend[l];
        uint2 sp_next = vspec[(size_t)(nch > 1 ? 1 : 0) * 16 + l], ve_next = vend[(size_t)(nch > 1 ? 1 : 0) * 16 + l];
        for (int c = 1; c < nch; c++) {
            const uint2 sp = sp_next, ve = ve_next;
            if (c + 1 < nch) {                         // prefetch the next chunk's vectors
                sp_next = vspec[(size_t)(c + 1) * 16 + l];
                ve_next = vend[(size_t)(c + 1) * 16 + l];
            }
            if (vitc_same_shape(cur, sp)) {
                cur = ve;
            } else {                                   // speculation missed: redo this chunk from the true metrics
                VitHalf<false> vh;
                vh.init(l);
                vh.E = cur.x;
                vh.O = cur.y;
                vitc_run<false>(vh, vin, a.len, total, c * CH_LEN, CH_LEN, c * CH_LEN, dec, true, l);
                cur = make_uint2(vh.E, vh.O);
                if (t == 0) hstate[c] = -1;            // its merge trace was made on discarded decisions
            }
        }
    }
    // first maximum in state order (reference src/conv_dec.c:310-317); lane l holds states 2l, 2l+32, 2l+1, 2l+33
    int v = (short)(cur.x & 0xffff), idx = 2 * l;
    const int w1 = (short)(cur.y & 0xffff);
    if (w1 > v) { v = w1; idx = 2 * l + 1; }
    int v2 = (short)(cur.x >> 16), idx2 = 2 * l + 32;
    const int w3 = (short)(cur.y >> 16);
    if (w3 > v2) { v2 = w3; idx2 = 2 * l + 33; }
    if (v2 > v) { v = v2; idx = idx2; }                    // equal values: the lower state index stays
#pragma unroll
    for (int o = 8; o; o >>= 1) {
        const int ov = __shfl_xor_sync(0xffffffffu, v, o, 16), oi = __shfl_xor_sync(0xffffffffu, idx, o, 16);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    __syncwarp();
    __threadfence_block();
    if (t == 0) {
        tbend[nch - 1] = idx;
        for (int c = nch - 1; c >= 1; c--) {
            int hs = hstate[c];
            if (hs < 0) {                              // rare: walk the whole chunk back from its known end state
                int state = tbend[c];
                const int hi = min(total, (c + 1) * CH_LEN), lo = c * CH_LEN;
                for (int q = hi - 1; q >= lo; q--) state = vitc_prev(state, dec, q);
                hs = state;
            }
            tbend[c - 1] = hs;
        }
    }
}

constexpr int VITC_EMIT_WARPS = 4;                     // one warp per chunk
constexpr int VITC_EMIT_STEPS = CH_LEN + VITC_HEAD;    // staged decisions per chunk

// Emit the decoded bits of a chunk from its known end state.  The chunk is cut
// into 32 segments of 32 steps, one per lane; each lane *guesses* its segment's
// end state by walking an arbitrary survivor back over the VITC_HEAD steps that
// follow the segment, emits its 32 bits, and the guesses are then checked
// against the neighbouring segment's start state from the (known) chunk end
// downwards; a wrong guess is repaired by re-walking that segment.
__global__ void __launch_bounds__(VITC_EMIT_WARPS * 32) k_vitc_emit(VitcArgs a)
{
    extern __shared__ __align__(16) unsigned char vitc_smem[];
    const int f = blockIdx.y;
    if (!a.ready[(size_t)f * a.ready_stride]) return;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int c = blockIdx.x * VITC_EMIT_WARPS + warp;
    if (c >= a.nch) return;
    const int total = a.len + 64;
    uint2 *sd = reinterpret_cast<uint2 *>(vitc_smem) + (size_t)warp * VITC_EMIT_STEPS;
    const uint2 *dec = a.dec + (size_t)f * a.dec_stride + (size_t)c * CH_LEN;
    const int lo = c * CH_LEN;
    const int n = min(total - lo, CH_LEN);             // steps of this chunk
    const int nst = min(total - lo, VITC_EMIT_STEPS);  // staged steps (incl. look-ahead into the next chunk)
    for (int i = lane; i < nst; i += 32) sd[i] = dec[i];
    __syncwarp();
    const int true_end = a.tbend[(size_t)f * a.nch + c];
    const int seg_end = 32 * lane + 31;                // local step index of this lane's last step
    const bool have = 32 * lane < n;
    auto walk = [&](int state, int from, int to) {     // state after local step `from` -> state after step `to`-... down to > to
        for (int q = from; q > to; q--) state = vitc_prev_head(state, sd, q);
        return state;
    };
    int g;                                             // state after local step seg_end
    if (!have) g = 0;
    else if (seg_end == n - 1) g = true_end;
    else if (seg_end + VITC_HEAD >= nst) g = walk(true_end, n - 1, seg_end);      // frame end inside the look-ahead (last chunk)
    else g = walk(0, seg_end + VITC_HEAD, seg_end);                               // guess
    unsigned word = 0;
    int b = 0;                                         // state before the segment's first step
    auto emit = [&](int gstate) {
        int state = gstate;
        unsigned w = 0;
        for (int k = 31; k >= 0; k--) {
            w |= (unsigned)((state >> 5) & 1) << k;
            state = vitc_prev_head(state, sd, 32 * lane + k);
        }
        word = w;
        b = state;
    };
    if (have) emit(g);
    // verification from the chunk end downwards
    const int nseg = (n + 31) / 32;
    for (;;) {
        const int bnext = __shfl_down_sync(0xffffffffu, b, 1);
        const bool bad = have && lane < nseg - 1 && g != bnext;
        const unsigned m = __ballot_sync(0xffffffffu, bad);
        if (!m) break;
        const int jj = 31 - __clz(m);                  // highest wrong segment: its right neighbour is already final
        if (lane == jj) { g = bnext; emit(g); }
    }
    const int widx = (lo >> 5) + lane - 1;             // frame bit 32*widx = step lo + 32*lane - 32
    if (have && widx >= 0 && widx < a.len / 32) a.bitsw[(size_t)f * (a.len / 32) + widx] = word;
}

// bits (one per byte) from the packed words — used by the stage-level test entry point
__global__ void k_vitc_unpack(const uint32_t *bitsw, uint8_t *out, size_t nbits)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nbits) out[i] = (uint8_t)((bitsw[i >> 5] >> (i & 31)) & 1u);
}

}  // namespace nb
