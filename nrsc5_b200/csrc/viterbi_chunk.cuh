// Kernels of the chunk-parallel exact Viterbi decoder; see viterbi_pack.cuh for the method.
#pragma once
#include "viterbi_pack.cuh"

namespace nb {

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
#if defined(NB_EMU)
    unsigned char *vitc_smem = emu::dyn_smem();
#else
    extern __shared__ __align__(16) unsigned char vitc_smem[];
#endif
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
