// L2 framing on the device: what the reference's frame_push / frame_process do with every L1
// PDU (reference src/frame.c:645-714, 516-643) - PCI extraction, the RS(255,247)-protected audio PDU headers
// (:158-196), packet locations (:315-326), header expansion fields (:198-265), PSD carried as HDLC frames with an
// FCS-16 (:328-391, :138-144), CRC-8 of every packet (:130-136, :617-620), and the fixed-data tail with its CCC and
// subchannel blocks (:393-514) - so that a frame's HDC packets, PSD messages and service changes leave the GPU as
// one REC_L2 record and the host has no L2 work left.
//
// One CTA per stream takes the frames of a pass in the order the reference would see them:
//   all threads : bit order swap + PCI removal -> PDU bytes in shared memory        (frame.c:692-710)
//   warp 0      : the walk over the PDU (frame_process).  Only the position of the next audio PDU is sequential;
//                 inside one, the 32 lanes share the work: RS(255,247) header decode (rs.cuh: syndromes over the
//                 lanes, all-zero early out), packet locations and their ordering checks by ballot, the PSD bytes'
//                 HDLC flag scan by ballot with whole segments appended at once, a closed HDLC frame pulled into
//                 shared memory before it is unescaped and checked, events written a word per lane.  The rare
//                 fixed-data tail (CCC, subchannel blocks) stays a byte-serial state machine on lane 0.
//   all threads : CRC-8 of the packets (one packet per thread), then the copy of events + PDU bytes into the log
// Per-stream state that outlives a frame (frame_t in the reference: service table, PSD assembly buffers, CCC and
// subchannel state) lives in L2State in device memory.
#pragma once
#include "common.cuh"
#include "rs.cuh"

namespace nbl2 {

constexpr int L2_THREADS = 128;
constexpr int L2_PDU_MAX = 18269;            // (146176 - 24) / 8, reference src/defines.h:63
constexpr int L2_AAS_MAX = 8212;             // reference src/frame.h:5
constexpr int L2_EV_CAP = 96 * 1024;         // staging for one frame's events
constexpr int L2_PK_MAX = 1024;              // packets whose CRC-8 the CTA computes in parallel (more: thread 0 does them)
constexpr int L2_RING = 64;                  // ELASTIC_BUFFER_LEN, reference src/defines.h:71
constexpr uint32_t REC_L2 = 20;              // include/nrsc5_b200.h
constexpr uint32_t EV_SERVICE = 16, EV_ALIGN = 17, EV_AAS = 18, EV_PACKET = 19;
constexpr uint32_t L2F_LOST = 1, L2F_EV_OVERFLOW = 2;
constexpr unsigned L2_NO_AUDIO = 0xffffffffu;    // fixed_tail: the frame is dropped (its CCC announces more than it holds)

struct L2Sub {                               // fixed_subchannel_t, reference src/frame.h:19-28
    unsigned length, fill;
    int idx;
    uint8_t blk[260];
    uint8_t data[L2_AAS_MAX];
};

struct L2Ccc {                               // ccc_data_t, reference src/frame.h:30-38
    unsigned width, count;
    int ccc_idx, ready;
    uint8_t ccc[32];
    L2Sub sub[4];
};

struct L2State {
    int svc[8][7];                           // access, type, codec mode, blend control, gain, common delay, latency
    int psd_idx[8];
    unsigned frames;                         // frames taken since the stream was (re)created
    L2Ccc ccc[3];
    uint8_t psd[8][L2_AAS_MAX];
    uint8_t ev[L2_EV_CAP];
};

// where the records go: a stream's log in the engine, a plain buffer in the stage entry point
struct L2Sink {
    uint8_t *base;
    size_t cap;
    unsigned *len;
    unsigned *overflow;
};

__device__ inline void l2_reset(L2State &z)                        // frame_reset, reference src/frame.c:716-742
{
    for (int p = 0; p < 8; p++) {
        for (int k = 0; k < 7; k++) z.svc[p][k] = -1;
        z.psd_idx[p] = -1;
    }
    for (int c = 0; c < 3; c++) {
        z.ccc[c].ready = 0;
        z.ccc[c].width = 0;
        z.ccc[c].count = 0;
        z.ccc[c].ccc_idx = -1;
    }
}

struct EvWriter {
    uint8_t *base;
    unsigned len, overflow;
};

// appends {type, plen, a, b} and returns the payload position (nullptr when the staging area is full)
__device__ inline uint8_t *ev_put(EvWriter &w, uint32_t type, const void *a, unsigned alen, const uint8_t *b, unsigned blen)
{
    const unsigned plen = alen + blen, need = 8 + ((plen + 3) & ~3u);
    if (w.overflow || w.len + need > (unsigned)L2_EV_CAP) {          // once full, stay full: the record keeps a prefix of the calls
        w.overflow = 1;
        return nullptr;
    }
    uint8_t *q = w.base + w.len;
    reinterpret_cast<uint32_t *>(q)[0] = type;
    reinterpret_cast<uint32_t *>(q)[1] = plen;
    const uint8_t *ab = static_cast<const uint8_t *>(a);
    for (unsigned i = 0; i < alen; i++) q[8 + i] = ab[i];
    for (unsigned i = 0; i < blen; i++) q[8 + alen + i] = b[i];
    for (unsigned i = plen; i < need - 8; i++) q[8 + i] = 0;
    w.len += need;
    return q + 8;
}

// warp versions (all 32 lanes call them with the same arguments; `w` lives in shared memory): a record of `nw`
// payload words held by every lane, and a record whose payload is `n` bytes of shared memory
__device__ inline uint8_t *ev_put_words(EvWriter &w, uint32_t type, const uint32_t *words, unsigned nw, int lane)
{
    const unsigned need = 8 + 4 * nw, at = w.len;
    const bool full = w.overflow || at + need > (unsigned)L2_EV_CAP;
    __syncwarp();
    if (full) {
        if (lane == 0) w.overflow = 1;
        __syncwarp();
        return nullptr;
    }
    uint32_t *q = reinterpret_cast<uint32_t *>(w.base + at);
    if (lane == 0) { q[0] = type; q[1] = 4 * nw; w.len = at + need; }
    if ((unsigned)lane < nw) q[2 + lane] = words[lane];
    __syncwarp();
    return w.base + at + 8;
}

__device__ inline void ev_put_bytes(EvWriter &w, uint32_t type, const uint8_t *src, unsigned n, int lane)
{
    const unsigned need = 8 + ((n + 3) & ~3u), at = w.len;
    const bool full = w.overflow || at + need > (unsigned)L2_EV_CAP;
    __syncwarp();
    if (full) {
        if (lane == 0) w.overflow = 1;
        __syncwarp();
        return;
    }
    uint8_t *q = w.base + at;
    if (lane == 0) {
        reinterpret_cast<uint32_t *>(q)[0] = type;
        reinterpret_cast<uint32_t *>(q)[1] = n;
        w.len = at + need;
    }
    for (unsigned i = lane; i < need - 8; i += 32) q[8 + i] = i < n ? src[i] : (uint8_t)0;
    __syncwarp();
}

// CRC tables of the frame being processed, in shared memory (built by the CTA at the start of l2_frame):
// fcs[x] = x pushed through 8 steps of the reflected polynomial 0x8408 (RFC 1662, reference src/frame.c:95-128),
// crc8[x] = x through 8 steps of x^8 + x^5 + x^4 + 1 (reference src/frame.c:60-93)
struct CrcTabs {
    uint16_t fcs[256];
    uint8_t crc8[256];
};

__device__ inline unsigned fcs16_of(const CrcTabs &tb, const uint8_t *p, int n)        // reference src/frame.c:138-144
{
    unsigned c = 0xFFFF;
    for (int i = 0; i < n; i++) c = (c >> 8) ^ tb.fcs[(c ^ p[i]) & 0xFFu];
    return c;
}

__device__ inline unsigned crc8_step(unsigned c, unsigned byte)     // x^8 + x^5 + x^4 + 1, reference src/frame.c:60-93
{
    c ^= byte;
#pragma unroll
    for (int k = 0; k < 8; k++) c = ((c << 1) ^ ((c & 0x80) ? 0x31u : 0u)) & 0xFFu;
    return c;
}

__device__ inline int hdlc_unescape(uint8_t *d, int n)              // reference src/frame.c:328-341
{
    int w = 0;
    for (int i = 0; i < n; i++) {
        if (d[i] == 0x7D) d[w++] = d[++i] | 0x20;
        else d[w++] = d[i];
    }
    return w;
}

// a complete HDLC frame of PSD / AAS data: aas_push, reference src/frame.c:343-367
__device__ inline void aas_frame(EvWriter &w, const CrcTabs &tb, uint8_t *d, int n)
{
    n = hdlc_unescape(d, n);
    if (n == 0) return;
    if (fcs16_of(tb, d, n) != 0xF0B8u) return;
    if (d[0] != 0x21) return;
    ev_put(w, EV_AAS, nullptr, 0, d + 1, (unsigned)(n - 3));
}

// a complete HDLC frame of the channel-configuration channel: process_fixed_ccc, reference src/frame.c:393-438
__device__ inline void ccc_frame(L2Ccc &c, const CrcTabs &tb, uint8_t *d, int n0)
{
    const unsigned n = (unsigned)hdlc_unescape(d, n0);
    if (n == 0 || c.ready) return;
    if (fcs16_of(tb, d, (int)n) != 0xF0B8u) return;
    for (unsigned i = 0; i < 4; i++) {
        L2Sub &s = c.sub[i];
        s.length = 0;
        if (5 + 4 * i <= n) {
            const unsigned mode = d[1 + 4 * i] | (d[2 + 4 * i] << 8), len = d[3 + 4 * i] | (d[4 + 4 * i] << 8);
            if (mode == 0) {
                s.length = len;
                s.fill = 0;
                s.idx = -1;
            }
        }
    }
    c.ready = 1;
}

// parse_hdlc, reference src/frame.c:369-391.  ccc != nullptr: frames go to ccc_frame, else to aas_frame
__device__ inline void hdlc_scan(EvWriter &w, const CrcTabs &tb, L2Ccc *ccc, uint8_t *acc, int *idx, int cap, const uint8_t *in,
                                 unsigned n)
{
    int k = *idx;
    for (unsigned i = 0; i < n; i++) {
        const uint8_t b = in[i];
        if (b == 0x7E) {
            if (k >= 0) {
                if (ccc) ccc_frame(*ccc, tb, acc, k);
                else aas_frame(w, tb, acc, k);
            }
            k = 0;
        } else if (k >= 0) {
            if (k == cap) k = -1;
            else acc[k++] = b;
        }
    }
    *idx = k;
}

// process_fixed_data, reference src/frame.c:448-514: returns where the audio part of the PDU ends
__device__ inline unsigned fixed_tail(L2State &z, EvWriter &w, const CrcTabs &tb, const uint8_t *pdu, unsigned length, unsigned lc)
{
    L2Ccc &c = z.ccc[lc];
    unsigned pos = length - 1;
    if (c.count < 2) {
        const unsigned b = pdu[pos];
        const unsigned wd = b == 0 ? 1u : ((b >> 4) == (b & 15u) ? (b & 15u) * 2u : 0u);
        c.count = (wd > 0 && c.width == wd) ? c.count + 1 : 0;
        c.width = wd;
        if (c.count < 2) return pos;
    }
    pos -= c.width;
    hdlc_scan(w, tb, &c, c.ccc, &c.ccc_idx, 32, pdu + pos, c.width);
    if (!c.ready) return pos;
    // a CCC that announces more subchannel bytes than the PDU holds (noise that passed the FCS-16) would make the
    // reference read in front of its buffer (frame.c:493-496); here such a frame is dropped whole - nothing is
    // consumed, no subchannel state changes, and it has no audio part (L2_NO_AUDIO: l2_walk returns at once)
    {
        unsigned total = 0;
        for (int i = 0; i < 4; i++) total += c.sub[i].length;
        if (total > pos) return L2_NO_AUDIO;
    }
    for (int i = 3; i >= 0; i--) {
        L2Sub &s = c.sub[i];
        if (s.length == 0) continue;
        pos -= s.length;
        for (unsigned j = 0; j < s.length; j++) {
            s.blk[s.fill++] = pdu[pos + j];
            if (s.fill == 4 && !(s.blk[0] == 0x7D && s.blk[1] == 0x3A && s.blk[2] == 0xE2 && s.blk[3] == 0x42)) {
                s.blk[0] = s.blk[1];                                  // not on a block boundary yet: slide by one byte
                s.blk[1] = s.blk[2];
                s.blk[2] = s.blk[3];
                s.fill = 3;
            }
            if (s.fill == 259) {
                hdlc_scan(w, tb, nullptr, s.data, &s.idx, L2_AAS_MAX, s.blk + 4, 255);
                s.fill = 0;
            }
        }
    }
    return pos;
}

struct Hef {
    unsigned prog, access, type;
};

// parse_hef, reference src/frame.c:198-265: bytes consumed (all n of them when the field list is cut short)
__device__ inline unsigned hef_walk(const uint8_t *b, unsigned n, Hef &h)
{
    unsigned i = 0;
    for (;;) {
        if (i >= n) return n;
        const unsigned v = b[i];
        switch ((v >> 4) & 7u) {
        case 1:
            h.prog = (v >> 1) & 7u;
            if (v & 1u) {
                if (i + 2 >= n) return n;
                i += 2;                                               // PDU length: not used downstream
            }
            break;
        case 2:
            if (i + 1 >= n) return n;
            h.access = (v >> 3) & 1u;
            h.type = ((v & 1u) << 7) | (b[i + 1] & 0x7Fu);
            i += 1;
            break;
        case 3: {
            const unsigned skip = (v & 8u) ? 4 : 3;
            if (i + skip >= n) return n;
            i += skip;
            break;
        }
        case 4: {
            const unsigned skip = (v & 8u) ? 3 : 1;                   // applied services + PDU marker: not used downstream
            if (i + skip >= n) return n;
            i += skip;
            break;
        }
        default:
            break;
        }
        if (!(b[i++] & 0x80u)) return i;
    }
}

struct PkEntry {
    unsigned start, cnt, ev;                 // first byte, payload bytes, staging offset of the packet's flags word
};

// a complete HDLC frame of PSD data (aas_push, reference src/frame.c:343-367), by one warp: the n accumulated bytes
// are pulled into shared memory in one sweep (plus the byte behind them, which a trailing escape reads), unescaped and
// checked there, and - like the reference, which unescapes in place - written back
__device__ inline void aas_frame_warp(EvWriter &w, const CrcTabs &tb, uint8_t *acc, int n, uint8_t *scratch, int lane)
{
    if (n == 0) return;
    for (int i = lane; i <= n; i += 32) scratch[i] = acc[i];           // acc[n] lies inside L2State (see hdlc_unescape)
    __syncwarp();
    int m = 0, ok = 0;
    if (lane == 0) {
        m = hdlc_unescape(scratch, n);
        ok = m != 0 && fcs16_of(tb, scratch, m) == 0xF0B8u && scratch[0] == 0x21;
    }
    m = __shfl_sync(0xffffffffu, m, 0);
    ok = __shfl_sync(0xffffffffu, ok, 0);
    for (int i = lane; i < m; i += 32) acc[i] = scratch[i];
    if (ok) ev_put_bytes(w, EV_AAS, scratch + 1, (unsigned)(m - 3), lane);
    __syncwarp();
}

// parse_hdlc (reference src/frame.c:369-391) over the PSD bytes of one audio PDU, by one warp: 32 bytes at a time,
// the 0x7E flags found by ballot, the bytes between two flags appended to the accumulator in one step
__device__ inline void hdlc_scan_psd_warp(EvWriter &w, const CrcTabs &tb, uint8_t *acc, int *idx, const uint8_t *in, unsigned n,
                                          uint8_t *scratch, int lane)
{
    int k = *idx;
    __syncwarp();
    for (unsigned base = 0; base < n; base += 32) {
        const unsigned i = base + (unsigned)lane;
        const bool valid = i < n;
        const unsigned b = valid ? in[i] : 0u;
        unsigned rem = __ballot_sync(0xffffffffu, valid && b == 0x7Eu);
        const int lim = (int)min(32u, n - base);
        int seg = 0;
        for (;;) {
            const int f = rem ? __ffs((int)rem) - 1 : lim;               // end of this run of data bytes (exclusive)
            const int cnt = f - seg;
            if (k >= 0 && cnt > 0) {
                // byte by byte: stored while k < cap; the byte that arrives at k == cap invalidates the frame
                const int room = L2_AAS_MAX - k;
                if (lane >= seg && lane < f && lane - seg < room) acc[k + lane - seg] = (uint8_t)b;
                k = cnt > room ? -1 : k + cnt;
            }
            if (!rem) break;
            if (k >= 0) {
                __syncwarp();
                aas_frame_warp(w, tb, acc, k, scratch, lane);
            }
            k = 0;
            rem &= rem - 1;
            seg = f + 1;
        }
    }
    __syncwarp();
    if (lane == 0) *idx = k;
    __syncwarp();
}

// the walk over one PDU (frame_process, reference src/frame.c:516-643) by warp 0: control flow is uniform (every lane
// reads the same PDU bytes and state), single writes are lane 0's, the pieces named above are shared by the lanes
__device__ inline void l2_walk(L2State &z, EvWriter &w, const CrcTabs &tb, const nb::GfTab &gf, uint8_t *pdu, unsigned length,
                               unsigned lc, uint32_t pci, PkEntry *pk, unsigned &npk, unsigned &flags, uint8_t *scratch, int lane)
{
    const uint32_t k = pci & 0xFFFFFCu;
    const bool fixed = k == (0xE3634Cu & 0xFFFFFCu) || k == (0x8D8D33u & 0xFFFFFCu) || k == (0x3634CEu & 0xFFFFFCu);
    unsigned end = length, off = 0;
    if (lc > 2u) return;                                              // L2Ccc[3]: P1, P3, P4 (callers check; never index past it)
    if (fixed) {
        // byte-serial state machine over CCC and subchannel blocks: lane 0 (rare; the others wait)
        if (lane == 0) end = fixed_tail(z, w, tb, pdu, length, lc);
        end = __shfl_sync(0xffffffffu, end, 0);
    }
    if (end == L2_NO_AUDIO) return;
    if (k == (0x3634CEu & 0xFFFFFCu)) return;                         // fixed data only: no audio
    while (off < end - 96u) {                                         // unsigned on purpose, as frame.c:527
        const unsigned start = off;
        uint8_t *h = pdu + off;
        if (!nb::rs8_fix_header_warp(gf, h, lane)) {
            if ((length == 18269u || length == 466u) && off == 0) flags |= L2F_LOST;     // frame.c:535-540
            return;
        }
        const unsigned codec = h[8] & 15u, stream = (h[8] >> 4) & 3u, pdu_seq = (h[8] >> 6) | ((h[9] & 1u) << 2);
        const unsigned blend = (h[9] >> 1) & 3u, gain = h[9] >> 3, common = h[10] & 0x3Fu;
        const unsigned latency = (h[10] >> 6) | ((h[11] & 1u) << 2), pfirst = (h[11] >> 1) & 1u, plast = (h[11] >> 2) & 1u;
        const unsigned seq0 = (h[11] >> 3) | ((h[12] & 1u) << 5), nop = (h[12] >> 1) & 0x3Fu, has_hef = h[12] >> 7;
        const unsigned la = h[13];
        off += 14;
        const bool narrow = (codec >= 1 && codec <= 3) ? stream == 0 : (codec == 10 || codec == 13);   // calc_lc_bits
        const unsigned lbits = narrow ? 12u : 16u, lbytes = (lbits * nop + 4) / 8;
        if (start + la + 1 < off + lbytes || start + la >= end) return;
        // packet locations: lane l holds those of packets l and l + 32 (nop <= 63); the reference gives up at the first
        // one that is not behind its predecessor or lies outside the audio part - nothing has been emitted by then
        unsigned loc2[2] = { 0, 0 };
        bool bad = false;
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const unsigned j = (unsigned)lane + 32u * hh;
            unsigned v = 0;
            if (j < nop) {
                const uint8_t *q = pdu + off;
                if (!narrow) v = q[2 * j] | (q[2 * j + 1] << 8);
                else {
                    const uint8_t *r = q + (j >> 1) * 3;
                    v = (j & 1u) ? ((unsigned)r[2] << 4) | (r[1] >> 4) : ((r[1] & 15u) << 8) | r[0];
                }
            }
            loc2[hh] = v;
        }
        const unsigned loc31 = __shfl_sync(0xffffffffu, loc2[0], 31);
        unsigned prev2[2];
        prev2[0] = __shfl_up_sync(0xffffffffu, loc2[0], 1);
        prev2[1] = __shfl_up_sync(0xffffffffu, loc2[1], 1);
        if (lane == 0) { prev2[0] = la; prev2[1] = loc31; }
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            const unsigned j = (unsigned)lane + 32u * hh;
            if (j < nop && (loc2[hh] <= prev2[hh] || start + loc2[hh] >= end)) bad = true;
        }
        if (__any_sync(0xffffffffu, bad)) return;
        const unsigned last_loc = nop ? (nop > 32 ? __shfl_sync(0xffffffffu, loc2[1], (int)(nop - 33)) : __shfl_sync(0xffffffffu, loc2[0], (int)(nop - 1)))
                                      : 0u;
        off += lbytes;
        if (stream >= 2) {                                            // MAX_STREAMS: skip this PDU
            if (nop == 0) return;                                     // (the reference would index locations[-1] here)
            off = start + last_loc + 1;
            continue;
        }
        Hef hef = { 0, 0, 0 };
        if (has_hef) off += hef_walk(pdu + off, end - off, hef);
        const unsigned prog = hef.prog;
        int *sv = z.svc[prog];
        const int now[7] = { (int)hef.access, (int)hef.type, (int)codec, (int)blend, (int)gain, (int)common, (int)latency };
        bool changed = false;
        for (int i = 0; i < 7; i++) changed |= sv[i] != now[i];
        __syncwarp();
        if (stream == 0 && changed) {
            if (lane < 7) sv[lane] = now[lane];
            const uint32_t r[8] = { prog, (uint32_t)now[0], (uint32_t)now[1], (uint32_t)now[2], (uint32_t)now[3],
                                    (uint32_t)(now[4] < 16 ? now[4] : now[4] - 32), (uint32_t)(now[5] * 4), (uint32_t)(now[6] * 2) };
            ev_put_words(w, EV_SERVICE, r, 8, lane);
        }
        unsigned avg;                                                 // calc_avg_packets, frame.c:289-313
        if (codec >= 1 && codec <= 3) avg = stream == 0 ? 4 : 32;
        else if (codec == 10) avg = stream == 0 ? 32 : 4;
        else avg = codec == 13 ? 4 : 32;
        const unsigned seq = (L2_RING + seq0 - pfirst) % L2_RING;
        unsigned out_off = (L2_RING + pdu_seq * avg - latency * 2) % L2_RING;
        if ((L2_RING + seq - out_off) % L2_RING >= L2_RING / 2) out_off = (out_off + L2_RING / 2) % L2_RING;
        {
            const uint32_t al[3] = { prog, stream, out_off };
            ev_put_words(w, EV_ALIGN, al, 3, lane);
        }
        // header expansion fields that run past la_location make the reference's byte count wrap (frame.c:608: it
        // would scan 4 GB); nothing is scanned here
        const unsigned psd_bytes = start + la + 1 >= off ? start + la + 1 - off : 0u;
        hdlc_scan_psd_warp(w, tb, z.psd[prog], &z.psd_idx[prog], pdu + off, psd_bytes, scratch, lane);
        off = start + la + 1;
        // the packets: packet j spans PDU bytes start + prev_j + 1 .. start + loc_j (payload + CRC byte); their
        // output_push events are equally long, so they are written side by side, a packet per lane
        {
            constexpr unsigned EVB = 8 + 28;
            const unsigned at = w.len;
            const unsigned fit = w.overflow ? 0u : min(nop, ((unsigned)L2_EV_CAP - at) / EVB);
            __syncwarp();
#pragma unroll
            for (int hh = 0; hh < 2; hh++) {
                const unsigned j = (unsigned)lane + 32u * hh;
                if (j >= nop) continue;
                const unsigned p0 = start + prev2[hh] + 1, cnt = loc2[hh] - prev2[hh] - 1;
                const unsigned shape = (j == 0 && pfirst) ? 3u : (j == nop - 1 && plast) ? 2u : 1u;      // output.h:28-31
                const bool in_table = npk + j < (unsigned)L2_PK_MAX;
                unsigned crc_flag = 0;
                if (!in_table) {                                      // table full: this lane checks the CRC itself
                    unsigned c = 0xFF;
                    for (unsigned i = 0; i <= cnt; i++) c = tb.crc8[c ^ pdu[p0 + i]];
                    crc_flag = c ? 1u : 0u;
                }
                if (j < fit) {
                    uint32_t *q = reinterpret_cast<uint32_t *>(w.base + at + j * EVB);
                    q[0] = EV_PACKET; q[1] = 28;
                    q[2] = prog; q[3] = stream; q[4] = (seq + j) % L2_RING; q[5] = shape; q[6] = crc_flag; q[7] = cnt; q[8] = p0;
                    if (in_table) {
                        pk[npk + j].start = p0;
                        pk[npk + j].cnt = cnt;
                        pk[npk + j].ev = at + j * EVB + 8 + 16;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                w.len = at + fit * EVB;
                if (fit < nop) w.overflow = 1;
            }
            npk = min((unsigned)L2_PK_MAX, npk + fit);
            __syncwarp();
        }
        off = nop ? start + last_loc + 1 : off;
    }
}

// frame geometry of frame_push, reference src/frame.c:651-690
__device__ inline bool l2_geometry(unsigned nbits, unsigned &first, unsigned &step, unsigned &npci)
{
    switch (nbits) {
    case 146176: first = 146176 - 30000; step = 1248; npci = 24; return true;
    case 4608: first = 120; step = 184; npci = 24; return true;
    case 2304: first = 120; step = 88; npci = 24; return true;
    case 3750: first = 120; step = 160; npci = 22; return true;
    case 24000: first = 120; step = 992; npci = 24; return true;
    case 30000: first = 120; step = 1240; npci = 24; return true;
    default: return false;
    }
}

// bit i of the frame after the per-byte order swap (frame.c:694-697); packed = the REC_FRAME bytes, MSB first
__device__ __forceinline__ unsigned swapped_bit(const uint8_t *packed, unsigned nbits, unsigned i)
{
    const unsigned base = i & ~7u, span = min(8u, nbits - base), src = base + span - 1 - (i & 7u);
    return (packed[src >> 3] >> (7 - (src & 7u))) & 1u;
}

// One L1 PDU through L2, by the whole CTA.  frame_off: what the REC_L2 record names as its frame (the engine passes
// the log offset of the frame's packed bits).
__device__ inline void l2_frame(L2State &z, const uint8_t *packed, unsigned nbits, unsigned lc, unsigned frame_off,
                                const L2Sink &sink)
{
    __shared__ __align__(16) uint8_t pdu[(L2_PDU_MAX + 15) & ~15];
    __shared__ PkEntry pk[L2_PK_MAX];
    __shared__ __align__(16) uint8_t scratch[(L2_AAS_MAX + 1 + 15) & ~15];     // one HDLC frame while it is unescaped and checked
    __shared__ CrcTabs tb;
    __shared__ nb::GfTab gf;
    __shared__ EvWriter w;
    __shared__ unsigned sh_pci, sh_npk, sh_flags, sh_evlen;
    __shared__ uint8_t *sh_out;
    const unsigned t = threadIdx.x, nt = blockDim.x;
    unsigned first, step, npci;
    if (!l2_geometry(nbits, first, step, npci)) return;
    const unsigned nout = (nbits - npci) / 8;
    if (t == 0) {
        sh_pci = 0;
        w.base = z.ev;
        w.len = 0;
        w.overflow = 0;
    }
    nb::gf_tab_load(gf, (int)t, (int)nt);
    for (unsigned x = t; x < 256; x += nt) {
        unsigned f = x;
#pragma unroll
        for (int k = 0; k < 8; k++) f = (f >> 1) ^ ((f & 1u) ? 0x8408u : 0u);
        tb.fcs[x] = (uint16_t)f;
        tb.crc8[x] = (uint8_t)crc8_step(0, x);
    }
    __syncthreads();
    if (t < npci) atomicOr(&sh_pci, swapped_bit(packed, nbits, first + step * t) << (23 - t));
    // PDU byte n = frame bits (after the swap) 8n .. 8n+7, counted without the PCI bits
    for (unsigned n = t; n < nout; n += nt) {
        const unsigned o0 = 8 * n, o7 = o0 + 7;
        const unsigned i0 = o0 + (o0 < first ? 0u : min(npci, (o0 - first) / (step - 1) + 1));
        const unsigned i7 = o7 + (o7 < first ? 0u : min(npci, (o7 - first) / (step - 1) + 1));
        unsigned v;
        if (i7 - i0 == 7 && i7 + 8 < (nbits & ~7u)) {                 // no PCI bit inside, whole source bytes
            const unsigned b = i0 >> 3, sft = i0 & 7u;
            const unsigned two = ((__brev((unsigned)packed[b]) >> 24) << 8) | (__brev((unsigned)packed[b + 1]) >> 24);
            v = (two >> (8 - sft)) & 0xFFu;
        } else {
            v = 0;
            for (unsigned o = o0; o <= o7; o++) {
                const unsigned i = o + (o < first ? 0u : min(npci, (o - first) / (step - 1) + 1));
                v = (v << 1) | swapped_bit(packed, nbits, i);
            }
        }
        pdu[n] = (uint8_t)v;
    }
    __syncthreads();
    if (t < 32) {                                                     // warp 0 walks the PDU
        unsigned npk = 0, flags = 0;
        l2_walk(z, w, tb, gf, pdu, nout, lc, sh_pci, pk, npk, flags, scratch, (int)t);
        __syncwarp();
        if (t == 0) {
            if (w.overflow) flags |= L2F_EV_OVERFLOW;
            sh_npk = npk;
            sh_flags = flags;
            sh_evlen = w.len;
            z.frames++;
        }
    }
    __syncthreads();
    for (unsigned j = t; j < sh_npk; j += nt) {                       // CRC-8 over payload + check byte: 0 when intact
        unsigned c = 0xFF;
        const uint8_t *q = pdu + pk[j].start;
        for (unsigned i = 0; i <= pk[j].cnt; i++) c = tb.crc8[c ^ q[i]];
        *reinterpret_cast<uint32_t *>(z.ev + pk[j].ev) = c ? 1u : 0u;
    }
    __syncthreads();
    const unsigned evlen = sh_evlen, body = 32 + evlen + ((nout + 3) & ~3u);
    if (t == 0) {
        uint8_t *out = nullptr;
        const unsigned need = 8 + body;
        if ((size_t)*sink.len + need > sink.cap) *sink.overflow = 1;
        else {
            out = sink.base + *sink.len;
            uint32_t *h = reinterpret_cast<uint32_t *>(out);
            h[0] = REC_L2; h[1] = body;
            h[2] = frame_off; h[3] = lc; h[4] = nbits; h[5] = sh_pci; h[6] = sh_flags; h[7] = nout; h[8] = evlen;
            h[9] = z.frames - 1;
            *sink.len += need;
        }
        sh_out = out;
    }
    __syncthreads();
    uint8_t *out = sh_out;
    if (!out) return;
    out += 40;
    for (unsigned i = t; i < evlen / 4; i += nt)
        reinterpret_cast<uint32_t *>(out)[i] = reinterpret_cast<const uint32_t *>(z.ev)[i];
    out += evlen;
    for (unsigned i = t; i < (nout + 3) / 4; i += nt) {
        const unsigned b = 4 * i;
        uint32_t v = pdu[b];
        if (b + 1 < nout) v |= (uint32_t)pdu[b + 1] << 8;
        if (b + 2 < nout) v |= (uint32_t)pdu[b + 2] << 16;
        if (b + 3 < nout) v |= (uint32_t)pdu[b + 3] << 24;
        reinterpret_cast<uint32_t *>(out)[i] = v;
    }
    __syncthreads();
}

}  // namespace nbl2
