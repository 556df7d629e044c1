// nrsc5_b200 engine: per-block control, acquisition, sync/equalise/demap,
// L1 decode kernels and the host-side C ABI (include/nrsc5_b200.h).
//
// This translation unit is compiled with -fmad=false: the acquisition and
// sync arithmetic keeps the reference's float operation order so that every
// discrete decision (timing arg-max, reference-subcarrier votes, rounded
// timing error) is taken on the same values as the reference computes, up to
// the GPU's libm.  The FFT-heavy demodulator lives in frontend.cu.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/nrsc5_b200.h"
#include "common.cuh"
#include "rs.cuh"
#include "viterbi.cuh"
#include "viterbi_chunk.cuh"

namespace nb {

void launch_demod(const DevPtrs &p, const EngineDims &d, cudaStream_t stream);
void launch_fft_test(const float2 *in, float2 *out, const float2 *twid, int nffts, cudaStream_t stream);
void launch_halfband_test(const uint8_t *cu8, long long npairs, short2 *out, cudaStream_t stream);

__device__ unsigned long long g_progress;      // bumped by every stream that processed a block

__constant__ int c_compat_mode[64];
__constant__ short c_bp_tap[32];               // coarse band-pass taps, tap[i] pairs w[i] and w[32-i]

// complex helpers with the reference's (gcc, no FMA) evaluation order
__device__ __forceinline__ float2 cmulf(float2 a, float2 b)
{
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cexp_j(float a)       // cexpf(I*a)
{
    float s, c;
    sincosf(a, &s, &c);
    return make_float2(c, s);
}

__device__ __forceinline__ int partitions_per_band(int psmi)
{
    switch (c_compat_mode[psmi & 63]) {
    case 2: return 11;
    case 3: return 12;
    case 5: case 6: case 11: return 14;
    default: return 10;
    }
}

// input_set_sync_state (reference src/input.c:172-188)
__device__ void set_state(const DevPtrs &p, const EngineDims &d, int s, int ns)
{
    StreamState &st = p.st[s];
    if (st.state == ns) return;
    if (st.state == ST_FINE) log_reserve(p, d, s, REC_LOST_SYNC, 0);
    if (ns == ST_FINE) {
        float fo = (float)(((double)st.prev_angle - 2 * M_PI * st.cfo) * 744187.5 / (2 * M_PI * NFFT));
        uint8_t *w = log_reserve(p, d, s, REC_SYNC, 8);
        if (w) {
            reinterpret_cast<float *>(w)[0] = fo;
            reinterpret_cast<int *>(w)[1] = st.psmi;
        }
    }
    st.state = ns;
}

// ===========================================================================
// k_prep: per stream, decide whether a full 33-symbol window is buffered, run
// coarse acquisition when not in FINE, and publish the block's NCO.
//   reference src/acquire.c:98-168 (+ src/sync.c:769-777, src/firdecim_q15.c:95-109,154-158)
// ===========================================================================
constexpr int PREP_THREADS = 512;

__global__ void __launch_bounds__(PREP_THREADS) k_prep(DevPtrs p, EngineDims d)
{
    const int s = blockIdx.x, t = threadIdx.x;
    StreamState &st = p.st[s];
    __shared__ int sh_active;
    __shared__ float2 sums[NSYM];
    __shared__ float red_mag[PREP_THREADS];
    __shared__ int red_idx[PREP_THREADS];
    __shared__ float2 red_v[PREP_THREADS];
    __shared__ int sh_samperr;
    __shared__ float sh_angle;

    if (t == 0) {
        if (st.force_state >= 0) {
            set_state(p, d, s, st.force_state);
            st.force_state = -1;
        }
        int act = st.in_avail >= 2 * (st.start + NACQ);
        st.active = act;
        sh_active = act;
        if (act) atomicAdd(&g_progress, 1ull);
    }
    __syncthreads();
    if (!sh_active) return;

    const uint8_t *iq = p.iq + (size_t)s * d.in_stride;
    const int state_in = st.state;

    if (state_in != ST_FINE) {
        short2 *y = p.ydec + (size_t)s * NACQ;
        float2 *tb = p.tbuf + (size_t)s * NACQ;
        for (int i = t; i < NACQ; i += PREP_THREADS) y[i] = halfband_at(iq, st.start + i);
        __syncthreads();
        // 32-tap symmetric Q15 band-pass; history = last 31 samples of the previous coarse window
        for (int i = t; i < NACQ; i += PREP_THREADS) {
            auto at = [&](int pos) -> short2 {
                if (pos >= 0) return y[pos];
                return make_short2(st.bp_hist[31 + pos][0], st.bp_hist[31 + pos][1]);
            };
            short accr = 0, acci = 0;
#pragma unroll 5
            for (int k = 1; k < 16; k++) {
                short2 a = at(i - 31 + k), b = at(i - 31 + 32 - k);
                accr = (short)(accr + ((((int)a.x + (int)b.x) * c_bp_tap[k]) >> 15));
                acci = (short)(acci + ((((int)a.y + (int)b.y) * c_bp_tap[k]) >> 15));
            }
            short2 c = at(i - 31 + 16);
            accr = (short)(accr + (((int)c.x * c_bp_tap[16]) >> 15));
            acci = (short)(acci + (((int)c.y * c_bp_tap[16]) >> 15));
            tb[i] = make_float2(__fdiv_rn((float)accr, 32767.0f), __fdiv_rn((float)acci, -32767.0f));
        }
        __syncthreads();
        if (t < 31) {
            short2 v = y[NACQ - 31 + t];
            st.bp_hist[t][0] = v.x;
            st.bp_hist[t][1] = v.y;
        }
        // cyclic-prefix correlation per sample offset (acquire.c:129-134)
        for (int i = t; i < NSYM; i += PREP_THREADS) {
            float2 acc = make_float2(0.f, 0.f);
            for (int j = 0; j < BLK; j++) {
                float2 a = tb[i + j * NSYM], b = tb[i + j * NSYM + NFFT];
                float2 bc = make_float2(b.x, -b.y);
                float2 pr = cmulf(a, bc);
                acc.x += pr.x;
                acc.y += pr.y;
            }
            sums[i] = acc;
        }
        __syncthreads();
        // pulse-shaped sliding sum and arg-max (acquire.c:136-151)
        float best = -1.0f;
        int besti = 0;
        float2 bestv = make_float2(0.f, 0.f);
        for (int i = t; i < NSYM; i += PREP_THREADS) {
            float2 v = make_float2(0.f, 0.f);
            for (int j = 0; j < NCP; j++) {
                int q = i + j;
                if (q >= NSYM) q -= NSYM;
                float2 sm = sums[q];
                float a = p.shape[j], b = p.shape[j + NFFT];
                v.x += (sm.x * a) * b;
                v.y += (sm.y * a) * b;
            }
            float mag = v.x * v.x + v.y * v.y;
            if (mag > best) { best = mag; besti = i; bestv = v; }
        }
        red_mag[t] = best; red_idx[t] = besti; red_v[t] = bestv;
        __syncthreads();
        for (int o = PREP_THREADS / 2; o; o >>= 1) {
            if (t < o) {
                float m2 = red_mag[t + o];
                int i2 = red_idx[t + o];
                if (m2 > red_mag[t] || (m2 == red_mag[t] && i2 < red_idx[t])) {
                    red_mag[t] = m2; red_idx[t] = i2; red_v[t] = red_v[t + o];
                }
            }
            __syncthreads();
        }
        if (t == 0) {
            float2 mv = red_v[0];
            float2 w = cmulf(mv, cexp_j(-st.prev_angle));
            float angle_diff = atan2f(w.y, w.x);
            float factor = (st.prev_angle != 0.0f) ? 0.25f : 1.0f;
            float angle = st.prev_angle + (angle_diff * factor);
            st.prev_angle = angle;
            sh_angle = angle;
            sh_samperr = (red_idx[0] + NSYM - 15) % NSYM;
            if (st.state == ST_NONE) st.state = ST_COARSE;
        }
    } else if (t == 0) {
        sh_samperr = NSYM / 2 + st.samperr;
        st.samperr = 0;
        float angle = st.prev_angle + (-st.angle);
        st.angle = 0;
        st.prev_angle = angle;
        sh_angle = angle;
    }
    __syncthreads();

    const int samperr = sh_samperr;
    const int adj = NSYM / 2 - samperr;
    if (adj != 0) {                                            // sync_adjust, sync.c:769-777
        float *cp = p.cphase + (size_t)s * NFFT;
        for (int i = t; i < SIDE; i += PREP_THREADS) {
            int bl = LB0 + i, bu = UB1 - i;
            cp[bl] = (float)((double)cp[bl] - (double)(adj * (bl - NFFT / 2) * 2) * M_PI / NFFT);
            cp[bu] = (float)((double)cp[bu] - (double)(adj * (bu - NFFT / 2) * 2) * M_PI / NFFT);
        }
    }
    __shared__ float sh_theta;
    if (t == 0) {
        float angle = sh_angle;
        angle = (float)((double)angle - 2 * M_PI * st.cfo);
        float pre = (float)(-adj) * angle / (float)NFFT;
        float2 ph = cmulf(st.phase, cexp_j(pre));
        float theta = angle / (float)NFFT;
        st.phase0 = ph;
        st.theta = theta;
        st.blk_samperr = samperr;
        st.blk_state_in = state_in;
        sh_theta = theta;
        // NCO phase after the 32 symbols of this block (acquire.c:250-252, closed form)
        double sn, cs;
        sincos((double)theta * (double)(NSYM * BLK), &sn, &cs);
        float2 pe = cmulf(ph, make_float2((float)cs, (float)sn));
        float nrm = sqrtf(pe.x * pe.x + pe.y * pe.y);
        st.phase = make_float2(pe.x / nrm, pe.y / nrm);
        uint8_t *w = log_reserve(p, d, s, REC_BLOCK, 32);
        if (w) {
            int *wi = reinterpret_cast<int *>(w);
            float *wf = reinterpret_cast<float *>(w);
            wi[0] = state_in; wi[1] = samperr; wf[2] = angle; wf[3] = ph.x; wf[4] = ph.y; wi[5] = st.cfo;
            wi[6] = (int)(unsigned)(st.start & 0xffffffffLL);
            wi[7] = (int)(st.start >> 32);
        }
    }
    __syncthreads();
    {
        const float theta = sh_theta;
        float2 *nco = p.nco + (size_t)s * NSYM;
        for (int j = t; j < NSYM; j += PREP_THREADS) {
            float2 e = cexp_j(theta * (float)j);
            float w = (j < NCP || j >= NFFT) ? p.shape[j] : 1.0f;
            nco[j] = make_float2(e.x * w, e.y * w);
        }
    }
}

// ===========================================================================
// k_sync: per stream and block — Costas loops on the reference subcarriers,
// COARSE->FINE decision, channel equalisation, timing/phase feedback, MER,
// soft demapping and the PIDS decode.      reference src/sync.c:90-610,
// src/decode.c:378-391,463-471
// ===========================================================================
constexpr int SYNC_THREADS = 256;
constexpr int ZLD = 33;                         // padded symbols-per-bin stride in shared memory
constexpr int MAXREF = 15;                      // reference subcarriers per sideband (14 partitions + 1)

struct SyncSmem {
    float2 z[NBINS * ZLD];                      // [bin][symbol]
    float phs[2 * MAXREF][BLK];                 // Costas phase per reference and symbol
    float smag[2 * MAXREF];
    float part_lb[BLK], part_ub[BLK];
    float mer_lb[8][BLK], mer_ub[8][BLK];
    float2 zero_row[BLK];
    float tmp_phs[32][BLK];                     // scratch phases for the CFO search
    int ref_ok[2 * MAXREF], ref_bc[2 * MAXREF], ref_psmi[2 * MAXREF];
    int offs[32];
    float mult_lb, mult_ub;
    int flag;
};

__device__ __forceinline__ int ref_bin(int slot, int nref)      // slot < nref: lower, else upper
{
    return slot < MAXREF ? LB0 + PW * slot : UB1 - PW * (slot - MAXREF);
}

// adjust_ref (sync.c:90-130) on one row of 32 symbols
__device__ void costas_row(float2 *z, int zstride, float *phs, float &cfreq, float &cphase, int cfo,
                           float alpha, float beta)
{
    const signed char pat[BLK] = { -1, 1, -1, -1, -1, 1, 1, 0, 1, -1, 0, 0, 0, -1, -1, 0,
                                   0, 0, 0, 0, -1, 1, -1, 0, 0, 0, 0, 0, 0, 0, 0, -1 };
    const float cfo_freq = (float)(2 * M_PI * cfo * NCP / NFFT);
    const float PI_F = 3.14159274101257324f;                  // smallest float above pi: (ph > M_PI) <=> (ph >= PI_F)
    float f = cfreq, ph = cphase;
    for (int n = 0; n < BLK; n++) {
        const float2 v = z[n * zstride];
        // u = v * exp(-j*ph); the loop error arg(v^2 * exp(-2j*ph)) / 2 equals arg(u^2) / 2
        const float2 u = cmulf(v, cexp_j(-ph));
        const float error = atan2f((u.x * u.y) * 2.0f, u.x * u.x - u.y * u.y) * 0.5f;
        phs[n] = ph;
        z[n * zstride] = u;
        f += beta * error;
        if (f > 0.5f) f = 0.5f;
        if (f < -0.5f) f = -0.5f;
        ph += (f + cfo_freq) + (alpha * error);
        if (ph >= PI_F) ph = (float)((double)ph - 2 * M_PI);
        if (ph <= -PI_F) ph = (float)((double)ph + 2 * M_PI);
    }
    float x = 0;
    for (int n = 0; n < BLK; n++) x += z[n * zstride].x * (float)pat[n];
    if (x < 0) {
        for (int n = 0; n < BLK; n++) {
            phs[n] = (float)((double)phs[n] + M_PI);
            float2 v = z[n * zstride];
            z[n * zstride] = make_float2(v.x * -1.0f, v.y * -1.0f);
        }
        ph = (float)((double)ph + M_PI);
    }
    cfreq = f;
    cphase = ph;
}

__device__ __forceinline__ int needle_bit(int n, unsigned rsid)       // -1 = don't care (sync.c:171-174)
{
    const signed char base[BLK] = { 0, 1, 0, 0, 0, 1, 1, -1, 1, 0, 0, 0, -1, 0, 0, -1,
                                    -1, -1, -1, -1, 0, 1, 0, -1, -1, -1, -1, -1, -1, -1, -1, 0 };
    if (n == 10) return (int)(rsid >> 1);
    if (n == 11) return (int)((rsid >> 1) ^ (rsid & 1));
    return base[n];
}

// find_ref_fm (sync.c:188-207): cyclic offset of the sync pattern, also trying the inverted bits
__device__ int ref_find(const float2 *z, int zstride, unsigned rsid)
{
    unsigned raw = 0;
    for (int n = 0; n < BLK; n++)
        if (!(z[n * zstride].x <= 0)) raw |= 1u << n;
    for (int pass = 0; pass < 2; pass++) {
        for (int n = 0; n < BLK; n++) {
            int i;
            for (i = 0; i < BLK; i++) {
                int nb_ = needle_bit(i, rsid);
                if (nb_ < 0) continue;
                if (nb_ != (int)((raw >> ((n + i) & 31)) & 1)) break;
            }
            if (i == BLK) return n;
        }
        raw = ~raw;
    }
    return -1;
}

__device__ __forceinline__ float half_pi_wrap(float a, float b)        // sync.c:284-290
{
    float dd = a - b;
    while ((double)dd > M_PI / 2) dd = (float)((double)dd - M_PI);
    while ((double)dd < -M_PI / 2) dd = (float)((double)dd + M_PI);
    return dd;
}

__device__ __forceinline__ int8_t soft_demap(float x, float mult)      // sync.c:69-73
{
    // lroundf semantics (round half away from zero) without the libm call: |v| <= 127 so v - trunc(v) is exact
    const float v = fmaxf(fminf(x, 1.0f), -1.0f) * mult;
    int r = __float2int_rz(v);
    const float f = v - (float)r;
    if (f >= 0.5f) r++;
    else if (f <= -0.5f) r--;
    return (int8_t)r;
}

__device__ void sync_block(const DevPtrs &p, const EngineDims &d, SyncSmem &sm, int s, int t)
{
    StreamState &st = p.st[s];

    float *cfreq = p.cfreq + (size_t)s * NFFT;
    float *cphase = p.cphase + (size_t)s * NFFT;
    const float loop_bw = 0.05f, damping = 0.70710678f;
    const float denom = 1 + (2 * damping * loop_bw) + (loop_bw * loop_bw);
    const float alpha = (4 * damping * loop_bw) / denom, beta = (4 * loop_bw * loop_bw) / denom;

    // stage the block's 32x534 spectrum as [bin][symbol]
    {
        const float2 *src = p.bins + (size_t)s * BLK * NBINS;
        constexpr int NEL = BLK * NBINS, U = 6;                   // 17088 = 256 * 66 + 192
        for (int i0 = t; i0 < NEL; i0 += SYNC_THREADS * U) {
            float2 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * SYNC_THREADS;
                if (i < NEL) v[u] = __ldg(&src[i]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * SYNC_THREADS;
                if (i < NEL) {
                    const int sym = i / NBINS, ci = i - sym * NBINS;
                    sm.z[ci * ZLD + sym] = v[u];
                }
            }
        }
        if (t < BLK) sm.zero_row[t] = make_float2(0.f, 0.f);
    }
    __syncthreads();

    int ppb = partitions_per_band(st.psmi);
    int nref = ppb + 1;
    // Costas loop on every reference subcarrier (sync.c:359-363)
    if (t < 2 * MAXREF) {
        int side = t / MAXREF, i = t - side * MAXREF;
        if (i < nref) {
            int b = ref_bin(t, nref);
            costas_row(&sm.z[compact_of_bin(b) * ZLD], 1, sm.phs[t], cfreq[b], cphase[b], 0, alpha, beta);
        }
    }
    __syncthreads();

    if (st.state == ST_COARSE) {                 // sync.c:366-421
        if (t < 2 * MAXREF) {
            int side = t / MAXREF, i = t - side * MAXREF;
            sm.ref_ok[t] = 0;
            if (i < nref) {
                const float2 *z = &sm.z[compact_of_bin(ref_bin(t, nref)) * ZLD];
                unsigned rsid = (unsigned)(30 - i) & 3;
                bool ok = true;
                unsigned raw = 0;
                for (int n = 0; n < BLK; n++) {
                    int nbit = needle_bit(n, rsid);
                    int pos = z[n].x > 0 ? 1 : 0;
                    if (nbit >= 0 && nbit != pos) ok = false;
                    if (!(z[n].x <= 0)) raw |= 1u << n;
                }
                unsigned dd = raw ^ (raw << 1);              // DBPSK decode, prev = 0 (sync.c:138-148)
                auto bit = [&](int n) { return (dd >> n) & 1u; };
                sm.ref_ok[t] = ok;
                sm.ref_bc[t] = (int)(bit(16) << 3 | bit(17) << 2 | bit(18) << 1 | bit(19));
                sm.ref_psmi[t] = (int)(bit(25) << 5 | bit(26) << 4 | bit(27) << 3 | bit(28) << 2 | bit(29) << 1 | bit(30));
            }
        }
        __syncthreads();
        __shared__ int sh_do_search;
        if (t == 0) {
            unsigned good = 0;
            for (int r = 0; r < 2 * MAXREF; r++)
                if (sm.ref_ok[r]) good++;
            sh_do_search = 0;
            if (good >= 4) {
                // strict majorities; the PSMI majority is only looked for among 0..15 (sync.c:396)
                int mbc = -1, mps = -1;
                for (int v = 0; v < 16; v++) {
                    unsigned nbc = 0, nps = 0;
                    for (int r = 0; r < 2 * MAXREF; r++) {
                        if (!sm.ref_ok[r]) continue;
                        nbc += sm.ref_bc[r] == v;
                        nps += sm.ref_psmi[r] == v;
                    }
                    if (nbc > good / 2) mbc = v;
                    if (nps > good / 2) mps = v;
                }
                if (mbc >= 0 && mps >= 0) {
                    st.bc = mbc;
                    st.psmi = mps;
                    set_state(p, d, s, ST_FINE);
                    st.started_pm = 0;                   // decode_reset (decode.c:556-565)
                }
            } else if (st.cfo_wait == 0) {
                sh_do_search = 1;
            } else {
                st.cfo_wait--;
            }
        }
        __syncthreads();
        if (sh_do_search && t < 32) {            // detect_cfo (sync.c:292-337), warp 0
            const int lane = t;
            for (int cfo = -2 * PW; cfo < 2 * PW; cfo++) {
                int off = -1;
                if (lane < 22) {
                    int i = lane >> 1, upper = lane & 1;
                    int b = upper ? cfo + UB1 - i * PW : cfo + LB0 + i * PW;
                    int ci = compact_of_bin(b);
                    float2 *row = ci >= 0 ? &sm.z[ci * ZLD] : sm.zero_row;
                    costas_row(row, 1, sm.tmp_phs[lane], cfreq[b], cphase[b], cfo, alpha, beta);
                    off = ref_find(row, 1, (unsigned)(30 - i) & 3);
                    for (int n = 0; n < BLK; n++)            // reset_ref (sync.c:132-136)
                        row[n] = cmulf(row[n], cexp_j(sm.tmp_phs[lane][n]));
                    if (ci < 0)
                        for (int n = 0; n < BLK; n++) row[n] = make_float2(0.f, 0.f);
                }
                sm.offs[lane] = off;
                __syncwarp();
                int found = 0;
                if (lane == 0) {
                    int best = -1;
                    unsigned bestn = 0;
                    for (int k = 0; k < BLK; k++) {
                        unsigned nv = 0;
                        for (int r = 0; r < 22; r++) nv += sm.offs[r] == k;
                        if (nv > bestn) { best = k; bestn = nv; }
                    }
                    if (best >= 0 && bestn >= 3) {
                        st.keep_extra = ((BLK - best) % BLK) * NSYM;
                        st.cfo += cfo;
                        st.cfo_wait = 8;
                        found = 1;
                    }
                }
                found = __shfl_sync(0xffffffffu, found, 0);
                if (found) break;
            }
        }
        __syncthreads();
        ppb = partitions_per_band(st.psmi);      // psmi may have changed with the FINE decision
        nref = ppb + 1;
    }

    if (st.state == ST_FINE) {
        // reference amplitude per subcarrier (calc_smag, sync.c:254-261)
        if (t < 2 * MAXREF) {
            int side = t / MAXREF, i = t - side * MAXREF;
            if (i < nref) {
                const float2 *z = &sm.z[compact_of_bin(ref_bin(t, nref)) * ZLD];
                float sum = 0;
                for (int n = 0; n < BLK; n++) sum += fabsf(z[n].x);
                sm.smag[t] = sum / BLK;
            }
        }
        __syncthreads();
        // adjust_data (sync.c:263-282): one thread per (partition, symbol)
        for (int item = t; item < 2 * ppb * BLK; item += SYNC_THREADS) {
            int n = item & (BLK - 1), pp = item >> 5;
            int upper = pp >= ppb, i = upper ? pp - ppb : pp;
            int lo_bin, slot_lo, slot_hi;
            if (!upper) { lo_bin = LB0 + PW * i; slot_lo = i; slot_hi = i + 1; }
            else { lo_bin = UB1 - PW * i - PW; slot_lo = MAXREF + i + 1; slot_hi = MAXREF + i; }
            float m0 = sm.smag[slot_lo], m19 = sm.smag[slot_hi];
            float2 up = cexp_j(sm.phs[slot_hi][n]);
            float2 lp = cexp_j(sm.phs[slot_lo][n]);
            float2 *zc = &sm.z[compact_of_bin(lo_bin) * ZLD + n];
            for (int k = 1; k < PW; k++) {
                float fa = (float)k * m19, fb = (float)(PW - k) * m0;
                float c = fa * up.x + fb * lp.x, dd = fa * up.y + fb * lp.y;
                const float rden = 19.0f / (c * c + dd * dd);
                // (19 + 19j) / (c + j dd)
                float2 C = make_float2((c + dd) * rden, (c - dd) * rden);
                zc[k * ZLD] = cmulf(zc[k * ZLD], C);
            }
        }
        // timing / phase feedback (sync.c:426-463)
        if (t == 0) {
            float samperr = 0, angle = 0, sum_xy = 0, sum_x2 = 0;
            for (int i = 0; i < ppb; i++) {
                samperr += half_pi_wrap(sm.phs[i][0], sm.phs[i + 1][0]);
                samperr += half_pi_wrap(sm.phs[MAXREF + i + 1][0], sm.phs[MAXREF + i][0]);
            }
            samperr = (float)((double)(samperr / (float)(ppb * 2) * (float)NFFT / (float)PW) / (2 * M_PI));
            for (int i = 0; i <= ppb; i++) {
                float x, y;
                x = (float)(LB0 + PW * i - NFFT / 2);
                y = cfreq[LB0 + PW * i];
                angle += y; sum_xy += x * y; sum_x2 += x * x;
                x = (float)(UB1 - PW * i - NFFT / 2);
                y = cfreq[UB1 - PW * i];
                angle += y; sum_xy += x * y; sum_x2 += x * x;
            }
            samperr = (float)((double)samperr - (double)((sum_xy / sum_x2) * (float)NFFT) / (2 * M_PI) * BLK);
            st.samperr = (int)roundf(samperr);
            angle /= (float)((ppb + 1) * 2);
            st.angle = angle;
            for (int i = 0; i <= ppb; i++) {
                cfreq[LB0 + PW * i] -= angle;
                cfreq[UB1 - PW * i] -= angle;
            }
        }
        __syncthreads();
        // modulation error (sync.c:465-488): 8 threads per symbol take the partitions round-robin, the
        // partial sums are then combined in a fixed order (per symbol, then over symbols)
        {
            const int n = t & (BLK - 1), g = t >> 5;           // SYNC_THREADS == 8 * BLK
            float e_lb = 0, e_ub = 0;
            for (int i = g; i < ppb; i += SYNC_THREADS / BLK)
                for (int j = 1; j < PW; j++) {
                    float2 c = sm.z[compact_of_bin(LB0 + PW * i + j) * ZLD + n];
                    float dx = (c.x >= 0 ? 1.0f : -1.0f) - c.x, dy = (c.y >= 0 ? 1.0f : -1.0f) - c.y;
                    e_lb += dx * dx + dy * dy;
                    c = sm.z[compact_of_bin(UB1 - PW * i - PW + j) * ZLD + n];
                    dx = (c.x >= 0 ? 1.0f : -1.0f) - c.x; dy = (c.y >= 0 ? 1.0f : -1.0f) - c.y;
                    e_ub += dx * dx + dy * dy;
                }
            sm.mer_lb[g][n] = e_lb;
            sm.mer_ub[g][n] = e_ub;
        }
        __syncthreads();
        if (t < BLK) {
            float e_lb = 0, e_ub = 0;
            for (int g = 0; g < SYNC_THREADS / BLK; g++) { e_lb += sm.mer_lb[g][t]; e_ub += sm.mer_ub[g][t]; }
            sm.part_lb[t] = e_lb;
            sm.part_ub[t] = e_ub;
        }
        __syncthreads();
        if (t == 0) {
            float e_lb = 0, e_ub = 0;
            for (int n = 0; n < BLK; n++) { e_lb += sm.part_lb[n]; e_ub += sm.part_ub[n]; }
            st.err_lb += e_lb;
            st.err_ub += e_ub;
            if (++st.mer_cnt == 16) {
                float signal = (float)(2 * BLK * (ppb * 18) * st.mer_cnt);
                uint8_t *w = log_reserve(p, d, s, REC_MER, 8);
                if (w) {
                    reinterpret_cast<float *>(w)[0] = 10 * log10f(signal / st.err_lb);
                    reinterpret_cast<float *>(w)[1] = 10 * log10f(signal / st.err_ub);
                }
                st.mer_cnt = 0;
                st.err_lb = 0;
                st.err_ub = 0;
            }
            const float mer_lb = 2.0f * BLK * (float)(ppb * 18) / e_lb;
            const float mer_ub = 2.0f * BLK * (float)(ppb * 18) / e_ub;
            sm.mult_lb = fmaxf(fminf(mer_lb * 10, 127.0f), 1.0f);
            sm.mult_ub = fmaxf(fminf(mer_ub * 10, 127.0f), 1.0f);
        }
        __syncthreads();
        // soft demap of the primary-main partitions (sync.c:509-536) into the interleaver matrix
        const int bc = st.bc;
        {
            int8_t *pm = p.pm + ((size_t)s * 16 + bc) * PM_BLOCK;
            const float mlb = sm.mult_lb, mub = sm.mult_ub;
            for (int col = t; col < 720; col += SYNC_THREADS) {      // one matrix column per thread, all 32 symbols
                const int part = col / 36, c = col - part * 36;
                const int j = 1 + (c >> 1);
                const int b = part < 10 ? LB0 + PW * part + j : (UB1 - 10 * PW) + PW * (part - 10) + j;
                const float2 *zc = &sm.z[compact_of_bin(b) * ZLD];
                const float mult = part < 10 ? mlb : mub;
                for (int n = 0; n < BLK; n++) {
                    const float2 v = zc[n];
                    pm[n * 720 + col] = soft_demap((c & 1) ? v.y : v.x, mult);
                }
            }
        }
        __syncthreads();
        if (d.emit_soft) {
            __shared__ uint8_t *sh_w;
            if (t == 0) {
                sh_w = log_reserve(p, d, s, REC_SOFT_PM, 4 + PM_BLOCK);
                if (sh_w) *reinterpret_cast<uint32_t *>(sh_w) = (uint32_t)bc;
            }
            __syncthreads();
            if (sh_w) {
                const int8_t *pm = p.pm + ((size_t)s * 16 + bc) * PM_BLOCK;
                for (int o = t; o < PM_BLOCK; o += SYNC_THREADS) sh_w[4 + o] = (uint8_t)pm[o];
            }
            __syncthreads();
        }
        // PIDS (decode.c:463-471): the 80-bit frame of this block is decoded by the next demodulator launch
        // (k_demod, symbol-0 CTA) so that its ~150 serial trellis steps stay off this kernel's critical
        // path; the record slot is reserved here to keep the stream's record order.
        if (t == 0) {
            uint8_t *w = log_reserve(p, d, s, REC_PIDS, 10);
            st.pids_rec = w ? (unsigned)(w - (p.log + (size_t)s * d.log_cap)) : 0xffffffffu;
            st.pids_bc = bc;
            st.pids_pending = 1;
            // P1 bookkeeping (decode.c:383-390)
            if (bc == 0) st.started_pm = 1;
            if (st.started_pm && bc == 15) st.p1_ready = 1;
            st.bc = (bc + 1) % 16;
        }
    }
    __syncthreads();
    if (t == 0) {                                // window overlap carry (acquire.c:259-262)
        int keep = NSYM + (NSYM / 2 - st.blk_samperr) + st.keep_extra;
        st.keep_extra = 0;
        st.start += NACQ - keep;
        st.blocks_done++;
    }
}

__global__ void __launch_bounds__(SYNC_THREADS) k_sync(DevPtrs p, EngineDims d)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SyncSmem &sm = *reinterpret_cast<SyncSmem *>(smem_raw);
    const int s = blockIdx.x, t = threadIdx.x;
    if (p.st[s].active) sync_block(p, d, sm, s, t);
}

// ===========================================================================
// P1 decode: for every stream whose interleaver matrix is complete —
//   k_p1_gather : interleaver I + depuncture 1,1,1,1,1,0  (decode.c:296-322)
//   k_vitc_fwd / k_vitc_ends / k_vitc_emit : K=7 tail-biting Viterbi (viterbi_chunk.cuh)
//   k_p1_fin    : channel BER (decode.c:234-265), descramble (:279-294), packing,
//                 and the L2 header predicate that feeds back into the sync state
//                 (frame.c:645-714,527-541,158-179; rs_decode.c)
// ===========================================================================
constexpr int P1_THREADS = 256;
constexpr int P1_NCH = (P1_STEPS + CH_LEN - 1) / CH_LEN;             // 143

__global__ void __launch_bounds__(256) k_p1_gather(DevPtrs p, EngineDims d)
{
    const int s = blockIdx.y;
    StreamState &st = p.st[s];
    if (!st.p1_ready) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // reserve the BER and FRAME records now so that they keep their place in the stream's record order
        uint8_t *w = log_reserve(p, d, s, REC_BER, 4);
        uint8_t *fw = log_reserve(p, d, s, REC_FRAME, 8 + P1_LEN / 8);
        st.p1_rec = (w && fw) ? (unsigned)(w - (p.log + (size_t)s * d.log_cap)) : 0xffffffffu;
        if (fw) {
            reinterpret_cast<uint32_t *>(fw)[0] = 0;            // P1 logical channel
            reinterpret_cast<uint32_t *>(fw)[1] = P1_LEN;
        }
        st.p1_errs = 0;
        st.p1_done = 0;
    }
    const int8_t *pm = p.pm + (size_t)s * 16 * PM_BLOCK;
    int8_t *vin = p.vit_in + (size_t)s * P1_VIT;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < P1_VIT; o += gridDim.x * blockDim.x) {
        const int q = o / 6, r = o - 6 * q;
        vin[o] = r == 5 ? (int8_t)0 : pm[p.p1_lut[5 * q + r]];
    }
}

constexpr int FIN_BYTES = 1024;                                      // packed PDU bytes per CTA
constexpr int FIN_CTAS = (P1_LEN / 8 + FIN_BYTES - 1) / FIN_BYTES;   // 18

__device__ __forceinline__ unsigned p1_bit(const uint32_t *bw, int i)
{
    return (bw[i >> 5] >> (i & 31)) & 1u;
}

__global__ void __launch_bounds__(P1_THREADS) k_p1_fin(DevPtrs p, EngineDims d)
{
    const int s = blockIdx.y, t = threadIdx.x;
    StreamState &st = p.st[s];
    if (!st.p1_ready) return;
    __shared__ int red[P1_THREADS];
    __shared__ int sh_last;
    __shared__ uint8_t hdr[96];
    __shared__ uint8_t blk[255];
    const int8_t *vin = p.vit_in + (size_t)s * P1_VIT;
    const uint32_t *bw = p.p1_bits + (size_t)s * (P1_LEN / 32);
    uint8_t *rec = st.p1_rec == 0xffffffffu ? nullptr : p.log + (size_t)s * d.log_cap + st.p1_rec;
    uint8_t *frame = rec ? rec + 4 + 8 + 8 : nullptr;                // BER payload (4) | FRAME header (8) | lc,nbits (8) | bytes
    const int byte0 = blockIdx.x * FIN_BYTES, byte1 = min(P1_LEN / 8, byte0 + FIN_BYTES);
    int errs = 0;
    for (int bi = byte0 + t; bi < byte1; bi += P1_THREADS) {
        // 14 decoded bits around this byte: bits 8*bi-6 .. 8*bi+7 (tail-biting wrap at the frame start)
        unsigned win = 0;
#pragma unroll
        for (int k = 0; k < 14; k++) {
            int idx = 8 * bi - 6 + k;
            if (idx < 0) idx += P1_LEN;
            win |= p1_bit(bw, idx) << k;
        }
        unsigned packed = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int i = 8 * bi + k;
            const unsigned reg = (win >> k) & 0x7f;              // bits i-6 .. i, newest at bit 6 (decode.c:243-249)
            const int8_t *c = vin + 3 * i;
            const int j = 3 * i;
            if ((j % 6) != 5 && ((c[0] > 0) != (int)(__popc(reg & 0133u) & 1))) errs++;
            if (((j + 1) % 6) != 5 && ((c[1] > 0) != (int)(__popc(reg & 0171u) & 1))) errs++;
            if (((j + 2) % 6) != 5 && ((c[2] > 0) != (int)(__popc(reg & 0165u) & 1))) errs++;
            packed |= (((win >> (k + 6)) & 1u) ^ p.pn[i]) << (7 - k);   // descramble (decode.c:279-294), MSB first
        }
        if (frame) frame[bi] = (uint8_t)packed;
    }
    red[t] = errs;
    __syncthreads();
    for (int o = P1_THREADS / 2; o; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
    }
    if (t == 0) {
        atomicAdd(&st.p1_errs, red[0]);
        __threadfence();
        sh_last = atomicAdd(&st.p1_done, 1) == FIN_CTAS - 1;
    }
    __syncthreads();
    if (!sh_last) return;
    // last CTA of this stream: BER record, and the L2 feedback predicate
    // (frame.c:645-714 PCI, :146-156 has_audio, :527-541 header RS check)
    __threadfence();
    if (t < 96) {
        // PDU byte n of frame_push() is the bit-reversed packed byte n (the reference swaps the bit order per byte)
        unsigned v = frame ? frame[t] : 0;
        hdr[t] = (uint8_t)(__brev(v) >> 24);
    }
    __syncthreads();
    if (t == 0) {
        if (rec) *reinterpret_cast<float *>(rec) = (float)atomicAdd(&st.p1_errs, 0) / (float)P1_ENC;
        unsigned pci = 0;
        for (int h = 0; h < 24; h++) {
            const unsigned i = 116176u + 1248u * h;
            const unsigned phys = (i & ~7u) + 7 - (i & 7);
            const unsigned bit = (p1_bit(bw, (int)phys) ^ p.pn[phys]) & 1u;
            pci |= bit << (23 - h);
        }
        const bool has_audio = (pci & 0xFFFFFC) != (0x3634CE & 0xFFFFFC);
        if (has_audio && !fix_header_96(hdr, blk)) set_state(p, d, s, ST_NONE);
        st.p1_ready = 0;
        st.p1_slow = 0;
        st.frames_done++;
    }
}

__host__ __device__ inline VitcArgs p1_vitc_args(const DevPtrs &dp)
{
    VitcArgs a;
    a.vin = dp.vit_in;
    a.dec = dp.vit_dec;
    a.vspec = dp.vspec;
    a.vend = dp.vend;
    a.hstate = dp.hstate;
    a.tbend = dp.tbend;
    a.bitsw = dp.p1_bits;
    a.ready = &dp.st[0].p1_ready;
    a.slow = &dp.st[0].p1_slow;
    a.ready_stride = (int)(sizeof(StreamState) / sizeof(int));
    a.len = P1_LEN;
    a.nch = P1_NCH;
    a.dec_stride = (size_t)P1_NCH * CH_LEN;
    return a;
}

constexpr size_t VITC_EMIT_SMEM = (size_t)VITC_EMIT_WARPS * VITC_EMIT_STEPS * sizeof(uint2);

// input_reset for a range of streams (reference src/input.c:126-138)
__global__ void k_reset(DevPtrs p, EngineDims d, int only)
{
    const int s = blockIdx.x, t = threadIdx.x;
    if (only >= 0 && s != only) return;
    for (int i = t; i < NFFT; i += blockDim.x) {
        p.cfreq[(size_t)s * NFFT + i] = 0.f;
        p.cphase[(size_t)s * NFFT + i] = 0.f;
    }
    if (t == 0) {
        StreamState &st = p.st[s];
        const long long avail = st.in_avail;
        StreamState z;
        memset(&z, 0, sizeof(z));
        z.phase = make_float2(1.0f, 0.0f);
        z.psmi = 1;
        z.state = ST_NONE;
        z.force_state = -1;
        z.in_avail = only == -2 ? avail : 0;     // -2: keep the attached input (rewind)
        st = z;
    }
}

// ---------------------------------------------------------------------------
// stage kernels for the parity tests
// ---------------------------------------------------------------------------
__global__ void k_viterbi_test(const int8_t *in, uint8_t *out, uint2 *dec, int len)
{
    const int f = blockIdx.x, t = threadIdx.x;
    const int8_t *vin = in + (size_t)f * 3 * len;
    uint2 *dd = dec + (size_t)f * (len + 64);
    int state = viterbi_forward(vin, len, dd, t);
    __syncwarp();
    if (t == 0) {
        uint8_t *o = out + (size_t)f * len;
        for (int q = len + 63; q >= 0; q--) {
            if (q >= 32 && q < 32 + len) o[q - 32] = (uint8_t)((state >> 5) & 1);
            state = vit_prev(state, dd[q]);
        }
    }
}

__global__ void k_rs_test(uint8_t *blocks, int *rc, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rc[i] = rs_decode_255_247(blocks + (size_t)i * 255);
}

}  // namespace nb

// ===========================================================================
// host side
// ===========================================================================
using namespace nb;

static size_t vitc_emit_smem() { return VITC_EMIT_SMEM; }

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) {                                                                   \
            fprintf(stderr, "nrsc5_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return NRSC5B_ECUDA;                                                                   \
        }                                                                                          \
    } while (0)

struct nrsc5b_engine {
    nrsc5b_config_t cfg;
    EngineDims dims;
    DevPtrs dp;
    cudaStream_t stream;
    cudaStream_t copy_stream;          // host->device input copies overlap with compute on `stream`
    uint8_t *iq_owned;                 // engine-owned cu8 buffer (null when attached)
    std::vector<long long> pushed;     // complex cu8 samples pushed per stream
    std::vector<unsigned> drained;     // log bytes already handed out per stream
    uint8_t *pinned;                   // staging for pushes
    size_t pinned_cap;
    cudaEvent_t pinned_free;
    long long *avail_ring;             // pinned, 4096 entries
    unsigned avail_pos;
    cudaEvent_t reset_done;            // copy stream waits for resets issued on the compute stream
    StreamState *h_state;              // pinned mirror for read-back
    nrsc5b_stats_t stats;
    unsigned long long last_progress;
    size_t sync_smem;
    std::vector<void *> allocs;
    int profiling;
    cudaEvent_t pev[5];
    double kernel_ms[4];
    unsigned long long kernel_n[4];
};

static const float k_bp_coeff[32] = {
    -0.000685643230099231f, 0.005636964458972216f, 0.009015781804919243f, -0.015486305579543114f,
    -0.035108357667922974f, 0.017446253448724747f, 0.08155813068151474f, 0.007995186373591423f,
    -0.13311293721199036f, -0.0727422907948494f, 0.15914097428321838f, 0.16498781740665436f,
    -0.1324498951435089f, -0.2484012246131897f, 0.051773931831121445f, 0.2821577787399292f,
    0.051773931831121445f, -0.2484012246131897f, -0.1324498951435089f, 0.16498781740665436f,
    0.15914097428321838f, -0.0727422907948494f, -0.13311293721199036f, 0.007995186373591423f,
    0.08155813068151474f, 0.017446253448724747f, -0.035108357667922974f, -0.015486305579543114f,
    0.009015781804919243f, 0.005636964458972216f, -0.000685643230099231f, 0.0f
};

static int upload_tables(int device)
{
    static int done_for = -1;
    if (done_for == device) return 0;
    uint8_t ex[256], lg[256];
    unsigned v = 1;
    lg[0] = 255;
    ex[255] = 0;
    for (int i = 0; i < 255; i++) {
        ex[i] = (uint8_t)v;
        lg[v] = (uint8_t)i;
        v <<= 1;
        if (v & 0x100) v ^= 0x11d;
    }
    CK(cudaMemcpyToSymbol(c_gf_exp, ex, 256));
    CK(cudaMemcpyToSymbol(c_gf_log, lg, 256));
    static const int compat[64] = {
        0, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5, 6, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5,
        6, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5, 6, 1, 2, 3, 1, 5, 6, 5, 6, 1, 2, 11, 1, 5, 6, 5
    };
    CK(cudaMemcpyToSymbol(c_compat_mode, compat, sizeof(compat)));
    short taps[32];
    for (int i = 0; i < 32; i++) taps[i] = (short)(k_bp_coeff[31 - i] * 32767.0f);
    CK(cudaMemcpyToSymbol(c_bp_tap, taps, sizeof(taps)));
    done_for = device;
    return 0;
}

template <typename T>
static int dev_alloc(nrsc5b_engine *e, T **ptr, size_t count, bool zero = true)
{
    void *q = nullptr;
    if (cudaMalloc(&q, count * sizeof(T)) != cudaSuccess) return NRSC5B_ENOMEM;
    if (zero && cudaMemset(q, 0, count * sizeof(T)) != cudaSuccess) return NRSC5B_ECUDA;
    e->allocs.push_back(q);
    *ptr = reinterpret_cast<T *>(q);
    return 0;
}

static std::vector<float2> make_twiddles()
{
    std::vector<float2> tw(NFFT);
    for (int k = 0; k < NFFT; k++) {
        double a = -2.0 * M_PI * (double)k / (double)NFFT;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    return tw;
}

static void init_state_host(StreamState &st)
{
    memset(&st, 0, sizeof(st));
    st.phase = make_float2(1.0f, 0.0f);
    st.psmi = 1;
    st.state = ST_NONE;
    st.force_state = -1;
}

extern "C" const char *nrsc5b_version(void) { return "nrsc5_b200 0.1 (sm_100a)"; }

extern "C" int nrsc5b_create(nrsc5b_engine_t **out, const nrsc5b_config_t *cfg)
{
    if (!out || !cfg || cfg->nstreams <= 0 || cfg->mode != NRSC5B_MODE_FM) return NRSC5B_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cfg->device >= ndev) {
        fprintf(stderr, "nrsc5_b200: no usable CUDA device (the engine has no CPU path)\n");
        return NRSC5B_ENODEV;
    }
    if (cudaSetDevice(cfg->device) != cudaSuccess) return NRSC5B_ENODEV;
    nrsc5b_engine *e = new (std::nothrow) nrsc5b_engine();
    if (!e) return NRSC5B_ENOMEM;
    e->cfg = *cfg;
    e->stream = 0;
    e->copy_stream = nullptr;
    e->iq_owned = nullptr;
    e->stats = nrsc5b_stats_t{};
    e->last_progress = 0;
    e->profiling = 0;
    for (int i = 0; i < 5; i++) e->pev[i] = nullptr;
    for (int i = 0; i < 4; i++) { e->kernel_ms[i] = 0; e->kernel_n[i] = 0; }
    e->pinned = nullptr; e->h_state = nullptr; e->pinned_free = nullptr; e->avail_ring = nullptr; e->avail_pos = 0; e->reset_done = nullptr;
    const int S = cfg->nstreams;
    e->dims.nstreams = S;
    e->dims.in_stride = (cfg->input_capacity + 63) & ~(size_t)63;
    e->dims.log_cap = cfg->log_capacity ? ((cfg->log_capacity + 15) & ~(size_t)15) : (1u << 20);
    e->dims.emit_soft = cfg->emit_soft;
    e->pushed.assign(S, 0);
    e->drained.assign(S, 0);
    int rc = upload_tables(cfg->device);
    if (rc) { delete e; return rc; }

    DevPtrs &dp = e->dp;
    memset(&dp, 0, sizeof(dp));
#define DA(field, type, count)                                              \
    do {                                                                    \
        type *tmp_ = nullptr;                                               \
        rc = dev_alloc(e, &tmp_, (count));                                  \
        if (rc) { nrsc5b_destroy(e); return rc; }                           \
        dp.field = tmp_;                                                    \
    } while (0)
    if (e->dims.in_stride) {
        uint8_t *tmp = nullptr;
        rc = dev_alloc(e, &tmp, (size_t)S * e->dims.in_stride + 64, false);
        if (rc) { nrsc5b_destroy(e); return rc; }
        cudaMemset(tmp, 0x7f, (size_t)S * e->dims.in_stride + 64);
        e->iq_owned = tmp;
        dp.iq = tmp;
    }
    DA(st, StreamState, S);
    DA(cfreq, float, (size_t)S * NFFT);
    DA(cphase, float, (size_t)S * NFFT);
    DA(nco, float2, (size_t)S * NSYM);
    DA(bins, float2, (size_t)S * BLK * NBINS);
    DA(pm, int8_t, (size_t)S * 16 * PM_BLOCK);
    DA(ydec, short2, (size_t)S * NACQ);
    DA(tbuf, float2, (size_t)S * NACQ);
    DA(vit_in, int8_t, (size_t)S * P1_VIT);
    DA(vit_dec, uint2, (size_t)S * P1_NCH * CH_LEN);
    DA(vspec, uint2, (size_t)S * P1_NCH * 16);
    DA(vend, uint2, (size_t)S * P1_NCH * 16);
    DA(tbend, int, (size_t)S * P1_NCH);
    DA(hstate, int, (size_t)S * P1_NCH);
    DA(p1_bits, uint32_t, (size_t)S * (P1_LEN / 32));
    DA(log, uint8_t, (size_t)S * e->dims.log_cap);
    {
        // tables
        std::vector<float> shape(NSYM);
        for (int i = 0; i < NSYM; i++) {
            if (i < NCP) shape[i] = sinf(M_PI / 2 * i / NCP);
            else if (i < NFFT) shape[i] = 1;
            else shape[i] = cosf(M_PI / 2 * (i - NFFT) / NCP);
        }
        float *dshape = nullptr;
        rc = dev_alloc(e, &dshape, NSYM);
        if (rc) { nrsc5b_destroy(e); return rc; }
        cudaMemcpy(dshape, shape.data(), NSYM * sizeof(float), cudaMemcpyHostToDevice);
        dp.shape = dshape;
        std::vector<float2> tw = make_twiddles();
        float2 *dtw = nullptr;
        rc = dev_alloc(e, &dtw, NFFT);
        if (rc) { nrsc5b_destroy(e); return rc; }
        cudaMemcpy(dtw, tw.data(), NFFT * sizeof(float2), cudaMemcpyHostToDevice);
        dp.twid = dtw;
        static const int PMV[20] = { 10, 2, 18, 6, 14, 8, 16, 0, 12, 4, 11, 3, 19, 7, 15, 9, 17, 1, 13, 5 };
        std::vector<uint32_t> lut(P1_ENC);
        for (unsigned i = 0; i < (unsigned)P1_ENC; i++) {
            unsigned part = (unsigned)PMV[i % 20];
            unsigned block = (i / 20 + part * 7) % 16;
            unsigned k = i / 320;
            unsigned row = (k * 11) % 32, col = (k * 11 + k / 288) % 36;
            lut[i] = (block * 32 + row) * 720 + part * 36 + col;
        }
        uint32_t *dlut = nullptr;
        rc = dev_alloc(e, &dlut, P1_ENC);
        if (rc) { nrsc5b_destroy(e); return rc; }
        cudaMemcpy(dlut, lut.data(), P1_ENC * sizeof(uint32_t), cudaMemcpyHostToDevice);
        dp.p1_lut = dlut;
        std::vector<uint8_t> pn(P1_LEN);
        unsigned reg = 0x3ff;
        for (int i = 0; i < P1_LEN; i++) {
            unsigned b = ((reg >> 9) ^ reg) & 1;
            reg |= b << 11;
            reg >>= 1;
            pn[i] = (uint8_t)b;
        }
        uint8_t *dpn = nullptr;
        rc = dev_alloc(e, &dpn, P1_LEN);
        if (rc) { nrsc5b_destroy(e); return rc; }
        cudaMemcpy(dpn, pn.data(), P1_LEN, cudaMemcpyHostToDevice);
        dp.pn = dpn;
    }
#undef DA
    e->pinned_cap = 8u << 20;
    if (cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { nrsc5b_destroy(e); return NRSC5B_ECUDA; }
    if (cudaMallocHost((void **)&e->pinned, e->pinned_cap) != cudaSuccess ||
        cudaMallocHost((void **)&e->h_state, sizeof(StreamState) * S) != cudaSuccess ||
        cudaMallocHost((void **)&e->avail_ring, sizeof(long long) * 4096) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->reset_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->pinned_free, cudaEventDisableTiming) != cudaSuccess) {
        nrsc5b_destroy(e);
        return NRSC5B_ENOMEM;
    }
    e->sync_smem = sizeof(SyncSmem);
    if (cudaFuncSetAttribute(k_sync, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->sync_smem) != cudaSuccess ||
        cudaFuncSetAttribute(k_vitc_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vitc_emit_smem()) != cudaSuccess) {
        nrsc5b_destroy(e);
        return NRSC5B_ECUDA;
    }
    *out = e;
    rc = nrsc5b_reset(e, -1);
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) rc = NRSC5B_ECUDA;
    if (rc) { nrsc5b_destroy(e); *out = nullptr; return rc; }
    return NRSC5B_OK;
}

extern "C" void nrsc5b_destroy(nrsc5b_engine_t *e)
{
    if (!e) return;
    cudaDeviceSynchronize();
    for (void *q : e->allocs) cudaFree(q);
    if (e->pinned) cudaFreeHost(e->pinned);
    if (e->h_state) cudaFreeHost(e->h_state);
    if (e->avail_ring) cudaFreeHost(e->avail_ring);
    if (e->reset_done) cudaEventDestroy(e->reset_done);
    if (e->pinned_free) cudaEventDestroy(e->pinned_free);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    for (int i = 0; i < 5; i++) if (e->pev[i]) cudaEventDestroy(e->pev[i]);
    delete e;
}

extern "C" int nrsc5b_set_cuda_stream(nrsc5b_engine_t *e, void *cuda_stream)
{
    if (!e) return NRSC5B_EINVAL;
    e->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return NRSC5B_OK;
}

extern "C" int nrsc5b_reset(nrsc5b_engine_t *e, int stream)
{
    if (!e || stream >= e->dims.nstreams) return NRSC5B_EINVAL;
    const int S = e->dims.nstreams;
    CK(cudaStreamSynchronize(e->copy_stream));       // no input copy of the old contents may still be in flight
    k_reset<<<S, 256, 0, e->stream>>>(e->dp, e->dims, stream < 0 ? -1 : stream);
    CK(cudaEventRecord(e->reset_done, e->stream));
    CK(cudaStreamWaitEvent(e->copy_stream, e->reset_done, 0));
    e->stats.kernel_launches += 1;
    for (int s = 0; s < S; s++) {
        if (stream >= 0 && s != stream) continue;
        e->pushed[s] = 0;
        e->drained[s] = 0;
    }
    CK(cudaGetLastError());
    return NRSC5B_OK;
}

/* Restart every stream from sample 0 of the input it already holds (benchmark loops). */
extern "C" int nrsc5b_rewind(nrsc5b_engine_t *e)
{
    if (!e) return NRSC5B_EINVAL;
    k_reset<<<e->dims.nstreams, 256, 0, e->stream>>>(e->dp, e->dims, -2);
    e->stats.kernel_launches += 1;
    for (int s = 0; s < e->dims.nstreams; s++) e->drained[s] = 0;
    CK(cudaGetLastError());
    return NRSC5B_OK;
}

static int publish_avail(nrsc5b_engine *e, int s, cudaStream_t on)
{
    // staged through a small pinned ring so that the asynchronous copy has a stable source
    if (e->avail_pos && (e->avail_pos & 4095) == 0) CK(cudaStreamSynchronize(on));   // ring wrap: let pending copies drain
    long long *slot = e->avail_ring + (e->avail_pos++ & 4095);
    *slot = e->pushed[s];
    CK(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(e->dp.st + s) + offsetof(StreamState, in_avail), slot, sizeof(*slot),
                       cudaMemcpyHostToDevice, on));
    return 0;
}

extern "C" int nrsc5b_push_cu8(nrsc5b_engine_t *e, int stream, const uint8_t *buf, size_t nbytes)
{
    if (!e || stream < 0 || stream >= e->dims.nstreams || (nbytes & 3) || !e->iq_owned) return NRSC5B_EINVAL;
    size_t off = (size_t)e->pushed[stream] * 2;
    if (off + nbytes > e->dims.in_stride) return NRSC5B_EFULL;
    uint8_t *dst = e->iq_owned + (size_t)stream * e->dims.in_stride + off;
    // page-locked caller memory is DMA'd directly; pageable memory goes through the engine's pinned staging buffer
    cudaPointerAttributes attr;
    bool pinned_src = cudaPointerGetAttributes(&attr, buf) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned_src) {
        CK(cudaMemcpyAsync(dst, buf, nbytes, cudaMemcpyHostToDevice, e->copy_stream));
    } else {
        size_t done = 0;
        while (done < nbytes) {
            size_t n = nbytes - done < e->pinned_cap ? nbytes - done : e->pinned_cap;
            CK(cudaEventSynchronize(e->pinned_free));
            memcpy(e->pinned, buf + done, n);
            CK(cudaMemcpyAsync(dst + done, e->pinned, n, cudaMemcpyHostToDevice, e->copy_stream));
            CK(cudaEventRecord(e->pinned_free, e->copy_stream));
            done += n;
        }
    }
    e->pushed[stream] += (long long)(nbytes / 2);
    // published on the copy stream, i.e. after the samples themselves have landed
    return publish_avail(e, stream, e->copy_stream);
}

extern "C" int nrsc5b_push_cu8_device(nrsc5b_engine_t *e, int stream, const void *dev_buf, size_t nbytes)
{
    if (!e || stream < 0 || stream >= e->dims.nstreams || (nbytes & 3) || !e->iq_owned) return NRSC5B_EINVAL;
    size_t off = (size_t)e->pushed[stream] * 2;
    if (off + nbytes > e->dims.in_stride) return NRSC5B_EFULL;
    CK(cudaMemcpyAsync(e->iq_owned + (size_t)stream * e->dims.in_stride + off, dev_buf, nbytes,
                       cudaMemcpyDeviceToDevice, e->stream));
    e->pushed[stream] += (long long)(nbytes / 2);
    return publish_avail(e, stream, e->stream);
}

extern "C" int nrsc5b_attach_device_input(nrsc5b_engine_t *e, const void *dev_buf, size_t stride, size_t nbytes)
{
    if (!e || !dev_buf || (nbytes & 3) || nbytes > stride) return NRSC5B_EINVAL;
    e->dp.iq = reinterpret_cast<const uint8_t *>(dev_buf);
    e->dims.in_stride = stride;
    for (int s = 0; s < e->dims.nstreams; s++) {
        e->pushed[s] = (long long)(nbytes / 2);
        int rc = publish_avail(e, s, e->stream);
        if (rc) return rc;
    }
    return NRSC5B_OK;
}

extern "C" int nrsc5b_attach_device_log(nrsc5b_engine_t *e, void *dev_buf, size_t stride)
{
    if (!e || !dev_buf || stride < 4096 || (stride & 15)) return NRSC5B_EINVAL;
    CK(cudaStreamSynchronize(e->stream));
    e->dp.log = reinterpret_cast<uint8_t *>(dev_buf);
    e->dims.log_cap = stride;
    return NRSC5B_OK;
}

static void launch_vitc(const VitcArgs &a, int nframes, cudaStream_t stream)
{
    dim3 gf((a.nch + 2 * VITC_FWD_WARPS - 1) / (2 * VITC_FWD_WARPS), nframes);
    k_vitc_fwd<<<gf, VITC_FWD_WARPS * 32, 0, stream>>>(a);
    k_vitc_ends<<<nframes, 32, 0, stream>>>(a);
    dim3 ge((a.nch + VITC_EMIT_WARPS - 1) / VITC_EMIT_WARPS, nframes);
    k_vitc_emit<<<ge, VITC_EMIT_WARPS * 32, vitc_emit_smem(), stream>>>(a);
}

static void launch_p1(nrsc5b_engine *e)
{
    const int S = e->dims.nstreams;
    k_p1_gather<<<dim3(16, S), 256, 0, e->stream>>>(e->dp, e->dims);
    launch_vitc(p1_vitc_args(e->dp), S, e->stream);
    k_p1_fin<<<dim3(FIN_CTAS, S), P1_THREADS, 0, e->stream>>>(e->dp, e->dims);
    e->stats.kernel_launches += 5;
}

static int launch_step(nrsc5b_engine *e, bool with_p1)
{
    const int S = e->dims.nstreams;
    const bool prof = e->profiling != 0;
    if (prof) cudaEventRecord(e->pev[0], e->stream);
    k_prep<<<S, PREP_THREADS, 0, e->stream>>>(e->dp, e->dims);
    if (prof) cudaEventRecord(e->pev[1], e->stream);
    launch_demod(e->dp, e->dims, e->stream);
    if (prof) cudaEventRecord(e->pev[2], e->stream);
    k_sync<<<S, SYNC_THREADS, e->sync_smem, e->stream>>>(e->dp, e->dims);
    if (prof) cudaEventRecord(e->pev[3], e->stream);
    if (with_p1) launch_p1(e);
    if (prof) {
        cudaEventRecord(e->pev[4], e->stream);
        cudaEventSynchronize(e->pev[4]);
        for (int i = 0; i < 4; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e->pev[i], e->pev[i + 1]);
            e->kernel_ms[i] += ms;
            e->kernel_n[i] += 1;
        }
    }
    e->stats.kernel_launches += 3;
    return 0;
}

/* Per-kernel device time (CUDA events around every launch; slows the run down, use a separate pass).
 * Order: prep, demod, sync, p1. */
extern "C" int nrsc5b_set_profiling(nrsc5b_engine_t *e, int on)
{
    if (!e) return NRSC5B_EINVAL;
    if (on && !e->pev[0])
        for (int i = 0; i < 5; i++) CK(cudaEventCreate(&e->pev[i]));
    e->profiling = on;
    for (int i = 0; i < 4; i++) { e->kernel_ms[i] = 0; e->kernel_n[i] = 0; }
    return NRSC5B_OK;
}

extern "C" int nrsc5b_get_kernel_times(nrsc5b_engine_t *e, double *ms4, unsigned long long *n4)
{
    if (!e || !ms4 || !n4) return NRSC5B_EINVAL;
    for (int i = 0; i < 4; i++) { ms4[i] = e->kernel_ms[i]; n4[i] = e->kernel_n[i]; }
    return NRSC5B_OK;
}

static int process_impl(nrsc5b_engine_t *e, bool wait_for_copies)
{
    if (!e) return NRSC5B_EINVAL;
    // Each block advances a stream's window by 69120 +- a few decimated samples, so the number of
    // blocks a stream can still take follows from its buffered samples and its window start.  Launch
    // that many steps, look at the device-side progress counter, and stop after a batch (always at
    // least one trailing step, which also flushes the deferred PIDS decode) made no progress.
    const int S = e->dims.nstreams;
    for (;;) {
        CK(cudaMemcpyAsync(e->h_state, e->dp.st, sizeof(StreamState) * S, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        long long most = 0;
        for (int s = 0; s < S; s++) {
            const long long avail_dec = e->pushed[s] / 2, start = e->h_state[s].start;
            if (avail_dec >= start + NACQ) {
                long long n = (avail_dec - start - NACQ) / (NSYM * BLK + 80) + 1;
                if (n > most) most = n;
            }
        }
        // The P1 decode kernels are only launched in the steps where a stream can complete its 16-block
        // interleaver matrix.  With the states just read back this is predictable for up to 15 steps: a
        // stream in FINE sync reaches block count 15 at a known step (a superset if it runs out of
        // samples or loses sync first), and a stream that is not in FINE sync now cannot finish a frame
        // in fewer than 16 blocks.  The last step of every batch launches them unconditionally, which
        // also covers streams that stalled with a frame pending.
        const int batch = (int)(most < 1 ? 1 : (most > 15 ? 15 : most));
        unsigned p1_steps = 0;
        for (int s = 0; s < S; s++) {
            const StreamState &hs = e->h_state[s];
            if (hs.p1_ready) p1_steps |= 1u;
            if (hs.state != ST_FINE) continue;
            for (int i = 0; i < batch; i++) {
                const int bc = (hs.bc + i) & 15;
                if (bc == 15 && (hs.started_pm || i >= ((16 - hs.bc) & 15))) p1_steps |= 1u << i;
            }
        }
        p1_steps |= 1u << (batch - 1);
        for (int i = 0; i < batch; i++) launch_step(e, e->profiling || ((p1_steps >> i) & 1u));
        unsigned long long prog = 0;
        CK(cudaMemcpyFromSymbolAsync(&prog, g_progress, sizeof(prog), 0, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaGetLastError());
        const unsigned long long delta = prog - e->last_progress;
        e->last_progress = prog;
        if (delta == 0) {
            // samples pushed asynchronously may still have been in flight: wait for them once, then retry
            if (!wait_for_copies || cudaStreamQuery(e->copy_stream) == cudaSuccess) break;
            CK(cudaStreamSynchronize(e->copy_stream));
        }
    }
    return NRSC5B_OK;
}

extern "C" int nrsc5b_process(nrsc5b_engine_t *e) { return process_impl(e, true); }
extern "C" int nrsc5b_process_available(nrsc5b_engine_t *e) { return process_impl(e, false); }

extern "C" int nrsc5b_synchronize(nrsc5b_engine_t *e)
{
    if (!e) return NRSC5B_EINVAL;
    CK(cudaStreamSynchronize(e->stream));
    return NRSC5B_OK;
}

extern "C" long nrsc5b_drain(nrsc5b_engine_t *e, int stream, uint8_t *out, size_t cap, size_t *needed)
{
    if (!e || stream < 0 || stream >= e->dims.nstreams) return NRSC5B_EINVAL;
    CK(cudaStreamSynchronize(e->stream));
    StreamState st;
    CK(cudaMemcpy(&st, e->dp.st + stream, sizeof(st), cudaMemcpyDeviceToHost));
    size_t avail = st.log_len - e->drained[stream];
    if (needed) *needed = avail;
    if (!out || cap < avail) return avail == 0 ? 0 : NRSC5B_EFULL;
    if (avail) {
        CK(cudaMemcpy(out, e->dp.log + (size_t)stream * e->dims.log_cap + e->drained[stream], avail,
                      cudaMemcpyDeviceToHost));
    }
    // the log is rewound once fully drained
    unsigned zero = 0;
    CK(cudaMemcpy(reinterpret_cast<uint8_t *>(e->dp.st + stream) + offsetof(StreamState, log_len), &zero, sizeof(zero),
                  cudaMemcpyHostToDevice));
    e->drained[stream] = 0;
    if (st.log_overflow) fprintf(stderr, "nrsc5_b200: stream %d output log overflowed (raise log_capacity)\n", stream);
    return (long)avail;
}

extern "C" int nrsc5b_set_sync_state(nrsc5b_engine_t *e, int stream, int state)
{
    if (!e || stream < 0 || stream >= e->dims.nstreams || state < 0 || state > 2) return NRSC5B_EINVAL;
    CK(cudaMemcpyAsync(reinterpret_cast<uint8_t *>(e->dp.st + stream) + offsetof(StreamState, force_state), &state,
                       sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return NRSC5B_OK;
}

extern "C" int nrsc5b_get_stats(nrsc5b_engine_t *e, nrsc5b_stats_t *out)
{
    if (!e || !out) return NRSC5B_EINVAL;
    CK(cudaStreamSynchronize(e->stream));
    const int S = e->dims.nstreams;
    CK(cudaMemcpy(e->h_state, e->dp.st, sizeof(StreamState) * S, cudaMemcpyDeviceToHost));
    unsigned long long frames = 0, samples = 0, blocks = 0;
    for (int s = 0; s < S; s++) {
        frames += e->h_state[s].frames_done;
        blocks += e->h_state[s].blocks_done;
        samples += (unsigned long long)(2 * e->h_state[s].start);
    }
    e->stats.p1_frames = frames;
    e->stats.blocks = blocks;
    e->stats.samples = samples;
    *out = e->stats;
    return NRSC5B_OK;
}

// ---- stage entry points ---------------------------------------------------
static int use_device(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) {
        fprintf(stderr, "nrsc5_b200: no usable CUDA device (the engine has no CPU path)\n");
        return NRSC5B_ENODEV;
    }
    if (cudaSetDevice(device) != cudaSuccess) return NRSC5B_ENODEV;
    return upload_tables(device);
}

extern "C" int nrsc5b_halfband_fm(int device, const uint8_t *cu8, size_t npairs, int16_t *out)
{
    int rc = use_device(device);
    if (rc) return rc;
    uint8_t *din = nullptr;
    short2 *dout = nullptr;
    CK(cudaMalloc(&din, 4 * npairs + 64));
    CK(cudaMalloc(&dout, npairs * sizeof(short2)));
    CK(cudaMemcpy(din, cu8, 4 * npairs, cudaMemcpyHostToDevice));
    launch_halfband_test(din, (long long)npairs, dout, 0);
    CK(cudaMemcpy(out, dout, npairs * sizeof(short2), cudaMemcpyDeviceToHost));
    cudaFree(din);
    cudaFree(dout);
    return NRSC5B_OK;
}

extern "C" int nrsc5b_viterbi_k7(int device, const int8_t *in, uint8_t *out, int len, int nframes)
{
    int rc = use_device(device);
    if (rc) return rc;
    if (len < 32 || nframes <= 0) return NRSC5B_EINVAL;
    int8_t *din = nullptr;
    uint8_t *dout = nullptr;
    size_t nin = (size_t)nframes * 3 * len;
    CK(cudaMalloc(&din, nin));
    CK(cudaMalloc(&dout, (size_t)nframes * len));
    CK(cudaMemcpy(din, in, nin, cudaMemcpyHostToDevice));
    if (len >= 2048 && (len % 32) == 0) {
        // chunk-parallel exact decoder (the engine's P1 path)
        VitcArgs a;
        a.len = len;
        a.nch = (len + 64 + CH_LEN - 1) / CH_LEN;
        a.dec_stride = (size_t)a.nch * CH_LEN;
        a.ready_stride = 1;
        int *dflags = nullptr;
        CK(cudaMalloc(&a.dec, (size_t)nframes * a.dec_stride * sizeof(uint2)));
        CK(cudaMalloc(&a.vspec, (size_t)nframes * a.nch * 16 * sizeof(uint2)));
        CK(cudaMalloc(&a.vend, (size_t)nframes * a.nch * 16 * sizeof(uint2)));
        CK(cudaMalloc(&a.tbend, (size_t)nframes * a.nch * sizeof(int)));
        CK(cudaMalloc(&a.hstate, (size_t)nframes * a.nch * sizeof(int)));
        uint32_t *dbw = nullptr;
        CK(cudaMalloc(&dbw, (size_t)nframes * (len / 32) * sizeof(uint32_t)));
        CK(cudaMalloc(&dflags, (size_t)nframes * 2 * sizeof(int)));
        std::vector<int> fl(2 * (size_t)nframes, 0);
        for (int i = 0; i < nframes; i++) fl[i] = 1;
        CK(cudaMemcpy(dflags, fl.data(), fl.size() * sizeof(int), cudaMemcpyHostToDevice));
        a.vin = din;
        a.bitsw = dbw;
        a.ready = dflags;
        a.slow = dflags + nframes;
        CK(cudaFuncSetAttribute(k_vitc_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vitc_emit_smem()));
        launch_vitc(a, nframes, 0);
        {
            size_t nb = (size_t)nframes * len;
            k_vitc_unpack<<<(unsigned)((nb + 255) / 256), 256>>>(dbw, dout, nb);
        }
        CK(cudaDeviceSynchronize());
        cudaFree(a.dec); cudaFree(a.vspec); cudaFree(a.vend); cudaFree(a.tbend); cudaFree(a.hstate); cudaFree(dflags); cudaFree(dbw);
    } else {
        uint2 *ddec = nullptr;
        CK(cudaMalloc(&ddec, (size_t)nframes * (len + 64) * sizeof(uint2)));
        k_viterbi_test<<<nframes, 32>>>(din, dout, ddec, len);
        CK(cudaDeviceSynchronize());
        cudaFree(ddec);
    }
    CK(cudaMemcpy(out, dout, (size_t)nframes * len, cudaMemcpyDeviceToHost));
    cudaFree(din);
    cudaFree(dout);
    return NRSC5B_OK;
}

extern "C" int nrsc5b_rs_decode(int device, uint8_t *blocks, int *rcs, int n)
{
    int rc = use_device(device);
    if (rc) return rc;
    uint8_t *db = nullptr;
    int *dr = nullptr;
    CK(cudaMalloc(&db, (size_t)n * 255));
    CK(cudaMalloc(&dr, (size_t)n * sizeof(int)));
    CK(cudaMemcpy(db, blocks, (size_t)n * 255, cudaMemcpyHostToDevice));
    k_rs_test<<<(n + 63) / 64, 64>>>(db, dr, n);
    CK(cudaMemcpy(blocks, db, (size_t)n * 255, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rcs, dr, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(db);
    cudaFree(dr);
    return NRSC5B_OK;
}

extern "C" int nrsc5b_fft2048(int device, const float *in, float *out, int nffts)
{
    int rc = use_device(device);
    if (rc) return rc;
    float2 *di = nullptr, *dout = nullptr, *dtw = nullptr;
    size_t n = (size_t)nffts * NFFT;
    CK(cudaMalloc(&di, n * sizeof(float2)));
    CK(cudaMalloc(&dout, n * sizeof(float2)));
    CK(cudaMalloc(&dtw, NFFT * sizeof(float2)));
    std::vector<float2> tw = make_twiddles();
    CK(cudaMemcpy(dtw, tw.data(), NFFT * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(di, in, n * sizeof(float2), cudaMemcpyHostToDevice));
    launch_fft_test(di, dout, dtw, nffts, 0);
    CK(cudaMemcpy(out, dout, n * sizeof(float2), cudaMemcpyDeviceToHost));
    cudaFree(di);
    cudaFree(dout);
    cudaFree(dtw);
    return NRSC5B_OK;
}
