// OFDM demodulator kernel: cu8 -> Q15 halfband decimation -> conjugate/scale ->
// NCO rotation + raised-sine window + cyclic-prefix fold -> 2048-point FFT ->
// the 534 sideband bins, for one OFDM symbol per CTA.
//
// Replaces, for the FINE/COARSE demodulation loop of one block:
//   reference src/input.c:52-94 (decimate_samples), src/firdecim_q15.c:137-165,
//   src/acquire.c:160-161 (cq15_to_cf_conj), :237-257 (rotate, window, fold, FFT,
//   fftshift) and src/sync.c:779-790 (sync_push bin selection).
// The NCO is applied in closed form: sample j of symbol i is multiplied by
// exp(j*theta*j) * window[j] (table `nco`, built once per block by k_prep) and
// the per-symbol factor phase0*exp(j*theta*2160*i) is applied to the 534 output
// bins (the FFT is linear), instead of the reference's per-sample recurrence.
#include "common.cuh"
#include "fft.cuh"
#include "viterbi_pack.cuh"

namespace nb {

constexpr int IN_BYTES = 4 * NSYM + 28 + 16 + 16;   // staged cu8 bytes per symbol (+ alignment slack)

__device__ __forceinline__ float2 sample_at(const uint32_t *sw, int j)
{
    // sw points at the 32-bit word holding input samples (2*base-14, 2*base-13);
    // word q of output j holds samples m = 2q (low half) and m = 2q+1 (high half)
    // of the 15-sample halfband window of y[base + j].
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 8; q++) w[q] = sw[j + q];
    const int tap[4] = { -134, 1078, -4417, 19864 };
    int ar = 0, ai = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t a = w[k], b = w[7 - k];
        int sr = (int)(a & 0xff) + (int)(b & 0xff) - 254;
        int si = (int)((a >> 8) & 0xff) + (int)((b >> 8) & 0xff) - 254;
        // ((64*sr) * tap) >> 15  ==  (sr * tap) >> 9   (exact)
        ar += (sr * tap[k]) >> 9;
        ai += (si * tap[k]) >> 9;
    }
    ar += ((int)((w[3] >> 16) & 0xff) - 127) * 64;
    ai += ((int)(w[3] >> 24) - 127) * 64;
    const float sc = 1.0f / 32767.0f;
    return make_float2((float)ar * sc, (float)ai * -sc);      // conj(x)/32767, acquire.c:160-161
}

// PIDS frame of the block the sync kernel just finished: interleaver II + depuncture
// (reference src/decode.c:324-342), K=7 tail-biting Viterbi (src/conv_dec.c), descramble
// (src/decode.c:279-294).  Run by the symbol-0 CTA of the next demodulator launch.
__device__ void pids_decode(const DevPtrs &p, const EngineDims &d, int s, int t)
{
    __shared__ int8_t vit[PIDS_LEN * 3];
    __shared__ uint2 dec[PIDS_LEN + 64];                     // 9 groups x 16 lanes of decision history
    StreamState &st = p.st[s];
    const int bc = st.pids_bc;
    const int8_t *pmall = p.pm + (size_t)s * 16 * PM_BLOCK;
    const int8_t PMV[20] = { 10, 2, 18, 6, 14, 8, 16, 0, 12, 4, 11, 3, 19, 7, 15, 9, 17, 1, 13, 5 };
    for (int o = t; o < PIDS_LEN * 3; o += FFT_THREADS) {
        int8_t v = 0;
        if (o % 6 != 5) {
            unsigned i = (unsigned)bc * 200 + (unsigned)(o - o / 6);
            unsigned part = (unsigned)PMV[i % 20];
            unsigned block = i / 200;
            unsigned k = (i / 20) % 10 + P1_ENC / 320;
            unsigned row = (k * 11) % 32, col = (k * 11 + k / 288) % 36;
            v = pmall[(block * 32 + row) * 720 + part * 36 + col];
        }
        vit[o] = v;
    }
    __syncthreads();
    if (t < 32) {
        // both half-warps decode the same frame (the packed kernel works on two chunks per warp); FM PIDS
        // soft bits are punctured 1,1,1,1,1,0, so the int16 metrics cannot saturate
        const int l = t & 15;
        VitHalf<false> vh;
        vh.init(l);
        vitc_run<false>(vh, vit, PIDS_LEN, PIDS_LEN + 64, 0, PIDS_LEN + 64, 0, dec, t < 16, l);
        __syncwarp();
        // first maximum in state order; lane l holds states 2l, 2l+32 (E) and 2l+1, 2l+33 (O)
        int v = (short)(vh.E & 0xffff), state = 2 * l;
        const int w1 = (short)(vh.O & 0xffff);
        if (w1 > v) { v = w1; state = 2 * l + 1; }
        int v2 = (short)(vh.E >> 16), idx2 = 2 * l + 32;
        const int w3 = (short)(vh.O >> 16);
        if (w3 > v2) { v2 = w3; idx2 = 2 * l + 33; }
        if (v2 > v) { v = v2; state = idx2; }
#pragma unroll
        for (int o = 8; o; o >>= 1) {
            const int ov = __shfl_xor_sync(0xffffffffu, v, o, 16), oi = __shfl_xor_sync(0xffffffffu, state, o, 16);
            if (ov > v || (ov == v && oi < state)) { v = ov; state = oi; }
        }
        if (t == 0) {
            uint8_t pk[10];
            for (int i = 0; i < 10; i++) pk[i] = 0;
            for (int q = PIDS_LEN + 63; q >= 0; q--) {
                if (q >= 32 && q < 32 + PIDS_LEN) {
                    const int i = q - 32;
                    const int bit = ((state >> 5) & 1) ^ p.pn[i];
                    pk[i >> 3] |= (uint8_t)(bit << (7 - (i & 7)));
                }
                state = vitc_prev_head(state, dec, q);
            }
            if (st.pids_rec != 0xffffffffu) {
                uint8_t *w = p.log + (size_t)s * d.log_cap + st.pids_rec;
                for (int i = 0; i < 10; i++) w[i] = pk[i];
            }
            st.pids_pending = 0;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FFT_THREADS) k_demod(DevPtrs p, EngineDims d)
{
    // blockIdx.x = stream, blockIdx.y = symbol: the symbol-0 CTAs (which also carry the deferred PIDS
    // decode) are scheduled first
    const int s = blockIdx.x, sym = blockIdx.y, t = threadIdx.x;
    const StreamState &st = p.st[s];
    if (sym == 0 && st.pids_pending) pids_decode(p, d, s, t);      // previous block's PIDS frame (block-uniform branch)
    if (!st.active) return;

    __shared__ __align__(16) uint8_t in[IN_BYTES];
    __shared__ float2 buf[FFT_SMEM_ELEMS];
    __shared__ float2 symphase;

    const long long base = st.start + st.blk_samperr + (long long)NSYM * sym;
    const long long b0 = 4 * base - 28;                 // first needed cu8 byte (may be < 0 at stream start)
    const long long b0a = b0 & ~15LL;
    const int off = (int)(b0 - b0a);
    const uint8_t *iq = p.iq + (size_t)s * d.in_stride;
    const long long avail_bytes = 4 * st.in_avail;      // 2 bytes per sample, in_avail counts complex cu8 samples... (I,Q)
    (void)avail_bytes;
    {
        const int nvec = (off + 4 * NSYM + 28 + 15) / 16;
        uint4 *dst = reinterpret_cast<uint4 *>(in);
        for (int v = t; v < nvec; v += FFT_THREADS) {
            long long a = b0a + 16LL * v;
            uint4 x;
            if (a >= 0)
                x = __ldg(reinterpret_cast<const uint4 *>(iq + a));
            else
                x = make_uint4(0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu, 0x7f7f7f7fu);
            dst[v] = x;
        }
    }
    if (t == 0) {
        double a = (double)st.theta * (double)(NSYM * sym);
        double sn, cs;
        sincos(a, &sn, &cs);
        float2 r = make_float2((float)cs, (float)sn);
        symphase = cmul(st.phase0, r);
    }
    __syncthreads();

    const uint32_t *sw = reinterpret_cast<const uint32_t *>(in + off);
    const float2 *nco = p.nco + (size_t)s * NSYM;
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) {
        const int j = n1 * 128 + t;
        v[n1] = cmul(sample_at(sw, j), __ldg(&nco[j]));
    }
    if (t < NCP) {                                       // fold the windowed tail onto the head (acquire.c:247-248)
        const int j = NFFT + t;
        v[0] = cadd(v[0], cmul(sample_at(sw, j), __ldg(&nco[j])));
    }
    float2 out[2][8];
    fft2048_block(v, out, buf, p.twid, t);

    float2 *dst = p.bins + ((size_t)s * BLK + sym) * NBINS;
    const float2 sp = symphase;
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int k3 = 0; k3 < 8; k3++) {
            const int k = t + 128 * h + 256 * k3;        // natural-order bin
            const int b = (k + NFFT / 2) & (NFFT - 1);   // fftshift (defines.h:123-138)
            const int ci = compact_of_bin(b);
            if (ci >= 0) dst[ci] = cmul(out[h][k3], sp);
        }
}

void launch_demod(const DevPtrs &p, const EngineDims &d, cudaStream_t stream)
{
    dim3 grid(d.nstreams, BLK);
    k_demod<<<grid, FFT_THREADS, 0, stream>>>(p, d);
}

// ---------------------------------------------------------------------------
// stand-alone stage kernels for the numerics / parity tests
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(FFT_THREADS) k_fft_test(const float2 *in, float2 *outp, const float2 *twid)
{
    __shared__ float2 buf[FFT_SMEM_ELEMS];
    const int t = threadIdx.x;
    const float2 *x = in + (size_t)blockIdx.x * NFFT;
    float2 v[16], out[2][8];
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) v[n1] = x[n1 * 128 + t];
    fft2048_block(v, out, buf, twid, t);
    float2 *y = outp + (size_t)blockIdx.x * NFFT;
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
        for (int k3 = 0; k3 < 8; k3++) y[t + 128 * h + 256 * k3] = out[h][k3];
}

void launch_fft_test(const float2 *in, float2 *out, const float2 *twid, int nffts, cudaStream_t stream)
{
    k_fft_test<<<nffts, FFT_THREADS, 0, stream>>>(in, out, twid);
}

__global__ void k_halfband_test(const uint8_t *cu8, long long npairs, short2 *out)
{
    long long dd = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (dd < npairs) out[dd] = halfband_at(cu8, dd);
}

void launch_halfband_test(const uint8_t *cu8, long long npairs, short2 *out, cudaStream_t stream)
{
    int th = 256;
    long long bl = (npairs + th - 1) / th;
    k_halfband_test<<<(unsigned)bl, th, 0, stream>>>(cu8, npairs, out);
}

}  // namespace nb
