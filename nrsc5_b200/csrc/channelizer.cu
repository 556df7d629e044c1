// Wideband channeliser: one cu8 capture at 32 x 744 187.5 = 23 814 000 S/s -> up to hundreds of
// FM channels at 744 187.5 S/s cs16, written straight into the receive engine's cs16 input buffers.  The reference has
// no such stage (its ingest is one narrowband device per handle, reference src/nrsc5.c:130-207, src/rtltcp.c); this is
// the step in front of input_push_cs16 (src/input.c:119-124) that makes "many channels per GPU" a physical workload.
//
// It is the one dense contraction of the whole system, and the one kernel here that runs on the 5th-generation tensor
// cores:  for channel k and output sample n
//
//     acc[k][n] = sum_{u < 256} W_k[u] * (x[32 n + u] - (127 + 127j))        W_k[u] = round(2^19 h[255-u] e^{-j 2 pi 50 m_k u / 11907})
//     y[k][n]   = sat16( ((acc + 2^12) >> 13) * conj(P[(1600 m_k n) mod 11907]) + 2^14 >> 15 )
//
// (m_k = the channel's offset from the capture centre in units of 100 kHz; 100 kHz / 23.814 MHz = 50 / 11907, so
// every phasor comes from ONE table P[i] = round(32767 e^{+j 2 pi i / 11907}); all arithmetic is integer, the rounding
// shifts are arithmetic - the definition is exact and tests/test_channelizer.py restates it in numpy.)
//
// As a GEMM:  D[n][r] = sum_kappa A[n][kappa] * B[r][kappa],  kappa = 2u + {0: real, 1: imaginary part of x}
//   A[n][.]  = the 512 raw capture bytes starting at byte 64 n           (unsigned 8-bit, straight from the capture:
//              the rows overlap - row n+1 starts 64 bytes after row n - so the capture is addressed as a [rows][64 B]
//              matrix and K chunk c of row n is matrix row n + c: eight TMA boxes per tile, no im2col pass)
//   B[r][.]  = per channel four rows: {real, imaginary} x {high, low byte} of the 16-bit taps (signed 8-bit; the
//              real row holds Wr, -Wi alternating, the imaginary row Wi, Wr)
//   D        = int32 in TMEM; acc = 256 * D_high + D_low - 127 * (sum of the row), the last term a per-channel constant
// tcgen05.mma.kind::i8, M = 128 output samples x N = 128 rows (32 channels) x K = 512 per tile; one persistent CTA per
// (channel group, tile slot): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..5 = epilogue (TMEM -> registers ->
// rounding, rotation, saturation -> coalesced cs16 stores).  The 32 channels' taps (64 KB) stay in shared memory for the
// CTA's lifetime; capture tiles are double-buffered (2 x 64 KB), accumulators double-buffered in TMEM (2 x 128 columns).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/nrsc5_b200.h"

namespace nbch {

constexpr int DECIM = 32, TAPS = 256, KBYTES = 2 * TAPS;          // 512 bytes of K per output row
constexpr int CHUNK = 64, NCHUNK = KBYTES / CHUNK;               // K chunks of 64 bytes (one capture-matrix row each)
constexpr int TILE_M = 128;                                       // output samples per tile
constexpr int GROUP = 32, TILE_N = 4 * GROUP;                    // channels per CTA, rows of B
constexpr int PERIOD = 11907;                                     // phasor table length (100 kHz / 23.814 MHz = 50 / 11907)
constexpr int SHIFT1 = 13, TAP_SCALE_LOG2 = 19;                   // unit DC gain -> 64 LSB per input LSB (the cu8 -> Q15 convention);
                                                                  // taps stay below 127 * 256 + 127: both bytes of the split are int8
constexpr int THREADS = 192;
constexpr uint32_t A_STAGE_BYTES = NCHUNK * TILE_M * CHUNK;      // 65536
constexpr uint32_t W_BYTES = NCHUNK * TILE_N * CHUNK;            // 65536
constexpr int HALF = PERIOD / 2 + 1;                              // phasor table entries 0 .. 5953; the rest are their conjugates
constexpr uint32_t EPI_BYTES = ((HALF * 4 + 15) & ~15) + GROUP * 4 + 2 * GROUP * 4;   // half table, rotation steps, offset corrections
constexpr uint32_t SMEM_BYTES = W_BYTES + 2 * A_STAGE_BYTES + 1024 /* alignment */ + 256 /* barriers */ + EPI_BYTES;

struct Params {
    int nch;                   // channels
    int ngroups;
    long long nout;            // output samples per channel
    long long tiles;           // ceil(nout / TILE_M)
    int16_t *out;              // [nch][out_stride] cs16 (I, Q interleaved): int16 pairs
    size_t out_stride;         // int16 values between channels
    const int *rot_step;       // [nch] (1600 m_k) mod 11907
    const long long *corr;     // [nch][2] 127 * (sum Wr - sum Wi), 127 * (sum Wi + sum Wr)   (already in acc units)
    const short2 *phasor;      // [PERIOD]
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// K-major operand tile in shared memory, 64-byte rows, SWIZZLE_64B: 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw64(const void *tile, uint32_t byte_offset)
{
    const uint32_t addr = smem_u32(tile) + byte_offset;
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);                 // start address
    d |= (uint64_t)1 << 16;                                    // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(512u >> 4) << 32;                          // stride byte offset: 8 rows x 64 B
    d |= (uint64_t)1 << 46;                                    // descriptor version (Blackwell)
    d |= (uint64_t)4 << 61;                                    // layout type SWIZZLE_64B
    return d;
}
// instruction descriptor, kind::i8: D = S32, A = unsigned 8-bit (capture bytes), B = signed 8-bit (taps), both K-major
constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(IDESC), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Barriers {
    uint64_t w_full, a_full[2], a_empty[2], d_full[2], d_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(THREADS, 1) k_channelize(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                                                           Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *smem_w = smem;                                        // [8 chunks][128 rows][64 B]
    uint8_t *smem_a = smem + W_BYTES;                              // [2 stages][8 chunks][128 rows][64 B]
    Barriers &bar = *reinterpret_cast<Barriers *>(smem + W_BYTES + 2 * A_STAGE_BYTES);
    // the epilogue's tables in shared memory: the first half of the phasor table (P[11907 - i] = conj(P[i]), made so on
    // the host), and this group's rotation steps and offset corrections
    short2 *ph_half = reinterpret_cast<short2 *>(smem + W_BYTES + 2 * A_STAGE_BYTES + 256);
    int *s_rot = reinterpret_cast<int *>(reinterpret_cast<uint8_t *>(ph_half) + ((HALF * 4 + 15) & ~15));
    uint32_t *s_corr = reinterpret_cast<uint32_t *>(s_rot + GROUP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = (int)blockIdx.x % p.ngroups, slot = (int)blockIdx.x / p.ngroups, nslots = (int)gridDim.x / p.ngroups;

    for (int i = threadIdx.x; i < HALF; i += THREADS) ph_half[i] = p.phasor[i];
    for (int i = threadIdx.x; i < GROUP; i += THREADS) {
        const int ch = min(((int)blockIdx.x % p.ngroups) * GROUP + i, p.nch - 1);
        s_rot[i] = p.rot_step[ch];
        s_corr[2 * i] = (uint32_t)p.corr[2 * ch];
        s_corr[2 * i + 1] = (uint32_t)p.corr[2 * ch + 1];
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar.w_full, 1);
        for (int i = 0; i < 2; i++) {
            mbar_init(&bar.a_full[i], 1);
            mbar_init(&bar.a_empty[i], 1);
            mbar_init(&bar.d_full[i], 1);
            mbar_init(&bar.d_empty[i], 4);                         // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                               // TMEM: 2 x 128 columns of int32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = bar.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(&bar.w_full, W_BYTES);
            for (int c = 0; c < NCHUNK; c++)
                tma_load_2d(smem_w + (size_t)c * TILE_N * CHUNK, &map_w, &bar.w_full, 0, (group * NCHUNK + c) * TILE_N);
            unsigned it = 0;
            for (long long tile = slot; tile < p.tiles; tile += nslots, it++) {
                const int s = it & 1;
                if (it >= 2) mbar_wait(&bar.a_empty[s], ((it >> 1) - 1) & 1);
                mbar_expect_tx(&bar.a_full[s], A_STAGE_BYTES);
                for (int c = 0; c < NCHUNK; c++)                 // K chunk c of output row n = capture-matrix row n + c
                    tma_load_2d(smem_a + (size_t)s * A_STAGE_BYTES + (size_t)c * TILE_M * CHUNK, &map_x, &bar.a_full[s], 0,
                                (int)(tile * TILE_M) + c);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            mbar_wait(&bar.w_full, 0);
            unsigned it = 0;
            for (long long tile = slot; tile < p.tiles; tile += nslots, it++) {
                const int s = it & 1;
                if (it >= 2) mbar_wait(&bar.d_empty[s], ((it >> 1) - 1) & 1);       // the epilogue has drained this accumulator
                mbar_wait(&bar.a_full[s], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t tmem_d = tmem_base + (uint32_t)s * TILE_N;
                for (int c = 0; c < NCHUNK; c++)
                    for (int k = 0; k < CHUNK / 32; k++) {       // UMMA K = 32 bytes
                        const uint64_t da = umma_desc_sw64(smem_a + (size_t)s * A_STAGE_BYTES + (size_t)c * TILE_M * CHUNK, 32u * k);
                        const uint64_t db = umma_desc_sw64(smem_w + (size_t)c * TILE_N * CHUNK, 32u * k);
                        umma_i8(tmem_d, da, db, (c | k) ? 1u : 0u);
                    }
                umma_commit(&bar.a_empty[s]);                     // shared-memory stage free once these MMAs have read it
                umma_commit(&bar.d_full[s]);                      // accumulator complete
            }
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = tile rows =====
        const int quarter = warp & 3;
        const int row = 32 * quarter + lane;
        unsigned it = 0;
        for (long long tile = slot; tile < p.tiles; tile += nslots, it++) {
            const int s = it & 1;
            mbar_wait(&bar.d_full[s], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const long long n = tile * TILE_M + row;
            const int nmod = (int)(n % PERIOD);
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)s * TILE_N;
#pragma unroll 1
            for (int c0 = 0; c0 < GROUP; c0 += 8) {              // 8 channels = 32 columns per load
                uint32_t v[32];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr + 4u * c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // 32-bit arithmetic throughout: 256 * hi + lo wraps, the filter output itself is bounded by
                // 128 * sum(|Wr| + |Wi|) < 2^28 (checked when the tables are made), so the wrapped sum is the value; after
                // the shift a component is below 2^15 and the rotation's two products stay below 2^31.  No branches: the
                // eight table reads of a load are independent and go out together.
                short2 ph[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const unsigned qi = ((unsigned)s_rot[c0 + q] * (unsigned)nmod) % (unsigned)PERIOD;   // < 11907^2 < 2^32
                    const bool up = qi >= (unsigned)HALF;
                    const short2 v2 = ph_half[up ? (unsigned)PERIOD - qi : qi];
                    ph[q] = make_short2(v2.x, up ? (short)-v2.y : v2.y);
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int ch = group * GROUP + c0 + q;
                    const int ar = (int)(256u * v[4 * q + 0] + v[4 * q + 2] - s_corr[2 * (c0 + q)]);
                    const int ai = (int)(256u * v[4 * q + 1] + v[4 * q + 3] - s_corr[2 * (c0 + q) + 1]);
                    const int vr = (ar + (1 << (SHIFT1 - 1))) >> SHIFT1, vi = (ai + (1 << (SHIFT1 - 1))) >> SHIFT1;
                    int zr = vr * ph[q].x + vi * ph[q].y, zi = vi * ph[q].x - vr * ph[q].y;    // v * conj(P)
                    zr = (zr + (1 << 14)) >> 15;
                    zi = (zi + (1 << 14)) >> 15;
                    zr = zr > 32767 ? 32767 : (zr < -32768 ? -32768 : zr);
                    zi = zi > 32767 ? 32767 : (zi < -32768 ? -32768 : zi);
                    const uint32_t packed = (uint32_t)(uint16_t)(int16_t)zr | ((uint32_t)(uint16_t)(int16_t)zi << 16);
                    if (ch < p.nch && n < p.nout) *reinterpret_cast<uint32_t *>(p.out + (size_t)ch * p.out_stride + 2 * n) = packed;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.d_empty[s]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256));
}

}  // namespace nbch

// ===========================================================================
// host side
// ===========================================================================
using namespace nbch;

struct nrsc5b_channelizer {
    int device, nch, ngroups;
    std::vector<int> offsets;             // m_k: channel offset from the capture centre in 100 kHz steps
    std::vector<int16_t> taps;            // [nch][TAPS][2] (Wr, Wi) of W_k[u]
    std::vector<short2> phasor;           // [PERIOD]
    int8_t *d_w;                          // [ngroups][8][128][64]
    int *d_rot;
    long long *d_corr;
    short2 *d_phasor;
    CUtensorMap map_w;
    PFN_cuTensorMapEncodeTiled_v12000 encode;
};

static double bessel_i0(double x)
{
    double s = 1, t = 1;
    for (int k = 1; k < 50; k++) {
        t *= (x / (2 * k)) * (x / (2 * k));
        s += t;
    }
    return s;
}

// The integer tables of the definition, on the host (no device needed): phasor[11907], taps[nch][256] = W_k[u] and, for
// the kernel, the B operand bytes, the rotation steps and the per-channel offset corrections.
static void make_tables(const int *offsets, int nch, std::vector<short2> &phasor, std::vector<int16_t> &taps, std::vector<int8_t> *w,
                        std::vector<int> *rot, std::vector<long long> *corr)
{
    phasor.resize(PERIOD);
    for (int i = 0; i <= PERIOD / 2; i++) {
        const double a = 2.0 * M_PI * i / PERIOD;
        phasor[i] = make_short2((short)lrint(32767.0 * cos(a)), (short)lrint(32767.0 * sin(a)));
    }
    for (int i = PERIOD / 2 + 1; i < PERIOD; i++)                          // exactly conjugate-symmetric: the kernel keeps half of it
        phasor[i] = make_short2(phasor[PERIOD - i].x, (short)-phasor[PERIOD - i].y);
    // prototype low-pass: Kaiser-windowed sinc, -6 dB at 372 kHz: flat over a hybrid FM channel (+-200 kHz), >= 55 dB down from 544 kHz on (what folds onto the channel after /32), unit DC gain
    std::vector<double> h(TAPS);
    {
        const double fc = 372000.0 / 23814000.0, beta = 5.65;
        double sum = 0;
        for (int t = 0; t < TAPS; t++) {
            const double x = t - (TAPS - 1) / 2.0;
            const double sinc = fabs(x) < 1e-12 ? 2 * fc : sin(2 * M_PI * fc * x) / (M_PI * x);
            const double r = 2.0 * t / (TAPS - 1) - 1.0;
            h[t] = sinc * bessel_i0(beta * sqrt(1 - r * r)) / bessel_i0(beta);
            sum += h[t];
        }
        for (int t = 0; t < TAPS; t++) h[t] /= sum;
    }
    const int ngroups = (nch + GROUP - 1) / GROUP;
    taps.assign((size_t)nch * TAPS * 2, 0);
    if (w) w->assign((size_t)ngroups * W_BYTES, 0);
    if (rot) rot->assign(nch, 0);
    if (corr) corr->assign((size_t)nch * 2, 0);
    for (int k = 0; k < nch; k++) {
        const int m = offsets[k];
        const long long step = (((long long)50 * m) % PERIOD + PERIOD) % PERIOD;
        if (rot) (*rot)[k] = (int)((((long long)1600 * m) % PERIOD + PERIOD) % PERIOD);
        long long swr = 0, swi = 0, sabs = 0;
        const int g = k / GROUP, cl = k % GROUP;
        for (int u = 0; u < TAPS; u++) {
            // W_k[u] = 2^19 h[255-u] e^{-j 2 pi 50 m u / 11907}, from the integer phasor table
            const short2 ph = phasor[(size_t)((step * u) % PERIOD)];
            const double g0 = ldexp(h[TAPS - 1 - u], TAP_SCALE_LOG2) / 32767.0;
            const int wr = (int)lrint(g0 * ph.x), wi = (int)lrint(-g0 * ph.y);
            taps[((size_t)k * TAPS + u) * 2 + 0] = (int16_t)wr;
            taps[((size_t)k * TAPS + u) * 2 + 1] = (int16_t)wi;
            swr += wr;
            swi += wi;
            sabs += (wr < 0 ? -wr : wr) + (wi < 0 ? -wi : wi);
            if (!w) continue;
            // rows of B: real = (Wr, -Wi) against (xr, xi); imaginary = (Wi, Wr); each 16-bit value as signed high and low bytes
            const int vals[2][2] = { { wr, -wi }, { wi, wr } };
            for (int part = 0; part < 2; part++)
                for (int comp = 0; comp < 2; comp++) {
                    const int v = vals[part][comp];
                    const int hi = (v + 128) >> 8, lo = v - 256 * hi;           // v = 256 hi + lo, both in [-128, 127]
                    if (hi < -128 || hi > 127) { fprintf(stderr, "nrsc5_b200: channeliser tap out of range\n"); abort(); }
                    const int kappa = 2 * u + comp, ck = kappa / CHUNK, b = kappa % CHUNK;
                    const size_t base = ((size_t)(g * NCHUNK + ck) * TILE_N) * CHUNK;
                    (*w)[base + (size_t)(4 * cl + part) * CHUNK + b] = (int8_t)hi;
                    (*w)[base + (size_t)(4 * cl + 2 + part) * CHUNK + b] = (int8_t)lo;
                }
        }
        // the kernel's epilogue computes in 32 bits (k_channelize): the filter output must stay below 2^28
        if (128 * sabs >= (1ll << 28)) { fprintf(stderr, "nrsc5_b200: channeliser taps too large for the 32-bit epilogue\n"); abort(); }
        if (corr) {
            (*corr)[2 * k + 0] = 127ll * (swr - swi);
            (*corr)[2 * k + 1] = 127ll * (swi + swr);
        }
    }
}

/* The definition's tables without a device: taps[nch][256][2], phasor[11907][2] (either may be NULL). */
extern "C" int nrsc5b_chan_make_tables(const int *offsets_100khz, int nch, int16_t *taps, int16_t *phasor)
{
    if (!offsets_100khz || nch <= 0 || nch > 4096) return NRSC5B_EINVAL;
    std::vector<short2> ph;
    std::vector<int16_t> tp;
    make_tables(offsets_100khz, nch, ph, tp, nullptr, nullptr, nullptr);
    if (taps) memcpy(taps, tp.data(), tp.size() * sizeof(int16_t));
    if (phasor) memcpy(phasor, ph.data(), PERIOD * sizeof(short2));
    return NRSC5B_OK;
}

extern "C" int nrsc5b_chan_create(nrsc5b_channelizer_t **out, int device, const int *offsets_100khz, int nch)
{
    if (!out || !offsets_100khz || nch <= 0 || nch > 4096) return NRSC5B_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) {
        fprintf(stderr, "nrsc5_b200: no usable CUDA device (the channeliser has no CPU path)\n");
        return NRSC5B_ENODEV;
    }
    if (cudaSetDevice(device) != cudaSuccess) return NRSC5B_ENODEV;
    nrsc5b_channelizer *c = new nrsc5b_channelizer();
    c->device = device;
    c->nch = nch;
    c->ngroups = (nch + GROUP - 1) / GROUP;
    c->offsets.assign(offsets_100khz, offsets_100khz + nch);
    c->d_w = nullptr; c->d_rot = nullptr; c->d_corr = nullptr; c->d_phasor = nullptr;
    // driver entry point for the tensor-map encoder (no link-time dependency on libcuda)
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { delete c; return NRSC5B_ECUDA; }
        c->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    std::vector<int8_t> w;
    std::vector<int> rot;
    std::vector<long long> corr;
    make_tables(offsets_100khz, nch, c->phasor, c->taps, &w, &rot, &corr);
    bool ok = cudaMalloc(&c->d_w, w.size()) == cudaSuccess && cudaMalloc(&c->d_rot, nch * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&c->d_corr, corr.size() * sizeof(long long)) == cudaSuccess &&
              cudaMalloc(&c->d_phasor, PERIOD * sizeof(short2)) == cudaSuccess;
    ok = ok && cudaMemcpy(c->d_w, w.data(), w.size(), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(c->d_rot, rot.data(), nch * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(c->d_corr, corr.data(), corr.size() * sizeof(long long), cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(c->d_phasor, c->phasor.data(), PERIOD * sizeof(short2), cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) {
        // taps as a [ngroups * 8 * 128 rows][64 B] matrix, boxes of 128 rows, 64-byte swizzle
        const cuuint64_t dims[2] = { CHUNK, (cuuint64_t)c->ngroups * NCHUNK * TILE_N };
        const cuuint64_t strides[1] = { CHUNK };
        const cuuint32_t box[2] = { CHUNK, TILE_N }, es[2] = { 1, 1 };
        ok = c->encode(&c->map_w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->d_w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    if (ok) ok = cudaFuncSetAttribute(k_channelize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) == cudaSuccess;
    if (!ok) {
        nrsc5b_chan_destroy(c);
        return NRSC5B_ECUDA;
    }
    *out = c;
    return NRSC5B_OK;
}

extern "C" void nrsc5b_chan_destroy(nrsc5b_channelizer_t *c)
{
    if (!c) return;
    cudaFree(c->d_w);
    cudaFree(c->d_rot);
    cudaFree(c->d_corr);
    cudaFree(c->d_phasor);
    delete c;
}

/* The integer tables the definition is made of (for the numpy restatement in the tests): taps[nch][256][2] = (Wr, Wi)
 * of W_k[u], phasor[11907][2]. */
extern "C" int nrsc5b_chan_tables(nrsc5b_channelizer_t *c, int16_t *taps, int16_t *phasor)
{
    if (!c) return NRSC5B_EINVAL;
    if (taps) memcpy(taps, c->taps.data(), c->taps.size() * sizeof(int16_t));
    if (phasor) memcpy(phasor, c->phasor.data(), PERIOD * sizeof(short2));
    return NRSC5B_OK;
}

/* How many output samples a capture of `nbytes` gives per channel: every output needs 256 input samples. */
extern "C" long long nrsc5b_chan_outputs(size_t nbytes)
{
    const long long samples = (long long)(nbytes / 2);
    return samples < TAPS ? 0 : (samples - TAPS) / DECIM + 1;
}

/* Device-resident capture (cu8, I/Q interleaved, 23 814 000 S/s; 64-byte aligned, nbytes of it valid) -> out[nch][out_stride]
 * cs16 on the device (out_stride in int16 values, >= 2 * outputs; 4-byte aligned rows).  Asynchronous on `cuda_stream`. */
extern "C" int nrsc5b_chan_run_device(nrsc5b_channelizer_t *c, const void *d_cu8, size_t nbytes, void *d_out, size_t out_stride,
                                      void *cuda_stream)
{
    if (!c || !d_cu8 || !d_out || ((uintptr_t)d_cu8 & 63) || (nbytes & 63) || ((uintptr_t)d_out & 3) || (out_stride & 1)) return NRSC5B_EINVAL;
    const long long nout = nrsc5b_chan_outputs(nbytes);
    if (nout <= 0) return NRSC5B_OK;
    if ((size_t)(2 * nout) > out_stride) return NRSC5B_EINVAL;
    if (cudaSetDevice(c->device) != cudaSuccess) return NRSC5B_ENODEV;
    // the capture as a [rows][64 B] matrix; rows past the end read as zero (only rows of outputs >= nout touch them)
    CUtensorMap map_x;
    const cuuint64_t dims[2] = { CHUNK, (cuuint64_t)(nbytes / CHUNK) };
    const cuuint64_t strides[1] = { CHUNK };
    const cuuint32_t box[2] = { CHUNK, TILE_M }, es[2] = { 1, 1 };
    if (c->encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(d_cu8), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return NRSC5B_ECUDA;
    Params p;
    p.nch = c->nch;
    p.ngroups = c->ngroups;
    p.nout = nout;
    p.tiles = (nout + TILE_M - 1) / TILE_M;
    p.out = reinterpret_cast<int16_t *>(d_out);
    p.out_stride = out_stride;
    p.rot_step = c->d_rot;
    p.corr = c->d_corr;
    p.phasor = c->d_phasor;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    long long slots = sms / c->ngroups;
    if (slots < 1) slots = 1;
    if (slots > p.tiles) slots = p.tiles;
    k_channelize<<<(unsigned)(slots * c->ngroups), THREADS, SMEM_BYTES, reinterpret_cast<cudaStream_t>(cuda_stream)>>>(map_x, c->map_w, p);
    return cudaGetLastError() == cudaSuccess ? NRSC5B_OK : NRSC5B_ECUDA;
}

/* Host convenience (tests): host capture in, host cs16 out[nch][2 * outputs]. */
extern "C" int nrsc5b_chan_run(nrsc5b_channelizer_t *c, const uint8_t *cu8, size_t nbytes, int16_t *out)
{
    if (!c || !cu8 || !out) return NRSC5B_EINVAL;
    nbytes &= ~(size_t)63;                                  // whole 64-byte rows (32 complex samples)
    const long long nout = nrsc5b_chan_outputs(nbytes);
    if (nout <= 0) return NRSC5B_OK;
    if (cudaSetDevice(c->device) != cudaSuccess) return NRSC5B_ENODEV;
    uint8_t *d_in = nullptr;
    int16_t *d_out = nullptr;
    const size_t padded = (nbytes + 63) & ~(size_t)63, stride = (size_t)(2 * nout);
    int rc = NRSC5B_ECUDA;
    if (cudaMalloc(&d_in, padded + 64) == cudaSuccess && cudaMalloc(&d_out, (size_t)c->nch * stride * sizeof(int16_t)) == cudaSuccess &&
        cudaMemset(d_in, 0, padded + 64) == cudaSuccess && cudaMemcpy(d_in, cu8, nbytes, cudaMemcpyHostToDevice) == cudaSuccess) {
        rc = nrsc5b_chan_run_device(c, d_in, nbytes, d_out, stride, nullptr);
        if (rc == NRSC5B_OK && cudaDeviceSynchronize() != cudaSuccess) {
            fprintf(stderr, "nrsc5_b200: channeliser kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            rc = NRSC5B_ECUDA;
        }
        if (rc == NRSC5B_OK && cudaMemcpy(out, d_out, (size_t)c->nch * stride * sizeof(int16_t), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = NRSC5B_ECUDA;
    }
    cudaFree(d_in);
    cudaFree(d_out);
    return rc;
}
