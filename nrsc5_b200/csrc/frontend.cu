// Stand-alone stage kernels (FFT, halfband) behind the C ABI's single-stage entry points; the product
// kernels live in front.cuh / viterbi_chunk.cuh / engine.cu.
#include "common.cuh"
#include "fft.cuh"

namespace nb {

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
