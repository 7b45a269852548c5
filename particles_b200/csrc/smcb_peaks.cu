// smcb_peaks.cu -- measured ceilings of THIS device for the roofline statements in bench.py:
// the fp64 FMA issue peak (the secondary bound of the step kernel) and a
// read+write streaming probe with the step kernel's access pattern (16-byte loads / stores).
// MEASURED_PEAKS.json (driver-written) stays the primary HBM denominator; these are builder-side
// cross-checks written by profiles/measure_peaks.py into profiles/*_peaks.json.
#include "smcb_common.cuh"

using namespace smcb;

namespace {

// ILP independent DFMA chains per thread, `iters` dependent links each
template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = fma(v[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
    if (s == 123.456) out[0] = s;          // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) k_stream(const double2 *__restrict__ in, double2 *__restrict__ out,
                                                int64_t n2) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 v = in[i];
        v.x += 1.0;
        out[i] = v;
    }
}

}  // namespace

// out[0] = fp64 TFLOP/s (2 flops per DFMA), out[1] = DFMA warp-instructions per cycle per SM at the
// SM clock implied by `sm_mhz` (0 -> not computed), out[2] = kernel ms.  Synchronises the stream.
extern "C" int smcb_measure_fp64_peak(smcb_ctx *c, double sm_mhz, double *out3) {
    SMCB_REQUIRE(c && out3, "smcb_measure_fp64_peak: NULL argument");
    SMCB_CUDA(cudaSetDevice(c->device));
    constexpr int ILP = 8;
    const int grid = kSMs * 8, iters = 20000;
    cudaEvent_t e0, e1;
    SMCB_CUDA(cudaEventCreate(&e0));
    SMCB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        SMCB_CUDA(cudaEventRecord(e0, c->stream));
        k_dfma<ILP><<<grid, 256, 0, c->stream>>>(c->ws, iters, 0.999999, 1e-7);
        SMCB_CUDA(cudaEventRecord(e1, c->stream));
        SMCB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SMCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 5;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SMCB_CUDA(cudaGetLastError());
    const double dfma = (double)grid * 256.0 * ILP * iters;
    out3[0] = 2.0 * dfma / (best * 1e-3) / 1e12;
    out3[1] = sm_mhz > 0 ? (dfma / 32.0) / (best * 1e-3 * sm_mhz * 1e6) / kSMs : 0.0;
    out3[2] = best;
    return SMCB_OK;
}

// out[0] = GB/s of a read + write pass over `bytes` (each way) with 16-byte accesses, best of 5.
extern "C" int smcb_measure_stream_peak(smcb_ctx *c, const double *in, double *out, int64_t n, double *out1) {
    SMCB_REQUIRE(c && in && out && out1 && n > 0 && (n & 1) == 0, "smcb_measure_stream_peak: bad argument");
    SMCB_CUDA(cudaSetDevice(c->device));
    cudaEvent_t e0, e1;
    SMCB_CUDA(cudaEventCreate(&e0));
    SMCB_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        SMCB_CUDA(cudaEventRecord(e0, c->stream));
        k_stream<<<kSMs * 8, 256, 0, c->stream>>>(reinterpret_cast<const double2 *>(in),
                                                   reinterpret_cast<double2 *>(out), n / 2);
        SMCB_CUDA(cudaEventRecord(e1, c->stream));
        SMCB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SMCB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 6;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SMCB_CUDA(cudaGetLastError());
    out1[0] = 16.0 * (double)n / (best * 1e-3) / 1e9;
    return SMCB_OK;
}
