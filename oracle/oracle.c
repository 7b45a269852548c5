/* Plain-C restatement of the serial kernels of the reference hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative
 * to /root/reference.
 */
#include <math.h>
#include <stdint.h>

/* particles/resampling.py:484-509 (numba inverse_cdf): for sorted su,
 *   A[n] = min{ j : W[0] + ... + W[j] >= su[n] },
 * sequential fp64 running sum.  The reference has no bounds check; we stop at
 * N-1 ("failure detection"). */
void oracle_inverse_cdf(const double *su, int64_t M, const double *W, int64_t N,
                        int64_t *A)
{
    int64_t j = 0;
    double s = W[0];
    for (int64_t n = 0; n < M; n++) {
        while (su[n] > s && j < N - 1) {
            j += 1;
            s += W[j];
        }
        A[n] = j;
    }
}

/* np.searchsorted(cdf, su, side='left') clipped to N-1: the definition the
 * CUDA search kernel is held to, bit-exactly, on the device's own CDF. */
void oracle_searchsorted_left(const double *cdf, int64_t N, const double *su,
                              int64_t M, int64_t *A)
{
    for (int64_t n = 0; n < M; n++) {
        int64_t lo = 0, hi = N;
        while (lo < hi) {
            int64_t mid = lo + ((hi - lo) >> 1);
            if (cdf[mid] < su[n]) lo = mid + 1; else hi = mid;
        }
        A[n] = lo < N - 1 ? lo : N - 1;
    }
}

/* particles/resampling.py:217-226 (Weights.__init__) with plain sequential
 * sums: out = {max, log_mean, ESS, sum w}.  NaN -> -inf written in place. */
void oracle_weights(double *lw, int64_t N, double *W, double *out)
{
    double m = -INFINITY;
    for (int64_t i = 0; i < N; i++) {
        if (isnan(lw[i])) lw[i] = -INFINITY;
        if (lw[i] > m) m = lw[i];
    }
    double s = 0.0;
    for (int64_t i = 0; i < N; i++) { W[i] = exp(lw[i] - m); s += W[i]; }
    double q = 0.0;
    for (int64_t i = 0; i < N; i++) { W[i] = W[i] / s; q += W[i] * W[i]; }
    out[0] = m; out[1] = m + log(s / (double)N); out[2] = 1.0 / q; out[3] = s;
}
