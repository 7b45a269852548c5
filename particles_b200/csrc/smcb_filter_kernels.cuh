#pragma once
// smcb_filter_kernels.cuh -- the fused SMC step (particles/core.py:369-383) for the 1-D
// Normal-kernel model family: one step = at most two kernels and no host sync.
//
//   k_scan_w   (resampling steps only; exits at once otherwise)
//       W_i = exp(lw_i - m)/s computed on the fly  ->  inclusive scan  ->  cdf      16 B/particle
//   k_move
//       resampling step:  su_k -> A_k = search(cdf) -> xp = X[A_k] -> x' ~ M_t(xp)
//                         -> lw' = logG  -> (max, sum exp, sum exp^2) partials        40 B/particle
//       otherwise:        xp = X_k -> x' -> lw' = lw + logG -> partials               32 B/particle
//       the block that retires last merges the partials in a fixed order and performs
//       compute_summaries (core.py:351-367) and the ESS test of the NEXT step
//       (core.py:181-183) on the device.
//
// The resample / no-resample decision, t, the ping-pong index and logLt live in a small
// device struct (FilterDev), so the launch arguments are identical for every step.
#include <string.h>

#include <new>

#include "smcb_common.cuh"
#include "smcb_math.cuh"
#include "smcb_models.cuh"
#include "smcb_reduce.cuh"
#include "smcb_scan.cuh"
#include "smcb_search.cuh"

using namespace smcb;

namespace smcb {

struct FilterDev {
    long long t;          // next step to run
    int cur;              // X[cur], lw[cur] hold the particles of step t-1
    int rs_flag;          // decision for step t, taken at the end of step t-1
    int last_rs;          // rs_flag of the step just completed
    int pad;
    double logLt, log_mean_w, ess;
    double wm, ws, wq;    // (max, sum exp, sum exp^2) of the inferential weights
    double am, as, aq;    // same for the auxiliary weights (APF); == w* otherwise
    double reset_c;       // APF: log_mean_exp(logetat, W), core.py:302
    long long t_stop;     // graph mode: the WHILE node keeps iterating while t < t_stop
    // exact global resampling over shards: shard r owns [goff[r], goff[r+1]) of the global CDF
    double goff[9], gpi[8];
    int sync_timeout;     // a bounded peer wait expired (diagnostic; results are then invalid)
    int pad2;
};

struct FilterArgs {
    double *X[2];
    double *lw[2];
    long long *A;
    double *cdf;
    double *su;           // multinomial: z = cumsum(-log u), (n + 1)
    const double *data;   // (T)
    const double *sc;     // (T) per-step model constants or NULL
    double *summaries;    // (T, 4)
    const double *z_in, *u_in;
    FilterDev *st;
    double *partials;
    unsigned int *ticket;
    ScanState scan, scan2;
    int64_t n, n_global, index_offset, T;
    int dy;               // observation dimension
    int world, rank;      // particle shards over `world` GPUs (1 = single device)
    int grid;             // blocks of the step kernel (= number of partials / tile prefixes)
    double *local_stats;  // world > 1: this rank's {w.m, w.s, w.q, 0, aux.m, aux.s, aux.q, 0}
    const double *gathered;  // world > 1: all ranks' local_stats, rank-major, after the all-gather
    // peer-memory exchange (NVLink P2P): every rank owns a mailbox of 2 x world x 16 doubles
    // ([parity][sender][8 stats, epoch, pad]); mail_peer[p] is rank p's mailbox mapped here.
    double *mail_local;
    double *mail_peer[8];
    int64_t chunk;        // pairs of particles per block (blocked assignment, multiple of kBlock)
    double *tile_pref;    // (grid + 1) exclusive prefixes of the blocks' normalised weight mass
    double essrmin;
    Philox key;
    // exact global resampling (world > 1): peers' particles and CDFs mapped over NVLink
    int rs_global;
    unsigned int *ticket2;
    double *stage_X, *stage_lw;
    const double *pX[8][2];
    const double *pcdf[8];
};

// bounded spin on a flag another GPU raises (epoch counters only grow); ~4 s at 2 GHz
__device__ __forceinline__ void wait_epoch(const volatile double *flag, double epoch, FilterDev *st) {
    const long long t0 = clock64();
    if (*reinterpret_cast<volatile int *>(&st->sync_timeout)) return;   // already broken: do not stall again
    while (*flag < epoch) {
        if (clock64() - t0 > 8000000000ll) { st->sync_timeout = 1; break; }
    }
    __threadfence_system();
}

__device__ __forceinline__ StepK step_consts(const FilterArgs &a, long long t) {
    StepK k;
    k.t = t;
#pragma unroll
    for (int i = 0; i < kMaxDy; i++) {
        k.yv[i] = (i < a.dy) ? a.data[t * a.dy + i] : 0.0;
        k.yn[i] = (i < a.dy && t + 1 < a.T) ? a.data[(t + 1) * a.dy + i] : 0.0;
    }
    k.y = k.yv[0];
    k.y_next = k.yn[0];
    k.sc0 = a.sc ? a.sc[t] : 0.0;
    return k;
}

// compute_summaries (core.py:351-367) + time_to_resample for the next step (core.py:181-183)
// executed by the whole last block; thread 0 owns the scalar work.
// w / aux: statistics over ALL particles (all ranks); wl / xl: over this rank's shard.
template <bool APF>
__device__ __forceinline__ void finalize_step(const FilterArgs &a, const Lse3 &w, const Lse3 &aux,
                                              const Lse3 &wl, const Lse3 &auxl) {
    __shared__ int s_next_flag;
    if (threadIdx.x == 0) {
        FilterDev *st = a.st;
        const long long t = st->t;
        const double N = (double)a.n_global;
        double log_mean, ess;
        weights_scalars(w, N, log_mean, ess);
        const bool fresh = (t == 0) || (st->rs_flag != 0);
        const double loglt = fresh ? log_mean : (log_mean - st->log_mean_w);   // core.py:355-358
        const double logLt = st->logLt + loglt;
        double *row = a.summaries + (size_t)t * SMCB_SUMMARY_STRIDE;
        row[0] = ess; row[1] = logLt; row[2] = (double)st->rs_flag; row[3] = log_mean;
        st->logLt = logLt; st->log_mean_w = log_mean; st->ess = ess;
        st->wm = w.m; st->ws = w.s; st->wq = w.q;
        st->last_rs = st->rs_flag;
        const Lse3 &x = APF ? aux : w;
        const Lse3 &xl = APF ? auxl : wl;
        // the CDF of a resampling step is built from this shard's own (auxiliary) weights
        st->am = xl.m; st->as = xl.s; st->aq = xl.q;
        double lm_aux, ess_aux;
        weights_scalars(x, N, lm_aux, ess_aux);
        // log-weight every resampled particle restarts from (minus logeta[A] for an APF):
        //   single device, non-APF : 0                       (Weights(), core.py:305)
        //   single device, APF     : log_mean_exp(logetat, W) (core.py:302) = LSE(aux) - LSE(w)
        //   sharded                : LSE_shard(aux) - LSE_all(w) + log(world): each shard resamples
        //                            locally and carries its share of the mass
        double rc = 0.0;
        if (a.rs_global)        // one global resampling: the reference's restart, from global sums
            rc = APF ? (log(x.s) + x.m) - (log(w.s) + w.m) : 0.0;
        else if (APF || a.world > 1)
            rc = (log(xl.s) + xl.m) - (log(w.s) + w.m) + log((double)a.world);
        st->reset_c = rc;
        int flag = (t + 1 < a.T) && (ess_aux < N * a.essrmin);    // strict <, NaN -> False
        st->rs_flag = flag;
        st->cur ^= 1;
        st->t = t + 1;
        s_next_flag = flag;
    }
    __syncthreads();
    if (s_next_flag) {
        // The blocks own contiguous particle ranges, so their partial sums ARE the tile
        // aggregates of the weight scan of the next step: turn them into exclusive prefixes
        // P_0 = 0 <= P_1 <= ... <= P_G here (fixed order, monotone), and the scan kernel needs
        // no look-back at all.
        __shared__ double s_w[kBlock / 32];
        const int G = a.grid, K = APF ? 2 : 1, j = APF ? 1 : 0;
        const double xm = a.st->am, xs = a.st->as;
        const int per = (G + kBlock - 1) / kBlock;
        const int b0 = threadIdx.x * per;
        double loc[8], run = 0.0;
        for (int i = 0; i < per && i < 8; i++) {
            const int b = b0 + i;
            double v = 0.0;
            if (b < G) {
                const volatile double *pp = a.partials + (size_t)b * 4 * K + 4 * j;
                v = pp[1] * fexp(pp[0] - xm) / xs;
            }
            run = run + v;
            loc[i] = run;
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const double iw = warp_scan_monotone(run, lane);
        if (lane == 31) s_w[warp] = iw;
        __syncthreads();
        double woff = 0.0;
        for (int w = 0; w < kBlock / 32; w++)
            if (w < warp) woff = woff + s_w[w];
        const double up = __shfl_up_sync(0xffffffffu, iw, 1);
        const double excl = (lane == 0) ? woff : (woff + up);
        const double cap = woff + iw;
        if (threadIdx.x == 0) a.tile_pref[0] = 0.0;
        for (int i = 0; i < per && i < 8; i++) {
            const int b = b0 + i;
            if (b < G) a.tile_pref[b + 1] = fmin(excl + loc[i], cap);
        }
        if (a.su) {   // multinomial: arm the look-back state of the spacings scan
            const int64_t tiles2 = (a.n + 1 + kScanTile - 1) / kScanTile;
            const int64_t words2 = 2 + tiles2 + tiles2 + 2;
            unsigned long long *q = reinterpret_cast<unsigned long long *>(a.scan2.ticket);
            for (int64_t i = threadIdx.x; i < words2; i += blockDim.x) q[i] = kNotReady;
        }
    }
}

template <int K>
__device__ __forceinline__ void publish_local(const FilterArgs &a, const Lse3 (&tot)[K]) {
    const double v[8] = {tot[0].m, tot[0].s, tot[0].q, 0.0, tot[K - 1].m, tot[K - 1].s, tot[K - 1].q, 0.0};
    if (a.mail_local == nullptr) {          // host-driven exchange: NCCL all-gather of local_stats
        if (threadIdx.x == 0)
            for (int i = 0; i < 8; i++) a.local_stats[i] = v[i];
        return;
    }
    // fused exchange: the step kernel's last CTA stores this shard's statistics straight into every
    // peer's mailbox over NVLink (one lane per peer), fences, then raises the epoch flag; k_finish
    // on each rank waits for `world` flags.  No host call, no collective launch on the step.
    const long long t = a.st->t;
    if ((int)threadIdx.x < a.world) {
        double *slot = a.mail_peer[threadIdx.x] + ((size_t)(t & 1) * a.world + a.rank) * 16;
        for (int i = 0; i < 8; i++) slot[i] = v[i];
        __threadfence_system();
        *reinterpret_cast<volatile double *>(slot + 8) = (double)(t + 1);
    }
}

__device__ __forceinline__ double fix_nan(double v) { return v != v ? -CUDART_INF : v; }  // resampling.py:220

// ---------------------------------------------------------------------------
// t = 0: generate_particles + reweight (core.py:315-324, 373-374)
// ---------------------------------------------------------------------------
template <class M, int FK>
__global__ void __launch_bounds__(kBlock) k_init(M model, FilterArgs a) {
    constexpr bool APF = FkTraits<FK>::apf;
    constexpr int K = APF ? 2 : 1;
    __shared__ Lse3 smem[kBlock / 32];
    const StepK k = step_consts(a, 0);
    double *Xo = a.X[0], *lwo = a.lw[0];
    Lse3 acc[K];
    acc[0] = lse3_empty();
    if (APF) acc[K - 1] = lse3_empty();
    const int64_t n = a.n, npairs = (n + 1) >> 1;
    const bool has_next = APF && a.T > 1;
    const int64_t pstart = (int64_t)blockIdx.x * a.chunk;
    const int64_t pend = pstart + a.chunk < npairs ? pstart + a.chunk : npairs;
    constexpr int D = M::D, NZ = M::NZ;
    for (int64_t p = pstart + threadIdx.x; p < pend; p += kBlock) {
        double z[2][NZ], x[2][D], l[2], av[2];
#pragma unroll
        for (int c = 0; c < NZ; c++) {
            if (a.z_in) {                                  // injected normals: (T, NZ, n)
                const double *zz = a.z_in + (size_t)c * n;
                z[0][c] = zz[2 * p];
                z[1][c] = (2 * p + 1 < n) ? zz[2 * p + 1] : 0.0;
            } else {
                normal_pair_fast(a.key, (uint64_t)((a.index_offset >> 1) + p), 0u, (uint32_t)c, z[0][c], z[1][c]);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double d;
            model_init<M, FK>(model, k, z[j], x[j], d);
            l[j] = fix_nan(d);
            av[j] = has_next ? fix_nan(l[j] + model_logeta<M>(model, k, x[j])) : -CUDART_INF;
        }
        if (2 * p + 1 < n) {
#pragma unroll
            for (int c = 0; c < D; c++) st2(Xo + (size_t)c * n + 2 * p, x[0][c], x[1][c]);
            st2(lwo + 2 * p, l[0], l[1]);
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) Xo[(size_t)c * n + 2 * p] = x[0][c];
            lwo[2 * p] = l[0];
            l[1] = -CUDART_INF; av[1] = -CUDART_INF;   // masked slot contributes exactly 0
        }
        lse3_add_batch<2>(acc[0], l);
        if (APF) lse3_add_batch<2>(acc[K - 1], av);
    }
    Lse3 tot[K];
    if (!grid_merge_lse3<kBlock, K>(acc, a.partials, a.ticket, smem, tot)) return;
    if (threadIdx.x == 0) a.st->cur = 1;   // finalize flips it to 0: step 0 wrote buffers [0]
    if (a.world > 1) { publish_local<K>(a, tot); return; }
    finalize_step<APF>(a, tot[0], tot[K - 1], tot[0], tot[K - 1]);
}

// ---------------------------------------------------------------------------
// resampling steps: normalised (auxiliary) weights -> CDF   (resampling.py:223-225 + scan)
// ---------------------------------------------------------------------------
template <class M, int FK>
struct LoadWeights {
    const double *lw, *X;
    int64_t ntot;  // particles on this device (SoA component stride)
    double m, s;
    M model;
    StepK kprev;   // step t-1 with y_next = data[t]: what logeta(t-1, X) needs
    __device__ __forceinline__ void operator()(int64_t i0, int64_t n, double (&v)[8]) const {
        constexpr bool APF = FkTraits<FK>::apf;
        constexpr int D = M::D;
        double l[8], x[8][APF ? D : 1];
        if (i0 + 8 <= n) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) { double2 t = ld2(lw + i0 + j); l[j] = t.x; l[j + 1] = t.y; }
            if (APF) {
#pragma unroll
                for (int c = 0; c < D; c++) {
#pragma unroll
                    for (int j = 0; j < 8; j += 2) {
                        double2 t = ld2(X + (size_t)c * ntot + i0 + j);
                        x[j][APF ? c : 0] = t.x; x[j + 1][APF ? c : 0] = t.y;
                    }
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                l[j] = (i0 + j < n) ? lw[i0 + j] : -CUDART_INF;
                if (APF) {
#pragma unroll
                    for (int c = 0; c < D; c++) x[j][APF ? c : 0] = (i0 + j < n) ? X[(size_t)c * ntot + i0 + j] : 0.0;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double e = l[j];
            if (APF) e = fix_nan(e + model_logeta<M>(model, kprev, x[j]));
            v[j] = (i0 + j < n) ? fexp(e - m) / s : 0.0;
        }
    }
};

template <class M, int FK>
__global__ void __launch_bounds__(kBlock) k_scan_w(M model, FilterArgs a) {
    const FilterDev *st = a.st;
    if (!st->rs_flag) return;
    __shared__ double s_warp[kBlock / 32];
    const long long t = st->t;
    LoadWeights<M, FK> load;
    load.lw = a.lw[st->cur];
    load.X = a.X[st->cur];
    load.ntot = a.n;
    load.m = st->am;
    load.s = st->as;
    load.model = model;
    load.kprev = step_consts(a, t - 1);
    // block b scans the particles it owns, [2 b chunk, 2 (b+1) chunk), from the exclusive
    // prefix P_b that finalize_step derived from the previous kernel's partial sums; every value
    // is clamped into [P_b, P_{b+1}], so the CDF is non-decreasing across blocks by construction
    const int64_t n = a.n;
    const int64_t e0 = 2 * (int64_t)blockIdx.x * a.chunk;
    const int64_t e1 = e0 + 2 * a.chunk < n ? e0 + 2 * a.chunk : n;
    const double p_b = a.tile_pref[blockIdx.x], p_next = a.tile_pref[blockIdx.x + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double carry = 0.0;
    for (int64_t base0 = e0; base0 < e1; base0 += kScanTile) {
        const int64_t i0 = base0 + (int64_t)tid * kScanItems;
        double r[kScanItems];
        load(i0, e1, r);
#pragma unroll
        for (int j = 1; j < kScanItems; j++) r[j] = r[j - 1] + r[j];
        const double iw = warp_scan_monotone(r[kScanItems - 1], lane);
        if (lane == 31) s_warp[warp] = iw;
        __syncthreads();
        double woff = 0.0, total = 0.0;
#pragma unroll
        for (int w = 0; w < kBlock / 32; w++) {
            if (w < warp) woff = woff + s_warp[w];
            total = total + s_warp[w];
        }
        const double incl = woff + iw;
        const double up = __shfl_up_sync(0xffffffffu, iw, 1);
        const double excl = (lane == 0) ? woff : (woff + up);
        const double b_i = p_b + carry;                       // base of this sub-tile
        const double carry_next = carry + total;
        const double b_next = fmin(p_b + carry_next, p_next); // base of the next one
        const double tb = b_i + excl;
        const double cap = fmin(b_i + incl, b_next);
        double o[kScanItems];
#pragma unroll
        for (int j = 0; j < kScanItems; j++) o[j] = fmin(tb + r[j], cap);
        if (i0 + kScanItems <= e1) {
            store_items(a.cdf, i0, o);
        } else {
#pragma unroll
            for (int j = 0; j < kScanItems; j++)
                if (i0 + j < e1) a.cdf[i0 + j] = o[j];
        }
        carry = carry_next;
        __syncthreads();
    }
    if (a.rs_global) {
        // tell every peer that this shard's CDF of step t is complete ("last block done" ticket,
        // then one lane per peer raises the scan epoch in that peer's mailbox)
        __shared__ bool s_last;
        __threadfence_system();
        __syncthreads();
        if (tid == 0) s_last = (atomicInc(a.ticket2, gridDim.x - 1) == gridDim.x - 1);
        __syncthreads();
        if (s_last && tid < a.world) {
            __threadfence_system();
            double *slot = a.mail_peer[tid] + ((size_t)(t & 1) * a.world + a.rank) * 16;
            *reinterpret_cast<volatile double *>(slot + 9) = (double)(t + 1);
        }
    }
}

// ---------------------------------------------------------------------------
// exact global resampling over particle shards (mode 2; resampling.py:599-610
// applied to the concatenation of all shards).  Output j of rank r is global offspring
// index_offset + j: its grid point su is located in the global CDF in two levels -- shard k with
// goff[k] <= su < goff[k+1] (offsets from the exchanged statistics, identical bits on every rank),
// then v = (su - goff[k]) / gpi[k] in shard k's own normalised CDF, read over NVLink -- and the
// ancestor's state is pulled from shard k's particle buffer.  Model-independent: the gathered
// ancestors land in stage_X / stage_lw and the step kernel then runs its streaming branch on them.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int shard_of(const double *goff, const double *gpi, int world, double su) {
    int k = 0;
    while (k + 1 < world && su >= goff[k + 1]) k++;
    while (k > 0 && !(gpi[k] > 0.0)) k--;
    while (k + 1 < world && !(gpi[k] > 0.0)) k++;
    return k;
}

template <int SCHEME>
static __global__ void __launch_bounds__(kBlock, 3) k_resample_global(FilterArgs a, int D) {
    FilterDev *st = a.st;
    if (!st->rs_flag) return;
    constexpr int kStage = 2048;
    __shared__ double s_su[2];
    __shared__ __align__(16) double s_cdf[kStage];
    __shared__ long long s_hi;
    __shared__ double s_goff[9], s_gpi[8];
    const long long t = st->t;
    const int cur = st->cur, world = a.world;
    if ((int)threadIdx.x < world)
        wait_epoch(a.mail_local + ((size_t)(t & 1) * world + threadIdx.x) * 16 + 9, (double)(t + 1), st);
    if ((int)threadIdx.x <= world) s_goff[threadIdx.x] = st->goff[threadIdx.x];
    if ((int)threadIdx.x < world) s_gpi[threadIdx.x] = st->gpi[threadIdx.x];
    __syncthreads();
    const int64_t n = a.n, npairs = n >> 1;                // sharded filters have even n
    const double M_ = (double)a.n_global;
    const double reset_c = st->reset_c;
    const double *uin = a.u_in ? a.u_in + (size_t)t * (n + 1) : nullptr;
    double u_sys = 0.0;
    if (SCHEME == SMCB_RS_SYSTEMATIC) {
        if (uin) u_sys = uin[0];
        else { double u1; uniform_pair(a.key, 0ull, (uint32_t)t, kPurposeUniform, u_sys, u1); }
    }
    const int64_t ntiles = (npairs + kBlock - 1) / kBlock;
    const int64_t per = a.chunk / kBlock;
    const int64_t tile_lo = (int64_t)blockIdx.x * per;
    const int64_t tile_hi = tile_lo + per < ntiles ? tile_lo + per : ntiles;
    int64_t lo = -1;
    int lo_shard = -1;
    for (int64_t tile = tile_lo; tile < tile_hi; tile++) {
        const int64_t p = tile * kBlock + threadIdx.x;
        const int64_t k0 = 2 * tile * kBlock;
        const int64_t k1 = (k0 + 2 * kBlock < n ? k0 + 2 * kBlock : n) - 1;
        double su[2] = {2.0, 2.0};
        if (p < npairs) {
            const double g0 = (double)(a.index_offset + 2 * p);
            if (SCHEME == SMCB_RS_SYSTEMATIC) {
                su[0] = (u_sys + g0) / M_;
                su[1] = (u_sys + (g0 + 1.0)) / M_;
            } else {
                double u0, u1;
                if (uin) { u0 = uin[2 * p]; u1 = uin[2 * p + 1]; }
                else uniform_pair(a.key, (uint64_t)((a.index_offset >> 1) + p), (uint32_t)t, kPurposeUniform, u0, u1);
                su[0] = (u0 + g0) / M_;
                su[1] = (u1 + (g0 + 1.0)) / M_;
            }
            if (2 * p == k0) s_su[0] = su[0];
            if (2 * p + 1 == k1) s_su[1] = su[1];
        }
        __syncthreads();
        const double su_first = s_su[0], su_last = s_su[1];
        const int kf = shard_of(s_goff, s_gpi, world, su_first);
        const int kl = shard_of(s_goff, s_gpi, world, su_last);
        int ks[2] = {kf, kf};
        int64_t an[2] = {0, 0};
        if (kf == kl) {
            // the whole tile draws from one shard: same staged search as the step kernel, on that
            // shard's CDF with the grid points mapped into its local scale
            const double *cdf = a.pcdf[kf];
            const double off = s_goff[kf], pi = s_gpi[kf];
            const double v0 = fmin((su[0] - off) / pi, 1.0), v1 = fmin((su[1] - off) / pi, 1.0);
            const double v_first = fmin((su_first - off) / pi, 1.0), v_last = fmin((su_last - off) / pi, 1.0);
            if (lo < 0 || lo_shard != kf) lo = block_lower_bound<kBlock>(cdf, 0, n, v_first);
            if (lo > n - 1) lo = n - 1;
            lo_shard = kf;
            const int64_t sbase = lo & ~(int64_t)1;
            const int cnt = (int)((n - sbase) < kStage ? (n - sbase) : kStage);
            for (int i = 2 * threadIdx.x; i < cnt; i += 2 * kBlock) {
                if (i + 1 < cnt) *reinterpret_cast<double2 *>(&s_cdf[i]) = ld2(cdf + sbase + i);
                else s_cdf[i] = cdf[sbase + i];
            }
            __syncthreads();
            const bool covered = (sbase + cnt >= n) || (s_cdf[cnt - 1] >= v_last);
            int64_t hi = lo;
            if (!covered) hi = block_lower_bound<kBlock>(cdf, lo, n, v_last);
            if (p < npairs) {
                if (covered) {
                    int l0 = (int)(lo - sbase), h0 = cnt;
                    while (l0 < h0) { const int mid = (l0 + h0) >> 1; if (s_cdf[mid] < v0) l0 = mid + 1; else h0 = mid; }
                    int l1 = l0, h1 = cnt;
#pragma unroll
                    for (int w = 0; w < 4; w++)
                        if (l1 < cnt && s_cdf[l1] < v1) l1++;
                    if (l1 < cnt && s_cdf[l1] < v1) {
                        while (l1 < h1) { const int mid = (l1 + h1) >> 1; if (s_cdf[mid] < v1) l1 = mid + 1; else h1 = mid; }
                    }
                    an[0] = sbase + l0;
                    an[1] = sbase + l1;
                    if (2 * p + 1 == k1) s_hi = an[1];
                } else {
                    const int64_t hi1 = hi < n ? hi + 1 : n;
                    an[0] = lower_bound(cdf, lo, hi1, v0);
                    an[1] = lower_bound(cdf, an[0], hi1, v1);
                }
            }
            __syncthreads();
            lo = covered ? (s_hi < n ? s_hi : n - 1) : (hi < n ? hi : n - 1);
        } else {
            // the tile straddles a shard boundary (at most world - 1 tiles per rank): plain searches
            if (p < npairs) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    ks[j] = shard_of(s_goff, s_gpi, world, su[j]);
                    const double v = fmin((su[j] - s_goff[ks[j]]) / s_gpi[ks[j]], 1.0);
                    an[j] = lower_bound(a.pcdf[ks[j]], 0, n, v);
                }
            }
            lo = -1;
        }
        if (p < npairs) {
            const int64_t a0 = an[0] < n - 1 ? an[0] : n - 1, a1 = an[1] < n - 1 ? an[1] : n - 1;
            const double *X0 = a.pX[ks[0]][cur], *X1 = a.pX[ks[1]][cur];
            for (int c = 0; c < D; c++)
                st2(a.stage_X + (size_t)c * n + 2 * p, __ldg(X0 + (size_t)c * n + a0), __ldg(X1 + (size_t)c * n + a1));
            // ancestors are GLOBAL particle indices
            *reinterpret_cast<longlong2 *>(a.A + 2 * p) = make_longlong2((long long)ks[0] * n + a0, (long long)ks[1] * n + a1);
            st2(a.stage_lw + 2 * p, reset_c, reset_c);
        }
        __syncthreads();       // s_su / s_hi are rewritten by the next tile
    }
}

// multinomial: exponential spacings z = cumsum(-log u), M + 1 of them (resampling.py:536)
struct LoadSpacings {
    Philox key; uint32_t t; const double *u_in;
    __device__ __forceinline__ void operator()(int64_t i0, int64_t n, double (&v)[8]) const {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            double u0, u1;
            if (u_in) {
                u0 = (i0 + j < n) ? u_in[i0 + j] : 1.0;
                u1 = (i0 + j + 1 < n) ? u_in[i0 + j + 1] : 1.0;
            } else {
                uniform_pair(key, (uint64_t)((i0 + j) >> 1), t, kPurposeUniform, u0, u1);
            }
            v[j] = (i0 + j < n) ? -log(u0) : 0.0;       // u may be 0 or injected: library log
            v[j + 1] = (i0 + j + 1 < n) ? -log(u1) : 0.0;
        }
    }
};

static __global__ void __launch_bounds__(kBlock) k_scan_spacings(FilterArgs a) {
    const FilterDev *st = a.st;
    if (!st->rs_flag) return;
    LoadSpacings load{a.key, (uint32_t)st->t,
                      a.u_in ? a.u_in + (size_t)st->t * (a.n + 1) : nullptr};
    scan_tiles_loop<double, LoadSpacings>(load, a.n + 1, a.su, a.scan2);
}

// ---------------------------------------------------------------------------
// the step kernel: resample_move + reweight_particles + compute_summaries
// (core.py:323-367)
// ---------------------------------------------------------------------------
#ifndef SMCB_KU
#define SMCB_KU 2
#endif
#ifndef SMCB_MINB
#define SMCB_MINB 3
#endif
// 1: the CDF slice of a resampling tile is brought in by ONE TMA bulk copy (cp.async.bulk + mbarrier),
//    double-buffered: the slice of tile i+1 is in flight while tile i propagates and reweights.
// 0: all threads stage it with 16-byte loads (kept for A/B timing, profiles/).
#ifndef SMCB_TMA_STAGE
#define SMCB_TMA_STAGE 1
#endif
// MODE 0: both branches, chosen at run time from FilterDev.rs_flag.  MODE 1 / 2 compile the identity /
// resampling branch alone (own register allocation); measured no faster than MODE 0 (DESIGN.md
// section 6), so only MODE 0 is instantiated.
#ifdef SMCB_TRACE
// per-CTA timeline of the LAST launch of the step kernel: {start, main loop done, exit, smid} in ns
__device__ unsigned long long g_trace[4 * 2048];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int smid() {
    unsigned int r;
    asm volatile("mov.u32 %0, %smid;" : "=r"(r));
    return r;
}
#endif
template <class M, int FK, int SCHEME, int MODE = 0>
__global__ void __launch_bounds__(kBlock, (M::D == 1 ? SMCB_MINB : 2)) k_move(M model, FilterArgs a) {
#ifdef SMCB_TRACE
    if (threadIdx.x == 0) { g_trace[4 * blockIdx.x] = gtimer(); g_trace[4 * blockIdx.x + 3] = smid(); }
#endif
    constexpr bool APF = FkTraits<FK>::apf;
    constexpr int K = APF ? 2 : 1;
    constexpr int kStage = (MODE == 1) ? 2 : 2048;     // doubles of CDF staged per output tile
    __shared__ Lse3 smem[kBlock / 32];
    __shared__ double s_su[2];
#if SMCB_TMA_STAGE
    __shared__ __align__(128) double s_cdf2[2][kStage];
    __shared__ __align__(8) uint64_t s_bar[2];
#else
    __shared__ __align__(16) double s_cdf[kStage];
#endif
    __shared__ long long s_hi;
    const FilterDev *st = a.st;
    const long long t = st->t;
    const int cur = st->cur;
    // after a global resampling the ancestors are already gathered (stage_X / stage_lw): stream them
    const bool staged = a.rs_global && (st->rs_flag != 0);
    const bool rs = (MODE == 0) ? (st->rs_flag != 0 && !a.rs_global) : (MODE == 2);
    const double reset_c = st->reset_c;
    const StepK k = step_consts(a, t);
    const StepK kprev = step_consts(a, t - 1);
    const double *__restrict__ Xi = staged ? a.stage_X : a.X[cur];
    const double *__restrict__ lwi = staged ? a.stage_lw : a.lw[cur];
    double *__restrict__ Xo = a.X[cur ^ 1];
    double *__restrict__ lwo = a.lw[cur ^ 1];
    const int64_t n = a.n, npairs = (n + 1) >> 1;
    const double *zin = a.z_in ? a.z_in + (size_t)t * M::NZ * n : nullptr;
    const bool last_apf = APF && (t + 1 < a.T);

    Lse3 acc[K];
    acc[0] = lse3_empty();
    if (APF) acc[K - 1] = lse3_empty();

    // propagate + reweight one pair of particles; writes x', lw'; returns lw' (and the
    // auxiliary log-weights of the next step for an APF), -inf in masked slots
    constexpr int D = M::D, NZ = M::NZ;
    auto do_pair = [&](int64_t p, const double (&xp)[2][D], const double (&base)[2], double *l,
                       double *av) {
        double z[2][NZ], x[2][D];
#pragma unroll
        for (int c = 0; c < NZ; c++) {
            if (zin) {                                   // injected normals: (T, NZ, n)
                const double *zz = zin + (size_t)c * n;
                z[0][c] = zz[2 * p];
                z[1][c] = (2 * p + 1 < n) ? zz[2 * p + 1] : 0.0;
            } else {
                normal_pair_fast(a.key, (uint64_t)((a.index_offset >> 1) + p), (uint32_t)t, (uint32_t)c,
                                 z[0][c], z[1][c]);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) {
            double d;
            model_move<M, FK>(model, k, xp[j], z[j], x[j], d);
            l[j] = fix_nan(base[j] + d);                          // Weights.add, resampling.py:241-244
            if (APF) av[j] = last_apf ? fix_nan(l[j] + model_logeta<M>(model, k, x[j])) : -CUDART_INF;
        }
        if (2 * p + 1 < n) {
#pragma unroll
            for (int c = 0; c < D; c++) st2(Xo + (size_t)c * n + 2 * p, x[0][c], x[1][c]);
            st2(lwo + 2 * p, l[0], l[1]);
        } else {
#pragma unroll
            for (int c = 0; c < D; c++) Xo[(size_t)c * n + 2 * p] = x[0][c];
            lwo[2 * p] = l[0];
            l[1] = -CUDART_INF;
            if (APF) av[1] = -CUDART_INF;
        }
    };

    if (MODE != 2 && !rs) {
        // A = arange(N), Xp = X (core.py:335-336): pure streaming pass, kU pairs in flight per thread
        constexpr int kU = SMCB_KU;
        constexpr int64_t stride = kBlock;
        const int64_t pstart = (int64_t)blockIdx.x * a.chunk;
        const int64_t pend = pstart + a.chunk < npairs ? pstart + a.chunk : npairs;
        for (int64_t p0 = pstart + threadIdx.x; p0 < pend; p0 += kU * stride) {
            double xp[kU][2][D], base[kU][2], l[2 * kU], av[APF ? 2 * kU : 1];
#pragma unroll
            for (int u = 0; u < kU; u++) {                        // all loads first (MLP)
                const int64_t p = p0 + u * stride;
                if (p < pend && 2 * p + 1 < n) {
                    double2 tl = ld2(lwi + 2 * p);
                    base[u][0] = tl.x; base[u][1] = tl.y;
#pragma unroll
                    for (int c = 0; c < D; c++) {
                        double2 tx = ld2(Xi + (size_t)c * n + 2 * p);
                        xp[u][0][c] = tx.x; xp[u][1][c] = tx.y;
                    }
                } else if (p < pend) {
                    base[u][0] = lwi[2 * p]; base[u][1] = 0.0;
#pragma unroll
                    for (int c = 0; c < D; c++) { xp[u][0][c] = Xi[(size_t)c * n + 2 * p]; xp[u][1][c] = 0.0; }
                }
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int64_t p = p0 + u * stride;
                if (p < pend) {
                    do_pair(p, xp[u], base[u], l + 2 * u, APF ? av + 2 * u : av);
                } else {
                    l[2 * u] = l[2 * u + 1] = -CUDART_INF;
                    if (APF) av[APF ? 2 * u : 0] = av[APF ? 2 * u + 1 : 0] = -CUDART_INF;
                }
            }
            lse3_add_batch<2 * kU>(acc[0], l);
            if (APF) lse3_add_batch<(APF ? 2 * kU : 1)>(acc[K - 1], av);
        }
    } else if (MODE != 1) {
        // A = resampling(scheme, aux.W, M=N); Xp = X[A]; reset_weights (core.py:329-333)
        const double M_ = (double)n;
        const double *uin = a.u_in ? a.u_in + (size_t)t * (n + 1) : nullptr;
        double u_sys = 0.0;
        if (SCHEME == SMCB_RS_SYSTEMATIC) {
            if (uin) u_sys = uin[0];
            else { double u1; uniform_pair(a.key, 0ull, (uint32_t)t, kPurposeUniform, u_sys, u1); }
        }
        const double zlast = (SCHEME == SMCB_RS_MULTINOMIAL) ? a.su[n] : 1.0;
        const int64_t ntiles = (npairs + kBlock - 1) / kBlock;
        const int64_t per = a.chunk / kBlock;                  // chunk is a multiple of kBlock
        const int64_t tile_lo = (int64_t)blockIdx.x * per;
        const int64_t tile_hi = tile_lo + per < ntiles ? tile_lo + per : ntiles;
        int64_t lo = -1;
#if SMCB_TMA_STAGE
        if (threadIdx.x == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_init_fence(); }
        __syncthreads();
        uint32_t phase0 = 0, phase1 = 0;
        int buf = 0;
        // thread 0: bring cdf[sb, sb + c) into buffer b (sb = lo_ rounded down to even => 16-byte aligned)
        auto issue = [&](int64_t lo_, int b) {
            const int64_t sb = lo_ & ~(int64_t)1;
            const int c = (int)((n - sb) < kStage ? (n - sb) : kStage);
            const uint32_t bytes = (uint32_t)(c & ~1) * 8u;
            if (c & 1) s_cdf2[b][c - 1] = a.cdf[sb + c - 1];      // odd tail (n odd, end of the array)
            if (bytes) {
                mbar_arrive_expect_tx(&s_bar[b], bytes);
                tma_bulk_g2s(&s_cdf2[b][0], a.cdf + sb, bytes, &s_bar[b]);
            } else {
                mbar_arrive(&s_bar[b]);
            }
        };
#endif
        for (int64_t tile = tile_lo; tile < tile_hi; tile++) {
            const int64_t p = tile * kBlock + threadIdx.x;
            const int64_t k0 = 2 * tile * kBlock;
            const int64_t k1 = (k0 + 2 * kBlock < n ? k0 + 2 * kBlock : n) - 1;
            double su[2] = {2.0, 2.0};
            if (p < npairs) {
                if (SCHEME == SMCB_RS_SYSTEMATIC) {                    // resampling.py:609
                    su[0] = (u_sys + (double)(2 * p)) / M_;
                    su[1] = (u_sys + (double)(2 * p + 1)) / M_;
                } else if (SCHEME == SMCB_RS_STRATIFIED) {             // resampling.py:602
                    double u0, u1;
                    if (uin) { u0 = uin[2 * p]; u1 = (2 * p + 1 < n) ? uin[2 * p + 1] : 0.0; }
                    else uniform_pair(a.key, (uint64_t)((a.index_offset >> 1) + p), (uint32_t)t,
                                      kPurposeUniform, u0, u1);
                    su[0] = (u0 + (double)(2 * p)) / M_;
                    su[1] = (u1 + (double)(2 * p + 1)) / M_;
                } else {                                               // resampling.py:537
                    su[0] = a.su[2 * p] / zlast;
                    su[1] = (2 * p + 1 < n) ? a.su[2 * p + 1] / zlast : 2.0;
                }
                if (2 * p == k0) s_su[0] = su[0];
                if (2 * p == k1) s_su[1] = su[0];
                if (2 * p + 1 == k1) s_su[1] = su[1];
            }
            __syncthreads();
            const double su_first = s_su[0], su_last = s_su[1];
#if SMCB_TMA_STAGE
            if (lo < 0) {                                      // first tile of this block
                lo = block_lower_bound<kBlock>(a.cdf, 0, n, su_first);
                lo = lo < n - 1 ? lo : n - 1;
                if (threadIdx.x == 0) issue(lo, buf);
                __syncthreads();
            }
            // the slice of the CDF this tile's outputs fall into was requested while the previous tile
            // was propagating (TMA bulk copy, 16 KB); su is sorted, so the slice starts at `lo`
            const int64_t sbase = lo & ~(int64_t)1;
            const int cnt = (int)((n - sbase) < kStage ? (n - sbase) : kStage);
            mbar_wait(&s_bar[buf], buf ? phase1 : phase0);
            if (buf) phase1 ^= 1u; else phase0 ^= 1u;
            const double *s_cdf = s_cdf2[buf];
#else
            if (lo < 0) lo = block_lower_bound<kBlock>(a.cdf, 0, n, su_first);
            // stage the slice of the CDF this tile's outputs fall into (16 KB, coalesced 16-byte
            // loads) and search it in shared memory; su is sorted, so the slice starts at `lo`
            const int64_t sbase = lo & ~(int64_t)1;
            const int cnt = (int)((n - sbase) < kStage ? (n - sbase) : kStage);
            for (int i = 2 * threadIdx.x; i < cnt; i += 2 * kBlock) {
                if (i + 1 < cnt) *reinterpret_cast<double2 *>(&s_cdf[i]) = ld2(a.cdf + sbase + i);
                else s_cdf[i] = a.cdf[sbase + i];
            }
            __syncthreads();
#endif
            const bool covered = (sbase + cnt >= n) || (s_cdf[cnt - 1] >= su_last);
            int64_t hi = lo;
            if (!covered) hi = block_lower_bound<kBlock>(a.cdf, lo, n, su_last);   // rare: sparse mass
            int64_t a0 = 0, a1 = 0;
            if (p < npairs) {
                if (covered) {
                    int l0 = (int)(lo - sbase), h0 = cnt;
                    while (l0 < h0) { const int mid = (l0 + h0) >> 1; if (s_cdf[mid] < su[0]) l0 = mid + 1; else h0 = mid; }
                    // su[1] >= su[0] and on average one CDF entry per output: walk forward a few
                    // entries before falling back to bisection (same result as searchsorted)
                    int l1 = l0, h1 = cnt;
#pragma unroll
                    for (int w = 0; w < 4; w++)
                        if (l1 < cnt && s_cdf[l1] < su[1]) l1++;
                    if (l1 < cnt && s_cdf[l1] < su[1]) {
                        while (l1 < h1) { const int mid = (l1 + h1) >> 1; if (s_cdf[mid] < su[1]) l1 = mid + 1; else h1 = mid; }
                    }
                    a0 = sbase + l0;
                    a1 = sbase + l1;
                    if (2 * p == k1) s_hi = a0;
                    if (2 * p + 1 == k1) s_hi = a1;
                } else {
                    const int64_t hi1 = hi < n ? hi + 1 : n;
                    a0 = lower_bound(a.cdf, lo, hi1, su[0]);
                    a1 = lower_bound(a.cdf, a0, hi1, su[1]);
                }
            }
#if SMCB_TMA_STAGE
            __syncthreads();                                   // s_hi published; buffer buf^1 is free
            const int64_t lo_next = covered ? (s_hi < n ? s_hi : n - 1) : (hi < n ? hi : n - 1);
            if (tile + 1 < tile_hi && threadIdx.x == 0) issue(lo_next, buf ^ 1);
#endif
            if (p < npairs) {
                a0 = a0 < n - 1 ? a0 : n - 1;
                a1 = a1 < n - 1 ? a1 : n - 1;
                double xp[2][D], base[2];
#pragma unroll
                for (int c = 0; c < D; c++) {                // Xp = X[A], component-wise (SoA)
                    xp[0][c] = __ldg(Xi + (size_t)c * n + a0);
                    xp[1][c] = __ldg(Xi + (size_t)c * n + a1);
                }
                if (APF) {   // core.py:302: lw = log_mean_exp(logetat, W) - logetat[A]
                    base[0] = reset_c - model_logeta<M>(model, kprev, xp[0]);
                    base[1] = reset_c - model_logeta<M>(model, kprev, xp[1]);
                } else {     // Weights() then add(delta): lw = 0 + delta (shard mass if sharded)
                    base[0] = reset_c; base[1] = reset_c;
                }
                if (2 * p + 1 < n) *reinterpret_cast<longlong2 *>(a.A + 2 * p) = make_longlong2(a0, a1);
                else a.A[2 * p] = a0;
                double l[2], av[2];
                do_pair(p, xp, base, l, av);
                lse3_add_batch<2>(acc[0], l);
                if (APF) lse3_add_batch<2>(acc[K - 1], av);
            }
#if SMCB_TMA_STAGE
            lo = lo_next;
            buf ^= 1;
#else
            __syncthreads();
            lo = covered ? (s_hi < n ? s_hi : n - 1) : hi;
#endif
        }
    }

    Lse3 tot[K];
#ifdef SMCB_TRACE
    if (threadIdx.x == 0) g_trace[4 * blockIdx.x + 1] = gtimer();
    const bool last_ = grid_merge_lse3<kBlock, K>(acc, a.partials, a.ticket, smem, tot);
    if (!last_) { if (threadIdx.x == 0) g_trace[4 * blockIdx.x + 2] = gtimer(); return; }
    finalize_step<APF>(a, tot[0], tot[K - 1], tot[0], tot[K - 1]);
    __syncthreads();
    if (threadIdx.x == 0) g_trace[4 * blockIdx.x + 2] = gtimer();
    return;
#endif
    if (!grid_merge_lse3<kBlock, K>(acc, a.partials, a.ticket, smem, tot)) return;
    if (a.world > 1) { publish_local<K>(a, tot); return; }
    finalize_step<APF>(a, tot[0], tot[K - 1], tot[0], tot[K - 1]);
}

// sharded runs: after the all-gather of the per-rank statistics, one block per rank merges
// them in rank order (identical bits on every rank) and runs the same finalize as above
template <bool APF>
__global__ void __launch_bounds__(kBlock) k_finish(FilterArgs a) {
    __shared__ Lse3 s_tot[4];
    const double *gath = a.gathered;
    if (a.mail_local != nullptr) {          // peer-memory exchange: wait for every sender's epoch
        const long long t = a.st->t;
        const double *box = a.mail_local + (size_t)(t & 1) * a.world * 16;
        if ((int)threadIdx.x < a.world) {
            const volatile double *flag = box + (size_t)threadIdx.x * 16 + 8;
            while (*flag != (double)(t + 1)) { }
            __threadfence_system();
        }
        __syncthreads();
        gath = box;                         // stride 16 doubles per sender
    }
    const int gstride = (a.mail_local != nullptr) ? 16 : 8;
    if (threadIdx.x == 0) {
        Lse3 w = lse3_empty(), x = lse3_empty();
        for (int r = 0; r < a.world; r++) {
            const volatile double *g = gath + (size_t)r * gstride;
            w = lse3_merge(w, Lse3{g[0], g[1], g[2]});
            x = lse3_merge(x, Lse3{g[4], g[5], g[6]});
        }
        if (a.rs_global) {      // shard r's share of the global (auxiliary) weight mass, rank order
            double run = 0.0;
            for (int r = 0; r < a.world; r++) {
                const volatile double *g = gath + (size_t)r * gstride;
                const double pi = (g[4] == -CUDART_INF) ? 0.0 : g[5] * fexp(g[4] - x.m) / x.s;
                a.st->goff[r] = run;
                a.st->gpi[r] = pi;
                run = run + pi;
            }
            a.st->goff[a.world] = run;
        }
        const volatile double *me = gath + (size_t)a.rank * gstride;
        s_tot[0] = w; s_tot[1] = x;
        s_tot[2] = Lse3{me[0], me[1], me[2]};
        s_tot[3] = Lse3{me[4], me[5], me[6]};
    }
    __syncthreads();
    finalize_step<APF>(a, s_tot[0], APF ? s_tot[1] : s_tot[0], s_tot[2], APF ? s_tot[3] : s_tot[2]);
}

// graph mode: first node of every WHILE iteration.  Publishes the two conditions of the iteration
// about to run: IF handle = resample or not (decided by the previous step's finalize), WHILE
// handle = whether another iteration follows this one.
static __global__ void k_cond(const FilterDev *st, cudaGraphConditionalHandle h_if,
                              cudaGraphConditionalHandle h_while) {
    if (threadIdx.x == 0) {
        cudaGraphSetConditional(h_if, st->rs_flag ? 1u : 0u);
        cudaGraphSetConditional(h_while, (st->t + 1 < st->t_stop) ? 1u : 0u);
    }
}

}  // namespace smcb

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct smcb_filter {
    smcb_ctx *ctx;
    smcb_filter_desc desc;
    FilterArgs args;
    FilterDev *st;
    double *sc_dev;
    void *scan_mem;
    int grid_move, grid_scan2;
    int blocks_per_sm;    // resident CTAs/SM of the step kernel (persistent grid = SMs x this)
    int64_t t_host;       // host mirror of FilterDev.t (one launch sequence per step)
    cudaEvent_t *timed_ev; // non-NULL inside smcb_filter_step_timed: event pairs per launch
    int *timed_kind;
    int (*launch_init)(smcb_filter *);
    int (*launch_step)(smcb_filter *);
    int (*launch_finish)(smcb_filter *);
    int (*build_graph)(smcb_filter *);
    // one CUDA graph for the whole step loop: WHILE(t < t_stop) { k_cond; IF(rs) {scan; move} ELSE {move} }
    cudaGraph_t graph;
    cudaGraphExec_t gexec;
    bool has_graph;
    long long t_stop_host;
};

template <class M, int FK, int SCHEME>
static int launch_step_t(smcb_filter *f) {
    M model;
    model.load(f->desc.params);
    cudaStream_t s = f->ctx->stream;
    cudaEvent_t *ev = f->timed_ev;
    int j = 0;
    auto before = [&]() { if (ev) cudaEventRecord(ev[2 * j], s); };
    auto after = [&](int kind) { if (ev) { cudaEventRecord(ev[2 * j + 1], s); f->timed_kind[j] = kind; j++; } };
    before();
    k_scan_w<M, FK><<<f->grid_move, kBlock, 0, s>>>(model, f->args);
    after(1);
    f->ctx->launches++;
    if (SCHEME == SMCB_RS_MULTINOMIAL) {
        before();
        k_scan_spacings<<<f->grid_scan2, kBlock, 0, s>>>(f->args);
        after(2);
        f->ctx->launches++;
    }
    if (f->args.rs_global) {
        if (SCHEME == SMCB_RS_STRATIFIED) k_resample_global<SMCB_RS_STRATIFIED><<<f->grid_move, kBlock, 0, s>>>(f->args, M::D);
        else k_resample_global<SMCB_RS_SYSTEMATIC><<<f->grid_move, kBlock, 0, s>>>(f->args, M::D);
        f->ctx->launches++;
    }
    before();
    k_move<M, FK, SCHEME><<<f->grid_move, kBlock, 0, s>>>(model, f->args);
    after(3);
    f->ctx->launches++;
    SMCB_CUDA(cudaGetLastError());
    return SMCB_OK;
}

// The whole step loop as ONE CUDA graph (conditional nodes, CUDA >= 12.4):
//   WHILE (h_while) { k_cond -> IF (h_if) { k_scan_w [-> k_scan_spacings] -> k_move<rs> } ELSE { k_move<id> } }
// The resample / no-resample decision never leaves the device and no launch is issued per step.
template <class M, int FK, int SCHEME>
static int build_graph_t(smcb_filter *f) {
    M model;
    model.load(f->desc.params);
    FilterArgs args = f->args;
    cudaGraph_t g = nullptr, body = nullptr;
    SMCB_CUDA(cudaGraphCreate(&g, 0));
    cudaGraphConditionalHandle h_while, h_if;
    SMCB_CUDA(cudaGraphConditionalHandleCreate(&h_while, g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams wp = {};
    wp.type = cudaGraphNodeTypeConditional;
    wp.conditional.handle = h_while;
    wp.conditional.type = cudaGraphCondTypeWhile;
    wp.conditional.size = 1;
    cudaGraphNode_t wnode;
    SMCB_CUDA(cudaGraphAddNode(&wnode, g, nullptr, 0, &wp));
    body = wp.conditional.phGraph_out[0];
    SMCB_CUDA(cudaGraphConditionalHandleCreate(&h_if, body, 0, 0));
    // k_cond
    cudaGraphNode_t ncond;
    {
        const FilterDev *stp = f->st;
        void *ka[] = {(void *)&stp, (void *)&h_if, (void *)&h_while};
        cudaKernelNodeParams kp = {};
        kp.func = (void *)k_cond;
        kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = ka;
        SMCB_CUDA(cudaGraphAddKernelNode(&ncond, body, nullptr, 0, &kp));
    }
    cudaGraphNodeParams ip = {};
    ip.type = cudaGraphNodeTypeConditional;
    ip.conditional.handle = h_if;
    ip.conditional.type = cudaGraphCondTypeIf;
    ip.conditional.size = 2;
    cudaGraphNode_t inode;
    SMCB_CUDA(cudaGraphAddNode(&inode, body, &ncond, 1, &ip));
    cudaGraph_t g_rs = ip.conditional.phGraph_out[0], g_id = ip.conditional.phGraph_out[1];
    void *ka2[] = {(void *)&model, (void *)&args};
    void *ka1[] = {(void *)&args};
    cudaGraphNode_t prev, cur;
    {   // IF body: weight scan [-> spacings scan] -> search / gather / move
        cudaKernelNodeParams kp = {};
        kp.func = (void *)k_scan_w<M, FK>;
        kp.gridDim = dim3(f->grid_move); kp.blockDim = dim3(kBlock); kp.kernelParams = ka2;
        SMCB_CUDA(cudaGraphAddKernelNode(&prev, g_rs, nullptr, 0, &kp));
        if (SCHEME == SMCB_RS_MULTINOMIAL) {
            cudaKernelNodeParams k2 = {};
            k2.func = (void *)k_scan_spacings;
            k2.gridDim = dim3(f->grid_scan2); k2.blockDim = dim3(kBlock); k2.kernelParams = ka1;
            SMCB_CUDA(cudaGraphAddKernelNode(&cur, g_rs, &prev, 1, &k2));
            prev = cur;
        }
        cudaKernelNodeParams k3 = {};
        k3.func = (void *)k_move<M, FK, SCHEME, 0>;
        k3.gridDim = dim3(f->grid_move); k3.blockDim = dim3(kBlock); k3.kernelParams = ka2;
        SMCB_CUDA(cudaGraphAddKernelNode(&cur, g_rs, &prev, 1, &k3));
    }
    {   // ELSE body: the streaming step
        cudaKernelNodeParams k4 = {};
        k4.func = (void *)k_move<M, FK, SCHEME, 0>;
        k4.gridDim = dim3(f->grid_move); k4.blockDim = dim3(kBlock); k4.kernelParams = ka2;
        SMCB_CUDA(cudaGraphAddKernelNode(&cur, g_id, nullptr, 0, &k4));
    }
    cudaGraphExec_t ex = nullptr;
    SMCB_CUDA(cudaGraphInstantiate(&ex, g, 0));
    f->graph = g;
    f->gexec = ex;
    f->has_graph = true;
    return SMCB_OK;
}

template <class M, int FK>
static int launch_init_t(smcb_filter *f) {
    M model;
    model.load(f->desc.params);
    k_init<M, FK><<<f->grid_move, kBlock, 0, f->ctx->stream>>>(model, f->args);
    f->ctx->launches++;
    SMCB_CUDA(cudaGetLastError());
    return SMCB_OK;
}

template <int FK>
static int launch_finish_t(smcb_filter *f) {
    k_finish<FkTraits<FK>::apf><<<1, kBlock, 0, f->ctx->stream>>>(f->args);
    f->ctx->launches++;
    SMCB_CUDA(cudaGetLastError());
    return SMCB_OK;
}

template <class M, int FK, int SCHEME>
static int bind_one(smcb_filter *f) {
    f->launch_init = launch_init_t<M, FK>;
    f->launch_step = launch_step_t<M, FK, SCHEME>;
    f->launch_finish = launch_finish_t<FK>;
    f->build_graph = build_graph_t<M, FK, SCHEME>;
    int nb = 0;
    SMCB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_move<M, FK, SCHEME>, kBlock, 0));
    f->blocks_per_sm = nb < 1 ? 1 : nb;
    return SMCB_OK;
}

template <class M, int FK>
static int bind_scheme(smcb_filter *f) {
    switch (f->desc.scheme) {
        case SMCB_RS_SYSTEMATIC: return bind_one<M, FK, SMCB_RS_SYSTEMATIC>(f);
        case SMCB_RS_STRATIFIED: return bind_one<M, FK, SMCB_RS_STRATIFIED>(f);
        case SMCB_RS_MULTINOMIAL: return bind_one<M, FK, SMCB_RS_MULTINOMIAL>(f);
        default:
            set_error("fused filter: resampling scheme %d is not fused (use systematic, stratified or "
                      "multinomial, or the unfused path)", f->desc.scheme);
            return SMCB_ENOSYS;
    }
}

template <class M>
static int bind_fk(smcb_filter *f) {
    switch (f->desc.fk) {
        case SMCB_FK_BOOTSTRAP: return bind_scheme<M, SMCB_FK_BOOTSTRAP>(f);
        case SMCB_FK_GUIDED:
            if (!M::has_proposal) break;
            return bind_scheme<M, SMCB_FK_GUIDED>(f);
        case SMCB_FK_APF:
            if (!M::has_proposal) break;
            return bind_scheme<M, SMCB_FK_APF>(f);
        case SMCB_FK_AUXBOOT:
            if (!M::has_proposal) break;
            return bind_scheme<M, SMCB_FK_AUXBOOT>(f);
        default:
            set_error("fused filter: unknown Feynman-Kac kind %d", f->desc.fk);
            return SMCB_EINVAL;
    }
    // the reference raises NotImplementedError from StateSpaceModel.proposal / logeta
    set_error("fused filter: model %d implements no proposal/logeta (Feynman-Kac kind %d)",
              f->desc.model, f->desc.fk);
    return SMCB_ENOSYS;
}

