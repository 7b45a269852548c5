/* smcb.h -- C-ABI of libsmcb.so: the B200 (sm_100a) SMC inner loop.
 *
 * Drop-in boundary for the per-step hot path of nchopin/particles
 * (particles.core.SMC: propagate -> log-weight -> normalise/ESS -> resample).
 * The reference is pure Python with no FFI of its own; each entry point below
 * names the reference function it replaces (paths relative to the reference
 * root), and INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - every array argument is a DEVICE pointer (fp64 / int64, contiguous) owned by
 *    the caller (torch tensors in the Python host layer); scalars results are
 *    written to device memory too, so no call synchronises the host;
 *  - work is enqueued on the context's stream (smcb_set_stream);
 *  - every function returns 0 on success, a negative SMCB_E* code otherwise,
 *    with a message available from smcb_last_error();
 *  - a context is bound to one device and is not thread-safe;
 *  - results are deterministic: same inputs + same seed -> same bits, whatever
 *    the grid size (all reductions and the scan use a fixed association order).
 */
#ifndef SMCB_H
#define SMCB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMCB_OK 0
#define SMCB_EINVAL (-1) /* bad argument (ValueError on the Python side)        */
#define SMCB_ECUDA (-2)  /* CUDA runtime error                                   */
#define SMCB_ENOSYS (-3) /* combination not implemented (NotImplementedError)    */

typedef struct smcb_ctx smcb_ctx;

const char *smcb_last_error(void);
int smcb_version(void);

/* one context per (device, stream); owns a small workspace + the Philox key */
int smcb_create(smcb_ctx **out, int device, uint64_t seed);
int smcb_destroy(smcb_ctx *ctx);
int smcb_set_stream(smcb_ctx *ctx, void *cuda_stream);
/* re-key the counter-based generator (replaces numpy.random.seed for this path) */
int smcb_seed(smcb_ctx *ctx, uint64_t seed);
/* number of kernels this context has launched so far (bench.py "gpu_launches") */
int64_t smcb_launch_count(const smcb_ctx *ctx);

/* ---------------------------------------------------------------------------
 * weights algebra  (particles/resampling.py)
 * ------------------------------------------------------------------------- */

/* Weights.__init__, resampling.py:217-226.  lw is modified in place (NaN -> -inf,
 * line 220).  W_out may be NULL.  stats_out[4] = {max lw, log_mean, ESS, sum w}. */
int smcb_normalise(smcb_ctx *ctx, double *lw, int64_t n, double *W_out, double *stats_out);
/* The second half alone, W = exp(lw - m) / s (resampling.py:223-225) with statistics the caller already holds
 * (stats[0] = m, stats[3] = s: the layout above; the fused filter's device state; for a sharded filter the GLOBAL
 * (m, s), so that W sums to one over all ranks). */
int smcb_weights_from_stats(smcb_ctx *ctx, const double *lw, int64_t n, const double *stats, double *W_out);

#define SMCB_LSE_SUM 0  /* log_sum_exp   resampling.py:247-270            */
#define SMCB_LSE_MEAN 1 /* log_mean_exp  resampling.py:291-317 (W optional) */
#define SMCB_LSE_ESSL 2 /* essl          resampling.py:166-188            */
int smcb_lse(smcb_ctx *ctx, int mode, const double *v, const double *W, int64_t n,
             double *out);

/* exp_and_normalise, resampling.py:138-163 */
int smcb_exp_and_normalise(smcb_ctx *ctx, const double *lw, int64_t n, double *W_out);

/* wmean_and_var, resampling.py:320-338; x is SoA (d, n); out = {mean[d], var[d]} */
int smcb_wmean_and_var(smcb_ctx *ctx, const double *W, const double *x, int64_t n, int d,
                       double *out);

/* ---------------------------------------------------------------------------
 * resampling  (particles/resampling.py)
 * ------------------------------------------------------------------------- */
#define SMCB_RS_MULTINOMIAL 0 /* resampling.py:540-558 */
#define SMCB_RS_STRATIFIED 1  /* resampling.py:599-603 */
#define SMCB_RS_SYSTEMATIC 2  /* resampling.py:606-610 */
#define SMCB_RS_RESIDUAL 3    /* resampling.py:613-627 */
#define SMCB_RS_SSP 4         /* resampling.py:630-677; sequential recursion, one device thread;
                                 u_in = N - 1 uniforms; synchronises once (ValueError check) */

/* Inclusive prefix sum of non-negative fp64 values (the CDF that inverse_cdf,
 * resampling.py:484-509, walks).  Single pass, decoupled look-back with a
 * fixed association order: deterministic and non-decreasing by construction. */
int smcb_cumsum(smcb_ctx *ctx, const double *w, int64_t n, double *cdf_out);

/* A[k] = min{ j : cdf[j] >= su[k] } clipped to n-1, for sorted su;
 * == np.searchsorted(cdf, su, 'left') bit-exactly (inverse_cdf, resampling.py:484-509) */
int smcb_searchsorted(smcb_ctx *ctx, const double *cdf, int64_t n, const double *su,
                      int64_t m, int64_t *A_out);

/* rs.resampling(scheme, W, M), resampling.py:464-481, 540-627.
 * u_in (device) = the uniforms to use, in the order the reference draws them
 * (systematic 1, stratified m, multinomial m+1, residual m+1, ssp n-1); NULL -> Philox.
 * scratch: at least smcb_resample_scratch_doubles(n, m) doubles (16-byte aligned).
 * Asynchronous on the context's stream except SMCB_RS_SSP, which synchronises once to report the
 * reference's "wrong size for output" ValueError (resampling.py:674-676) as SMCB_EINVAL. */
int64_t smcb_resample_scratch_doubles(int64_t n, int64_t m);
int smcb_resample(smcb_ctx *ctx, int scheme, const double *W, int64_t n, int64_t m,
                  int64_t *A_out, const double *u_in, double *scratch);

/* Xp = X[A]  (core.py:332); X is SoA (d, n), Xp is SoA (d, m) */
int smcb_gather(smcb_ctx *ctx, const double *X, int64_t n, const int64_t *A, int64_t m,
                int d, double *Xp);
/* same for row-major (n, d) particles, the layout user closures index as xp[:, i] */
int smcb_gather_rows(smcb_ctx *ctx, const double *X, int64_t n, const int64_t *A, int64_t m,
                     int d, double *Xp);

/* ---------------------------------------------------------------------------
 * distributions  (particles/distributions.py)
 * an array argument may be NULL, in which case the scalar next to it is used
 * ------------------------------------------------------------------------- */
/* Normal.rvs, distributions.py:270-271; z_in = injected N(0,1) draws or NULL */
int smcb_normal_rvs(smcb_ctx *ctx, const double *loc, double loc0, const double *scale,
                    double scale0, const double *z_in, double *out, int64_t n);
/* Normal.logpdf, distributions.py:273-274 */
int smcb_normal_logpdf(smcb_ctx *ctx, const double *x, double x0, const double *loc,
                       double loc0, const double *scale, double scale0, double *out,
                       int64_t n);
/* other univariate log-densities (distributions.py): kind 0 Student(df = p0, loc, scale) :417-433, 1 Gamma(a = p0,
 * rate) :336-356, 2 Laplace(loc, scale) :399-414, 3 Logistic(loc, scale) :381-396; p1 / p2 arrays or scalars as for
 * Normal; c0 = the host-computed lgamma constant of the density (0 for kinds 2, 3) */
int smcb_logpdf1(smcb_ctx *ctx, int kind, const double *x, double x0, double p0, double c0, const double *p1,
                 double p10, const double *p2, double p20, double *out, int64_t n);
/* MvNormal.rvs / logpdf, distributions.py:946-969; SoA (d, n), d <= 32; L = host (d,d) lower
 * Cholesky factor of cov, row-major; loc/scale: array (d, n), or NULL + host vector[d].
 * d <= 8: factor in the kernel parameters, two particles per thread; 8 < d <= 32: factor in shared
 * memory, one particle per thread (CUDA cores: HBM-bound at every d <= 32, see smcb_api.cu) */
int smcb_mvnormal_rvs(smcb_ctx *ctx, const double *loc, const double *loc0,
                      const double *scale, const double *scale0, const double *L, int d,
                      const double *z_in, double *out, int64_t n);
int smcb_mvnormal_logpdf(smcb_ctx *ctx, const double *x, const double *loc,
                         const double *loc0, const double *scale, const double *scale0,
                         const double *L, int d, double *out, int64_t n);
/* standard normals / uniforms from the context's Philox stream */
int smcb_standard_normal(smcb_ctx *ctx, double *out, int64_t n);
int smcb_uniform(smcb_ctx *ctx, double *out, int64_t n);

/* ---------------------------------------------------------------------------
 * SMC samplers: tempering / waste-free move step  (particles/smc_samplers.py)
 * theta is (n, d) row-major, d <= 32
 * ------------------------------------------------------------------------- */
/* Tempering.current_target, smc_samplers.py:836-845, for the logistic-regression static model
 * (book/smc_samplers/logistic_reg.py:60-67): prior MvNormal(0, prior_scale^2 I_d);
 * llik = sum_t -log(1 + exp(-theta . data[t])) (StaticModel.loglik, 263-284); lpost = lprior + epn*llik */
int smcb_logistic_target(smcb_ctx *ctx, const double *theta, int64_t n, int d, const double *data,
                         int64_t n_data, double prior_scale, double epn, double *lprior,
                         double *llik, double *lpost);
/* ArrayRandomWalk.proposal, smc_samplers.py:624-629: prop = theta + z @ L.T; L_dev = device (d, d)
 * row-major lower factor; z_in = injected N(0,1) (n, d) or NULL */
int smcb_rw_propose(smcb_ctx *ctx, const double *theta, int64_t n, int d, const double *L_dev,
                    const double *z_in, double *prop);
/* ArrayMetropolis.step, smc_samplers.py:601-611: accept where u < exp(min(lpost' - lpost, 0)) and
 * copy the proposal's fields in place; mean_acc (device scalar) = mean acceptance probability */
int smcb_mh_accept(smcb_ctx *ctx, int64_t n, int d, double *theta, double *lprior, double *llik,
                   double *lpost, const double *theta_p, const double *lprior_p, const double *llik_p,
                   const double *lpost_p, const double *u_in, double *mean_acc);

/* MCMCSequenceWF.__call__, smc_samplers.py:672-683, fused for the logistic model + random-walk
 * Metropolis: ONE launch runs the P-1 Metropolis steps of all M chains.  Inputs: the M resampled
 * particles; outputs: the P*M particles of the next generation in concatenate(xs) order and the
 * (P-1, M) acceptance probabilities.  z_in (P-1, M, d) / u_in (P-1, M): injected noise or NULL. */
int smcb_logistic_wf_move(smcb_ctx *ctx, int64_t M, int d, int P, const double *theta0,
                          const double *lprior0, const double *llik0, const double *lpost0,
                          const double *data, int64_t n_data, double prior_scale, double epn,
                          const double *L_dev, const double *z_in, const double *u_in,
                          double *theta_out, double *lprior_out, double *llik_out, double *lpost_out,
                          double *pb_out);

/* AdaptiveTempering's control plane on the device (no host round trip inside a tempering step):
 * next_annealing_epn, smc_samplers.py:876-895: the exponent at which ESS(delta * llik) = alpha * n, by an 11-pass
 * 16-way bracketing search whose state stays in device memory; out_dev[0] = new exponent (1.0 if the whole step fits) */
int smcb_next_annealing_epn(smcb_ctx *ctx, const double *llik, int64_t n, double epn, double alpha, double *out_dev);
/* ArrayRandomWalk.calibrate, smc_samplers.py:617-622 (rs.wmean_and_cov, resampling.py:341-358): L_out (d, d)
 * row-major = scale * chol(weighted covariance of the rows of theta), d <= 20 */
int smcb_rw_calibrate(smcb_ctx *ctx, const double *W, const double *theta, int64_t n, int d, double scale,
                      double *L_out);

/* the same control plane in pieces, for a tempering run sharded over ranks (the host layer all-reduces between them):
 * raw ESS sums of one root-find pass; raw weighted-moment sums; the factor from all-reduced sums */
int smcb_essl_grid(smcb_ctx *ctx, const double *llik, int64_t n, double lo, double hi, const double *max_dev,
                   double *out32_dev);
int smcb_wcov_sums(smcb_ctx *ctx, const double *W, const double *theta, int64_t n, int d, const double *mean_dev,
                   double *out_dev);
int smcb_chol_from_sums(smcb_ctx *ctx, const double *tri_dev, const double *sw_dev, int d, double scale, double *L_out);

/* test hook: the kernels' own fp64 exp / log / sincos (csrc/smcb_math.cuh) on an array;
 * fn: 0 exp, 1 log (x > 0, normal), 2 sin(2 pi x), 3 cos(2 pi x), x in [0, 1)   -- polynomial family;
 *     4 exp, 5 log, 6 sin(2 pi x), 7 cos(2 pi x), 8 sqrt (x > 0, normal)        -- table family (step kernels) */
int smcb_device_math(smcb_ctx *ctx, int fn, const double *x, double *out, int64_t n);

/* ---------------------------------------------------------------------------
 * fused filter: the whole step of core.py:369-383 for a recognised model
 * ------------------------------------------------------------------------- */
#define SMCB_FK_BOOTSTRAP 0 /* state_space_models.py:299-349 */
#define SMCB_FK_GUIDED 1    /* state_space_models.py:352-398 */
#define SMCB_FK_APF 2       /* state_space_models.py:406-428 */
#define SMCB_FK_AUXBOOT 3   /* state_space_models.py:431-438 */

#define SMCB_MODEL_STOCHVOL 0      /* state_space_models.py:446-498             */
#define SMCB_MODEL_LINGAUSS 1      /* kalman.py:397-452 (also README ToySSM)    */
#define SMCB_MODEL_GORDON 2        /* state_space_models.py:546-577             */
#define SMCB_MODEL_THETALOGISTIC 3 /* state_space_models.py:657-689             */
#define SMCB_MODEL_BEARINGS 4      /* state_space_models.py:580-608 (d = 4)     */
#define SMCB_MODEL_MVLINGAUSS 5    /* kalman.py:296-394 (d <= 8)                */
#define SMCB_MODEL_DISCRETECOX 6   /* state_space_models.py:611-630             */
#define SMCB_MODEL_STOCHVOLLEV 7   /* state_space_models.py:501-543             */

#define SMCB_MAX_PARAMS 256
#define SMCB_SUMMARY_STRIDE 4 /* per step: ESS, logLt, rs_flag, log_mean_w */

typedef struct smcb_filter smcb_filter;

typedef struct {
    int32_t model, fk, scheme, dim;  /* dim = state dimension d                     */
    int32_t dy, n_params;
    int32_t world, rank;             /* particle shards over `world` GPUs (0/1 = one device) */
    int64_t n;                       /* particles on this device                    */
    int64_t n_global;                /* particles over all ranks (== n if 1 GPU)    */
    int64_t index_offset;            /* global index of local particle 0 (Philox)   */
    int64_t T;                       /* number of data points                       */
    double essrmin;
    uint64_t seed;
    double params[SMCB_MAX_PARAMS];  /* model constants, layout per model */
    /* device buffers, all caller-owned */
    double *X[2];      /* ping-pong state, SoA (d, n)                               */
    double *lw[2];     /* ping-pong log-weights (n)                                 */
    int64_t *A;        /* ancestors of the last resampling step (n)                 */
    double *cdf;       /* (n) scratch: CDF of the last resampling step              */
    double *data;      /* (T, dy) observations                                      */
    double *summaries; /* (T, SMCB_SUMMARY_STRIDE)                                  */
    const double *z_in; /* NULL, or injected N(0,1): (T, n_noise, n)                */
    const double *u_in; /* NULL, or injected uniforms: (T, n + 1)                   */
    double *scratch;    /* NULL, or n + 2 doubles (multinomial: exponential spacings) */
    const double *step_consts; /* NULL, or (T) host-computed per-step model constants */
    double *local_stats;       /* world > 1: 16 doubles, this rank's weight statistics    */
    const double *gathered;    /* world > 1: world x 16 doubles, filled by the all-gather */
    double *mail_local;        /* world > 1, optional: this rank's peer mailbox (smcb_p2p_alloc,
                                  2 * world * 32 doubles) -> statistics are exchanged by the
                                  kernels themselves over NVLink and smcb_filter_step works */
    double *mail_peer[8];      /* every rank's mailbox as mapped in THIS process (own = mail_local) */
    /* world > 1, optional: EXACT global resampling (mode 2).  X[0], X[1] and cdf
       of every rank live in peer-mapped memory (smcb_p2p_alloc); on a resampling step each rank
       searches the global CDF (shard offsets from the exchanged statistics + the owning shard's
       local CDF) and pulls the selected ancestors over NVLink.  Needs mail_local; systematic or
       stratified. */
    int32_t rs_global, reserved0;
    double *moments;           /* NULL, or (T, 8): per step the weighted mean [0..3] and variance [4..7] of the
                                  state components (collectors.Moments with the default wmean_and_var,
                                  collectors.py:301-317, resampling.py:320-338), accumulated by the step kernel */
    double *reserved1;
    const double *peer_X0[8];  /* rank r's X[0] / X[1] / cdf as mapped in THIS process    */
    const double *peer_X1[8];
    const double *peer_cdf[8];
} smcb_filter_desc;

int smcb_filter_create(smcb_ctx *ctx, const smcb_filter_desc *desc, smcb_filter **out);
int smcb_filter_destroy(smcb_filter *f);
/* enqueue nsteps steps of SMC.__next__ (core.py:369-383); no host sync.  ONE kernel launch per step (its
 * prologue finalises the previous step) plus one single-CTA launch that finalises the last step of the batch. */
int smcb_filter_step(smcb_filter *f, int64_t nsteps);
/* sharded filter: particles are partitioned over `world` GPUs, one
 * process each.  With the host-driven exchange a step is step_local (this rank's kernels, ending
 * with its (max, sum exp, sum exp^2 [, moments]) in desc.local_stats), ONE all-gather of 16 doubles
 * per rank into desc.gathered done by the host layer on the same stream (NCCL), and step_finish
 * (global log-normaliser, ESS, logLt recursion and the resampling decision, identical on every
 * rank).  Resampling is per shard with the shard's mass carried in the restart log-weight. */
int smcb_filter_step_local(smcb_filter *f);
int smcb_filter_step_finish(smcb_filter *f);
/* peer memory for the fused exchange: alloc (zeroed) + 64-byte IPC handle to hand to the other
 * ranks (any transport), open a peer's handle, close / free */
int smcb_p2p_alloc(smcb_ctx *ctx, int64_t bytes, void **dev_ptr, unsigned char *handle64);
int smcb_p2p_open(smcb_ctx *ctx, const unsigned char *handle64, void **dev_ptr);
int smcb_p2p_close(void *peer_ptr);
int smcb_p2p_free(void *dev_ptr);
/* same, with a CUDA-event pair around every kernel launch; synchronises at the end.
 * out[0..3] = summed device ms of {init, step kernels of resampling steps, tail, step kernels of
 * non-resampling steps}, out[4..7] = launches of each (bench.py "roofline") */
int smcb_filter_step_timed(smcb_filter *f, int64_t nsteps, double *out8);
/* host-visible snapshot (synchronises the stream):
 * out[0]=t, [1]=cur buffer index, [2]=rs_flag of last step, [3]=logLt, [4]=ESS,
 * [5]=log_mean_w, [6]=max lw, [7]=sum w */
int smcb_filter_state(smcb_filter *f, double *out8);

/* ---------------------------------------------------------------------------
 * measured ceilings of this device (bench.py "roofline.secondary"; no reference counterpart)
 * ------------------------------------------------------------------------- */
/* fp64 FMA issue peak: out3 = {TFLOP/s, DFMA warp-instructions / cycle / SM at sm_mhz (0: skip), ms} */
int smcb_measure_fp64_peak(smcb_ctx *ctx, double sm_mhz, double *out3_host);
/* read + write streaming probe with 16-byte accesses over n doubles (n even): out1 = {GB/s} */
int smcb_measure_stream_peak(smcb_ctx *ctx, const double *in, double *out, int64_t n, double *out1_host);

#ifdef __cplusplus
}
#endif
#endif /* SMCB_H */
