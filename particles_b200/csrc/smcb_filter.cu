// smcb_filter.cu -- host side of the fused filter (C-ABI entry points) + the 1-D model family.
// The kernels live in smcb_filter_kernels.cuh; the d-dimensional models are instantiated in
// smcb_filter_nd.cu so that the two translation units compile in parallel.
#include <stdlib.h>

#include "smcb_filter_kernels.cuh"

int smcb_bind_nd(smcb_filter *f);      // smcb_filter_nd.cu
int smcb_bind_1d_more(smcb_filter *f); // smcb_filter_1d.cu

static int smcb_bind_1d(smcb_filter *f) {
    const smcb_filter_desc *d = &f->desc;
    switch (d->model) {
#ifdef SMCB_BENCH_ONLY   // experiment builds: only the config-2 instantiation (fast compile)
        case SMCB_MODEL_STOCHVOL:
            return (d->fk == SMCB_FK_BOOTSTRAP && d->scheme == SMCB_RS_SYSTEMATIC)
                       ? bind_one<StochVolM, SMCB_FK_BOOTSTRAP, SMCB_RS_SYSTEMATIC>(f) : SMCB_ENOSYS;
#else
        case SMCB_MODEL_STOCHVOL: return bind_fk<StochVolM>(f);
        case SMCB_MODEL_LINGAUSS: return bind_fk<LinGaussM>(f);
        case SMCB_MODEL_GORDON:
        case SMCB_MODEL_THETALOGISTIC:
        case SMCB_MODEL_DISCRETECOX:
        case SMCB_MODEL_STOCHVOLLEV: return smcb_bind_1d_more(f);
#endif
        default:
            set_error("fused filter: model id %d is not available in the fused 1-D family", d->model);
            return SMCB_ENOSYS;
    }
}

static int filter_setup(smcb_filter *f, smcb_ctx *c, const smcb_filter_desc *d);

extern "C" int smcb_filter_create(smcb_ctx *c, const smcb_filter_desc *d, smcb_filter **out) {
    SMCB_REQUIRE(c && d && out, "smcb_filter_create: NULL argument");
    SMCB_REQUIRE(d->n >= 1 && d->T >= 1, "smcb_filter_create: need n >= 1 and T >= 1");
    SMCB_REQUIRE(d->X[0] && d->X[1] && d->lw[0] && d->lw[1] && d->A && d->cdf && d->data && d->summaries,
                 "smcb_filter_create: NULL device buffer");
    SMCB_REQUIRE((d->index_offset & 1) == 0, "smcb_filter_create: index_offset must be even");
    SMCB_REQUIRE(d->dim >= 1 && d->dim <= 4, "smcb_filter_create: fused kernels exist for state dimension 1..4 (dim=%d)", d->dim);
    SMCB_REQUIRE(d->dy >= 1 && d->dy <= kMaxDy, "smcb_filter_create: observation dimension must be 1..%d", kMaxDy);
    SMCB_REQUIRE(d->essrmin >= 0.0 && d->essrmin <= 1.0, "smcb_filter_create: ESSrmin must be in [0, 1]");
    smcb_filter *f = new (std::nothrow) smcb_filter();
    SMCB_REQUIRE(f != nullptr, "smcb_filter_create: out of host memory");
    f->ctx = c;
    f->desc = *d;
    f->t_host = 0;
    f->timed_ev = nullptr;
    f->timed_kind = nullptr;
    f->graph = nullptr; f->gexec = nullptr; f->has_graph = false; f->t_stop_host = 0;
    f->scan_mem = nullptr;
    const int rc = filter_setup(f, c, d);
    if (rc) {                       // nothing of a half-built filter survives an error
        if (f->scan_mem) cudaFree(f->scan_mem);
        delete f;
        return rc;
    }
    *out = f;
    return SMCB_OK;
}

static int filter_setup(smcb_filter *f, smcb_ctx *c, const smcb_filter_desc *d) {
    int rc = (d->dim == 1) ? smcb_bind_1d(f) : smcb_bind_nd(f);
    if (rc) return rc;

    const int64_t n = d->n;
    const size_t sb1 = (scan_state_bytes(n) + 63) & ~(size_t)63;
    const size_t sb2 = (scan_state_bytes(n + 1) + 63) & ~(size_t)63;
    const size_t part = (size_t)kMaxGrid * 8 * sizeof(double);
    char *mem;
    const size_t tpb = (size_t)(kMaxGrid + 8) * sizeof(double);
    constexpr size_t kHdr = 512;       // FilterDev, then the two "last block done" tickets
    static_assert(sizeof(FilterDev) <= 448, "FilterDev outgrew its header slot");
    SMCB_CUDA(cudaMalloc(&mem, kHdr + part + sb1 + sb2 + tpb + 64));
    f->scan_mem = mem;
    SMCB_CUDA(cudaMemsetAsync(mem, 0, kHdr + part, c->stream));
    SMCB_CUDA(cudaMemsetAsync(mem + kHdr + part, 0xFF, sb1 + sb2, c->stream));
    f->st = reinterpret_cast<FilterDev *>(mem);
    FilterArgs &a = f->args;
    memset(&a, 0, sizeof(a));
    a.X[0] = d->X[0]; a.X[1] = d->X[1]; a.lw[0] = d->lw[0]; a.lw[1] = d->lw[1];
    a.A = reinterpret_cast<long long *>(d->A);
    a.cdf = d->cdf;
    a.su = (d->scheme == SMCB_RS_MULTINOMIAL) ? d->scratch : nullptr;
    SMCB_REQUIRE(d->scheme != SMCB_RS_MULTINOMIAL || d->scratch != nullptr,
                 "smcb_filter_create: multinomial needs desc.scratch of n + 2 doubles");
    a.data = d->data;
    a.sc = d->step_consts;
    a.summaries = d->summaries;
    a.z_in = d->z_in; a.u_in = d->u_in;
    a.st = f->st;
    a.partials = reinterpret_cast<double *>(mem + kHdr);
    a.ticket = reinterpret_cast<unsigned int *>(mem + 448);
    a.ticket2 = reinterpret_cast<unsigned int *>(mem + 456);
    char *sp = mem + kHdr + part;
    a.scan.ticket = reinterpret_cast<unsigned int *>(sp);
    a.scan.agg = reinterpret_cast<unsigned long long *>(sp + 16);
    a.scan.cpref = a.scan.agg + scan_tiles(n);
    sp += sb1;
    a.scan2.ticket = reinterpret_cast<unsigned int *>(sp);
    a.scan2.agg = reinterpret_cast<unsigned long long *>(sp + 16);
    a.scan2.cpref = a.scan2.agg + scan_tiles(n + 1);
    a.dy = d->dy;
    a.n = n; a.n_global = d->n_global > 0 ? d->n_global : n;
    a.index_offset = d->index_offset; a.T = d->T;
    a.essrmin = d->essrmin;
    a.key = key_of(d->seed);
    a.tile_pref = reinterpret_cast<double *>(mem + kHdr + part + sb1 + sb2);
    {   // persistent grid: one wave of resident CTAs, each owning a contiguous range of pairs
        int dev = 0, sms = kSMs;
        SMCB_CUDA(cudaGetDevice(&dev));
        SMCB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int64_t npairs = (n + 1) / 2;
        const int64_t unit = (int64_t)kBlock * SMCB_KU;            // pairs per block iteration
        int64_t g = (int64_t)sms * f->blocks_per_sm;
        if (g > kMaxGrid) g = kMaxGrid;
        int64_t chunk = ((npairs + g - 1) / g + unit - 1) / unit * unit;
        g = (npairs + chunk - 1) / chunk;
        a.chunk = chunk;
        f->grid_move = (int)(g < 1 ? 1 : g);
        a.grid = f->grid_move;
        a.world = d->world > 1 ? d->world : 1;
        a.rank = d->world > 1 ? d->rank : 0;
        a.local_stats = d->local_stats;
        a.gathered = d->gathered;
        a.mail_local = (a.world > 1) ? d->mail_local : nullptr;
        for (int r = 0; r < 8; r++) a.mail_peer[r] = (a.world > 1 && r < a.world) ? d->mail_peer[r] : nullptr;
        SMCB_REQUIRE(a.world == 1 || (a.rank >= 0 && a.rank < a.world), "smcb_filter_create: bad rank");
        SMCB_REQUIRE(a.world == 1 || a.mail_local || (d->local_stats && d->gathered),
                     "smcb_filter_create: world > 1 needs either a peer mailbox or local_stats / gathered");
        if (a.mail_local) {
            SMCB_REQUIRE(a.world <= 8, "smcb_filter_create: the peer mailbox supports at most 8 ranks");
            for (int r = 0; r < a.world; r++)
                SMCB_REQUIRE(a.mail_peer[r] != nullptr, "smcb_filter_create: mail_peer[%d] is NULL", r);
        }
        a.rs_global = (a.world > 1 && d->rs_global) ? 1 : 0;
        if (a.rs_global) {
            SMCB_REQUIRE(a.mail_local != nullptr, "smcb_filter_create: global resampling needs the peer mailbox");
            SMCB_REQUIRE((n & 1) == 0, "smcb_filter_create: global resampling needs an even shard size");
            SMCB_REQUIRE(d->stage_X && d->stage_lw, "smcb_filter_create: global resampling needs stage_X / stage_lw");
            if (d->fk == SMCB_FK_APF || d->fk == SMCB_FK_AUXBOOT) {
                set_error("fused filter: global resampling over shards is built for Feynman-Kac kinds without "
                          "auxiliary weights (bootstrap, guided)");
                return SMCB_ENOSYS;
            }
            if (d->scheme != SMCB_RS_SYSTEMATIC && d->scheme != SMCB_RS_STRATIFIED) {
                set_error("fused filter: global resampling over shards supports systematic and stratified");
                return SMCB_ENOSYS;
            }
            a.stage_X = d->stage_X; a.stage_lw = d->stage_lw;
            for (int r = 0; r < a.world; r++) {
                SMCB_REQUIRE(d->peer_X0[r] && d->peer_X1[r] && d->peer_cdf[r],
                             "smcb_filter_create: peer_X0 / peer_X1 / peer_cdf[%d] is NULL", r);
                a.pX[r][0] = d->peer_X0[r]; a.pX[r][1] = d->peer_X1[r]; a.pcdf[r] = d->peer_cdf[r];
            }
        }
    }
    {
        int64_t t2 = scan_tiles(n + 1);
        f->grid_scan2 = (int)(t2 < kMaxGrid ? t2 : kMaxGrid);
    }
    // Optional: the whole step loop as ONE CUDA graph (WHILE + IF/ELSE conditional nodes, branch-
    // specialised kernels).  Measured on this driver (profiles/graph_vs_loop.py) a conditional-node
    // iteration costs ~7 us more than the two plain launches it replaces (14 -> 21 us per step at
    // N = 1e3, 121 -> 130 us at N = 1e7), so the launch-per-step loop stays the default and the
    // graph is opt-in (SMCB_GRAPH=1).
    if (a.world == 1 && getenv("SMCB_GRAPH") != nullptr) {
        if (f->build_graph(f) != SMCB_OK) {
            cudaGetLastError();
            f->has_graph = false;
            set_error("");
        }
    }
    return SMCB_OK;
}

extern "C" int smcb_filter_destroy(smcb_filter *f) {
    if (!f) return SMCB_OK;
    cudaStreamSynchronize(f->ctx->stream);
    if (f->gexec) cudaGraphExecDestroy(f->gexec);
    if (f->graph) cudaGraphDestroy(f->graph);
    cudaFree(f->scan_mem);
    delete f;
    return SMCB_OK;
}

// sharded filters: one step = step_local (kernels up to the per-rank statistics), an all-gather
// of 8 doubles per rank done by the host layer (NCCL, same stream), then step_finish
extern "C" int smcb_filter_step_local(smcb_filter *f) {
    SMCB_REQUIRE(f != nullptr && f->args.world > 1, "smcb_filter_step_local: not a sharded filter");
    SMCB_REQUIRE(f->t_host < f->desc.T, "smcb_filter_step_local: all steps already done");
    return (f->t_host == 0) ? f->launch_init(f) : f->launch_step(f);
}

extern "C" int smcb_filter_step_finish(smcb_filter *f) {
    SMCB_REQUIRE(f != nullptr && f->args.world > 1, "smcb_filter_step_finish: not a sharded filter");
    int rc = f->launch_finish(f);
    if (rc) return rc;
    f->t_host++;
    return SMCB_OK;
}

extern "C" int smcb_filter_step(smcb_filter *f, int64_t nsteps) {
    SMCB_REQUIRE(f != nullptr, "smcb_filter_step: NULL filter");
    SMCB_REQUIRE(f->args.world == 1 || f->args.mail_local != nullptr,
                 "smcb_filter_step: sharded filters without a peer mailbox use step_local / step_finish");
    if (f->has_graph && nsteps > 0) {
        SMCB_REQUIRE(f->t_host + nsteps <= f->desc.T, "smcb_filter_step: all %lld steps already done (StopIteration)",
                     (long long)f->desc.T);
        if (f->t_host == 0) {                      // step 0 (generate_particles) is a plain launch
            int rc = f->launch_init(f);
            if (rc) return rc;
            f->t_host++;
            nsteps--;
        }
        if (nsteps > 0) {                          // steps t_host .. t_host + nsteps - 1: ONE graph launch
            f->t_stop_host = f->t_host + nsteps;
            SMCB_CUDA(cudaMemcpyAsync(&f->st->t_stop, &f->t_stop_host, sizeof(long long), cudaMemcpyHostToDevice,
                                      f->ctx->stream));
            SMCB_CUDA(cudaGraphLaunch(f->gexec, f->ctx->stream));
            f->ctx->launches += 2 * nsteps;        // k_cond + k_move per step (+1 scan per resampling step)
            f->t_host += nsteps;
        }
        return SMCB_OK;
    }
    // host mirror of t: the device advances by exactly one per launched step
    for (int64_t i = 0; i < nsteps; i++) {
        if (f->t_host >= f->desc.T) {
            set_error("smcb_filter_step: all %lld steps already done (StopIteration)", (long long)f->desc.T);
            return SMCB_EINVAL;
        }
        int rc = (f->t_host == 0) ? f->launch_init(f) : f->launch_step(f);
        if (rc) return rc;
        if (f->args.world > 1 && (rc = f->launch_finish(f))) return rc;   // exchange happens on device
        f->t_host++;
    }
    return SMCB_OK;
}

// Same as smcb_filter_step, with a CUDA-event pair around every kernel launch (on the
// launching stream).  out[0..3] = summed device milliseconds of {init, scan, spacings, move},
// out[4..7] = number of launches of each.  Synchronises once, at the end.
extern "C" int smcb_filter_step_timed(smcb_filter *f, int64_t nsteps, double *out8) {
    SMCB_REQUIRE(f && out8, "smcb_filter_step_timed: NULL argument");
    SMCB_REQUIRE(nsteps >= 0 && nsteps <= 100000, "smcb_filter_step_timed: nsteps out of range");
    SMCB_REQUIRE(f->t_host + nsteps <= f->desc.T, "smcb_filter_step_timed: past the last step");
    SMCB_REQUIRE(f->args.world == 1, "smcb_filter_step_timed: single-device filters only");
    cudaStream_t s = f->ctx->stream;
    const bool multi = f->desc.scheme == SMCB_RS_MULTINOMIAL;
    const int per = multi ? 3 : 2;
    const size_t nev = (size_t)nsteps * per * 2 + 2;
    cudaEvent_t *ev = new (std::nothrow) cudaEvent_t[nev];
    SMCB_REQUIRE(ev != nullptr, "smcb_filter_step_timed: out of host memory");
    for (size_t i = 0; i < nev; i++) SMCB_CUDA(cudaEventCreate(&ev[i]));
    int *kind = new int[nev / 2];
    size_t k = 0;
    // re-implements launch_step with events in between: the kernels are the same objects
    for (int64_t i = 0; i < nsteps; i++) {
        if (f->t_host == 0) {
            SMCB_CUDA(cudaEventRecord(ev[2 * k], s));
            int rc = f->launch_init(f);
            if (rc) return rc;
            SMCB_CUDA(cudaEventRecord(ev[2 * k + 1], s));
            kind[k++] = 0;
        } else {
            f->timed_ev = ev + 2 * k;
            f->timed_kind = kind + k;
            int rc = f->launch_step(f);
            f->timed_ev = nullptr;
            if (rc) return rc;
            k += per;
        }
        f->t_host++;
    }
    SMCB_CUDA(cudaStreamSynchronize(s));
    for (int j = 0; j < 8; j++) out8[j] = 0.0;
    for (size_t i = 0; i < k; i++) {
        float ms = 0.f;
        SMCB_CUDA(cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]));
        out8[kind[i]] += ms;
        out8[4 + kind[i]] += 1.0;
    }
    for (size_t i = 0; i < nev; i++) cudaEventDestroy(ev[i]);
    delete[] ev;
    delete[] kind;
    return SMCB_OK;
}

extern "C" int smcb_filter_state(smcb_filter *f, double *out8) {
    SMCB_REQUIRE(f && out8, "smcb_filter_state: NULL argument");
    FilterDev h;
    SMCB_CUDA(cudaMemcpyAsync(&h, f->st, sizeof(h), cudaMemcpyDeviceToHost, f->ctx->stream));
    SMCB_CUDA(cudaStreamSynchronize(f->ctx->stream));
    out8[0] = (double)h.t; out8[1] = (double)h.cur; out8[2] = (double)h.last_rs; out8[3] = h.logLt;
    out8[4] = h.ess; out8[5] = h.log_mean_w; out8[6] = h.wm; out8[7] = h.ws;
    if (h.sync_timeout) {
        set_error("sharded filter: a wait on a peer GPU's flag timed out (a rank died or fell out of step)");
        return SMCB_ECUDA;
    }
    return SMCB_OK;
}

// ---------------------------------------------------------------------------
// peer memory for the fused statistics exchange (one process per GPU, same node)
// ---------------------------------------------------------------------------
extern "C" int smcb_p2p_alloc(smcb_ctx *c, int64_t bytes, void **dev_ptr, unsigned char *handle64) {
    SMCB_REQUIRE(c && dev_ptr && handle64 && bytes > 0, "smcb_p2p_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    SMCB_CUDA(cudaMalloc(&p, (size_t)bytes));
    SMCB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    SMCB_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return SMCB_OK;
}

extern "C" int smcb_p2p_open(smcb_ctx *c, const unsigned char *handle64, void **dev_ptr) {
    SMCB_REQUIRE(c && dev_ptr && handle64, "smcb_p2p_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SMCB_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SMCB_OK;
}

extern "C" int smcb_p2p_close(void *peer_ptr) {
    if (peer_ptr) SMCB_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return SMCB_OK;
}

extern "C" int smcb_p2p_free(void *dev_ptr) {
    if (dev_ptr) SMCB_CUDA(cudaFree(dev_ptr));
    return SMCB_OK;
}

#ifdef SMCB_TRACE
// debug builds only (profiles/build_variant.sh trace -DSMCB_TRACE): timeline of the last step-kernel launch
extern "C" int smcb_debug_trace(unsigned long long *host_out, int n_words) {
    SMCB_CUDA(cudaDeviceSynchronize());
    SMCB_CUDA(cudaMemcpyFromSymbol(host_out, smcb::g_trace, sizeof(unsigned long long) * n_words));
    return SMCB_OK;
}
#endif
