// smcb_filter_nd.cu -- instantiations of the fused step kernels for the d-dimensional models
// (SoA state): BearingsOnly (d = 4) and MVLinearGauss (dx = 2..4).  See smcb_step.cuh.
#include "smcb_step.cuh"

int smcb_bind_nd(smcb_filter *f) {
    const smcb_filter_desc *d = &f->desc;
#if defined(SMCB_BENCH_ND)      // experiment builds with the two config-3 instantiations (profiles/build_variant.sh)
    if (d->model == SMCB_MODEL_BEARINGS && d->fk == SMCB_FK_BOOTSTRAP && d->scheme == SMCB_RS_STRATIFIED)
        return bind_one<BearingsM, SMCB_FK_BOOTSTRAP, SMCB_RS_STRATIFIED>(f);
    if (d->model == SMCB_MODEL_MVLINGAUSS && d->dim == 4 && d->fk == SMCB_FK_GUIDED && d->scheme == SMCB_RS_STRATIFIED)
        return bind_one<MvLinGaussM<4>, SMCB_FK_GUIDED, SMCB_RS_STRATIFIED>(f);
    set_error("experiment build: only the config-3 instantiations are compiled in");
    return SMCB_ENOSYS;
#elif defined(SMCB_BENCH_ONLY)
    set_error("experiment build: d-dimensional models are not compiled in");
    return SMCB_ENOSYS;
#else
    switch (d->model) {
        case SMCB_MODEL_BEARINGS:
            if (d->dim != 4) { set_error("fused BearingsOnly: state dimension must be 4"); return SMCB_EINVAL; }
            return bind_fk<BearingsM>(f);
        case SMCB_MODEL_MVLINGAUSS:
            if (d->dim == 2) return bind_fk<MvLinGaussM<2>>(f);
            if (d->dim == 3) return bind_fk<MvLinGaussM<3>>(f);
            if (d->dim == 4) return bind_fk<MvLinGaussM<4>>(f);
            set_error("fused MVLinearGauss: dx must be 2, 3 or 4 (got %d)", d->dim);
            return SMCB_ENOSYS;
        default:
            set_error("fused filter: model id %d has no d-dimensional kernel", d->model);
            return SMCB_ENOSYS;
    }
#endif
}
