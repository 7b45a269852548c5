"""Generate golden vectors from the LIVE reference (nchopin/particles @ /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python tests/golden/make_golden.py

Everything written here is an OUTPUT OF THE REFERENCE's own code on seeded
inputs; tests compare the oracle (bit-for-bit) and the CUDA path (bit-exact for
integer work, stated tolerances for fp64) against these files.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import particles  # noqa: E402
from particles import distributions as dists  # noqa: E402
from particles import kalman  # noqa: E402
from particles import resampling as rs  # noqa: E402
from particles import state_space_models as ssm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def adversarial_lw():
    """log-weight vectors covering the edge cases of SURVEY.md section 8(c)."""
    r = np.random.RandomState(123)
    out = {}
    out["gauss_1000"] = r.randn(1000) * 3.0
    out["equal_257"] = np.zeros(257)
    out["dominant_513"] = np.concatenate([[0.0], np.full(512, -800.0)])
    v = r.randn(777) * 5
    v[::7] = -np.inf
    out["neginf_777"] = v
    v = r.randn(300)
    v[5] = np.nan
    v[77] = np.nan
    out["nan_300"] = v
    out["single_1"] = np.array([-3.25])
    out["wide_4099"] = r.randn(4099) * 50.0 - 1000.0
    out["tiny_2"] = np.array([-1e-300, 0.0])
    return out


def gen_weights(g):
    for name, lw in adversarial_lw().items():
        g[f"w/{name}/lw_in"] = lw.copy()
        w = rs.Weights(lw=lw.copy())
        g[f"w/{name}/W"] = w.W
        g[f"w/{name}/stats"] = np.array([w.lw.max(), w.log_mean, w.ESS])
        fin = lw.copy()
        fin[np.isnan(fin)] = -np.inf
        g[f"w/{name}/lse"] = np.array([rs.log_sum_exp(fin), rs.log_mean_exp(fin),
                                      rs.essl(fin)])
        g[f"w/{name}/exp_and_normalise"] = rs.exp_and_normalise(fin)
        Wn = rs.exp_and_normalise(np.cos(np.arange(fin.shape[0]) * 0.37))
        g[f"w/{name}/log_mean_exp_W"] = np.array([rs.log_mean_exp(fin, W=Wn)])
        g[f"w/{name}/Wn"] = Wn
    g["w/lse_ab"] = np.array([rs.log_sum_exp_ab(-3.0, 2.5), rs.log_sum_exp_ab(700.0, -2.0)])


def gen_resampling(g):
    r = np.random.RandomState(7)
    cases = {
        "dirichlet_1000": (r.dirichlet(np.ones(1000)), 1000),
        "skewed_513": (rs.exp_and_normalise(r.randn(513) * 6.0), 513),
        "M_lt_N": (rs.exp_and_normalise(r.randn(100)), 10),
        "M_gt_N": (rs.exp_and_normalise(r.randn(100)), 250),
        "zeros_300": (rs.exp_and_normalise(np.where(np.arange(300) % 3 == 0, -np.inf,
                                                    r.randn(300))), 300),
        "dominant_64": (rs.exp_and_normalise(np.concatenate([[0.0], np.full(63, -40.0)])), 64),
        "equal_1025": (np.full(1025, 1.0 / 1025), 1025),
        "n7": (rs.exp_and_normalise(r.randn(7)), 7),
    }
    for name, (W, M) in cases.items():
        g[f"rs/{name}/W"] = W
        g[f"rs/{name}/M"] = np.array([M])
        for scheme in ["systematic", "stratified", "multinomial", "residual"]:
            np.random.seed(99)
            A = rs.resampling(scheme, W, M=M)
            assert A.dtype == np.int64
            g[f"rs/{name}/{scheme}/A"] = A
            # the uniforms that call consumed, in order
            np.random.seed(99)
            if scheme == "systematic":
                u = np.random.rand(1)
            elif scheme == "stratified":
                u = np.random.rand(M)
            elif scheme == "multinomial":
                u = np.random.rand(M + 1)
            else:
                sres = M - int(np.sum(np.floor(M * W)))
                u = np.random.rand(sres + 1) if sres > 0 else np.zeros(0)
            g[f"rs/{name}/{scheme}/u"] = u
        su = np.sort(r.rand(M))
        g[f"rs/{name}/su"] = su
        g[f"rs/{name}/inverse_cdf"] = rs.inverse_cdf(su, W)
    try:
        rs.resampling("bogus", cases["n7"][0])
    except ValueError as e:
        g["rs/bogus_error"] = np.frombuffer(str(e).encode(), dtype=np.uint8)


def gen_dists(g):
    r = np.random.RandomState(11)
    x = r.randn(500) * 2
    loc = r.randn(500)
    scale = np.exp(r.randn(500) * 0.3)
    g["d/normal/x"], g["d/normal/loc"], g["d/normal/scale"] = x, loc, scale
    g["d/normal/logpdf"] = dists.Normal(loc=loc, scale=scale).logpdf(x)
    g["d/normal/logpdf_scalar"] = dists.Normal(loc=0.3, scale=1.7).logpdf(x)
    np.random.seed(5)
    g["d/normal/rvs"] = dists.Normal(loc=loc, scale=scale).rvs(size=500)
    np.random.seed(5)
    g["d/normal/rvs_z"] = np.random.standard_normal(500)
    # MvNormal, d = 4
    d = 4
    Amat = r.randn(d, d)
    cov = Amat @ Amat.T + d * np.eye(d)
    locs = r.randn(200, d)
    xs = r.randn(200, d) * 2
    sc = np.exp(r.randn(d) * 0.2)
    g["d/mvn/cov"], g["d/mvn/loc"], g["d/mvn/x"], g["d/mvn/scale"] = cov, locs, xs, sc
    g["d/mvn/logpdf"] = dists.MvNormal(loc=locs, cov=cov).logpdf(xs)
    g["d/mvn/logpdf_scaled"] = dists.MvNormal(loc=locs, scale=sc, cov=cov).logpdf(xs)
    np.random.seed(6)
    g["d/mvn/rvs"] = dists.MvNormal(loc=locs, scale=sc, cov=cov).rvs(size=200)
    np.random.seed(6)
    g["d/mvn/rvs_z"] = np.random.standard_normal((200, d))
    # IndepProd(Normal, Normal, Dirac, Dirac) as in BearingsOnly
    xp = r.randn(50, 4)
    law = ssm.BearingsOnly().PX(1, xp)
    np.random.seed(8)
    xn = law.rvs(size=50)
    g["d/indep/xp"], g["d/indep/rvs"] = xp, xn
    g["d/indep/logpdf"] = law.logpdf(xn)
    with np.errstate(all="ignore"):
        g["d/indep/bearing_logpdf"] = ssm.BearingsOnly().PY(1, xp, xn).logpdf(np.array([0.7]))


class ToySSM(ssm.StateSpaceModel):
    """README.md:58-66 of the reference."""
    default_params = {"sigma": 0.2}

    def PX0(self):
        return dists.Normal()

    def PX(self, t, xp):
        return dists.Normal(loc=xp)

    def PY(self, t, xp, x):
        return dists.Normal(loc=x, scale=self.sigma)


def flat(y):
    return np.array([np.asarray(v).reshape(-1) for v in y]).squeeze()


def run_case(g, key, fk, N, scheme, essrmin, seed, keepX=True):
    np.random.seed(seed)
    pf = particles.SMC(fk=fk, N=N, resampling=scheme, ESSrmin=essrmin)
    pf.run()
    g[f"run/{key}/logLt"] = np.array([pf.logLt])
    g[f"run/{key}/ESSs"] = np.array(pf.summaries.ESSs)
    g[f"run/{key}/logLts"] = np.array(pf.summaries.logLts)
    g[f"run/{key}/rs_flags"] = np.array(pf.summaries.rs_flags)
    g[f"run/{key}/meta"] = np.array([N, essrmin, seed])
    if keepX:
        g[f"run/{key}/X"] = pf.X
        g[f"run/{key}/A"] = pf.A
        g[f"run/{key}/lw"] = pf.wgts.lw
    return pf


def gen_runs(g):
    # C2-shaped data (T shortened for the bit-exact runs)
    np.random.seed(1)
    sv = ssm.StochVol()
    _, ys = sv.simulate(1000)
    g["data/sv_seed1_T1000"] = flat(ys)
    y60 = ys[:60]
    for scheme in ["systematic", "stratified", "multinomial", "residual"]:
        run_case(g, f"sv_boot_{scheme}", ssm.Bootstrap(ssm=sv, data=y60), 2000, scheme, 0.5, 42)
    run_case(g, "sv_boot_ess1", ssm.Bootstrap(ssm=sv, data=y60), 2000, "systematic", 1.0, 43)
    run_case(g, "sv_guided", ssm.GuidedPF(ssm=sv, data=y60), 2000, "systematic", 0.5, 44)
    run_case(g, "sv_apf", ssm.AuxiliaryPF(ssm=sv, data=y60), 2000, "multinomial", 0.5, 45)
    # C1: ToySSM N=1000 T=200 seed 0
    np.random.seed(0)
    toy = ToySSM(sigma=0.2)
    _, yt = toy.simulate(200)
    g["data/toy_seed0_T200"] = flat(yt)
    run_case(g, "toy_c1", ssm.Bootstrap(ssm=toy, data=yt), 1000, "systematic", 0.5, 0,
             keepX=False)
    # linear Gaussian: bootstrap / guided / APF + exact Kalman
    np.random.seed(2)
    lg = kalman.LinearGauss(sigmaX=1.0, sigmaY=0.2, rho=0.9)
    _, yl = lg.simulate(100)
    g["data/lg_seed2_T100"] = flat(yl)
    kf = kalman.Kalman(ssm=lg, data=yl)
    kf.filter()
    g["kalman/lg_logpyt"] = np.array(kf.logpyt).squeeze()
    run_case(g, "lg_boot", ssm.Bootstrap(ssm=lg, data=yl), 1500, "stratified", 0.5, 46)
    run_case(g, "lg_guided", ssm.GuidedPF(ssm=lg, data=yl), 1500, "stratified", 0.5, 47)
    run_case(g, "lg_apf", ssm.AuxiliaryPF(ssm=lg, data=yl), 1500, "systematic", 0.5, 48)
    # Gordon et al / theta-logistic (non-linear 1-D kernels)
    np.random.seed(3)
    go = ssm.Gordon_etal()
    _, yg = go.simulate(50)
    g["data/gordon_seed3_T50"] = flat(yg)
    run_case(g, "gordon_boot", ssm.Bootstrap(ssm=go, data=yg), 1500, "systematic", 0.5, 49)
    np.random.seed(4)
    tl = ssm.ThetaLogistic()
    _, ytl = tl.simulate(50)
    g["data/thetalogistic_seed4_T50"] = flat(ytl)
    run_case(g, "thetalogistic_boot", ssm.Bootstrap(ssm=tl, data=ytl), 1500, "residual", 0.5, 50)
    # discrete Cox (Poisson observations) and stochastic volatility with leverage
    np.random.seed(6)
    dc = ssm.DiscreteCox(mu=0.5, sigma=0.5, phi=0.9)
    _, yc = dc.simulate(60)
    g["data/cox_seed6_T60"] = flat(yc).astype(np.float64)
    run_case(g, "cox_boot", ssm.Bootstrap(ssm=dc, data=yc), 1500, "systematic", 0.5, 55)
    np.random.seed(7)
    svl = ssm.StochVolLeverage(phi=-0.6)
    _, yv = svl.simulate(60)
    g["data/svlev_seed7_T60"] = flat(yv)
    run_case(g, "svlev_boot", ssm.Bootstrap(ssm=svl, data=yv), 1500, "stratified", 0.5, 56)
    # bearings-only, 4-D IndepProd state (C3 (i))
    np.random.seed(0)
    bo = ssm.BearingsOnly()
    _, yb = bo.simulate(40)
    g["data/bearings_seed0_T40"] = flat(yb)
    with np.errstate(all="ignore"):
        run_case(g, "bearings_boot", ssm.Bootstrap(ssm=bo, data=yb), 1000, "stratified", 0.5, 51)
    # 4-D MvNormal (C3 (ii)): Guarniero et al, bootstrap / guided / APF + Kalman
    np.random.seed(5)
    mv = kalman.MVLinearGauss_Guarniero_etal(alpha=0.4, dx=4)
    _, ym = mv.simulate(30)
    g["data/mvlg_seed5_T30"] = np.array([np.asarray(v).reshape(-1) for v in ym])
    kf = kalman.Kalman(ssm=mv, data=ym)
    kf.filter()
    g["kalman/mvlg_logpyt"] = np.array(kf.logpyt).squeeze()
    run_case(g, "mvlg_boot", ssm.Bootstrap(ssm=mv, data=ym), 800, "stratified", 0.5, 52)
    run_case(g, "mvlg_guided", ssm.GuidedPF(ssm=mv, data=ym), 800, "stratified", 0.5, 53)
    run_case(g, "mvlg_apf", ssm.AuxiliaryPF(ssm=mv, data=ym), 800, "stratified", 0.5, 54)


def gen_stats(g):
    """Monte-Carlo anchors: mean/std of the reference's logLt over repeated runs
    (the sigma of the 3-sigma test the north_star asks for)."""
    ys = g["data/sv_seed1_T1000"]
    sv = ssm.StochVol()
    for N, nrep in [(10000, 24), (100000, 8)]:
        ll, nrs = [], []
        for r in range(nrep):
            np.random.seed(1000 + r)
            pf = particles.SMC(fk=ssm.Bootstrap(ssm=sv, data=ys), N=N)
            pf.run()
            ll.append(pf.logLt)
            nrs.append(int(np.sum(pf.summaries.rs_flags)))
        g[f"stat/sv_T1000_N{N}/logLt"] = np.array(ll)
        g[f"stat/sv_T1000_N{N}/n_resample"] = np.array(nrs)
    yl = g["data/lg_seed2_T100"]
    lg = kalman.LinearGauss(sigmaX=1.0, sigmaY=0.2, rho=0.9)
    for name, cls in [("boot", ssm.Bootstrap), ("guided", ssm.GuidedPF), ("apf", ssm.AuxiliaryPF)]:
        ll = []
        for r in range(20):
            np.random.seed(2000 + r)
            pf = particles.SMC(fk=cls(ssm=lg, data=yl), N=10000)
            pf.run()
            ll.append(pf.logLt)
        g[f"stat/lg_T100_N10000_{name}/logLt"] = np.array(ll)
    ym = g["data/mvlg_seed5_T30"]
    mv = kalman.MVLinearGauss_Guarniero_etal(alpha=0.4, dx=4)
    for name, cls in [("boot", ssm.Bootstrap), ("guided", ssm.GuidedPF), ("apf", ssm.AuxiliaryPF)]:
        ll = []
        for r in range(12):
            np.random.seed(3000 + r)
            pf = particles.SMC(fk=cls(ssm=mv, data=list(ym)), N=10000, resampling="stratified")
            pf.run()
            ll.append(pf.logLt)
        g[f"stat/mvlg_T30_N10000_{name}/logLt"] = np.array(ll)
    yb = g["data/bearings_seed0_T40"]
    bo = ssm.BearingsOnly()
    ll = []
    with np.errstate(all="ignore"):
        for r in range(12):
            np.random.seed(4000 + r)
            pf = particles.SMC(fk=ssm.Bootstrap(ssm=bo, data=list(yb.reshape(-1, 1))), N=20000,
                               resampling="stratified")
            pf.run()
            ll.append(pf.logLt)
    g["stat/bearings_T40_N20000_boot/logLt"] = np.array(ll)


if __name__ == "__main__":
    g = {}
    gen_weights(g)
    gen_resampling(g)
    gen_dists(g)
    gen_runs(g)
    np.savez_compressed(os.path.join(HERE, "golden_exact.npz"), **g)
    s = {"data/sv_seed1_T1000": g["data/sv_seed1_T1000"], "data/lg_seed2_T100": g["data/lg_seed2_T100"],
         "data/mvlg_seed5_T30": g["data/mvlg_seed5_T30"],
         "data/bearings_seed0_T40": g["data/bearings_seed0_T40"]}
    gen_stats(s)
    np.savez_compressed(os.path.join(HERE, "golden_stats.npz"), **s)
    print("wrote", len(g), "exact arrays,", len(s), "stat arrays")
