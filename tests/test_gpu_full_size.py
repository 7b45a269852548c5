"""Full-size runs of the other BASELINE configs (3 and 5) on one GPU, checked through
size-independent properties and known answers (config 2 at N = 1e7 is in test_gpu_filter.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import samplers_numpy as sp  # noqa: E402


def host(t):
    return t.detach().cpu().numpy()


def test_config3_mvnormal_guided_full_size(golden, golden_stats):
    """Config 3 (ii): 4-D MvNormal guided and auxiliary filters at N = 1e6, T = 500 (the 30 golden
    observations, cycled) -- exact Kalman log-likelihood of the same cycled data from the oracle."""
    import particles_b200 as pb
    from particles_b200 import kalman, state_space_models as ssm
    from oracle import smc_numpy as orc
    ym = list(np.tile(golden_stats["data/mvlg_seed5_T30"], (17, 1))[:500])
    exact = float(np.sum(orc.MVLinearGauss_Guarniero_etal(0.4, 4).kalman_loglik(ym)))
    mv = kalman.MVLinearGauss_Guarniero_etal(alpha=0.4, dx=4)
    for cls, tol in [(ssm.GuidedPF, 0.05), (ssm.AuxiliaryPF, 0.05), (ssm.Bootstrap, 0.25)]:
        pf = pb.SMC(fk=cls(ssm=mv, data=ym), N=1_000_000, resampling="stratified", seed=5)
        assert pf.fused
        pf.run()
        assert abs(pf.logLt - exact) < tol, (cls.__name__, pf.logLt, exact)
        ess = np.array(pf.summaries.ESSs)
        assert np.all((ess >= 1) & (ess <= 1_000_000 * (1 + 1e-12)))
        assert pf.X.shape == (1_000_000, 4)


def test_config3_bearings_full_size(golden_stats):
    """Config 3 (i): BearingsOnly bootstrap, N = 1e6, stratified: ancestors sorted and in range, the Dirac
    components obey x2' = x0 + x2 exactly for the resampled parents, weights normalise."""
    import particles_b200 as pb
    from particles_b200 import state_space_models as ssm
    yb = list(golden_stats["data/bearings_seed0_T40"].reshape(-1, 1))
    pf = pb.SMC(fk=ssm.Bootstrap(ssm=ssm.BearingsOnly(), data=yb), N=1_000_000, resampling="stratified", seed=1)
    checked = 0
    for t in range(len(yb)):
        next(pf)
        if t > 0 and pf.rs_flag and checked < 3:
            A, X, Xp = host(pf.A), host(pf.X), host(pf.Xp)
            assert A.min() >= 0 and A.max() < 1_000_000 and np.all(np.diff(A) >= 0)
            assert np.array_equal(X[:, 2], Xp[:, 0] + Xp[:, 2]) and np.array_equal(X[:, 3], Xp[:, 1] + Xp[:, 3])
            checked += 1
    assert checked == 3
    assert abs(float(pf.W.sum().item()) - 1) < 1e-10
    ref = golden_stats["stat/bearings_T40_N20000_boot/logLt"]
    assert abs(pf.logLt - ref.mean()) < 4 * ref.std(ddof=1) + 0.05


def test_config5_tempering_full_size():
    """Config 5: 20-D logistic regression, n_data = 1000, waste-free adaptive tempering with 1e4 chains x 100
    = 1e6 particles: log evidence against an oracle run at 300 x 20 (sd ~ 1.5), exponents increasing to 1,
    acceptance rate of the calibrated random walk near 0.25, path-sampling identity d log Z = E[llik] d epn."""
    import particles_b200 as pb
    from particles_b200 import smc_samplers as ssp
    data = sp.synthetic_logistic(1000, 20, seed=0)
    model = ssp.LogisticRegression(data=data, prior_scale=5.0)
    pf = pb.SMC(fk=ssp.AdaptiveTempering(model=model, ESSrmin=0.5, wastefree=True, len_chain=100), N=10_000,
                ESSrmin=1.0, seed=4)
    pf.run()
    epn = np.array(pf.X.shared["exponents"])
    assert pf.X.N == 1_000_000 and epn[0] == 0 and epn[-1] == 1 and np.all(np.diff(epn) > 0)
    np.random.seed(1)
    ref = sp.run_tempering(sp.LogisticModel(data), 300, 20, 0.5)
    assert abs(pf.logLt - ref["logLt"]) < 6.0, (pf.logLt, ref["logLt"])
    acc = np.concatenate([host(a) for a in pf.X.shared["acc_rates"]])
    assert 0.15 < acc.mean() < 0.40
    ess = np.array(pf.summaries.ESSs)
    assert np.all(ess[:-1] > 0.45 * 1_000_000) and np.all(ess[:-1] < 0.55 * 1_000_000)   # ESS target of the bisection
    W = pf.W
    post_mean = host((W[:, None] * pf.X.theta).sum(0) / W.sum())
    ref_mean = np.average(ref["X"].theta, weights=ref["W"], axis=0)
    assert np.max(np.abs(post_mean - ref_mean)) < 0.35
