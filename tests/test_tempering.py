"""Waste-free adaptive tempering (BASELINE config 5).

CPU: the oracle (oracle/samplers_numpy.py) against a seeded run of the LIVE reference
(tests/golden/golden_tempering.npz, made by tests/golden/make_golden_tempering.py).
GPU: the device kernels against the oracle on identical inputs, then full runs against the
reference's own Monte-Carlo spread."""
import numpy as np
import torch
import pytest

from oracle import samplers_numpy as sp
from oracle import smc_numpy as orc


@pytest.fixture(scope="module")
def gt():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_tempering.npz"))


def test_oracle_reproduces_reference_run(gt):
    """Same seed, same stream order -> same trajectory (BLAS / LAPACK in the reference's covariance,
    Cholesky and matmul leave last-bit differences: 1e-12)."""
    N, P, seed = (int(v) for v in gt["exact/meta"])
    np.random.seed(seed)
    out = sp.run_tempering(sp.LogisticModel(gt["exact/data"]), N, P, 0.5)
    assert len(out["exponents"]) == len(gt["exact/exponents"])
    np.testing.assert_allclose(out["exponents"], gt["exact/exponents"], rtol=1e-11)
    np.testing.assert_allclose(out["logLts"], gt["exact/logLts"], rtol=1e-12)
    np.testing.assert_allclose(out["ESSs"], gt["exact/ESSs"], rtol=1e-10)
    np.testing.assert_allclose(out["X"].theta, gt["exact/theta"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(out["X"].lpost, gt["exact/lpost"], rtol=1e-11)
    np.testing.assert_allclose(out["path_sampling"], gt["exact/path_sampling"][0], rtol=1e-12)
    assert out["exponents"][-1] == 1.0 and out["exponents"][0] == 0.0


def test_synthetic_logistic_is_deterministic():
    a, b = sp.synthetic_logistic(50, 5, seed=1), sp.synthetic_logistic(50, 5, seed=1)
    assert np.array_equal(a, b) and a.shape == (50, 5) and np.all(np.abs(a[:, 0]) == 1.0)


# ------------------------------------------------------------------------------------ GPU
gpu = pytest.mark.gpu


def host(t):
    return t.detach().cpu().numpy()


@gpu
@pytest.mark.parametrize("d,n_data,n", [(4, 150, 1001), (20, 1000, 4096), (7, 33, 2), (32, 64, 513)])
def test_logistic_target_kernel_vs_oracle(d, n_data, n):
    """lprior / llik / lpost of the fused target kernel against the oracle's row-by-row NumPy loop."""
    torch = pytest.importorskip("torch")
    from particles_b200 import smc_samplers as ssp
    data = sp.synthetic_logistic(n_data, d, seed=d)
    r = np.random.RandomState(1)
    theta = r.randn(n, d) * 2.0
    theta[0] *= 40.0                                     # saturated logits
    m = sp.LogisticModel(data)
    mdev = ssp.LogisticRegression(data=data, prior_scale=5.0)
    for epn in (0.0, 0.37, 1.0):
        x = ssp.ThetaParticles(theta=torch.from_numpy(theta).cuda())
        mdev.target(x, epn)
        ll, lp = m.loglik(theta), m.prior.logpdf(theta)
        np.testing.assert_allclose(host(x.llik), ll, rtol=1e-13, atol=1e-11)
        np.testing.assert_allclose(host(x.lprior), lp, rtol=1e-13)
        np.testing.assert_allclose(host(x.lpost), lp + epn * ll if epn > 0 else lp, rtol=1e-13, atol=1e-11)


@gpu
def test_rw_proposal_and_accept_vs_oracle():
    """One Metropolis step with injected normals / uniforms: same proposals (1e-14), same accept
    decisions, same state afterwards."""
    torch = pytest.importorskip("torch")
    from particles_b200 import smc_samplers as ssp
    d, n = 6, 3001
    data = sp.synthetic_logistic(200, d, seed=2)
    r = np.random.RandomState(3)
    theta = r.randn(n, d)
    W = orc.exp_and_normalise(r.randn(n))
    m = sp.LogisticModel(data)
    fk = sp.AdaptiveTemperingWF(m)
    xo = sp.ThetaParticles(theta=theta.copy())
    fk.target(0.6)(xo)
    fk.calibrate(W, xo)
    z, u = r.standard_normal((n, d)), r.rand(n)
    # oracle step with the same draws
    xprop = sp.ThetaParticles(theta=xo.theta + z @ xo.shared["chol_cov"].T)
    fk.target(0.6)(xprop)
    pb = np.exp(np.clip(xprop.lpost - xo.lpost, None, 0.0))
    acc = u < pb
    ref = xo.copy()
    ref.copyto(xprop, where=acc)
    # device
    mdev = ssp.LogisticRegression(data=data)
    x = ssp.ThetaParticles(theta=torch.from_numpy(theta).cuda())
    mdev.target(x, 0.6)
    rw = ssp.ArrayRandomWalk()
    rw.calibrate(torch.from_numpy(W).cuda(), x)
    np.testing.assert_allclose(host(x.shared["chol_cov"]), xo.shared["chol_cov"], rtol=1e-10, atol=1e-14)
    mean_acc = rw.step(x, lambda xx: mdev.target(xx, 0.6), noise=(z, u))
    np.testing.assert_allclose(float(mean_acc.item()), pb.mean(), rtol=1e-10)
    same = np.isclose(host(x.lpost), ref.lpost, rtol=1e-10)
    assert same.mean() > 0.999                       # a draw within 1e-10 of its threshold may flip
    np.testing.assert_allclose(host(x.theta)[same], ref.theta[same], rtol=1e-9, atol=1e-12)
    assert 0.05 < acc.mean() < 0.95


@gpu
def test_device_root_find_matches_brentq():
    """next_annealing_epn on the device (16-way bracketing, 11 passes) against the reference's formulation (brentq
    on the host with the device essl) and against the oracle's NumPy version."""
    pytest.importorskip("torch")
    from particles_b200 import smc_samplers as ssp
    r = np.random.RandomState(7)
    for n, scale, epn in [(100_000, 40.0, 0.0), (50_001, 300.0, 0.013), (4000, 5.0, 0.4), (20_000, 0.01, 0.2)]:
        lw = -np.abs(r.randn(n)) * scale - 3.0
        lwd = torch.from_numpy(lw).cuda()
        got = ssp.next_annealing_epn(epn, 0.5, lwd)
        want = ssp.next_annealing_epn_host(epn, 0.5, lwd)
        assert abs(got - want) <= 1e-9 * max(1.0, want) + 1e-12, (n, got, want)
        ref = sp.next_annealing_epn(epn, 0.5, lw)
        assert abs(got - ref) <= 1e-9 * max(1.0, ref) + 1e-11, (n, got, ref)
        if got < 1.0:       # it IS the root: ESS at the new exponent is alpha N
            np.testing.assert_allclose(orc.essl((got - epn) * lw), 0.5 * n, rtol=1e-8)


@gpu
def test_adaptive_tempering_vs_reference_runs(gt):
    """Full waste-free adaptive-tempering runs (N = 400 chains x P = 25) against 12 runs of the
    reference on the same data: log evidence within 3 sigma, posterior mean within Monte-Carlo
    error, same number of tempering steps."""
    pytest.importorskip("torch")
    import particles_b200 as pb
    from particles_b200 import smc_samplers as ssp
    data = gt["stat/data"]
    N, P = (int(v) for v in gt["stat/meta"])
    ref_ll, ref_mean = gt["stat/logLt"], gt["stat/post_mean"]
    mu, sd = ref_ll.mean(), ref_ll.std(ddof=1)
    lls, means = [], []
    for s in range(6):
        model = ssp.LogisticRegression(data=data, prior_scale=5.0)
        fk = ssp.AdaptiveTempering(model=model, ESSrmin=0.5, wastefree=True, len_chain=P)
        pf = pb.SMC(fk=fk, N=N, ESSrmin=1.0, seed=40 + s)
        pf.run()
        assert pf.X.N == N * P and pf.X.shared["exponents"][-1] == 1.0
        assert abs(len(pf.summaries.ESSs) - int(np.median(gt["stat/nsteps"]))) <= 1
        W = pf.W
        lls.append(pf.logLt)
        means.append(host((W[:, None] * pf.X.theta).sum(0) / W.sum()))
    lls, means = np.array(lls), np.array(means)
    assert abs(lls.mean() - mu) < 3 * sd * np.sqrt(1 / 6 + 1 / len(ref_ll)) + 1e-6, (lls, mu, sd)
    assert 0.25 * sd < lls.std(ddof=1) < 4 * sd
    msd = ref_mean.std(axis=0, ddof=1)
    assert np.all(np.abs(means.mean(0) - ref_mean.mean(0)) < 4 * msd * np.sqrt(1 / 6 + 1 / 12) + 1e-3)


@gpu
def test_sharded_sampler_world1_vs_reference_runs(gt):
    """The sharded waste-free sampler (global resampling through the two-level CDF, device control plane) with a
    single rank: log evidence / posterior mean against the reference's runs, same number of tempering steps."""
    pytest.importorskip("torch")
    from particles_b200 import smc_samplers as ssp
    from particles_b200.sharded_samplers import ShardedAdaptiveTempering
    data = gt["stat/data"]
    N, P = (int(v) for v in gt["stat/meta"])
    ref_ll, ref_mean = gt["stat/logLt"], gt["stat/post_mean"]
    mu, sd = ref_ll.mean(), ref_ll.std(ddof=1)
    lls, means = [], []
    for s in range(6):
        sm = ShardedAdaptiveTempering(model=ssp.LogisticRegression(data=data, prior_scale=5.0), M_local=N, len_chain=P,
                                      ESSrmin=0.5, seed=70 + s).run()
        assert sm.exponents[-1] == 1.0 and sm.X.N == N * P
        assert abs(len(sm.exponents) - 1 - int(np.median(gt["stat/nsteps"]))) <= 1
        lls.append(sm.logLt)
        means.append(sm.posterior_mean())
    lls, means = np.array(lls), np.array(means)
    assert abs(lls.mean() - mu) < 3 * sd * np.sqrt(1 / 6 + 1 / len(ref_ll)) + 1e-6, (lls, mu, sd)
    msd = ref_mean.std(axis=0, ddof=1)
    assert np.all(np.abs(means.mean(0) - ref_mean.mean(0)) < 4 * msd * np.sqrt(1 / 6 + 1 / 12) + 1e-3)


@gpu
def test_fixed_tempering_and_standard_move():
    """Tempering with a fixed exponent ladder and the non-waste-free move runs and agrees with the
    adaptive waste-free estimate of the same evidence."""
    pytest.importorskip("torch")
    import particles_b200 as pb
    from particles_b200 import smc_samplers as ssp
    data = sp.synthetic_logistic(120, 3, seed=9)
    model = ssp.LogisticRegression(data=data)
    a = pb.SMC(fk=ssp.AdaptiveTempering(model=model, wastefree=True, len_chain=20), N=500, ESSrmin=1.0, seed=1)
    a.run()
    ladder = np.linspace(0, 1, 25)[1:] ** 3
    b = pb.SMC(fk=ssp.Tempering(model=model, wastefree=False, len_chain=6, exponents=list(ladder)), N=10_000,
               ESSrmin=1.0, seed=2)
    b.run()
    assert b.t == len(ladder) and abs(a.logLt - b.logLt) < 0.5


@gpu
def test_fused_wastefree_move_vs_oracle():
    """The one-launch waste-free move (P-1 Metropolis steps of every chain) with injected noise against
    the oracle's MCMCSequenceWF: same accept/reject path for (almost) every chain, same output order."""
    torch = pytest.importorskip("torch")
    from particles_b200 import smc_samplers as ssp
    d, M, P = 5, 700, 9
    data = sp.synthetic_logistic(180, d, seed=5)
    r = np.random.RandomState(7)
    theta = r.randn(M, d)
    W = orc.exp_and_normalise(r.randn(M))
    z, u = r.standard_normal((P - 1, M, d)), r.rand(P - 1, M)
    m = sp.LogisticModel(data)
    fk = sp.AdaptiveTemperingWF(m, len_chain=P)
    xo = sp.ThetaParticles(theta=theta.copy())
    fk.target(0.45)(xo)
    fk.calibrate(W, xo)
    xs, x = [xo], xo
    for s in range(P - 1):                                  # MCMCSequenceWF with the injected draws
        x = x.copy()
        xprop = sp.ThetaParticles(theta=x.theta + z[s] @ x.shared["chol_cov"].T)
        fk.target(0.45)(xprop)
        pb = np.exp(np.clip(xprop.lpost - x.lpost, None, 0.0))
        x.copyto(xprop, where=u[s] < pb)
        xs.append(x)
    ref = sp.ThetaParticles.concatenate(*xs)
    mdev = ssp.LogisticRegression(data=data)
    xd = ssp.ThetaParticles(theta=torch.from_numpy(theta).cuda())
    mdev.target(xd, 0.45)
    ssp.ArrayRandomWalk().calibrate(torch.from_numpy(W).cuda(), xd)
    out = mdev.wf_move(xd, 0.45, P, noise=(z, u))
    assert out.theta.shape == (P * M, d)
    lp = host(out.lpost).reshape(P, M)
    same_chain = np.all(np.isclose(lp, ref.lpost.reshape(P, M), rtol=1e-9), axis=0)
    assert same_chain.mean() > 0.99                      # a draw within 1e-10 of its threshold may flip a chain
    th = host(out.theta).reshape(P, M, d)
    np.testing.assert_allclose(th[:, same_chain], ref.theta.reshape(P, M, d)[:, same_chain], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(host(out.llik).reshape(P, M)[:, same_chain], ref.llik.reshape(P, M)[:, same_chain],
                               rtol=1e-11, atol=1e-10)
    assert np.array_equal(th[0], theta)                  # generation row 0 = the resampled particles
