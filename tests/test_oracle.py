"""Independent pins for the CPU oracle (the reference ships no tests or golden vectors for this path; the pin
against the reference's own sources is tests/test_reference_golden.py).  Each restated function is checked against an INDEPENDENT
implementation: sklearn's Matern(nu=1.5) (the cross-check the reference's author used,
train/RandomFourierMaternFeatureMapper.py:69-82), sklearn's GaussianProcessRegressor, scipy's
multivariate normal, mpmath at 50 digits, finite differences, and the closed-form toy optimum."""
import math

import numpy as np
import pytest
import scipy.linalg
import scipy.stats

from oracle import gp_oracle as O


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def test_matern32_matches_sklearn():
    from sklearn.gaussian_process.kernels import Matern
    rs = np.random.RandomState(0)
    X, Y = rs.randn(300, 5), rs.randn(200, 5)
    ell = np.exp(rs.randn(5) * 0.5)
    got = O.matern32_kernel(O.NP, X, Y, stddev=ell)
    want = Matern(nu=1.5, length_scale=ell)(X, Y)
    assert rel(got, want) < 1e-12
    # self-Gram: the expanded-distance form leaves ~1e-8 distance noise on the diagonal (clamp 1e-32)
    K = O.matern32_kernel(O.NP, X, X, stddev=ell)
    assert np.max(np.abs(np.diag(K) - 1.0)) < 1e-12


def test_matern32_high_precision():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    rs = np.random.RandomState(1)
    X, Y = rs.randn(6, 3), rs.randn(4, 3)
    ell = np.array([0.7, 1.3, 2.0])
    got = O.matern32_kernel(O.NP, X, Y, stddev=ell)
    for i in range(6):
        for j in range(4):
            d2 = sum((mp.mpf(X[i, k]) / mp.mpf(ell[k]) - mp.mpf(Y[j, k]) / mp.mpf(ell[k])) ** 2 for k in range(3))
            v = mp.sqrt(3) * mp.sqrt(d2)
            want = (1 + v) * mp.exp(-v)
            assert abs(got[i, j] - float(want)) < 1e-14


def test_nll_matches_scipy_multivariate_normal():
    rs = np.random.RandomState(2)
    N, d = 120, 3
    theta, y = rs.randn(N, d), rs.randn(N, 1)
    p = O.build_params(d, "params", likhd_var_init=1e-2)
    p["log_scale"][:] = 0.3
    p["mean"][:] = -0.2
    p["log_bw"][:] = rs.randn(d) * 0.2
    b = O.TrainBatch(h_params=theta, obs=y)
    loss, L, k, ks = O.neg_log_marg_likhd(O.NP, p, "params", b, return_all=True)
    want = -scipy.stats.multivariate_normal(mean=np.full(N, -0.2), cov=ks).logpdf(y[:, 0])
    assert abs(float(loss) - want) < 1e-9 * abs(want)
    assert rel(L @ L.T, ks) < 1e-13
    assert np.allclose(np.diag(ks) - np.diag(k), 1e-2 + 1e-6, rtol=1e-12)      # sigma^2 + jitter


def test_posterior_matches_sklearn_gp():
    """Posterior mean / variance against sklearn's GaussianProcessRegressor with the same fixed
    kernel (no optimiser).  The reference does NOT add the noise to the predictive variance
    (log_gaus_likelihood.py:187: var = scale - colsum(A^2))."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import ConstantKernel, Matern
    rs = np.random.RandomState(3)
    N, d, M = 150, 2, 60
    theta, y, cand = rs.randn(N, d), rs.randn(N, 1), rs.randn(M, d)
    p = O.build_params(d, "params", likhd_var_init=0.05)
    p["log_scale"][:] = 0.4
    p["log_bw"][:] = [0.2, -0.1]
    b = O.TrainBatch(h_params=theta, obs=y)
    mean, var = O.predict(O.NP, p, "params", b, cand)
    ker = ConstantKernel(math.exp(0.4), "fixed") * Matern(nu=1.5, length_scale=np.exp(p["log_bw"]),
                                                         length_scale_bounds="fixed")
    gpr = GaussianProcessRegressor(kernel=ker, alpha=0.05 + 1e-6, optimizer=None).fit(theta, y[:, 0])
    m2, s2 = gpr.predict(cand, return_std=True)
    assert rel(mean[:, 0], m2) < 1e-9
    assert np.max(np.abs(var[:, 0] - s2 ** 2)) < 1e-9


def test_posterior_high_precision_small():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    rs = np.random.RandomState(4)
    N, d, M = 12, 2, 5
    theta, y, cand = rs.randn(N, d), rs.randn(N, 1), rs.randn(M, d)
    p = O.build_params(d, "params", likhd_var_init=1e-2)
    b = O.TrainBatch(h_params=theta, obs=y)
    mean, var = O.predict(O.NP, p, "params", b, cand)

    def k(a, c):
        r = mp.sqrt(sum((mp.mpf(a[i]) - mp.mpf(c[i])) ** 2 for i in range(d)))
        return (1 + mp.sqrt(3) * r) * mp.exp(-mp.sqrt(3) * r)
    K = mp.matrix(N, N)
    for i in range(N):
        for j in range(N):
            K[i, j] = k(theta[i], theta[j]) + ((mp.mpf("0.01") + mp.mpf("0.000001")) if i == j else 0)
    Kinv = K ** -1
    yv = mp.matrix([mp.mpf(v) for v in y[:, 0]])
    for c in range(M):
        ks = mp.matrix([k(theta[i], cand[c]) for i in range(N)])
        m = (ks.T * Kinv * yv)[0]
        v = 1 - (ks.T * Kinv * ks)[0]
        assert abs(mean[c, 0] - float(m)) < 1e-10
        assert abs(var[c, 0] - float(v)) < 1e-10


def test_acquisitions_closed_form():
    rs = np.random.RandomState(5)
    mean, std = rs.randn(1000, 1), np.abs(rs.randn(1000, 1)) + 1e-3
    y_max, xi = 0.3, 0.01
    ei = O.acq_ei(mean, std, y_max, xi)
    # E[max(f - y_max - xi, 0)] by quadrature for a few entries
    from scipy.integrate import quad
    for i in range(0, 1000, 97):
        m, s = float(mean[i, 0]), float(std[i, 0])
        want = quad(lambda f: max(f - y_max - xi, 0.0) * scipy.stats.norm.pdf(f, m, s), m - 12 * s, m + 12 * s,
                    points=[y_max + xi] if abs(y_max + xi - m) < 12 * s else None)[0]
        assert abs(ei[i] - want) < 1e-8
    assert np.array_equal(O.acq_ucb(mean, std, 2.58), (mean + 2.58 * std)[:, 0])
    assert np.array_equal(O.acq_lcb(mean, std, 2.58), (mean - 2.58 * std)[:, 0])
    assert np.array_equal(O.std_from_var(np.array([-1.0, 0.0, 4.0])), [1e-16, 1e-16, 2.0])


@pytest.mark.parametrize("kernel", ["params", "manual", "dist"])
def test_autograd_gradient_matches_finite_differences(kernel):
    rs = np.random.RandomState(6)
    T, per, d = 3, 8, 2
    N = T * per
    theta, y = rs.randn(N, d), rs.randn(N, 1)
    kw = {}
    b = O.TrainBatch(h_params=theta, obs=y, sizes=[per] * T)
    if kernel == "manual":
        p = O.build_params(d, "manual", in_dim_meta=4, likhd_var_init=0.05)
        b.embed = rs.rand(T, 4)
    elif kernel == "dist":
        p = O.build_params(d, "dist", in_dim_x=3, in_dim_y=1, weight_dim_x=[5, 4], weight_dim_y=[5, 3],
                           model_type="regression", embed_op="joint", likhd_var_init=0.05, seed=1)
        b.batch_X, b.batch_y, b.embed_sizes = rs.randn(30, 3), rs.rand(30, 1), [10, 12, 8]
        kw = dict(embed_op="joint", model_type="regression")
    else:
        p = O.build_params(d, "params", likhd_var_init=0.05)
    loss, grads = O.loss_and_grads(p, kernel, b, **kw)
    assert abs(loss - float(O.neg_log_marg_likhd(O.NP, p, kernel, b, **kw))) < 1e-10 * abs(loss)
    for name in p:
        flat = np.asarray(p[name], dtype=np.float64).reshape(-1)
        for idx in range(0, flat.size, max(1, flat.size // 3)):
            h = 1e-6
            vals = []
            for sgn in (+1, -1):
                q = {k: np.array(v, dtype=np.float64, copy=True) for k, v in p.items()}
                q[name].reshape(-1)[idx] += sgn * h
                vals.append(float(O.neg_log_marg_likhd(O.NP, q, kernel, b, **kw)))
            fd = (vals[0] - vals[1]) / (2 * h)
            g = np.asarray(grads[name]).reshape(-1)[idx]
            assert abs(fd - g) < 1e-5 * max(1.0, abs(g)), (name, idx, fd, g)


def test_segment_mean_and_embedding_layout():
    """psi_t = mean over the dataset's rows of vec(phi_x phi_y^T), index i*q + j; classification
    appends the class ratios (distbo_net.py:33-48)."""
    rs = np.random.RandomState(7)
    sizes = [5, 1, 9]
    X = rs.randn(15, 4)
    Y = np.eye(2)[rs.randint(0, 2, 15)]
    p = O.build_params(1, "dist", in_dim_x=4, in_dim_y=2, weight_dim_x=[6, 3], model_type="classification",
                       embed_op="joint", seed=3)
    E = O.distbo_net(O.NP, X, Y, sizes, p, embed_op="joint", model_type="classification")
    assert E.shape == (3, 3 * 2 + 2)
    phi = np.tanh(X @ p["weights_x_1"] + p["bias_x_1"]) @ p["weights_x_2"]
    off = np.r_[0, np.cumsum(sizes)]
    for t in range(3):
        rows = slice(off[t], off[t + 1])
        want = np.mean([np.outer(phi[r], Y[r]).reshape(-1) for r in range(off[t], off[t + 1])], axis=0)
        assert np.allclose(E[t, :6], want, rtol=1e-13, atol=1e-15)
        assert np.allclose(E[t, 6:], Y[rows].mean(0), rtol=1e-13)


def test_tf_adam_first_steps_by_hand():
    adam = O.TFAdam(0.01)
    x = {"w": np.array([1.0, -2.0])}
    g = {"w": np.array([0.5, -4.0])}
    x = adam.step(x, g)
    # t=1: m = 0.1 g, v = 0.001 g^2, lr_t = lr*sqrt(1-0.999)/(1-0.9); update = lr_t*m/(sqrt(v)+1e-8)
    lr_t = 0.01 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = np.array([1.0, -2.0]) - lr_t * (0.1 * g["w"]) / (np.sqrt(0.001 * g["w"] ** 2) + 1e-8)
    assert np.allclose(x["w"], want, rtol=1e-15)
    # global-norm clipping at 5 (train/train.py:33-35)
    adam2 = O.TFAdam(0.01)
    big = {"w": np.array([30.0, 40.0])}
    a = adam2.step({"w": np.zeros(2)}, big, clip_norm=5.0)
    b = O.TFAdam(0.01).step({"w": np.zeros(2)}, {"w": np.array([3.0, 4.0])})
    assert np.allclose(a["w"], b["w"], rtol=1e-15)


def test_toy_known_answer_end_to_end():
    """Known-answer check on the closed-form toy objective (model_embed/models.py:239-242): with
    observations of exp(-(theta - mu)^2 / 2), the GP posterior mean's arg-max over a dense grid
    lands next to mu."""
    rs = np.random.RandomState(8)
    mu = 1.7
    theta = rs.uniform(-8, 8, size=(40, 1))
    y = np.exp(-0.5 * (theta - mu) ** 2)
    from sklearn.preprocessing import StandardScaler
    sc = StandardScaler().fit(theta)
    p = O.build_params(1, "params", likhd_var_init=1e-4)
    p["log_bw"][:] = math.log(0.3)
    b = O.TrainBatch(h_params=sc.transform(theta), obs=y)
    grid = np.linspace(-8, 8, 3201)[:, None]
    mean, var = O.predict(O.NP, p, "params", b, sc.transform(grid))
    assert abs(grid[np.argmax(mean), 0] - mu) < 0.15


def test_concat_operator_matches_kernel_ridge_form():
    """'concat' pooling (distbo_net.py:49-89): the reference evaluates the conditional embedding operator through
    the Woodbury form; the pin is the n x n kernel-ridge form Psi^T (Phi Phi^T + lambda I)^-1 Phi quoted in its
    comments (Song, Fukumizu, Gretton 2013), evaluated independently here."""
    rs = np.random.RandomState(3)
    sizes = [40, 9, 3]
    S = sum(sizes)
    X, Y = rs.randn(S, 5), rs.rand(S, 1)
    p = O.build_params(2, "dist", in_dim_x=5, in_dim_y=1, weight_dim_x=[8, 4], weight_dim_y=[8, 3],
                       model_type="regression", embed_op="concat", seed=1)
    p["log_reg"] = np.asarray(-1.3)
    assert p["log_bw_embed"].shape == (4 * (3 + 1),)
    got = O.distbo_net(O.NP, X, Y, sizes, p, embed_op="concat", model_type="regression")
    lam = np.exp(-1.3)
    b = np.r_[0, np.cumsum(sizes)]
    for t, (s, e) in enumerate(zip(b[:-1], b[1:])):
        phi = O.nn_net(O.NP, p, X[s:e], "x")
        psi = O.nn_net(O.NP, p, Y[s:e], "y")
        ridge = psi.T @ np.linalg.solve(phi @ phi.T + lam * np.eye(e - s), phi)      # [q, p]
        np.testing.assert_allclose(got[t, :4], phi.mean(0), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(got[t, 4:], ridge.reshape(-1), rtol=1e-9, atol=1e-12)


def test_multi_task_kernel_restatement():
    """kernel 'multi': fill_triangular order is TF's documented example ([1..6] -> [[4,0,0],[6,5,0],[3,2,1]]);
    the task covariance has unit diagonal and is positive semi-definite; its autograd gradient agrees with
    central finite differences."""
    import torch
    idx, mask = O.fill_triangular_index(3)
    x = np.arange(1.0, 7.0)
    np.testing.assert_array_equal(x[idx] * mask, [[4, 0, 0], [6, 5, 0], [3, 2, 1]])
    for n in (1, 2, 7):
        idx, mask = O.fill_triangular_index(n)
        assert sorted(idx[mask > 0]) == list(range(n * (n + 1) // 2))         # a bijection onto the triangle
    rs = np.random.RandomState(0)
    T = 6
    x = rs.normal(size=T * (T + 1) // 2)
    k = O.task_kernel_matrix(O.NP, {"log_triangular_task": x}, T)
    np.testing.assert_allclose(np.diag(k), 1.0, rtol=0, atol=1e-15)
    np.testing.assert_allclose(k, k.T, atol=1e-15)
    assert np.linalg.eigvalsh(k).min() > -1e-12
    assert (k > 0).all()                                                       # "positive correlation" (:74)
    H = rs.randn(T, T)
    xt = torch.tensor(x, requires_grad=True)
    (O.task_kernel_matrix(O.TH(), {"log_triangular_task": xt}, T) * torch.from_numpy(H)).sum().backward()
    fd = np.zeros_like(x)
    for i in range(len(x)):
        d = np.zeros_like(x)
        d[i] = 1e-6
        fd[i] = ((O.task_kernel_matrix(O.NP, {"log_triangular_task": x + d}, T) * H).sum()
                 - (O.task_kernel_matrix(O.NP, {"log_triangular_task": x - d}, T) * H).sum()) / 2e-6
    np.testing.assert_allclose(xt.grad.numpy(), fd, rtol=1e-6, atol=1e-8)
    # expanded to observations: row/column t(i) of the task matrix; candidates sit in the last task
    sizes = [2, 1, 3, 1, 1, 2]
    t = np.repeat(np.arange(T), sizes)
    np.testing.assert_array_equal(O.task_kernel(O.NP, {"log_triangular_task": x}, sizes), k[np.ix_(t, t)])
    np.testing.assert_array_equal(O.task_kernel_pred(O.NP, {"log_triangular_task": x}, sizes, 4),
                                  np.repeat(k[t, -1][:, None], 4, axis=1))


def test_blr_restatement_pins():
    """algorithm='BLR' (log_gaus_likelihood_feature.py): the loss is -2 log N(y | 0, sigma^2 I + alpha Phi Phi^T)
    - N log 2 pi (scipy), the posterior is the textbook weight-space posterior, and the random Fourier map
    converges to sklearn's Matern(nu=1.5) kernel (the author's own check, RandomFourierMaternFeatureMapper.py:69-82)."""
    from scipy.stats import multivariate_normal
    from sklearn.gaussian_process.kernels import Matern
    from oracle import blr_oracle as BO
    rs = np.random.RandomState(4)
    N, d, T = 40, 3, 4
    sizes = [10] * T
    theta, y = rs.randn(N, d), rs.randn(N, 1)
    p, rff_in = BO.build_params_blr(d, "multi", [16, 12, 9], n_tasks=T, likhd_var_init=0.05, seed=3)
    assert rff_in == 9 and p["weights__1"].shape == (T + d, 16) and p["weights__3"].shape == (12, 9)
    p["log_alpha"] = np.asarray(0.3)
    p["bias__1"] = rs.randn(16) * 0.1
    p["bias__2"] = rs.randn(12) * 0.1
    batch = O.TrainBatch(h_params=theta, obs=y, sizes=sizes)
    loss = float(BO.blr_loss(O.NP, p, "multi", batch, rff_in, 0, n_tasks=T))
    fea = BO.train_feature(O.NP, p, "multi", batch, n_tasks=T)
    np.testing.assert_array_equal(fea[:, :T], np.eye(T)[np.repeat(np.arange(T), 10)])
    phi = BO.nn_features(O.NP, p, fea, rff_in, 0)
    sig2, alpha = np.exp(2 * p["log_sigma"]), np.exp(p["log_alpha"])
    cov = sig2 * np.eye(N) + alpha * phi @ phi.T
    want = -2.0 * multivariate_normal(np.zeros(N), cov).logpdf(y.ravel()) - N * np.log(2 * np.pi)
    assert loss == pytest.approx(want, rel=1e-10)
    # posterior: weights ~ N(0, alpha I)  =>  Sigma_w = (I/alpha + Phi^T Phi / sigma^2)^-1, m_w = Sigma_w Phi^T y / sigma^2
    cand = np.concatenate([np.eye(T)[[T - 1] * 6], rs.randn(6, d)], 1)
    phi_c = BO.nn_features(O.NP, p, cand, rff_in, 0)
    mean, var = BO.blr_predict(O.NP, p, phi, y, phi_c)
    Sw = np.linalg.inv(np.eye(9) / alpha + phi.T @ phi / sig2)
    np.testing.assert_allclose(mean, phi_c @ (Sw @ phi.T @ y / sig2), rtol=1e-9)
    np.testing.assert_allclose(var.ravel(), np.einsum("ij,jk,ik->i", phi_c, Sw, phi_c), rtol=1e-9)
    # random Fourier map: E[phi(x) phi(x')^T] = Matern-3/2(x, x')
    x = rs.randn(30, 2)
    feat = BO.nn_features(O.NP, {"log_bw": np.zeros(2)}, x, 2, 40000)
    assert feat.dtype == np.float64 and feat.shape == (30, 40000)
    assert np.max(np.abs(feat @ feat.T - Matern(length_scale=1.0, nu=1.5)(x))) < 0.03
