"""CPU: the STAR oracle (oracle/star.py) -- manual backward vs torch.autograd in float64, PartitionedNorm statistics,
zero-debiased moving averages, layout order."""
import numpy as np
import torch

from oracle.star import OracleStar, StarSpec, init_star_weights, PN_EPS, PN_MOMENTUM


def _problem(b=37, dtype=np.float64, seed=0):
    rng = np.random.default_rng(seed)
    spec = StarSpec(50, 40, 4, (8, 8, 4), (12, 6))
    w = init_star_weights(spec, 3)
    for i, n in enumerate(spec.names):   # move off ones / zeros so every product term is exercised
        if n.startswith(('gamma', 'beta', 'bias')) or n == 'out_bias':
            w[i] = w[i] + rng.standard_normal(w[i].shape).astype(np.float32) * 0.1
    ut = rng.standard_normal((50, 8)).astype(np.float32) * 0.3
    it = rng.standard_normal((40, 8)).astype(np.float32) * 0.3
    m = OracleStar(spec, w, ut, it, dtype=dtype)
    uid, pid = rng.integers(0, 50, b), rng.integers(0, 40, b)
    y = (rng.random(b) < 0.4).astype(np.float64)
    return spec, m, uid, pid, y, ut, it


def test_backward_matches_autograd():
    spec, m, uid, pid, y, ut, it = _problem()
    dom = 2
    loss, p, grads = m.gradients(uid, pid, dom, y, update_stats=False)
    W = {n: torch.tensor(w, dtype=torch.float64, requires_grad=True) for n, w in zip(spec.names, m.weights)}
    b = len(uid)
    X = torch.cat([torch.tensor(ut, dtype=torch.float64)[uid], torch.tensor(it, dtype=torch.float64)[pid]], dim=1)
    mu = X.mean(0)
    var = ((X - mu) ** 2).mean(0)
    xhat_ui = (X - mu) / torch.sqrt(var + PN_EPS)
    xhat = torch.cat([xhat_ui, torch.zeros(b, 4, dtype=torch.float64)], dim=1)   # constant E_d columns -> 0
    gamma = W['gamma_shared'] * W['gamma_specific'][dom]
    beta = W['beta_shared'] + W['beta_specific'][dom]
    H = xhat * gamma + beta
    for l in range(2):
        Wl = W['kernel_shared%d' % l] * W['kernel_specific%d' % l][dom]
        bl = W['bias_shared%d' % l] + W['bias_specific%d' % l][dom]
        H = torch.relu(H @ Wl + bl)
    z = (H @ W['out_kernel'])[:, 0] + W['out_bias'][0]
    pt = torch.sigmoid(z)
    yt = torch.tensor(y)
    lt = -(yt * torch.log(pt) + (1 - yt) * torch.log(1 - pt)).mean()
    lt.backward()
    assert abs(loss - lt.item()) < 1e-10
    for n, g in zip(spec.names, grads):
        ref = W[n].grad.numpy() if W[n].grad is not None else np.zeros_like(g)
        np.testing.assert_allclose(g, ref, rtol=1e-9, atol=1e-12, err_msg=n)
    # only the batch's domain slice of the specific tensors gets a gradient
    assert np.all(grads[spec.names.index('kernel_specific0')][[0, 1, 3]] == 0)
    assert np.all(grads[spec.names.index('domain_emb')] == 0)


def test_partitioned_norm_statistics_and_moving_average():
    spec, m, uid, pid, y, ut, it = _problem(dtype=np.float32)
    dom = 1
    H, p, c = m.forward(uid, pid, dom, train=True)
    X = np.concatenate([ut[uid], it[pid]], axis=1)
    np.testing.assert_allclose(c['mu'][:16], X.mean(0), rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(c['var'][:16], X.var(0), rtol=1e-4, atol=1e-8)
    assert np.all(c['var'][16:] == 0) and np.all(c['xhat'][:, 16:] == 0)
    for t in range(1, 4):
        m.gradients(uid, pid, dom, y)
        # zero-debiased EMA of a CONSTANT input equals the input from the first step on
        np.testing.assert_allclose(m.moving_mean[dom][:16], X.mean(0), rtol=1e-4, atol=1e-6)
        assert m.pn_steps[dom] == t and m.pn_steps[0] == 0
    np.testing.assert_allclose(m.biased_mean[dom][:16], X.mean(0) * (1 - PN_MOMENTUM ** 3), rtol=1e-3, atol=1e-7)
    # inference uses the moving statistics of the batch's domain; other domains still hold (0, 1)
    assert np.all(m.moving_var[0] == 1) and np.all(m.moving_mean[0] == 0)
    l0, a0 = m.evaluate(uid, pid, dom, y, batch_size=16)
    assert np.isfinite(l0) and 0.0 <= a0 <= 1.0


def test_trainable_weight_order_and_sizes_taobao20():
    spec = StarSpec(58190, 16319, 20)
    assert spec.names[:5] == ['domain_emb', 'gamma_specific', 'beta_specific', 'gamma_shared', 'beta_shared']
    n = sum(int(np.prod(s)) for s in spec.shapes)
    assert n == 20 * 128 + 2 * 20 * 384 + 2 * 384 + 21 * (384 * 256 + 256) + 21 * (256 * 128 + 128) + 21 * (128 * 64 + 64) + 65
