"""The oracle's manual backward against torch.autograd in float64, finite differences,
and the TF-Adam formula.  (parity unpinned vs TF itself -- see oracle/__init__.py.)"""
import numpy as np
import torch

from oracle import philox
from oracle.mlp import AdamState, MLPSpec, OracleMLP


def _problem(emb_trainable, dtype=np.float64, b=37, seed=0):
    rng = np.random.default_rng(seed)
    spec = MLPSpec(n_uid=50, n_pid=40, n_domain=3, emb_dim=(8, 8, 8), hidden=(16, 12, 8), dropout=0.5,
                   emb_trainable=emb_trainable)
    ws = []
    for name, shape in zip(spec.names, spec.shapes):
        scale = 0.5 if 'kernel' in name else 0.3
        ws.append(rng.standard_normal(shape) * scale)
    ut, it = rng.standard_normal((50, 8)) * 0.3, rng.standard_normal((40, 8)) * 0.3
    m = OracleMLP(spec, ws, None if emb_trainable else ut, None if emb_trainable else it, dtype=dtype)
    uid = rng.integers(0, 50, b)
    uid[:5] = uid[0]  # duplicates exercise the scatter-add
    pid = rng.integers(0, 40, b)
    y = (rng.random(b) < 0.4).astype(np.float64)
    return spec, m, uid, pid, y, ut, it


def _torch_loss(spec, weights, ut, it, uid, pid, dom, y, masks):
    W = dict(zip(spec.names, weights))
    Eu = W['user_emb'] if spec.emb_trainable else torch.tensor(ut)
    Ei = W['item_emb'] if spec.emb_trainable else torch.tensor(it)
    b = len(uid)
    X = torch.cat([Eu[uid], Ei[pid], W['domain_emb'][dom].expand(b, -1)], dim=1)
    H = X
    for l in range(len(spec.hidden)):
        H = torch.relu(H @ W['kernel%d' % l] + W['bias%d' % l]) * torch.tensor(masks[l])
    s = (H @ W['dense_kernel'])[:, 0] + W['global_bias'][0]
    p = torch.sigmoid(s)
    ph = torch.clamp(p, 1e-7, 1 - 1e-7)
    lg = torch.log(ph / (1 - ph))
    yt = torch.tensor(y)
    bce = torch.clamp(lg, min=0) - lg * yt + torch.log1p(torch.exp(-lg.abs()))
    reg = spec.l2_emb * (W['domain_emb'] ** 2).sum()
    if spec.emb_trainable:
        reg = reg + spec.l2_emb * ((Eu ** 2).sum() + (Ei ** 2).sum())
    return bce.mean() + reg


def _check_against_autograd(emb_trainable):
    spec, m, uid, pid, y, ut, it = _problem(emb_trainable)
    masks = [philox.dropout_mask(len(uid), h, 1024 + l, 0, 0.5, np.float64) for l, h in enumerate(spec.hidden)]
    loss, p, grads = m.gradients(uid, pid, 1, y, masks=masks)
    tw = [torch.tensor(w, requires_grad=True) for w in m.weights]
    tl = _torch_loss(spec, tw, ut, it, torch.tensor(uid), torch.tensor(pid), 1, y, masks)
    tl.backward()
    const = 0.0 if emb_trainable else m.frozen_reg
    assert abs(loss - (tl.item() + const)) < 1e-12
    for name, g, t in zip(spec.names, grads, tw):
        np.testing.assert_allclose(g, t.grad.numpy(), rtol=1e-10, atol=1e-13, err_msg=name)


def test_backward_matches_autograd_frozen():
    _check_against_autograd(False)


def test_backward_matches_autograd_trainable_tables():
    _check_against_autograd(True)


def test_finite_differences():
    spec, m, uid, pid, y, ut, it = _problem(False)
    masks = [philox.dropout_mask(len(uid), h, 1024 + l, 0, 0.5, np.float64) for l, h in enumerate(spec.hidden)]
    _, _, grads = m.gradients(uid, pid, 2, y, masks=masks)
    rng = np.random.default_rng(1)
    for ti, name in enumerate(spec.names):
        w = m.weights[ti]
        for _ in range(3):
            idx = tuple(rng.integers(0, s) for s in w.shape)
            old = w[idx]
            eps = 1e-6
            w[idx] = old + eps
            lp = m.loss_from_p(m.forward(uid, pid, 2, True, masks)[1], y)
            w[idx] = old - eps
            lm = m.loss_from_p(m.forward(uid, pid, 2, True, masks)[1], y)
            w[idx] = old
            assert abs((lp - lm) / (2 * eps) - grads[ti][idx]) < 1e-6, name


def test_clip_kills_gradient():
    spec, m, uid, pid, y, ut, it = _problem(False, dtype=np.float32)
    m.w('global_bias')[0] = 40.0  # p == 1.0 in fp32 -> outside [1e-7, 1-1e-7]
    _, p, grads = m.gradients(uid, pid, 0, y)
    assert np.all(p > np.float32(1 - 1e-7))
    assert grads[spec.names.index('global_bias')][0] == 0.0
    assert np.all(grads[spec.names.index('kernel0')] == 0.0)


def test_adam_tf_formula():
    rng = np.random.default_rng(0)
    w = [rng.standard_normal((5, 3)).astype(np.float32)]
    w0 = w[0].copy()
    st = AdamState(w, lr=1e-3)
    m = np.zeros_like(w0, dtype=np.float64)
    v = np.zeros_like(w0, dtype=np.float64)
    ref = w0.astype(np.float64)
    for t in range(1, 6):
        g = rng.standard_normal((5, 3)).astype(np.float32)
        st.apply(w, [g])
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g.astype(np.float64) ** 2
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
        ref = ref - lr_t * m / (np.sqrt(v) + 1e-8)   # epsilon-hat form
        np.testing.assert_allclose(w[0], ref, rtol=2e-5, atol=1e-7)
    assert st.step == 5
    assert abs(float(st.b1pow) - 0.9 ** 6) < 1e-6


def test_first_adam_step_is_sign_update():
    w = [np.ones((4,), dtype=np.float32)]
    st = AdamState(w, lr=1e-3)
    st.apply(w, [np.asarray([0.5, -2.0, 1e-3, 0.0], dtype=np.float32)])
    np.testing.assert_allclose(w[0], [1 - 1e-3, 1 + 1e-3, 1 - 1e-3, 1.0], rtol=1e-5)


def test_evaluate_mean_of_batch_means_and_ragged_tail():
    spec, m, uid, pid, y, ut, it = _problem(False, dtype=np.float32, b=23)
    loss, auc = m.evaluate(uid, pid, 0, y, batch_size=10)
    parts = []
    for s in (0, 10, 20):
        _, p = m.forward(uid[s:s + 10], pid[s:s + 10], 0, False)
        parts.append(m.loss_from_p(p, y[s:s + 10].astype(np.float32)))
    assert abs(loss - np.mean(parts)) < 1e-7
    assert 0.0 <= auc <= 1.0
