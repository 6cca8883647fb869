"""CPU: the multi-task tower oracle (oracle/mtl.py) -- manual backward vs torch.autograd in float64 for MMOE, PLE (CGC) and
SharedBottom; sub-model reachability and the shared-optimizer semantics of deep_mtl_ctr.py:53-65."""
import numpy as np
import pytest
import torch

from oracle.mtl import MTLSpec, OracleMTL, init_mtl_weights


def _spec(kind, trainable=True):
    return MTLSpec(30, 25, 3, kind=kind, emb_dim=(8, 8, 4), expert_hidden=(12, 8), tower_hidden=(8,), gate_hidden=(4,),
                   num_experts=3, specific_expert_num=2, shared_expert_num=2, dropout=0.5, emb_trainable=trainable)


def _problem(kind, trainable=True, b=29, seed=0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    spec = _spec(kind, trainable)
    w = init_mtl_weights(spec, 5)
    for i, n in enumerate(spec.names):
        if n.endswith('_emb'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.3).astype(np.float32)
        elif 'bias' in n:
            w[i] = (rng.standard_normal(w[i].shape) * 0.1).astype(np.float32)
    ut = (rng.standard_normal((30, 8)) * 0.3).astype(np.float32)
    it = (rng.standard_normal((25, 8)) * 0.3).astype(np.float32)
    m = OracleMTL(spec, w, None if trainable else ut, None if trainable else it, dtype=dtype)
    uid, pid = rng.integers(0, 30, b), rng.integers(0, 25, b)
    y = (rng.random(b) < 0.4).astype(np.float64)
    return spec, m, uid, pid, y, ut, it


def _masks(spec, t, b, rng):
    mk = {}
    for e in spec.expert_sets[t]:
        for l, n in enumerate(spec.expert_hidden):
            mk[('expert', e, l)] = (rng.random((b, n)) < 0.5) * 2.0
    for l, n in enumerate(spec.gate_hidden):
        mk[('gate', t, l)] = (rng.random((b, n)) < 0.5) * 2.0
    for l, n in enumerate(spec.tower_hidden):
        mk[('tower', t, l)] = (rng.random((b, n)) < 0.5) * 2.0
    return mk


def _torch_loss(spec, W, ut, it, uid, pid, y, t, mk):
    T = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    Eu, Ei = (W['user_emb'], W['item_emb']) if spec.emb_trainable else (T(ut), T(it))
    b = len(uid)
    X = torch.cat([Eu[uid], Ei[pid], W['domain_emb'][t].expand(b, -1)], dim=1)

    def dnn(prefix, kind, idx, H, widths):
        for l in range(len(widths)):
            H = torch.relu(H @ W['%s_kernel%d' % (prefix, l)] + W['%s_bias%d' % (prefix, l)]) * T(mk[(kind, idx, l)])
        return H
    outs = [dnn('expert%d' % e, 'expert', e, X, spec.expert_hidden) for e in spec.expert_sets[t]]
    if spec.has_gate:
        a = torch.softmax(dnn('gate%d' % t, 'gate', t, X, spec.gate_hidden) @ W['gate%d_out' % t], dim=1)
        mix = sum(a[:, j:j + 1] * outs[j] for j in range(spec.k))
    else:
        mix = outs[0]
    z = (dnn('tower%d' % t, 'tower', t, mix, spec.tower_hidden) @ W['tower%d_out' % t])[:, 0] + W['bias%d' % t][0]
    p = torch.sigmoid(z)
    yt = T(y)
    loss = -(yt * torch.log(p) + (1 - yt) * torch.log(1 - p)).mean()
    reg = (W['domain_emb'] ** 2).sum()
    reg = reg + ((W['user_emb'] ** 2).sum() + (W['item_emb'] ** 2).sum() if spec.emb_trainable else (T(ut) ** 2).sum() + (T(it) ** 2).sum())
    return loss + spec.l2_emb * reg


@pytest.mark.parametrize("kind,trainable", [("mmoe", True), ("ple", True), ("shared_bottom", True), ("mmoe", False)])
def test_backward_matches_autograd(kind, trainable):
    spec, m, uid, pid, y, ut, it = _problem(kind, trainable)
    t = 1
    mk = _masks(spec, t, len(uid), np.random.default_rng(7))
    loss, p, g = m.gradients(uid, pid, t, y, masks=mk)
    W = {n: torch.tensor(w, dtype=torch.float64, requires_grad=True) for n, w in zip(spec.names, m.weights)}
    lt = _torch_loss(spec, W, ut, it, uid, pid, y, t, mk)
    lt.backward()
    assert abs(loss - lt.item()) < 1e-10
    reach = set(spec.reachable(t))
    assert set(g.keys()) == reach
    for n in spec.names:
        if n in reach:
            np.testing.assert_allclose(g[n], W[n].grad.numpy(), rtol=1e-9, atol=1e-12, err_msg=n)
        else:   # not reachable from output t: no gradient at all
            assert W[n].grad is None or not W[n].grad.abs().max() > 0, n


def test_expert_sets_and_layout():
    sp = _spec('ple')
    assert sp.n_experts == 2 + 3 * 2 and sp.k == 4
    assert sp.expert_sets[1] == [4, 5, 0, 1]          # specific experts first, then the shared ones
    names = sp.names
    # domain block t is contiguous: specific experts, gate, gate out, tower, tower out, bias
    i0, i1 = names.index('expert4_kernel0'), names.index('bias1')
    assert all(n.startswith(('expert4_', 'expert5_', 'gate1_', 'tower1_')) for n in names[i0:i1])
    sp = _spec('mmoe')
    assert sp.expert_sets[2] == [0, 1, 2] and sp.k == 3
    sp = _spec('shared_bottom')
    assert not sp.has_gate and not any(n.startswith('gate') for n in sp.names)


def test_shared_optimizer_touches_only_the_sub_model():
    spec, m, uid, pid, y, ut, it = _problem('ple', dtype=np.float32)
    before = m.get_weights()
    m.train_on_batch(uid, pid, 0, y)
    reach = set(spec.reachable(0))
    for n, w0, w1, mm_, vv in zip(spec.names, before, m.weights, m.adam.m, m.adam.v):
        if n in reach:
            assert np.any(w0 != w1) or n.endswith('bias0') or 'bias' in n, n
        else:
            assert np.array_equal(w0, w1) and not mm_.any() and not vv.any(), n
    assert m.adam.step == 1 and m.adam.b1pow == np.float32(0.9) * np.float32(0.9)
    # a step on another domain advances the same beta powers and leaves domain 0's private slots alone
    m0 = [x.copy() for x in m.adam.m]
    m.train_on_batch(uid, pid, 2, y)
    assert m.adam.step == 2
    i = spec.names.index('gate0_kernel0')
    assert np.array_equal(m.adam.m[i], m0[i])
    j = spec.names.index('expert0_kernel0')            # shared expert: trained by both
    assert np.any(m.adam.m[j] != m0[j])


def test_eval_has_no_dropout_and_softmax_rows_sum_to_one():
    spec, m, uid, pid, y, ut, it = _problem('mmoe', dtype=np.float32)
    c, p = m.forward(uid, pid, 0, train=False)
    np.testing.assert_allclose(c['a'].sum(axis=1), 1.0, rtol=1e-6)
    c2, p2 = m.forward(uid, pid, 0, train=False)
    assert np.array_equal(p, p2)
    loss, auc = m.evaluate(uid, pid, 0, y, batch_size=16)
    assert np.isfinite(loss) and 0.0 <= auc <= 1.0
