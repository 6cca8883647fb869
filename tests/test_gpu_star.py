"""-m gpu: BASELINE config #4 -- the STAR tower (PartitionedNorm + StarFCN, model_zoo/Star/*.py) on the fp32 path,
through the C-ABI (mamdr_star_train_step / mamdr_star_eval_step) against the CPU oracle (oracle/star.py); joint
training and under the MAMDR wrapper with meta_parms = ["emb", "kernel_shared", "bias_shared"].
Tolerances: single-batch gradients rel 2e-5 (summation order), parameters rel 1e-4 after N steps, AUC 1e-3."""
import numpy as np
import pytest
import torch

from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule
from oracle.meta import MetaSubset, OracleMAMDR, joint_train_epoch
from oracle.star import OracleStar, StarSpec

pytestmark = pytest.mark.gpu


def _cfg(**over):
    kw = {"model.name": "star", "model.norm": "pn", "model.dense": "star", "dataset.synthetic.scale": 0.05, "b200.precision": "fp32"}
    kw.update(over)
    return make_config(**kw)


def _weights(m):
    return m.layout.unpack(m.params.cpu().numpy())


def _oracle(base, weights=None, dtype=np.float32):
    spec = StarSpec(base.n_uid, base.n_pid, base.n_domain, (128, 128, 128), (256, 128, 64))
    assert spec.names == base.layout.names and [tuple(s) for s in spec.shapes] == base.layout.shapes
    return OracleStar(spec, weights if weights is not None else _weights(base.model), base.dataset.user_table,
                      base.dataset.item_table, lr=base.train_config['learning_rate'], dtype=dtype)


def _perturb(base, seed=0):
    m = base.model
    rng = np.random.default_rng(seed)
    w = _weights(m)
    for i, n in enumerate(m.layout.names):
        if n.startswith(('gamma', 'beta', 'bias')) or n == 'out_bias':
            w[i] = (w[i] + rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    m.params.copy_(torch.from_numpy(m.layout.pack(w)))
    return w


@pytest.mark.parametrize("rows", [1024, 333])
def test_star_train_step_matches_oracle(rows):
    import run
    base = run.build(_cfg())
    m = base.model
    w = _perturb(base)
    o = _oracle(base, w)
    dom = 3
    data = base.dataset.train_dataset[dom]['data']
    rows = min(rows, data.n_data)
    order = Schedule(1).batch_order(dom, data.n_data)
    data.set_order(order)
    loss = torch.zeros(1, device="cuda")
    m._train_step(data, 0, rows, loss)
    torch.cuda.synchronize()
    h = data.host
    sel = order[:rows]
    ol, _, og = o.gradients(h['uid'][sel], h['pid'][sel], dom, h['label'][sel])
    assert abs(loss.item() - ol) < 2e-5 * abs(ol)
    g = m.layout.unpack(m.grads.cpu().numpy())
    for name, a, b in zip(m.layout.names, g, og):
        if np.max(np.abs(b)) == 0:
            assert np.max(np.abs(a)) == 0, name          # untouched domain slices and domain_emb: exactly zero
        else:
            assert rel_err(a, b) < 2e-5, (name, rel_err(a, b))
    mm_, mv_ = m.moving_stats()
    np.testing.assert_allclose(mm_.cpu().numpy(), o.moving_mean, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(mv_.cpu().numpy(), o.moving_var, rtol=1e-4, atol=1e-8)
    o.adam.apply(o.weights, og)
    for name, a, b in zip(m.layout.names, _weights(m), o.weights):
        assert rel_err(a, b) < 1e-5, (name, rel_err(a, b))


def test_star_joint_training_and_eval_match_oracle():
    import run
    c = _cfg(**{"dataset.synthetic.scale": 0.1})
    base = run.build(c)
    m = base.model
    m.reset_optimizer()
    o = _oracle(base)
    seed = c['dataset']['seed']
    data = base.dataset.host_splits()
    base.schedule, osched = Schedule(seed), Schedule(seed)
    seq_g, seq_o = list(range(base.n_domain)), list(range(base.n_domain))
    for epoch in range(2):
        seq_g = base.schedule.shuffle_sequence(seq_g)
        base.stage_epoch_orders(list(seq_g))
        for idx in seq_g:
            m.reset_states()
            base.run_train_pass(idx)
        seq_o = joint_train_epoch(o, data, base.dataset.batch_size, osched, seq_o)
    for name, a, b in zip(m.layout.names, _weights(m), o.weights):
        assert rel_err(a, b) < 1e-4, (name, rel_err(a, b))
    l, a, dl, da = base.val_and_test("val")
    for k in range(base.n_domain):
        hv = data['val'][k]
        ol, oa = o.evaluate(hv['uid'], hv['pid'], k, hv['label'], base.dataset.batch_size)
        assert abs(dl[k] - ol) < 1e-4 * abs(ol) and abs(da[k] - oa) < 1e-3, (k, dl[k], ol, da[k], oa)


def test_star_under_mamdr_matches_oracle():
    """star_meta_mamdr_finetune: only emb / kernel_shared / bias_shared are meta parameters; the specific tensors, the
    PartitionedNorm tensors and the output layer train continuously outside DN / DR."""
    import run
    c = _cfg(**{"model.name": "star_meta_mamdr_finetune", "train.meta_parms": ["emb", "kernel_shared", "bias_shared"],
                "dataset.synthetic.scale": 0.05})
    wrapper = run.build(c)
    wrapper.prepare()
    base = wrapper.base_model
    meta_names = [p.name for p in wrapper.model_meta_parms]
    assert len(meta_names) == 1 + 2 * 3 and all(("emb" in n) or ("_shared" in n) for n in meta_names)
    assert not any("gamma" in n or "beta" in n for n in meta_names)
    # the oracle's MAMDR restates the list-of-arrays algebra on the meta subset only
    o = _oracle(base, wrapper.model.layout.unpack(wrapper.model.params.cpu().numpy()))
    idx = [i for i, p in enumerate(wrapper.model.trainable_weights) if p.name in meta_names]
    om = OracleMAMDR(MetaSubset(o, idx), base.dataset.host_splits(), c['train'], 1024, Schedule(123),
                     {k: [v.numpy()[i] for i in idx] for k, v in wrapper.domain_weights.items()}, name=c['model']['name'])
    om.meta_weights = [wrapper.meta_weights.numpy()[i].copy() for i in idx]
    base.schedule = Schedule(123)
    names = wrapper.model.layout.names
    # theta after one meta-step: rel 5e-4 (kernels ~1e-7; the 64-wide bias_shared2, values ~1e-3, sits at 1.2e-4).  theta_d / the live non-meta tensors: 5e-2, and 1e-1
    # after two meta-steps -- STAR's effective kernels are products of two Glorot draws, so the deep pre-activations are
    # ~1e-4 with half of the gates closed; tests/diag_trace_star.py shows the live weights tracking the fp64 oracle at 1e-7
    # for 40 passes until ONE ReLU gate of 25 152 (pre-activation ~1e-9) opens on one side only, which changes that
    # step's gradients by ~1 % (single-step and joint-training parity above stay at 2e-5 / 1e-4).
    for e, tol_theta, tol in ((0, 5e-4, 5e-2),):   # one meta-step: after the event the specific tensors (sparse per-domain
        # Adam-normalised updates) diverge quickly -- 21 % after a second meta-step -- so parity is judged here
        wrapper.train_epoch(e)
        om.train_epoch()
        for k, i in enumerate(idx):
            if names[i] == 'domain_emb':
                continue   # zero gradient by construction
            assert rel_err(wrapper.meta_weights.numpy()[i], om.meta_weights[k]) < tol_theta, ("theta", e, names[i])
            for d in om.domain_weights:
                assert rel_err(wrapper.domain_weights[d].numpy()[i], om.domain_weights[d][k]) < tol, ("theta_%d" % d, e, names[i])
        live = _weights(wrapper.model)
        for i, n_ in enumerate(names):
            if i not in idx:
                assert rel_err(live[i], o.weights[i]) < tol, ("live non-meta tensor", e, n_, rel_err(live[i], o.weights[i]))
    l, a, dl, da = wrapper.val_and_test("val")
    ol, oa, odl, oda = om.val_and_test("val")
    assert abs(a - oa) < 5e-3 and abs(l - ol) < 1e-3 * abs(ol)


def _pn_pack(o):
    """OracleStar's non-trainable PartitionedNorm state in the device layout [moving_mean | moving_var | biased_mean | biased_var]."""
    return np.concatenate([o.moving_mean.ravel(), o.moving_var.ravel(), o.biased_mean.ravel(), o.biased_var.ravel()]).astype(np.float32)


def test_star_teacher_forced_meta_step_tracks_the_oracle_pass_by_pass():
    """star_meta_mamdr_finetune, one whole meta-step (10 DN passes + 10 x 3 x 2 DR passes): before EVERY pass the oracle's state --
    all variables (meta and non-meta), Adam slots / beta powers / step, the PartitionedNorm moving statistics and their update
    counters -- is loaded into the device model; after the pass every variable must agree with the oracle's to 1e-5 (kernels) / 1e-4
    (the small PartitionedNorm / bias / output tensors).  A pass beyond that bound is a ReLU-gate event (STAR's deep pre-activations
    are ~1e-4 with half of the gates closed): it must stay within 5e-2 and be rare.  This is the evidence behind the loose
    free-running bars of test_star_under_mamdr_matches_oracle."""
    import run
    import oracle.meta as ometa
    c = _cfg(**{"model.name": "star_meta_mamdr_finetune", "train.meta_parms": ["emb", "kernel_shared", "bias_shared"],
                "dataset.synthetic.scale": 0.1, "train.sample_num": 2})
    wrapper = run.build(c)
    wrapper.prepare()
    base, m = wrapper.base_model, wrapper.base_model.model
    lo = m.layout
    meta_names = [p.name for p in wrapper.model_meta_parms]
    o = _oracle(base, lo.unpack(m.params.cpu().numpy()))
    idx = [i for i, p in enumerate(m.trainable_weights) if p.name in meta_names]
    om = OracleMAMDR(MetaSubset(o, idx), base.dataset.host_splits(), c['train'], 1024, Schedule(123),
                     {k: [v.numpy()[i] for i in idx] for k, v in wrapper.domain_weights.items()}, name=c['model']['name'])
    om.meta_weights = [wrapper.meta_weights.numpy()[i].copy() for i in idx]
    passes = []
    orig = ometa.train_pass

    def tp(mdl, d, domain, order, batch_size, max_steps=0, optimizer='adam', sgd_lr=None):
        ad = o.adam
        rec = {"domain": domain, "order": np.asarray(order).copy(), "w0": [x.copy() for x in o.weights], "m0": [x.copy() for x in ad.m],
               "v0": [x.copy() for x in ad.v], "opt0": (ad.step, float(ad.b1pow), float(ad.b2pow)), "pn0": _pn_pack(o), "pns0": o.pn_steps.copy()}
        r = orig(mdl, d, domain, order, batch_size, max_steps, optimizer, sgd_lr)
        rec.update({"w1": [x.copy() for x in o.weights], "pn1": _pn_pack(o), "steps": r[2]})
        passes.append(rec)
        return r
    ometa.train_pass = tp
    try:
        om.train_epoch()
    finally:
        ometa.train_pass = orig
    assert len(passes) == 10 + 10 * 3 * 2
    worst_clean, flagged = 0.0, []
    for k, rec in enumerate(passes):
        data = base.dataset.train_dataset[rec["domain"]]['data']
        m.params.copy_(torch.from_numpy(lo.pack(rec["w0"])))
        m.m.copy_(torch.from_numpy(lo.pack(rec["m0"])))
        m.v.copy_(torch.from_numpy(lo.pack(rec["v0"])))
        m.set_opt_words(torch.tensor(rec["opt0"], dtype=torch.float64))
        pn_f, pn_steps = m.pn_parts()
        pn_f.copy_(torch.from_numpy(rec["pn0"]))
        pn_steps[:base.n_domain].copy_(torch.from_numpy(rec["pns0"].astype(np.int32)))
        data.set_order(rec["order"])
        m.fit_pass(data, rec["steps"])
        torch.cuda.synchronize()
        errs = {n_: rel_err(a, b) for n_, a, b in zip(lo.names, lo.unpack(m.params.cpu().numpy()), rec["w1"]) if float(np.max(np.abs(b))) > 0}
        errs["pn_state"] = rel_err(m.pn_parts()[0].cpu().numpy(), rec["pn1"])
        bad = [n_ for n_, e in errs.items() if e > (1e-5 if n_.startswith("kernel") else 1e-4)]
        if bad:
            assert max(errs.values()) < 5e-2, (k, errs)
            flagged.append((k, max(errs.values())))
        else:
            worst_clean = max(worst_clean, max(errs.values()))
    print("STAR: %d passes teacher-forced; worst clean error %.2e; gate events %s" % (len(passes), worst_clean, flagged))
    assert len(flagged) <= len(passes) // 10, flagged


def test_star_sgd_step_matches_oracle():
    """The finetune stage's plain SGD (specific_base_model.py:118-122) on the STAR tower: one mini-batch, every variable vs the
    oracle (the gradient arena is fully written: zero outside the batch's domain slices)."""
    import run
    base = run.build(_cfg())
    m = base.model
    w = _perturb(base)
    o = _oracle(base, w)
    dom = 4
    data = base.dataset.train_dataset[dom]['data']
    rows = min(1024, data.n_data)
    order = Schedule(2).batch_order(dom, data.n_data)
    data.set_order(order)
    m.compile(optimizer="sgd", lr=0.001)
    loss = torch.zeros(1, device="cuda")
    m._train_step(data, 0, rows, loss)
    torch.cuda.synchronize()
    h, sel = data.host, order[:rows]
    ol, _ = o.train_on_batch(h['uid'][sel], h['pid'][sel], dom, h['label'][sel], optimizer='sgd', sgd_lr=0.001)
    assert abs(loss.item() - ol) < 2e-5 * abs(ol)
    for n_, a, b in zip(m.layout.names, _weights(m), o.weights):
        assert rel_err(a, b) < 1e-5, (n_, rel_err(a, b))
    m.compile(optimizer="adam")
