"""-m gpu: the mlp tower mini-batch (fp32 mode) and the DN / MAMDR meta loops against the CPU oracle on
identical synthetic inputs, seeds, schedules and dropout masks.

Tolerances (BASELINE.json north_star): parameters rel 1e-4 after N meta-steps in fp32 mode, AUC within
1e-3; single-batch quantities are held tighter (rel 2e-5) because only summation order differs.
"""
import numpy as np
import pytest
import torch

from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule
from oracle import philox
from oracle.meta import OracleDN, OracleMAMDR
from oracle.mlp import MLPSpec, OracleMLP

pytestmark = pytest.mark.gpu


def _build(config):
    import run
    return run.build(config)


def _oracle_for(wrapper, weights=None, dtype=np.float32):
    base = wrapper.base_model if hasattr(wrapper, "base_model") else wrapper
    mc = base.model_config
    spec = MLPSpec(base.n_uid, base.n_pid, base.n_domain, (mc['user_dim'], mc['item_dim'], mc['domain_dim']),
                   tuple(mc['hidden_dim']), dropout=mc['dropout'], emb_trainable=base.emb_trainable)
    w = weights if weights is not None else base.layout.unpack(base.model.params.cpu().numpy())
    return OracleMLP(spec, w, base.dataset.user_table, base.dataset.item_table, lr=base.train_config['learning_rate'],
                     dtype=dtype)


def _weights(model):
    return model.layout.unpack(model.params.cpu().numpy())


PRECISIONS = ["fp32", "tf32x3"]


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("rows", [1024, 977, 1, 33])
def test_forward_eval_matches_oracle(rows, prec):
    base = _build(make_config(**{"model.name": "mlp", "dataset.synthetic.scale": 0.05, "b200.precision": prec}))
    o = _oracle_for(base)
    data = base.dataset.val_dataset[0]['data']
    rows = min(rows, data.n_data)
    probs, loss = base.model.predict(data, 0, rows)
    h = data.host
    _, p = o.forward(h['uid'][:rows], h['pid'][:rows], 0, train=False)
    np.testing.assert_allclose(probs.cpu().numpy(), p, rtol=2e-5, atol=1e-7)
    assert abs(loss.item() - o.loss_from_p(p, h['label'][:rows])) < 2e-5 * abs(loss.item())


@pytest.mark.parametrize("prec", PRECISIONS)
def test_dropout_mask_bits_match_oracle(prec):
    """With kernel = 0 and bias = 1 every hidden unit is relu(1) * M = M: the activations ARE the mask."""
    base = _build(make_config(**{"model.name": "mlp", "dataset.synthetic.scale": 0.05, "b200.precision": prec}))
    m = base.model
    w = _weights(m)
    names = m.layout.names
    for l in range(3):
        w[names.index('kernel%d' % l)][...] = 0
        w[names.index('bias%d' % l)][...] = 1
    m.params.copy_(torch.from_numpy(m.layout.pack(w)))
    data = base.dataset.train_dataset[1]['data']
    rows = min(1000, data.n_data)
    loss = torch.zeros(1, device="cuda")
    for step in range(3):
        m._train_step(data, 0, rows, loss)
        # read H_1..H_3 back from the workspace through the documented layout: easier -- compare dZ-free
        # quantity: grads of bias via oracle with the same masks
        o = _oracle_for(base, weights=w)
        o.adam.step = step
        masks = [philox.dropout_mask(rows, h, 1024 + l, step, 0.5) for l, h in enumerate((256, 128, 64))]
        h = data.host
        _, _, og = o.gradients(h['uid'][:rows], h['pid'][:rows], 1, h['label'][:rows], masks=masks)
        g = m.layout.unpack(m.grads.cpu().numpy())
        # dense_kernel grad = sum_r H3[r,c]*ds[r] = sum_r M3[r,c]*ds[r]: sensitive to every mask bit
        np.testing.assert_allclose(g[names.index('dense_kernel')], og[names.index('dense_kernel')], rtol=1e-4, atol=1e-8)
        # undo the Adam update so the next step sees the same weights but a new step counter
        m.params.copy_(torch.from_numpy(m.layout.pack(w)))


@pytest.mark.parametrize("prec", PRECISIONS)
@pytest.mark.parametrize("rows", [1024, 977, 130])
def test_train_step_gradients_match_oracle(rows, prec):
    base = _build(make_config(**{"model.name": "mlp", "dataset.synthetic.scale": 0.05, "b200.precision": prec}))
    m = base.model
    # move off the symmetric init so every gradient path is exercised
    rng = np.random.default_rng(0)
    w = _weights(m)
    names = m.layout.names
    for i, n in enumerate(names):
        if n.startswith('bias') or n == 'global_bias':
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
        if n == 'domain_emb':
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    m.params.copy_(torch.from_numpy(m.layout.pack(w)))
    o = _oracle_for(base, weights=w)
    o64 = _oracle_for(base, weights=w, dtype=np.float64)
    data = base.dataset.train_dataset[2]['data']
    rows = min(rows, data.n_data)
    order = Schedule(1).batch_order(2, data.n_data)
    data.set_order(order)
    loss = torch.zeros(1, device="cuda")
    m._train_step(data, 0, rows, loss)
    h = data.host
    sel = order[:rows]
    ol, _, og = o.gradients(h['uid'][sel], h['pid'][sel], 2, h['label'][sel])
    _, _, og64 = o64.gradients(h['uid'][sel], h['pid'][sel], 2, h['label'][sel])
    assert abs(loss.item() - ol) < 2e-5 * abs(ol)
    g = m.layout.unpack(m.grads.cpu().numpy())
    for name, a, b, b64 in zip(names, g, og, og64):
        e_gpu = rel_err(a, b64)
        e_np = rel_err(b, b64)
        tol = 2e-5 if prec == "fp32" else 5e-5   # 3xTF32 keeps ~2^-21 per product (mlp_pass.cu)
        assert rel_err(a, b) < tol, (name, rel_err(a, b))
        # vs fp64 truth: the SIMT path is as accurate as numpy fp32; 3xTF32 stays within 1e-5
        assert e_gpu < (max(4 * e_np, 2e-6) if prec == "fp32" else 1e-5), (name, e_gpu, e_np)
    # one Adam step later the parameters agree too
    o.adam.apply(o.weights, og)
    for name, a, b in zip(names, _weights(m), o.weights):
        assert rel_err(a, b) < 1e-5, name


def _param_tol(prec, name, specific=False):
    """Parameter tolerance after N meta-steps.  fp32 mode: the north-star bar, rel 1e-4.  3xTF32 keeps ~2^-21 per
    product, so its pre-activations sit ~10x further from the fp32 ones; every ~50 steps one ReLU gate with a
    pre-activation within that distance of zero flips, a discrete change of one Adam-normalised update (the live
    weights track the fp64 oracle to 1e-6..1e-5 between such events: tests/diag_trace.py).  The stated bar for the
    tensor-core mode is therefore 1e-2 on theta and 5e-2 on the (difference-valued, small) theta_d / domain_emb."""
    if prec == "fp32":
        return 1e-4
    return 5e-2 if (specific or name == 'domain_emb') else 1e-2


def _run_both(config, kind, epochs):
    wrapper = _build(config)
    base = wrapper.base_model
    seed = config['dataset']['seed']
    data = base.dataset.host_splits()
    tc = config['train']
    if kind == "mamdr":
        wrapper.prepare()
        w0 = wrapper.meta_weights.numpy()
        dw0 = {k: v.numpy() for k, v in wrapper.domain_weights.items()}
        o = _oracle_for(wrapper, weights=w0)
        om = OracleMAMDR(o, data, tc, base.dataset.batch_size, Schedule(seed), dw0, name=config['model']['name'])
    else:
        wrapper._get_model_meta_parms()
        wrapper.meta_weights = wrapper._get_meta_weights()
        wrapper.model.reset_optimizer()
        wrapper.meta_sequence = wrapper.build_meta_data_split()
        o = _oracle_for(wrapper, weights=wrapper.meta_weights.numpy())
        om = OracleDN(o, data, tc, base.dataset.batch_size, Schedule(seed))
    base.schedule = Schedule(seed)
    for e in range(epochs):
        wrapper.train_epoch(e)
        om.train_epoch()
    return wrapper, om


@pytest.mark.parametrize("use_graphs,prec", [(True, "fp32"), (False, "fp32"), (True, "tf32x3")])
def test_dn_epochs_match_oracle(use_graphs, prec):
    c = make_config(**{"model.name": "mlp_meta_domain_negotiation_finetune", "dataset.synthetic.scale": 0.1,
                       "b200.cuda_graphs": use_graphs, "b200.precision": prec})
    wrapper, om = _run_both(c, "dn", 2)
    for name, a, b in zip(wrapper.model.layout.names, wrapper.meta_weights.numpy(), om.meta_weights):
        assert rel_err(a, b) < _param_tol(prec, name), (name, rel_err(a, b))
    step, b1, b2 = wrapper.model.read_step()
    assert step == om.model.adam.step and np.float32(b1) == om.model.adam.b1pow
    l, a, dl, da = wrapper.val_and_test("val")
    ol, oa, odl, oda = om.val_and_test("val")
    assert abs(a - oa) < 1e-3 and abs(l - ol) < 1e-4 * abs(ol)
    for k in da:
        assert abs(da[k] - oda[k]) < 1e-3


@pytest.mark.parametrize("name,merged,prec", [("mlp_meta_mamdr_finetune", "plus", "fp32"),
                                              ("mlp_meta_mamdr_batch", "plus", "fp32"),
                                              ("mlp_meta_mamdr_finetune", "times", "fp32"),
                                              ("mlp_meta_mamdr_finetune", "plus", "tf32x3")])
def test_mamdr_epochs_match_oracle(name, merged, prec):
    c = make_config(**{"model.name": name, "train.merged_method": merged, "dataset.synthetic.scale": 0.05,
                       "b200.precision": prec})
    wrapper, om = _run_both(c, "mamdr", 2)
    names = wrapper.model.layout.names
    for n_, a, b in zip(names, wrapper.meta_weights.numpy(), om.meta_weights):
        assert rel_err(a, b) < _param_tol(prec, n_), ("theta", n_, rel_err(a, b))
    for d in om.domain_weights:
        for n_, a, b in zip(names, wrapper.domain_weights[d].numpy(), om.domain_weights[d]):
            assert rel_err(a, b) < _param_tol(prec, n_, specific=True), ("theta_%d" % d, n_, rel_err(a, b))
    l, a, dl, da = wrapper.val_and_test("val")
    ol, oa, odl, oda = om.val_and_test("val")
    assert abs(a - oa) < 1e-3
    for k in da:
        assert abs(da[k] - oda[k]) < 1e-3, (k, da[k], oda[k])
    # early-stop bookkeeping + test split with the best snapshots
    assert wrapper.early_stop_step(a) == om.early_stop_step(oa)
    tl, ta, _, tda = wrapper.val_and_test("test")
    otl, ota, _, otda = om.val_and_test("test")
    assert abs(ta - ota) < 1e-3


@pytest.mark.parametrize("prec", PRECISIONS)
def test_replicated_runs_are_bit_identical(prec):
    """No float atomics: two independent runs of the same schedule give identical bits
    (what keeps replicated DN phases on several GPUs in lock-step without communication)."""
    outs = []
    for _ in range(2):
        c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 0.01, "train.sample_num": 1,
                           "b200.precision": prec})
        wrapper = _build(c)
        wrapper.prepare()
        wrapper.base_model.schedule = Schedule(5)
        wrapper.train_epoch(0)
        outs.append(torch.cat([wrapper.meta_weights.flat] + [wrapper.domain_weights[d].flat for d in range(10)]).cpu())
    assert torch.equal(outs[0], outs[1])


def test_tf32_single_pass_is_close_but_not_fp32():
    """The 1-pass TF32 speed mode: logits within 1e-3 (the AUC-level bar), clearly worse than 3xTF32."""
    errs = {}
    for prec in ("tf32", "tf32x3"):
        base = _build(make_config(**{"model.name": "mlp", "dataset.synthetic.scale": 0.05, "b200.precision": prec}))
        o = _oracle_for(base, dtype=np.float64)
        data = base.dataset.val_dataset[0]['data']
        rows = min(1024, data.n_data)
        probs, _ = base.model.predict(data, 0, rows)
        h = data.host
        _, p = o.forward(h['uid'][:rows], h['pid'][:rows], 0, train=False)
        errs[prec] = float(np.max(np.abs(probs.cpu().numpy() - p)))
    assert errs["tf32"] < 1e-3
    assert errs["tf32x3"] < 1e-6 and errs["tf32x3"] < errs["tf32"]


@pytest.mark.parametrize("name,prec", [("mlp_meta_reptile", "fp32"), ("mlp_meta_reptile_batch", "fp32"), ("mlp_meta_reptile", "tf32x3")])
def test_reptile_epochs_match_oracle(name, prec):
    """`Reptile.train` (reptile.py:45-99,127-142) on the same kernels -- two epochs, theta vs the oracle."""
    from oracle.meta import OracleReptile
    c = make_config(**{"model.name": name, "dataset.synthetic.scale": 0.05, "b200.precision": prec})
    wrapper = _build(c)
    base = wrapper.base_model
    seed = c['dataset']['seed']
    wrapper.prepare()
    o = _oracle_for(wrapper, weights=wrapper.meta_weights.numpy())
    om = OracleReptile(o, base.dataset.host_splits(), c['train'], base.dataset.batch_size, Schedule(seed), name=name)
    base.schedule = Schedule(seed)
    for e in range(2):
        wrapper.train_epoch(e)
        om.train_epoch()
    assert wrapper.train_sequence == om.sequence
    for n_, a, b in zip(wrapper.model.layout.names, wrapper.meta_weights.numpy(), om.meta_weights):
        assert rel_err(a, b) < _param_tol(prec, n_), (n_, rel_err(a, b))
    for n_, a, b in zip(wrapper.model.layout.names, _weights(wrapper.model), om.meta_weights):
        assert rel_err(a, b) < _param_tol(prec, n_), ("live model == theta", n_, rel_err(a, b))
    step, b1, _ = wrapper.model.read_step()
    assert step == om.model.adam.step and np.float32(b1) == om.model.adam.b1pow
    _, a, _, da = wrapper.val_and_test("val")
    _, oa, _, oda = om.val_and_test("val")
    assert abs(a - oa) < 1e-3


@pytest.mark.parametrize("name,prec", [("mlp_meta_mamdr_finetune", "fp32"), ("mlp_meta_domain_negotiation_finetune", "tf32x3"),
                                       ("mlp_meta_reptile_batch", "fp32")])
def test_save_state_resume_is_bit_identical(tmp_path, name, prec):
    """SURVEY.md 8(f) row f3 (what the reference lacks): theta, theta_d[], the Adam slots / beta powers, the domain sequence
    and the schedule's RNG survive a save / load -- a resumed run continues bit-identically."""
    c = make_config(**{"model.name": name, "dataset.synthetic.scale": 0.03, "b200.precision": prec})

    def fresh():
        w = _build(c)
        if hasattr(w, "prepare"):
            w.prepare()
        else:
            w._get_model_meta_parms()
            w.meta_weights = w._get_meta_weights()
            w.model.reset_optimizer()
            w.meta_sequence = w.build_meta_data_split()
        w.base_model.schedule = Schedule(c['dataset']['seed'])
        return w

    a = fresh()
    a.train_epoch(0)
    path = a.save_state(str(tmp_path / "state.pt"), epoch=0)
    a.train_epoch(1)
    torch.cuda.synchronize()
    b = fresh()
    assert b.load_state(path) == 0
    b.train_epoch(1)
    torch.cuda.synchronize()
    assert torch.equal(a.meta_weights.flat, b.meta_weights.flat) and torch.equal(a.model.params, b.model.params)
    assert torch.equal(a.model.m, b.model.m) and torch.equal(a.model.v, b.model.v) and a.model.read_step() == b.model.read_step()
    if getattr(a, "domain_weights", None):
        for d in a.domain_weights:
            assert torch.equal(a.domain_weights[d].flat, b.domain_weights[d].flat), d
