"""-m gpu: BASELINE config #2 -- the joint `mlp` baseline with TRAINABLE embedding tables (Amazon shape,
`load_pretrain_emb: false`): gather from the arena tables, sparse embedding backward (sort + segment-sum dedup, K6)
and the fused L2 + non-lazy Adam sweep over every table row (K7), against the CPU oracle.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule
from oracle.meta import joint_train_epoch
from test_gpu_mlp import _build, _oracle_for, _weights

pytestmark = pytest.mark.gpu


def _amazon(scale=0.002, **over):
    kw = {"model.name": "mlp", "train.load_pretrain_emb": False, "dataset.name": "Amazon",
          "dataset.synthetic.shape": "Amazon-6", "dataset.synthetic.scale": scale, "b200.precision": "fp32"}
    kw.update(over)
    return make_config(**kw)


@pytest.mark.parametrize("prec", ["fp32", "tf32x3"])
@pytest.mark.parametrize("rows", [1024, 333])
def test_trainable_step_matches_oracle(rows, prec):
    """fp32: the per-mini-batch SIMT path.  tf32x3: the tcgen05 pass kernel (tables gathered from the arena, tower + dX on the
    tensor cores, dense apply in-kernel) followed by the same de-duplication and table sweeps."""
    base = _build(_amazon(**{"b200.precision": prec}))
    m = base.model
    assert m.emb_trainable and base.layout.names[:2] == ['user_emb', 'item_emb']
    # lift the tables off their 1e-4 init so the data gradient and the l2 term are both visible
    rng = np.random.default_rng(0)
    w = _weights(m)
    for i, n in enumerate(m.layout.names):
        if n.endswith('_emb'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    m.params.copy_(torch.from_numpy(m.layout.pack(w)))
    o = _oracle_for(base, weights=w)
    data = base.dataset.train_dataset[0]['data']
    rows = min(rows, data.n_data)
    order = Schedule(1).batch_order(0, data.n_data)
    data.set_order(order)
    loss = torch.zeros(1, device="cuda")
    m._train_step(data, 0, rows, loss)
    torch.cuda.synchronize()
    h = data.host
    sel = order[:rows]
    # de-duplicated ids are bit-exact: sorted unique of the batch ids
    for t, col in enumerate(("uid", "pid")):
        ids, srows, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        wsb = m.pass_ws if m.pass_kernel else m.ws
        getter = m.ctx.lib.mamdr_mlp_pass_sparse_grads if m.pass_kernel else m.ctx.lib.mamdr_mlp_sparse_grads
        assert getter(C.byref(m.desc), rows, C.c_void_p(wsb.data_ptr()), t, C.byref(ids), C.byref(srows), C.byref(cnt)) == 0
        off = ids.value - wsb.data_ptr()
        noff = cnt.value - wsb.data_ptr()
        n_u = int(wsb[noff:noff + 4].view(torch.int32).item())
        got = wsb[off:off + 4 * n_u].view(torch.int32).cpu().numpy()
        assert np.array_equal(got, np.unique(h[col][sel]))
    ol, _, og = o.gradients(h['uid'][sel], h['pid'][sel], 0, h['label'][sel])
    assert abs(loss.item() - ol) < 2e-5 * abs(ol)
    o.adam.apply(o.weights, og)
    for name, a, b in zip(m.layout.names, _weights(m), o.weights):
        assert rel_err(a, b) < 1e-5, (name, rel_err(a, b))
    # every row moved (non-lazy Adam + l2 on all rows), and the slot maps were left clean
    assert not np.array_equal(_weights(m)[0], w[0])
    for _, _, _, slot in m._tables:
        assert int((slot != -1).sum().item()) == 0
    step, b1, b2 = m.read_step()
    assert step == 1 and np.float32(b1) == o.adam.b1pow


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("tf32x3", 3e-2)])
def test_joint_training_epoch_matches_oracle(prec, tol):
    """`DeepCTR.train` (deepctr.py:63-93): shuffled domains, one full pass each, one Adam -- two epochs.  tf32x3: the
    free-running bar of the tensor-core mode (ReLU-gate events; measured 1.0e-2 on kernel0 here)."""
    c = _amazon(scale=0.001, **{"b200.precision": prec})
    base = _build(c)
    m = base.model
    # Free-running from the raw init (tables N(0, 1e-4^2), zero biases) every pre-activation sits within ~1e-3 of zero and
    # Adam's first steps amplify last-bit differences into ReLU-gate flips: the outcome then depends on the summation order
    # of the ORACLE's CPU GEMMs, i.e. on the host's core count (seen: within 1e-4 on the 1-GPU box, 1.5e-3 on the 2-GPU box, same GPU code).  The
    # trajectory test starts from a lifted state (as tests/test_gpu_sharded.py does); the raw init is covered step by step
    # in test_trainable_step_matches_oracle and tests/test_gpu_mtl.py::test_mtl_trajectory_step_by_step.
    rng = np.random.default_rng(0)
    w = _weights(m)
    for i, n in enumerate(m.layout.names):
        if n.endswith('_emb') or n.startswith('bias'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    m.params.copy_(torch.from_numpy(m.layout.pack(w)))
    m.reset_optimizer()
    o = _oracle_for(base, weights=w)
    seed = c['dataset']['seed']
    data = base.dataset.host_splits()
    base.schedule = Schedule(seed)
    osched = Schedule(seed)
    seq_g, seq_o = list(range(base.n_domain)), list(range(base.n_domain))
    for epoch in range(2):
        seq_g = base.schedule.shuffle_sequence(seq_g)
        base.stage_epoch_orders(list(seq_g))
        for idx in seq_g:
            m.reset_states()
            base.run_train_pass(idx)
        seq_o = joint_train_epoch(o, data, base.dataset.batch_size, osched, seq_o)
        assert seq_g == seq_o
    for name, a, b in zip(m.layout.names, _weights(m), o.weights):
        assert rel_err(a, b) < tol, (name, rel_err(a, b))
    d = base.dataset.val_dataset[0]
    loss, auc = m.evaluate(d['data'], d['n_step'])
    hv = d['data'].host
    ol, oa = o.evaluate(hv['uid'], hv['pid'], 0, hv['label'], base.dataset.batch_size)
    assert abs(loss - ol) < 2e-5 * abs(ol) and abs(auc - oa) < 1e-3


def test_table_sweep_properties_at_amazon6_size():
    """Size-independent properties of the fused table sweep at the real Amazon-6 user table (445 789 x 128):
    rows without a sparse gradient see g = 2*l2*E exactly; a second call with n_uniq = 0 touches nothing sparse;
    sum(E^2) returned through the loss slot matches torch's fp64 reduction."""
    from gpu_util import ctx, ptr, stream
    c = ctx()
    rows, dim = 445789, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.randn(rows, dim, device="cuda", generator=g) * 0.05
    p0 = p.clone()
    mm, vv = torch.zeros_like(p), torch.zeros_like(p)
    state = torch.zeros(c.lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device="cuda")
    c.call("mamdr_opt_state_init", ptr(state), 0.9, 0.999, stream())
    slot = torch.full((rows,), -1, dtype=torch.int32, device="cuda")
    ws = torch.zeros(c.lib.mamdr_adam_table_workspace_bytes(), dtype=torch.uint8, device="cuda")
    ids = torch.tensor([5, 77, 445788], dtype=torch.int32, device="cuda")
    srows = torch.ones(3, dim, device="cuda")
    cnt = torch.tensor([3], dtype=torch.int32, device="cuda")
    loss = torch.zeros(1, device="cuda")
    c.call("mamdr_adam_table_step", ptr(p), ptr(mm), ptr(vv), rows, dim, ptr(ids), ptr(srows), ptr(cnt), 3, ptr(slot), 1e-5,
           ptr(state), 1e-3, 0.9, 0.999, 1e-8, ptr(loss), ptr(ws), ws.numel(), stream())
    torch.cuda.synchronize()
    assert int((slot != -1).sum().item()) == 0
    ref = 1e-5 * float((p0.double() ** 2).sum().item())
    assert abs(loss.item() - ref) < 1e-6 * ref
    # first Adam step from zero slots: p -= lr * g / (|g| + eps')  ~  lr * sign(g); touched rows have g ~ +1
    delta = (p - p0)
    assert torch.all(delta[ids.long()] < 0)
    untouched = torch.ones(rows, dtype=torch.bool, device="cuda")
    untouched[ids.long()] = False
    gl2 = (2e-5 * p0[untouched])
    expect = -1e-3 * gl2 / (gl2.abs() + 1e-8 / (1 - 0.999) ** 0.5)
    assert torch.allclose(delta[untouched], expect, rtol=2e-3, atol=1e-9)


@pytest.mark.parametrize("prec", ["fp32", "tf32x3"])
def test_trainable_sgd_steps_match_oracle(prec):
    """The finetune stage's plain SGD (specific_base_model.py:120, base_model.py:69) on a model with TRAINABLE tables
    (config/Amazon_6/deepctr_DN+DR.json is such a `_finetune` name): every table row moves by (2 l2 E + sparse) * lr
    (`mamdr_sgd_table_step`), the dense variables by their gradient * lr -- three steps vs the oracle (dropout 0: the oracle's
    SGD does not advance the dropout step counter)."""
    base = _build(_amazon(**{"b200.precision": prec, "model.dropout": 0.0}))
    m = base.model
    rng = np.random.default_rng(3)
    w = _weights(m)
    for i, n in enumerate(m.layout.names):
        if n.endswith('_emb'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    m.params.copy_(torch.from_numpy(m.layout.pack(w)))
    o = _oracle_for(base, weights=w)
    m.compile(optimizer="sgd", lr=0.05)
    data = base.dataset.train_dataset[1]['data']
    order = Schedule(2).batch_order(1, data.n_data)
    data.set_order(order)
    h = data.host
    rows = min(700, data.n_data)
    loss = torch.zeros(1, device="cuda")
    m_before, v_before = m.m.clone(), m.v.clone()
    for s in range(3):
        off = (s * 37) % max(1, data.n_data - rows)
        loss.zero_()
        m._train_step(data, off, rows, loss)
        sel = order[off:off + rows]
        ol, _ = o.train_on_batch(h['uid'][sel], h['pid'][sel], 1, h['label'][sel], optimizer='sgd', sgd_lr=0.05)
        assert abs(loss.item() - ol) < 2e-5 * abs(ol), (s, loss.item(), ol)
    for name, a, b in zip(m.layout.names, _weights(m), o.weights):
        assert rel_err(a, b) < (1e-5 if prec == "fp32" else 5e-5), (name, rel_err(a, b))
    assert torch.equal(m.m, m_before) and torch.equal(m.v, v_before)      # SGD leaves the Adam slots alone
    for _, _, _, slot in m._tables:
        assert int((slot != -1).sum().item()) == 0
    m.compile(optimizer="adam")


@pytest.mark.parametrize("prec", ["tf32x3", "fp32"])
def test_trainable_mamdr_finetune_name_runs_end_to_end(tmp_path, prec):
    """`mlp_meta_mamdr_finetune` with trainable tables (the shape of config/Amazon_6/deepctr_DN+DR.json) through run.main:
    meta-training, test, reload, the per-domain SGD finetune stage, result files."""
    import os
    import run
    c = _amazon(scale=0.001, **{"model.name": "mlp_meta_mamdr_finetune", "b200.precision": prec, "train.epoch": 2, "train.sample_num": 1})
    c["train"]["result_save_path"], c["train"]["checkpoint_path"] = str(tmp_path / "result"), str(tmp_path / "ckpt")
    avg_loss, avg_auc, domain_loss, domain_auc = run.main(c)
    assert np.isfinite(avg_loss) and 0.0 <= avg_auc <= 1.0 and len(domain_auc) == 6
    assert any(f == "result.json" for _, _, fs in os.walk(str(tmp_path / "result")) for f in fs)
