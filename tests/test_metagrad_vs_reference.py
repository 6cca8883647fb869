"""CPU: MAML / MLDG / PCGrad against the reference's own training loops EXECUTED in the build
container (tests/golden/reference_metagrad_v1.npz, make_reference_metagrad.py):

  * the oracle (`oracle/meta.py: OracleMAML / OracleMLDG / OraclePCGrad`, `pcgrad_project`) over a toy model replays the executed
    loops bit for bit: train steps, gradient calls, the live model, the kept checkpoint, the early-stop state;
  * the PRODUCT's wrappers (`mamdr_b200/{maml,mldg,pcgrad}.py`, the real code) driven on the CPU through stand-ins for their
    two seams -- `model.ctx.call(<C-ABI entry point>)` interpreted by numpy with the kernels' formulas and a toy model for
    `fit_pass` / `grads_on_batch` -- land on the same bits.  A test harness, not a product path (no CPU fallback exists).
"""
import ctypes as C
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT
from mamdr_b200.schedule import Schedule
from oracle import meta as ometa

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_reference_golden as mrg  # noqa: E402
import make_reference_metagrad as mg  # noqa: E402
from test_product_loops_vs_reference import _NumpyMetaOps, _ToyDeviceModel, _base, _bits, _flat, _view  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_metagrad_v1.npz"))


# ---- the oracle ------------------------------------------------------------------------------------------------------
class _ToyOracleModel(object):
    dtype = np.dtype(np.float32)

    def __init__(self):
        self.weights = mrg.toy_init(0)
        self.steps, self.grad_calls = [], []
        self.auc = types.SimpleNamespace(reset_states=lambda: None, update_state=lambda y, p: None)

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, ws):
        for a, b in zip(self.weights, ws):
            a[...] = b

    def train_on_batch(self, uid, pid, domain, label, optimizer='adam', sgd_lr=None):
        mrg.toy_step(self.weights, domain)
        self.steps.append(domain)
        return 0.0, 0.5

    def gradients(self, uid, pid, domain, label, masks=None, train=True):
        assert train is False, "the accumulating K.function runs the inference-mode forward"
        self.grad_calls.append(domain)
        return 0.0, np.zeros(len(uid), np.float32), mg.toy_grad(self.weights, domain)

    def evaluate(self, uid, pid, domain, label, batch_size):
        return mrg.toy_eval(self.weights, domain)


def _toy_data():
    col = lambda n: {'uid': np.zeros(n, np.int32), 'pid': np.zeros(n, np.int32), 'label': np.zeros(n, np.float32)}   # noqa: E731
    return {'train': {d: col(mg.N_DATA[d]) for d in sorted(mrg.N_STEP)},
            'val': {d: col(2 + d) for d in sorted(mrg.N_STEP)}, 'test': {d: col(3 + d) for d in sorted(mrg.N_STEP)}}


@pytest.mark.parametrize("i", range(len(mg.CASES)))
def test_oracle_replays_the_reference_metagrad_loops(i):
    kind, name, over = mg.CASES[i]
    tc = dict(mg.TC)
    tc.update(over)
    model = _ToyOracleModel()
    cls = {"maml": ometa.OracleMAML, "mldg": ometa.OracleMLDG, "pcgrad": ometa.OraclePCGrad}[kind]
    om = cls(model, _toy_data(), tc, mg.BATCH, Schedule(mrg.LOOP_SEED), name=name)
    for epoch in range(tc["epoch"]):
        om.train_epoch()
        _, val_auc, _, val_domain_auc = om.val()
        if om.early_stop_step(val_domain_auc[tc["target_domain"]] if tc["target_domain"] >= 0 else val_auc):
            break
        om.val_and_test("test")      # reloads the best checkpoint into the live model, like the reference (base_model.py:121)
    key = "case%d|" % i
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), G[key + "steps"])
    np.testing.assert_array_equal(np.array(model.grad_calls, dtype=np.int32), G[key + "grad_calls"])
    np.testing.assert_array_equal(_bits(mrg.flat_any(model.weights)), _bits(G[key + "live"]))
    np.testing.assert_array_equal(_bits(mrg.flat_any(om.best_weights)), _bits(G[key + "best"]))
    np.testing.assert_array_equal(np.array([om.es.counter, om.es.best_metric], dtype=np.float64), G[key + "es"])


def _proj_lists():
    shapes = [tuple(int(x) for x in s if x > 0) for s in G["proj|shapes"]]

    def split(flat):
        out, off = [], 0
        for s in shapes:
            n = int(np.prod(s))
            out.append(np.array(flat[off:off + n], dtype=np.float32).reshape(s))
            off += n
        return out
    return shapes, split


def test_oracle_pcgrad_projection_equals_the_executed_reference():
    """`PCGrad.PCGrad` (pcgrad.py:152-160) executed on 2-D, 1-D, [n, 1] and [1] variables, two support domains in a row."""
    _, split = _proj_lists()
    cur = split(G["proj|current"])
    for k in range(2):
        ometa.pcgrad_project(cur, cur, split(G["proj|aux%d" % k]))
        np.testing.assert_array_equal(_bits(mrg.flat_any(cur)), _bits(G["proj|final%d" % k]))


# ---- the product's wrappers on the CPU harness -------------------------------------------------------------------------
class _MetaGradOps(_NumpyMetaOps):
    """+ the entry points MAML / MLDG / PCGrad add: the second Adam over arena ranges and the PCGrad projection."""

    def __init__(self):
        _NumpyMetaOps.__init__(self)

    def call(self, name, *a):
        f32 = np.float32
        if name == "mamdr_adam_ranges_step":
            params, m, v, g, begin, length, n, state, lr, b1, b2, eps = a[:12]
            self.calls.append(name)
            pw = _view(state, 4)                                              # [step (int64) | b1pow | b2pow] like the device state
            b1pow, b2pow = f32(pw[2]), f32(pw[3])
            one = f32(1.0)
            alpha = f32(lr) * np.sqrt(one - b2pow) / (one - b1pow)            # oracle.mlp.AdamState.apply == TF ApplyAdam
            for k in range(n):
                off, ln = int(begin[k]), int(length[k])
                sl = lambda p: _view(C.c_void_p(p.value + 4 * off), ln)       # noqa: E731
                w, mm, vv, gg = sl(params), sl(m), sl(v), sl(g)
                mm += (gg - mm) * (one - f32(b1))
                vv += (gg * gg - vv) * (one - f32(b2))
                w -= (mm * alpha) / (np.sqrt(vv) + f32(eps))
            pw[2], pw[3] = f32(b1pow * f32(b1)), f32(b2pow * f32(b2))
        elif name == "mamdr_pcgrad_project":
            final, aux, rows, cols = a[:4]
            self.calls.append(name)
            shape = (int(rows), int(cols)) if rows > 1 else (int(cols),)
            f = _view(final, rows * cols).reshape(shape)
            x = np.array(_view(aux, rows * cols)).reshape(shape)
            ometa.pcgrad_project([f], [f], [x])
        elif name == "mamdr_opt_state_init":
            self.calls.append(name)
        else:
            _NumpyMetaOps.call(self, name, *a)


def _metagrad_base(name, tc):
    base, model = _base(name, "plus")
    base.train_config = dict(tc, meta_parms=["all"])
    model.ctx = _MetaGradOps()
    model.grads = torch.zeros_like(model.params)
    model.opt_state = torch.zeros(16, dtype=torch.uint8)
    model.beta1, model.beta2, model.eps = 0.9, 0.999, 1e-8
    model.grad_calls = []
    base.dataset.batch_size = mg.BATCH
    for d, rec in base.dataset.train_dataset.items():
        rec["n_data"] = mg.N_DATA[d]
        rec["data"] = types.SimpleNamespace(domain=d, n_data=mg.N_DATA[d], batch_size=mg.BATCH, device=None, uid=None, pid=None,
                                            label=None, set_order=lambda order: None)

    def grads_on_batch(data, offset, rows, loss_slot, with_auc=True):
        g = mg.toy_grad(model.views(), data.domain)
        model.grads.copy_(torch.from_numpy(model.layout.pack(g)))
        model.grad_calls.append(data.domain)
    model.grads_on_batch = grads_on_batch

    def new_optimizer_slots():
        st = torch.zeros(4, dtype=torch.float32)                              # mamdr_opt_state_init: beta powers start at beta
        st[2], st[3] = model.beta1, model.beta2
        return torch.zeros_like(model.params), torch.zeros_like(model.params), st.view(torch.uint8)
    model.new_optimizer_slots = new_optimizer_slots

    def fit_pass(data, steps=None, order=None):
        for _ in range(steps):
            mrg.toy_step(model.views(), data.domain)
            model.steps.append(data.domain)
    model.fit_pass = fit_pass
    return base, model


@pytest.mark.parametrize("i", range(len(mg.CASES)))
def test_product_wrappers_replay_the_reference_metagrad_loops(i):
    from mamdr_b200.maml import MAML
    from mamdr_b200.mldg import MLDG
    from mamdr_b200.pcgrad import PCGrad
    kind, name, over = mg.CASES[i]
    tc = dict(mg.TC)
    tc.update(over)
    base, model = _metagrad_base(name, tc)
    wrapper = {"maml": MAML, "mldg": MLDG, "pcgrad": PCGrad}[kind](base)
    wrapper.train()
    key = "case%d|" % i
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), G[key + "steps"])
    np.testing.assert_array_equal(np.array(model.grad_calls, dtype=np.int32), G[key + "grad_calls"])
    np.testing.assert_array_equal(_bits(_flat(model, model.params)), _bits(G[key + "live"]))
    np.testing.assert_array_equal(_bits(_flat(model, base.saved)), _bits(G[key + "best"]))
    np.testing.assert_array_equal(np.array([base.counter, base.best_metric], dtype=np.float64), G[key + "es"])
    used = set(model.ctx.calls)
    assert "mamdr_adam_ranges_step" in used and ("mamdr_pcgrad_project" in used) == (kind == "pcgrad")


def test_split_view_windows():
    """`engine.SplitView`: take / skip windows of the three meta_split modes (maml.py:296-318) as sample-id orders."""
    from mamdr_b200.engine import SplitView
    data = types.SimpleNamespace(domain=3, batch_size=4, device=None, uid=None, pid=None, label=None)
    perm = np.array([4, 0, 3, 1, 2], dtype=np.int32)
    tr, mv = SplitView(data, 0, 6), SplitView(data, 6, 11)                       # exclusive: first 6 / last 5 samples
    assert (tr.n_data, tr.n_step, mv.n_data, mv.n_step) == (6, 2, 5, 2)
    np.testing.assert_array_equal(mv.window_order(perm), perm + 6)
    a, b = SplitView(data, 0, 5, slice(0, 3)), SplitView(data, 0, 5, slice(3, 5))   # shuffle-then-take / -skip
    assert (a.n_data, b.n_data) == (3, 2)
    np.testing.assert_array_equal(a.window_order(perm), [4, 0, 3])
    np.testing.assert_array_equal(b.window_order(perm), [1, 2])


@pytest.mark.parametrize("i", [0, 1, 3, 5])
def test_metagrad_save_state_resume_on_the_cpu_harness(tmp_path, i):
    """`save_state` after the first epoch (theta, live model, the accumulators, BOTH optimizers' slots, the schedule),
    `load_state` into a freshly prepared wrapper, second epoch: the same steps, gradient calls, theta and live arena as the
    uninterrupted run."""
    from mamdr_b200.maml import MAML
    from mamdr_b200.mldg import MLDG
    from mamdr_b200.pcgrad import PCGrad
    kind, name, over = mg.CASES[i]
    tc = dict(mg.TC)
    tc.update(over)

    def fresh():
        base, model = _metagrad_base(name, tc)
        w = {"maml": MAML, "mldg": MLDG, "pcgrad": PCGrad}[kind](base)
        w.prepare()
        return w, model

    a, model_a = fresh()
    a.train_epoch(0)
    path = a.save_state(str(tmp_path / "state.pt"), epoch=0)
    first = (len(model_a.steps), len(model_a.grad_calls))
    a.train_epoch(1)
    b, model_b = fresh()
    assert b.load_state(path) == 0
    b.train_epoch(1)
    assert model_b.steps == model_a.steps[first[0]:] and model_b.grad_calls == model_a.grad_calls[first[1]:]
    assert torch.equal(model_a.params, model_b.params) and torch.equal(a.meta_weights.flat, b.meta_weights.flat)
    assert torch.equal(a._meta_m, b._meta_m) and torch.equal(a._meta_v, b._meta_v) and torch.equal(a._meta_opt_state, b._meta_opt_state)
