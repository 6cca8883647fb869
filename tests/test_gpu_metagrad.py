"""-m gpu: MAML / MLDG / PCGrad on the CUDA path, through the C-ABI, against the CPU oracle on
identical synthetic inputs, seeds and schedules -- and `mamdr_pcgrad_project` against vectors produced by EXECUTING the
reference's `PCGrad.PCGrad` (tests/golden/reference_metagrad_v1.npz).

Tolerances: parameters rel 1e-4 after two epochs in fp32 mode (BASELINE.json north_star), AUC 1e-3; one gradient call 2e-5;
the projection 2e-6 (only the order of the per-row sums differs from numpy's pairwise reduction)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, make_config, rel_err
from mamdr_b200.schedule import Schedule

pytestmark = pytest.mark.gpu

from gpu_util import ctx, dev, ptr, stream  # noqa: E402
from test_gpu_mlp import _build, _oracle_for, _param_tol, _weights  # noqa: E402

G = np.load(os.path.join(ROOT, "tests", "golden", "reference_metagrad_v1.npz"))


def test_pcgrad_projection_kernel_vs_the_executed_reference():
    c = ctx()
    shapes = [tuple(int(x) for x in s if x > 0) for s in G["proj|shapes"]]
    final = dev(G["proj|current"])
    for k in range(2):
        aux = dev(G["proj|aux%d" % k])
        off = 0
        for s in shapes:
            n = int(np.prod(s))
            cols = s[-1]
            c.call("mamdr_pcgrad_project", ptr(final[off:off + n]), ptr(aux[off:off + n]), n // cols, cols, stream())
            off += n
        got, want = final.cpu().numpy(), G["proj|final%d" % k]
        off = 0
        for s in shapes:
            n = int(np.prod(s))
            assert rel_err(got[off:off + n], want[off:off + n]) < 2e-6, (k, s)
            off += n
    # rows whose dot product is <= 0 are added unprojected: bit-exact there
    cur, aux = np.array([[1.0, 2.0], [3.0, -1.0]], np.float32), np.array([[-1.0, 0.25], [1.0, 3.0]], np.float32)   # dots: -0.5, 0
    f = dev(cur.reshape(-1))
    c.call("mamdr_pcgrad_project", ptr(f), ptr(dev(aux.reshape(-1))), 2, 2, stream())
    np.testing.assert_array_equal(f.cpu().numpy().reshape(2, 2), cur + aux)


@pytest.mark.parametrize("prec", ["fp32", "tf32x3"])
def test_gradient_call_is_the_inference_mode_gradient(prec):
    """`grads_on_batch` == the K.function of maml.py:196-233: no dropout, no optimizer apply, gradients of every variable."""
    base = _build(make_config(**{"model.name": "mlp", "dataset.synthetic.scale": 0.05, "b200.precision": prec}))
    m = base.model
    o = _oracle_for(base)
    data = base.dataset.train_dataset[2]['data']
    order = Schedule(5).batch_order(2, data.n_data)
    data.set_order(order)
    rows = min(1000, data.n_data)
    before = m.params.clone()
    loss = torch.zeros(1, device="cuda")
    m.grads_on_batch(data, 0, rows, loss)
    h, sel = data.host, order[:rows]
    ol, _, og = o.gradients(h['uid'][sel], h['pid'][sel], 2, h['label'][sel], train=False)
    assert torch.equal(before, m.params)
    assert abs(loss.item() - ol) < 2e-5 * abs(ol)
    for name, a, b in zip(m.layout.names, m.layout.unpack(m.grads.cpu().numpy()), og):
        assert rel_err(a, b) < 2e-5, (name, rel_err(a, b))


CASES = [("maml", "mlp_meta_maml", "fp32", {}), ("maml", "mlp_meta_maml_batch", "fp32", {"meta_split": "train-train"}),
         ("maml", "mlp_meta_maml", "tf32x3", {"meta_split": "meta-train/val-no-exclusive"}),
         ("mldg", "mlp_meta_mldg", "fp32", {}), ("mldg", "mlp_meta_mldg_batch", "fp32", {"meta_train_step": 1, "average_meta_grad": "mean"}),
         ("pcgrad", "mlp_pcgrad", "fp32", {"sample_num": 2})]


GATE_EPS = 5e-7     # |pre-activation| below this can round to either side of the ReLU (tests/test_gpu_trajectory.py)
CLEAN = {"kernel": 1e-5, "small": 1e-4}
LOOSE = 5e-2        # a domain step in which a gate flipped (the meta Adam normalises the changed gradient entries to full steps)


def _poke_opt(state, step, b1pow, b2pow):
    state[:8].view(torch.int64)[0] = int(step)
    state[8:12].view(torch.float32)[0] = float(b1pow)
    state[12:16].view(torch.float32)[0] = float(b2pow)


def _teacher_force(wrapper, om):
    """Load the oracle's whole state (live model, theta, both optimizers, accumulators) into the device wrapper."""
    m, o, lo = wrapper.model, om.model, wrapper.model.layout
    put = lambda dst, ws: dst.copy_(torch.from_numpy(lo.pack(ws)))   # noqa: E731
    put(m.params, o.weights)
    put(m.m, o.adam.m)
    put(m.v, o.adam.v)
    m.set_opt_words(torch.tensor([o.adam.step, float(o.adam.b1pow), float(o.adam.b2pow)], dtype=torch.float64))
    put(wrapper.meta_weights.flat, om.meta_weights)
    put(wrapper._meta_m, om.meta_adam.m)
    put(wrapper._meta_v, om.meta_adam.v)
    _poke_opt(wrapper._meta_opt_state, om.meta_adam.step, om.meta_adam.b1pow, om.meta_adam.b2pow)
    put(wrapper.accum_grads, om.accum)


@pytest.mark.parametrize("kind,name,prec,over", CASES)
def test_metagrad_epochs_match_oracle(kind, name, prec, over):
    """Two epochs, teacher-forced per domain step (the oracle's state is loaded before every step, like
    tests/test_gpu_trajectory.py): after the step the live model and theta agree to the clean per-step bound; a step that
    does not must coincide with a near-zero pre-activation in the oracle's own forwards (a ReLU gate that rounds to the other
    side flips one gradient column, and the meta Adam turns that into full-size steps) and stay within the loose bound."""
    from oracle.meta import OracleMAML, OracleMLDG, OraclePCGrad
    keys = {"model.name": name, "dataset.synthetic.scale": 0.05, "b200.precision": prec, "train.meta_split": "meta-train/val",
            "train.meta_split_ratio": 0.8, "train.average_meta_grad": "none", "train.meta_learning_rate": 1e-3}
    keys.update({"train." + k: v for k, v in over.items()})
    c = make_config(**keys)
    wrapper = _build(c)
    assert type(wrapper).__name__ == {"maml": "MAML", "mldg": "MLDG", "pcgrad": "PCGrad"}[kind]
    base = wrapper.base_model
    seed = c['dataset']['seed']
    wrapper.prepare()
    o = _oracle_for(wrapper, weights=wrapper.meta_weights.numpy())
    om = {"maml": OracleMAML, "mldg": OracleMLDG, "pcgrad": OraclePCGrad}[kind](
        o, base.dataset.host_splits(), c['train'], base.dataset.batch_size, Schedule(seed), name=name)
    base.schedule = Schedule(seed)
    names = wrapper.model.layout.names
    flagged, worst, steps = [], 0.0, 0
    for e in range(2):
        wrapper.train_sequence = base.schedule.shuffle_sequence(wrapper.train_sequence)
        om.sequence = om.schedule.shuffle_sequence(om.sequence)
        assert wrapper.train_sequence == om.sequence
        for idx in om.sequence:
            _teacher_force(wrapper, om)
            o.preact_log = []
            om.domain_step(idx)
            log, o.preact_log = np.asarray(o.preact_log), None
            wrapper.domain_step(idx)
            steps += 1
            errs = {n_: rel_err(a, b) for n_, a, b in zip(names, _weights(wrapper.model), o.weights)}
            if kind != "pcgrad":
                errs.update({"theta/" + n_: rel_err(a, b) for n_, a, b in zip(names, wrapper.meta_weights.numpy(), om.meta_weights)})
            tol = CLEAN if prec == "fp32" else {"kernel": 5e-5, "small": 5e-4}
            bad = [n_ for n_, x in errs.items() if x > (tol["kernel"] if "kernel" in n_ else tol["small"])]
            if bad:
                assert np.sum(log < GATE_EPS) > 0, ("domain step diverged without a near-zero pre-activation", e, idx, errs, log.min())
                assert max(errs.values()) < LOOSE, (e, idx, errs)
                flagged.append((e, idx, max(errs.values()), float(log.min())))
            else:
                worst = max(worst, max(errs.values()))
            step, b1, _ = wrapper.model.read_step()
            assert step == o.adam.step and np.float32(b1) == o.adam.b1pow
        _teacher_force(wrapper, om)
        wrapper.finish_epoch()
        om.finish_epoch()
        for n_, a, b in zip(names, _weights(wrapper.model), o.weights):
            assert rel_err(a, b) < 1e-4, ("end of epoch", n_, rel_err(a, b))
    print("%s %s %s: %d domain steps teacher-forced, worst clean error %.2e, gate events (epoch, domain, error, min |pre-activation|) %s"
          % (kind, name, prec, steps, worst, flagged))
    assert len(flagged) <= 3, flagged
    assert float(wrapper.accum_grads.abs().max().item()) == 0.0
    _, a, _, _ = wrapper.val_and_test("val")
    _, oa, _, _ = om.val_and_test("val")
    assert abs(a - oa) < 1e-3
