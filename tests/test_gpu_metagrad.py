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


@pytest.mark.parametrize("kind,name,prec,over", CASES)
def test_metagrad_epochs_match_oracle(kind, name, prec, over):
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
    for e in range(2):
        wrapper.train_epoch(e)
        om.train_epoch()
    assert wrapper.train_sequence == om.sequence
    names = wrapper.model.layout.names
    for n_, a, b in zip(names, _weights(wrapper.model), o.weights):
        assert rel_err(a, b) < _param_tol(prec, n_), ("live model", n_, rel_err(a, b))
    if kind != "pcgrad":
        for n_, a, b in zip(names, wrapper.meta_weights.numpy(), om.meta_weights):
            assert rel_err(a, b) < _param_tol(prec, n_), ("theta", n_, rel_err(a, b))
    # the meta optimizer's slots and the accumulators (cleared) agree as well
    for n_, a, b in zip(names, wrapper.model.layout.unpack(wrapper._meta_m.cpu().numpy()), om.meta_adam.m):
        assert rel_err(a, b) < (1e-4 if prec == "fp32" else 1e-2), ("meta Adam m", n_, rel_err(a, b))
    assert float(wrapper.accum_grads.abs().max().item()) == 0.0
    _, a, _, _ = wrapper.val_and_test("val")
    _, oa, _, _ = om.val_and_test("val")
    assert abs(a - oa) < 1e-3
