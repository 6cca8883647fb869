"""Diagnostic (not a test): multi-step pass kernel vs per-step kernels vs the CPU oracle, per tensor."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule
from oracle.meta import train_pass
from test_gpu_mlp import _build, _oracle_for, _weights
from test_gpu_pass import _perturb


def main(prec="tf32x3", bs=512, scale=0.3):
    outs = {}
    for use_pass in (True, False):
        base = _build(make_config(**{"model.name": "mlp", "dataset.synthetic.scale": scale, "b200.precision": prec,
                                     "b200.pass_kernel": use_pass, "dataset.batch_size": bs}))
        m = base.model
        w = _perturb(base)
        data = base.dataset.train_dataset[0]['data']
        order = Schedule(3).batch_order(0, data.n_data)
        data.set_order(order)
        m.reset_states()
        losses = m.fit_pass(data)
        torch.cuda.synchronize()
        outs[use_pass] = (_weights(m), losses.cpu().numpy().copy())
    o = _oracle_for(base, weights=w)
    h = data.host
    ol = train_pass(o, {"uid": h['uid'], "pid": h['pid'], "label": h['label']}, 0, order, bs)
    o64 = _oracle_for(base, weights=w, dtype=np.float64)
    train_pass(o64, {"uid": h['uid'], "pid": h['pid'], "label": h['label']}, 0, order, bs)
    print("steps", data.n_step, "n", data.n_data)
    print("%-14s %12s %12s %12s %12s" % ("tensor", "pass-np32", "step-np32", "pass-step", "np32-f64"))
    for i, n in enumerate(base.model.layout.names):
        print("%-14s %12.3e %12.3e %12.3e %12.3e" % (n, rel_err(outs[True][0][i], o.weights[i]), rel_err(outs[False][0][i], o.weights[i]),
                                              rel_err(outs[True][0][i], outs[False][0][i]), rel_err(o.weights[i], o64.weights[i])))
    print("losses pass", outs[True][1][:4], "step", outs[False][1][:4])


if __name__ == "__main__":
    for prec in sys.argv[1:] or ["tf32x3", "tf32"]:
        print("=====", prec)
        main(prec)
