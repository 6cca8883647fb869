"""Diagnostic (not a test): per-tensor drift of the GPU paths and of the fp32 oracle against the fp64 oracle."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
from conftest import make_config, rel_err
import run
from mamdr_b200.schedule import Schedule
from oracle.meta import OracleMAMDR
from oracle.mlp import MLPSpec, OracleMLP


def oracle(wrapper, dtype):
    base = wrapper.base_model
    spec = MLPSpec(base.n_uid, base.n_pid, base.n_domain, (128, 128, 128), (256, 128, 64), dropout=0.5)
    o = OracleMLP(spec, wrapper.meta_weights.numpy(), base.dataset.user_table, base.dataset.item_table, lr=1e-3, dtype=dtype)
    return OracleMAMDR(o, base.dataset.host_splits(), base.train_config, 1024, Schedule(123),
                       {k: v.numpy() for k, v in wrapper.domain_weights.items()}, name=base.model_config['name'])


def main(scale=0.05, epochs=2):
    ws = {}
    for prec in ("fp32", "tf32x3", "tf32"):
        c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": scale, "b200.precision": prec})
        w = run.build(c)
        w.prepare()
        w.base_model.schedule = Schedule(123)
        ws[prec] = w
    o32, o64 = oracle(ws["fp32"], np.float32), oracle(ws["fp32"], np.float64)
    names = ws["fp32"].model.layout.names
    for e in range(epochs):
        for w in ws.values():
            w.train_epoch(e)
        o32.train_epoch()
        o64.train_epoch()
        print("epoch", e, "adam steps", o64.model.adam.step)
        print("%-14s %10s %10s %10s %10s | %10s" % ("tensor", "np32-f64", "gpu32-f64", "x3-f64", "tf32-f64", "gpu32-np32"))
        for i, n in enumerate(names):
            t64 = o64.meta_weights[i]
            row = [rel_err(o32.meta_weights[i], t64)] + [rel_err(ws[p].meta_weights.numpy()[i], t64) for p in ("fp32", "tf32x3", "tf32")]
            row.append(rel_err(ws["fp32"].meta_weights.numpy()[i], o32.meta_weights[i]))
            print("%-14s %10.2e %10.2e %10.2e %10.2e | %10.2e" % ((n,) + tuple(row)))
        worst = {p: 0.0 for p in ("np32", "fp32", "tf32x3", "tf32")}
        for d in o64.domain_weights:
            for i in range(len(names)):
                t64 = o64.domain_weights[d][i]
                worst["np32"] = max(worst["np32"], rel_err(o32.domain_weights[d][i], t64))
                for p in ("fp32", "tf32x3", "tf32"):
                    worst[p] = max(worst[p], rel_err(ws[p].domain_weights[d].numpy()[i], t64))
        print("theta_d worst:", {k: "%.2e" % v for k, v in worst.items()})


if __name__ == "__main__":
    main(float(sys.argv[1]) if len(sys.argv) > 1 else 0.05, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
