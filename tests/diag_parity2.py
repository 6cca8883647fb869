"""Diagnostic (not a test): drift of the tcgen05 pass path vs the fp64 oracle, DN-only and MAMDR, per tensor."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from conftest import make_config, rel_err
import run
from mamdr_b200.schedule import Schedule
from oracle.meta import OracleDN, OracleMAMDR
from oracle.mlp import MLPSpec, OracleMLP


def omlp(base, w, dtype):
    spec = MLPSpec(base.n_uid, base.n_pid, base.n_domain, (128, 128, 128), (256, 128, 64), dropout=0.5)
    return OracleMLP(spec, w, base.dataset.user_table, base.dataset.item_table, lr=1e-3, dtype=dtype)


def dn(prec, scale, epochs):
    c = make_config(**{"model.name": "mlp_meta_domain_negotiation_finetune", "dataset.synthetic.scale": scale, "b200.precision": prec})
    w = run.build(c)
    base = w.base_model
    w._get_model_meta_parms()
    w.meta_weights = w._get_meta_weights()
    w.model.reset_optimizer()
    w.meta_sequence = w.build_meta_data_split()
    o64 = OracleDN(omlp(base, w.meta_weights.numpy(), np.float64), base.dataset.host_splits(), c['train'], 1024, Schedule(123))
    base.schedule = Schedule(123)
    names = w.model.layout.names
    for e in range(epochs):
        w.train_epoch(e)
        o64.train_epoch()
        errs = [rel_err(a, b) for a, b in zip(w.meta_weights.numpy(), o64.meta_weights)]
        print("DN %s epoch %d steps %d: " % (prec, e, o64.model.adam.step) + " ".join("%s=%.1e" % (n[:7], x) for n, x in zip(names, errs)))


def mamdr(prec, scale, epochs):
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": scale, "b200.precision": prec})
    w = run.build(c)
    w.prepare()
    base = w.base_model
    o64 = OracleMAMDR(omlp(base, w.meta_weights.numpy(), np.float64), base.dataset.host_splits(), c['train'], 1024, Schedule(123),
                      {k: v.numpy() for k, v in w.domain_weights.items()}, name=c['model']['name'])
    base.schedule = Schedule(123)
    names = w.model.layout.names
    for e in range(epochs):
        w.train_epoch(e)
        o64.train_epoch()
        errs = [rel_err(a, b) for a, b in zip(w.meta_weights.numpy(), o64.meta_weights)]
        print("MAMDR %s epoch %d steps %d theta: " % (prec, e, o64.model.adam.step) + " ".join("%s=%.1e" % (n[:7], x) for n, x in zip(names, errs)))
        worst = np.zeros(len(names))
        for d in o64.domain_weights:
            worst = np.maximum(worst, [rel_err(a, b) for a, b in zip(w.domain_weights[d].numpy(), o64.domain_weights[d])])
        print("MAMDR %s epoch %d theta_d worst: " % (prec, e) + " ".join("%s=%.1e" % (n[:7], x) for n, x in zip(names, worst)))


if __name__ == "__main__":
    for prec in ("fp32", "tf32x3", "tf32"):
        dn(prec, 0.1, 5)
    for prec in ("fp32", "tf32x3"):
        mamdr(prec, 0.05, 2)
