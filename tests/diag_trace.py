"""Diagnostic (not a test): pass-by-pass drift of the GPU model weights vs the fp64 oracle over a MAMDR meta-step."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from conftest import make_config, rel_err
import run
from mamdr_b200.schedule import Schedule
import oracle.meta as ometa
from oracle.meta import OracleMAMDR
from oracle.mlp import MLPSpec, OracleMLP


def main(prec, scale=0.05):
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": scale, "b200.precision": prec})
    w = run.build(c)
    w.prepare()
    base = w.base_model
    spec = MLPSpec(base.n_uid, base.n_pid, base.n_domain, (128, 128, 128), (256, 128, 64), dropout=0.5)
    o = OracleMLP(spec, w.meta_weights.numpy(), base.dataset.user_table, base.dataset.item_table, lr=1e-3, dtype=np.float64)
    om = OracleMAMDR(o, base.dataset.host_splits(), c['train'], 1024, Schedule(123),
                     {k: v.numpy() for k, v in w.domain_weights.items()}, name=c['model']['name'])
    base.schedule = Schedule(123)
    gpu_trace, ora_trace = [], []
    orig_rtp = base.run_train_pass

    def rtp(domain_idx, steps=None):
        r = orig_rtp(domain_idx, steps)
        m = base.model
        gpu_trace.append((domain_idx, m.layout.unpack(m.params.cpu().numpy()), m.layout.unpack(m.grads.cpu().numpy())))
        return r
    base.run_train_pass = rtp
    orig_tp = ometa.train_pass

    def tp(model, d, domain, order, batch_size, max_steps=0, optimizer='adam', sgd_lr=None):
        r = orig_tp(model, d, domain, order, batch_size, max_steps, optimizer, sgd_lr)
        ora_trace.append((domain, [x.copy() for x in model.weights], len(order)))
        return r
    ometa.train_pass = tp
    w.train_epoch(0)
    om.train_epoch()
    names = base.model.layout.names
    print(prec, "passes", len(gpu_trace), len(ora_trace))
    for k, (g, o_) in enumerate(zip(gpu_trace, ora_trace)):
        assert g[0] == o_[0]
        errs = [rel_err(a, b) for a, b in zip(g[1], o_[1])]
        print("pass %2d dom %d n=%5d " % (k, g[0], o_[2]) + " ".join("%s=%.1e" % (n[:5], x) for n, x in zip(names, errs)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
