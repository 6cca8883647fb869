"""Diagnostic (not a test): pass-by-pass drift of the STAR model weights vs the fp64 oracle over a MAMDR meta-step."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import make_config, rel_err
import run
from mamdr_b200.schedule import Schedule
import oracle.meta as ometa
from oracle.meta import MetaSubset, OracleMAMDR
from oracle.star import OracleStar, StarSpec

c = make_config(**{"model.name": "star_meta_mamdr_finetune", "model.norm": "pn", "model.dense": "star", "train.meta_parms": ["emb", "kernel_shared", "bias_shared"],
                   "dataset.synthetic.scale": 0.05, "b200.precision": "fp32", "b200.cuda_graphs": len(sys.argv) < 2})
w = run.build(c)
w.prepare()
base = w.base_model
m = base.model
spec = StarSpec(base.n_uid, base.n_pid, base.n_domain)
o = OracleStar(spec, m.layout.unpack(m.params.cpu().numpy()), base.dataset.user_table, base.dataset.item_table, lr=1e-3, dtype=np.float64)
meta_names = [p.name for p in w.model_meta_parms]
idx = [i for i, p in enumerate(m.trainable_weights) if p.name in meta_names]
om = OracleMAMDR(MetaSubset(o, idx), base.dataset.host_splits(), c['train'], 1024, Schedule(123),
                 {k: [v.numpy()[i] for i in idx] for k, v in w.domain_weights.items()}, name=c['model']['name'])
om.meta_weights = [w.meta_weights.numpy()[i].astype(np.float64) for i in idx]
base.schedule = Schedule(123)
gpu_trace, ora_trace = [], []
orig_rtp = base.run_train_pass
def rtp(domain_idx, steps=None):
    r = orig_rtp(domain_idx, steps)
    gpu_trace.append((domain_idx, m.layout.unpack(m.params.cpu().numpy())))
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
names = m.layout.names
show = [names.index(n) for n in ('kernel_shared0', 'kernel_specific0', 'bias_shared0', 'gamma_shared', 'kernel_shared2', 'out_kernel')]
for k, (g, o_) in enumerate(zip(gpu_trace, ora_trace)):
    assert g[0] == o_[0]
    print("pass %2d dom %d n=%5d " % (k, g[0], o_[2]) + " ".join("%s=%.1e" % (names[i][:9] + names[i][-1], rel_err(g[1][i], o_[1][i])) for i in show))
print("theta  :", " ".join("%s=%.1e" % (names[i][:9] + names[i][-1], rel_err(w.meta_weights.numpy()[i], om.meta_weights[k])) for k, i in enumerate(idx)))
for d in range(10):
    print("theta_%d:" % d, " ".join("%s=%.1e" % (names[i][:9] + names[i][-1], rel_err(w.domain_weights[d].numpy()[i], om.domain_weights[d][k])) for k, i in enumerate(idx)))
print("sequence", w.train_sequence, om.sequence)
