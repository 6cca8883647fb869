"""GPU diagnostic: MAML stage-by-stage vs the oracle (first domain of the first epoch)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule
from test_gpu_mlp import _build, _oracle_for, _weights
from oracle.meta import OracleMAML, train_pass

keys = {"model.name": "mlp_meta_maml", "dataset.synthetic.scale": 0.05, "b200.precision": "fp32", "train.meta_split": "meta-train/val",
        "train.meta_split_ratio": 0.8, "train.average_meta_grad": "none", "train.meta_learning_rate": 1e-3}
c = make_config(**keys)
w = _build(c)
base = w.base_model
w.prepare()
o = _oracle_for(w, weights=w.meta_weights.numpy())
om = OracleMAML(o, base.dataset.host_splits(), c['train'], base.dataset.batch_size, Schedule(123), name="mlp_meta_maml")
base.schedule = Schedule(123)
names = w.model.layout.names

def cmp(tag, A, B):
    print(tag, " ".join("%s=%.1e" % (n, rel_err(a, b)) for n, a, b in zip(names, A, B)))

import sys
LOCKSTEP = "--lockstep" in sys.argv
if not LOCKSTEP:
    for e in range(2):
        w.train_epoch(e); om.train_epoch()
        cmp("epoch %d live " % e, _weights(w.model), o.weights)
        cmp("epoch %d theta" % e, w.meta_weights.numpy(), om.meta_weights)
    sys.exit(0)
for e in range(2):
  seq = base.schedule.shuffle_sequence(w.train_sequence); seq2 = om.schedule.shuffle_sequence(om.sequence)
  w.train_sequence = seq; om.sequence = seq2
  assert seq == seq2
  for idx in seq:
      d = w.meta_data_split[idx]; sp = om.split[idx]
      w._set_model_meta_parms(w.meta_weights); o.set_weights(om.meta_weights)
      w._init_iter(d['train_iter']); w._init_iter(d['meta_iter'])
      ot = om._order(idx, sp['train']); omv = om._order(idx, sp['meta'])
      assert np.array_equal(d['train_iter'].order.cpu().numpy()[:len(ot)], ot) and np.array_equal(d['meta_iter'].order.cpu().numpy()[:len(omv)], omv)
      print("domain", idx, "train_step", d['train_step'], sp['train_step'], "meta", d['meta_val_step'], sp['meta_val_step'], "n", d['train_iter'].n_data, d['meta_iter'].n_data)
      w.run_view_train_pass(d['train_iter'], d['train_step'])
      train_pass(o, om.data['train'][idx], idx, ot, om.bs, sp['train_step'])
      cmp("  after inner  ", _weights(w.model), o.weights)
      w.meta_train_pass(d['meta_iter'], d['meta_val_step'])
      om.meta_train_pass(idx, omv, sp['meta_val_step'])
      cmp("  accum        ", w.model.layout.unpack(w.accum_grads.cpu().numpy()), om.accum)
      if max(rel_err(a, b) for a, b in zip(w.model.layout.unpack(w.accum_grads.cpu().numpy()), om.accum)) > 1e-4 and not globals().get("_gate_done"):
          _gate_done = True
          # gate check: at the ORACLE's live weights, the smallest |pre-activation| of the meta batch in fp64, and the oracle
          # gradient with that one gate flipped vs the product's accumulated gradient
          dd = om.data['train'][idx]; sel = omv[:om.bs]
          o64 = _oracle_for(w, weights=[x.astype(np.float64) for x in o.weights], dtype=np.float64)
          H, p = o64.forward(dd['uid'][sel], dd['pid'][sel], idx, train=False)
          for l in range(3):
              Z = H[l] @ o64.w('kernel%d' % l) + o64.w('bias%d' % l)
              r, j = np.unravel_index(np.argmin(np.abs(Z)), Z.shape)
              print("   layer", l, "min |Z| = %.3e at (row %d, unit %d), typical |Z| %.3e" % (abs(Z[r, j]), r, j, np.mean(np.abs(Z))))
          got = w.model.layout.unpack(w.accum_grads.cpu().numpy())
          _, _, g32 = o.gradients(dd['uid'][sel], dd['pid'][sel], idx, dd['label'][sel], train=False)
          diff = got[names.index('bias1')] - g32[names.index('bias1')]
          print("   bias1 gradient difference is concentrated in units", np.argsort(-np.abs(diff))[:3], np.sort(-np.abs(diff))[:3], "max |g|", np.abs(g32[names.index('bias1')]).max())
          diff = got[names.index('bias0')] - g32[names.index('bias0')]
          print("   bias0 gradient difference: units", np.argsort(-np.abs(diff))[:3], np.sort(-np.abs(diff))[:3], "max |g|", np.abs(g32[names.index('bias0')]).max())
      w._set_model_meta_parms(w.meta_weights); o.set_weights(om.meta_weights)
      w.meta_weights = w._meta_train_step(); om.meta_weights = om._meta_train_step()
      cmp("  after meta   ", w.meta_weights.numpy(), om.meta_weights)
      cmp("  meta m       ", w.model.layout.unpack(w._meta_m.cpu().numpy()), om.meta_adam.m)
      cmp("  meta v       ", w.model.layout.unpack(w._meta_v.cpu().numpy()), om.meta_adam.v)
