"""Diagnostic (not a test): step-by-step divergence of the MTL path vs the oracle over a joint epoch."""
import sys

import numpy as np
import torch

import conftest  # noqa: F401 (puts the repo root on sys.path)
import run
from conftest import rel_err  # noqa: E402
from mamdr_b200.schedule import Schedule
from test_gpu_mtl import MMOE, SMALL, _cfg, _oracle_for, _weights

SYNC = "--sync" in sys.argv
name, arch = (sys.argv[1], {"MMOE": MMOE, "SMALL": SMALL}[sys.argv[2]]) if len(sys.argv) > 2 and not sys.argv[1].startswith('--') else ("mmoe", MMOE)
c = _cfg(name, arch, True, scale=0.0006)
base = run.build(c)
m = base.model
m.reset_optimizer()
o = _oracle_for(base)
sched = Schedule(5)
bs = base.dataset.batch_size
step = 0
seen = set()
for idx in [3, 0, 7, 0, 3]:
    data = base.dataset.train_dataset[idx]['data']
    order = sched.batch_order(idx, data.n_data)
    data.set_order(order)
    h = data.host
    for s in range(data.n_step):
        rows = min(bs, data.n_data - s * bs)
        loss = torch.zeros(1, device="cuda")
        if SYNC:   # teacher forcing: start every step from the oracle's state
            m.params.copy_(torch.from_numpy(m.layout.pack(o.weights)))
            m.m.copy_(torch.from_numpy(m.layout.pack(o.adam.m)))
            m.v.copy_(torch.from_numpy(m.layout.pack(o.adam.v)))
        m._train_step(data, s * bs, rows, loss)
        sel = order[s * bs:s * bs + rows]
        ol, _ = o.train_on_batch(h['uid'][sel], h['pid'][sel], idx, h['label'][sel])
        torch.cuda.synchronize()
        errs = {n: rel_err(a, b) for n, a, b in zip(m.layout.names, _weights(m), o.weights)}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print("step %d dom %d rows %d loss %.6f / %.6f  worst %s" % (step, idx, rows, loss.item(), ol,
              ", ".join("%s %.2e" % kv for kv in worst)))
        if step <= 2:
            for tn, col in (("user_emb", "uid"), ("expert2_bias0", None)):
                i = m.layout.index(tn)
                a, b = _weights(m)[i], o.weights[i]
                e = np.abs(a - b)
                k = np.unravel_index(np.argmax(e), e.shape)
                print("   ", tn, "argmax", k, "gpu %.6e oracle %.6e" % (a[k], b[k]), "n bad(>1e-5 abs)", int((e > 1e-5).sum()), "of", e.size,
                      ("row in batch: %s, in earlier batches: %s" % (k[0] in set(h[col][sel].tolist()), k[0] in seen)) if col else "")
                mi = m.layout.views(m.m)[i].cpu().numpy()
                vi = m.layout.views(m.v)[i].cpu().numpy()
                print("      m gpu %.6e oracle %.6e   v gpu %.6e oracle %.6e" % (mi[k], o.adam.m[i][k], vi[k], o.adam.v[i][k]))
        seen |= set(h['uid'][sel].tolist())
        step += 1
