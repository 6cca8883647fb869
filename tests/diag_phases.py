"""Diagnostic (not a test): per-phase timing of the pass kernel from in-kernel time stamps (globaltimer across
CTAs, clock64 inside a CTA)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from conftest import make_config
import run
from mamdr_b200.engine import _ptr

GHZ = 1.965


def main(prec="tf32"):
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 1.0, "b200.precision": prec})
    w = run.build(c)
    w.prepare()
    base = w.base_model
    m = base.model
    idx = max(base.dataset.train_dataset, key=lambda i: base.dataset.train_dataset[i]['n_step'])
    d = base.dataset.train_dataset[idx]
    steps = min(d['n_step'], 6)
    G = m.ctx.sm_count
    nph = 2
    buf = torch.zeros(steps * nph * G * 16, dtype=torch.int64, device=m.device)
    for _ in range(3):
        m.fit_pass(d['data'], steps)
    m.ctx.call("mamdr_debug_pass_timing", _ptr(buf), buf.numel())
    m.fit_pass(d['data'], steps)
    torch.cuda.synchronize()
    m.ctx.call("mamdr_debug_pass_timing", None, 0)
    t = buf.cpu().numpy().reshape(steps, nph, G, 16).astype(np.float64)
    names = ["chain", "dW+update"]
    print(prec, "steps", steps, "G", G)
    s = 2
    base_t = t[s, 0, :, 0].min()
    print("step duration %.2f us" % ((t[s + 1, 0, :, 0].min() - base_t) / 1e3))
    for p in range(nph):
        st, en = t[s, p, :, 0], t[s, p, :, 1]
        busy = (en - st) / 1e3
        nxt = t[s, p + 1, :, 0].min() if p + 1 < nph else t[s + 1, 0, :, 0].min()
        print("%-16s start %.2f  work max %.2f med %.2f us  barrier %.2f" % (names[p], (st.min() - base_t) / 1e3, busy.max(), np.median(busy), (nxt - en.max()) / 1e3))
        if p < 2:
            # the 3 busiest CTAs: clock64 deltas from phase start (us)
            pick = list(np.argsort(-busy)[:3]) + ([0, 5, 16] if p == 1 else [])   # dW phase: + a domain job, the column-sum job
            for cta in pick:
                c0 = t[s, p, cta, 7]
                rel = [(t[s, p, cta, k] - c0) / GHZ / 1e3 if t[s, p, cta, k] > 0 else float('nan') for k in (2, 3, 4, 5, 6)]
                ep = " ".join("%d:%.2f" % (k, (t[s, p, cta, k] - c0) / GHZ / 1e3) for k in range(8, 16) if t[s, p, cta, k] > 0)
                print("    cta %3d busy %.2f us | tma first-issue %.2f all-issued %.2f | mma first-full %.2f all-issued %.2f | done-seen %.2f | stamps (chain: segment epilogues done; dW: 8 partial published, 9 tile complete, 10 slice applied; domain job: 11 db0, 12 E_d applied, 13 done; 14 column-sum job done) %s" % ((cta, busy[cta]) + tuple(rel) + (ep,)))


if __name__ == "__main__":
    for p in sys.argv[1:] or ["tf32", "tf32x3"]:
        main(p)
