"""Diagnostic (not a test): per-chunk stamps of ONE chain segment of the pass kernel.  Needs a library built with
MAMDR_NVCC_EXTRA=-DPASS_DBG_SEG=<s>: slots 2..6 = producer issue time of chunks 0, 2, 4, 6, 7 of the segment, slots 8..15 =
time the MMA warp saw chunks 0..7 landed."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from conftest import make_config
import run
from mamdr_b200.engine import _ptr

GHZ = 1.965


def main(prec="tf32x3"):
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
    s = 2
    for cta in (0, 1, 17, 40, 63):
        c0 = t[s, 0, cta, 7]
        prod = " ".join("%.2f" % ((t[s, 0, cta, k] - c0) / GHZ / 1e3) for k in (2, 3, 4, 5, 6))
        mma = " ".join("%.2f" % ((t[s, 0, cta, k] - c0) / GHZ / 1e3) for k in range(8, 16))
        print("cta %3d | producer issue of chunks 0 2 4 6 7: %s | MMA saw chunks 0..7 landed: %s" % (cta, prod, mma))


if __name__ == "__main__":
    main(*sys.argv[1:])
