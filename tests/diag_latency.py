"""Diagnostic (not a test): where a mini-batch's time goes -- graph replay of one domain pass per precision,
host time of a meta-step (enqueue only) vs device time."""
import sys
import time

import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from conftest import make_config
import run


def main():
    for prec in sys.argv[1:] or ("fp32", "tf32", "tf32x3"):
        c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 1.0,
                           "train.sample_num": 5, "b200.precision": prec})
        w = run.build(c)
        w.prepare()
        base = w.base_model
        m = base.model
        # biggest domain pass
        idx = max(base.dataset.train_dataset, key=lambda i: base.dataset.train_dataset[i]['n_step'])
        d = base.dataset.train_dataset[idx]
        for _ in range(3):
            m.fit_pass(d['data'], d['n_step'])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            m.fit_pass(d['data'], d['n_step'])
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / reps / d['n_step']
        print("%s: graph replay of a %d-step pass: %.1f us / mini-batch" % (prec, d['n_step'], us))
        for _ in range(2):
            w.train_epoch(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w.train_epoch(0)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print("%s: meta-step host enqueue %.1f ms, total %.1f ms, samples %d" % (prec, 1e3 * (t1 - t0), 1e3 * (t2 - t0), base.samples_trained // 3))


if __name__ == "__main__":
    main()
