"""Diagnostic (not a test): per-launch durations of the persistent kernel inside one meta-step and the idle gaps between them."""
import sys
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import json
import bench
import run


def main():
    c = bench.load_config("Taobao-10")
    c["b200"]["precision"] = "tf32x3"
    w = run.build(c)
    w.prepare()
    m = w.base_model.model
    for _ in range(3):
        w.train_epoch(0)
    torch.cuda.synchronize()
    m.launch_times = []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    w.train_epoch(0)
    w.train_epoch(0)
    b.record()
    torch.cuda.synchronize()
    lt, m.launch_times = m.launch_times, None
    tot = a.elapsed_time(b)
    print("two meta-steps: %.2f ms; launches %d" % (tot, len(lt)))
    prev_end = a
    for k, (s, e, n) in enumerate(lt):
        print("launch %2d: gap before %.3f ms, duration %.3f ms, %4d mini-batches, %.1f us per mini-batch" % (k, prev_end.elapsed_time(s), s.elapsed_time(e), n, 1e3 * s.elapsed_time(e) / max(n, 1)))
        prev_end = e
    print("tail gap %.3f ms" % prev_end.elapsed_time(b))


if __name__ == "__main__":
    main()
