"""Diagnostic (not a test): the HBM-bound kernels named by the north star at HBM scale, each launched a few times -- the target of
`ncu --set full -k regex:<kernel>` captures (profiles/r2_micro_ncu.json) and of CUDA-event timing.
  gather_kernel (K1), adam_kernel (K7), meta_kernel (DN update sweep, K9), adam_table_kernel (K6+K7 fused), lg_window_sum_kernel
  (the one pass over the gradient rows of the multi-CTA scatter-add, K6)."""
import sys
import numpy as np
import torch
sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from gpu_util import ctx, ptr, stream


def main(reps=3):
    c = ctx()
    dev = "cuda"
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)   # noqa: E731

    def timed(name, fn, gb):
        for _ in range(2):
            fn()
        a, b = ev(), ev()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        out[name] = {"ms": ms, "alg_GB": gb, "GBps": gb / (ms * 1e-3)}
        print("%-18s %.3f ms  %.2f GB algorithmic  %.0f GB/s" % (name, ms, gb, gb / (ms * 1e-3)))

    rows, dim, n = 1 << 21, 128, 1 << 22
    table = torch.empty(rows, dim, device=dev).normal_()
    ids = torch.randint(0, rows, (n,), dtype=torch.int32, device=dev)
    dst = torch.empty(n, dim, device=dev)
    timed("gather", lambda: c.call("mamdr_gather_f32", ptr(table), rows, dim, ptr(ids), n, ptr(dst), dim, stream()), 2.0 * n * dim * 4 / 1e9)
    del table, ids, dst
    P = 79301152
    p = torch.empty(P, device=dev).normal_()
    m = torch.zeros(P, device=dev)
    v = torch.zeros(P, device=dev)
    g = torch.empty(P, device=dev).normal_()
    state = torch.zeros(c.lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device=dev)
    c.call("mamdr_opt_state_init", ptr(state), 0.9, 0.999, stream())
    timed("adam", lambda: c.call("mamdr_adam_step", ptr(p), ptr(m), ptr(v), ptr(g), P, ptr(state), 1e-3, 0.9, 0.999, 1e-8, stream()), 28.0 * P / 1e9)
    # DN update sweep: theta += (model - theta) * beta, model <- theta : 2 reads + 2 writes per parameter
    timed("dn_update", lambda: c.call("mamdr_dn_update", ptr(m), ptr(p), 0.1, P, ptr(p), stream()), 16.0 * P / 1e9)
    del g
    rows_t, dim_t = 445789, 128
    n_el = rows_t * dim_t
    tp, tm, tv = p[:n_el], m[:n_el], v[:n_el]
    slot = torch.full((rows_t,), -1, dtype=torch.int32, device=dev)
    tws = torch.zeros(c.lib.mamdr_adam_table_workspace_bytes(), dtype=torch.uint8, device=dev)
    uids = torch.unique(torch.randint(0, rows_t, (1024,), dtype=torch.int32, device=dev))
    urows = torch.randn(1024, dim_t, device=dev)
    ucnt = torch.tensor([uids.numel()], dtype=torch.int32, device=dev)
    targs = (ptr(tp), ptr(tm), ptr(tv), rows_t, dim_t, ptr(uids), ptr(urows), ptr(ucnt), 1024, ptr(slot), 1e-5,
             ptr(state), 1e-3, 0.9, 0.999, 1e-8, None, ptr(tws), tws.numel(), stream())
    timed("table_adam", lambda: c.call("mamdr_adam_table_step", *targs), (24.0 * n_el + 4.0 * rows_t) / 1e9)
    del p, m, v
    n_s, dim_s = 1 << 21, 128
    zi = torch.from_numpy((np.random.default_rng(5).zipf(1.05, n_s) - 1).clip(0, 445788).astype("int32")).to(dev)
    grows = torch.randn(n_s, dim_s, device=dev)
    uo = torch.empty(n_s, dtype=torch.int32, device=dev)
    ro = torch.empty(n_s, dim_s, device=dev)
    nu = torch.zeros(1, dtype=torch.int32, device=dev)
    sws = torch.zeros(c.lib.mamdr_scatter_large_workspace_bytes(n_s, dim_s), dtype=torch.uint8, device=dev)
    sargs = (ptr(zi), ptr(grows), dim_s, n_s, dim_s, ptr(uo), ptr(ro), ptr(nu), ptr(sws), sws.numel(), stream())
    c.call("mamdr_scatter_dedup_large_f32", *sargs)
    torch.cuda.synchronize()
    u = int(nu.item())
    timed("scatter (whole)", lambda: c.call("mamdr_scatter_dedup_large_f32", *sargs), (n_s + u) * dim_s * 4.0 / 1e9)
    print("scatter: %d rows -> %d unique ids" % (n_s, u))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
