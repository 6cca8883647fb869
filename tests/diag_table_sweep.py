"""Diagnostic (not a test): grid-size A/B of the fused table sweep (mamdr_adam_table_step) on the Amazon-13 user table."""
import os
import sys

import torch

import conftest  # noqa: F401
from gpu_util import ctx, ptr, stream

c = ctx()
rows, dim = 502222, 128
p = torch.randn(rows, dim, device="cuda") * 0.05
m, v = torch.zeros_like(p), torch.zeros_like(p)
state = torch.zeros(c.lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device="cuda")
c.call("mamdr_opt_state_init", ptr(state), 0.9, 0.999, stream())
slot = torch.full((rows,), -1, dtype=torch.int32, device="cuda")
ws = torch.zeros(c.lib.mamdr_adam_table_workspace_bytes(), dtype=torch.uint8, device="cuda")
ids = torch.unique(torch.randint(0, rows, (1024,), dtype=torch.int32, device="cuda"))
srows = torch.randn(1024, dim, device="cuda")
cnt = torch.tensor([ids.numel()], dtype=torch.int32, device="cuda")
args = (ptr(p), ptr(m), ptr(v), rows, dim, ptr(ids), ptr(srows), ptr(cnt), 1024, ptr(slot), 1e-5, ptr(state), 1e-3, 0.9, 0.999, 1e-8,
        None, ptr(ws), ws.numel(), stream())
gb = (24.0 * rows * dim + 4.0 * rows) / 1e9
for per_sm in [int(x) for x in (sys.argv[1:] or ["8", "3", "6", "9", "12", "16", "8"])]:
    os.environ["MAMDR_TABLE_CTAS_PER_SM"] = str(per_sm)
    for _ in range(3):
        c.call("mamdr_adam_table_step", *args)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        c.call("mamdr_adam_table_step", *args)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print("ctas/SM %2d  grid %5d  %.1f us  %.0f GB/s algorithmic" % (per_sm, 148 * per_sm, 1e3 * ms, gb / (ms * 1e-3)))
