"""bench.py -- MAMDR meta-train throughput on the synthetic Taobao-10 shape (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload Taobao-10]

A "step" is ONE MAMDR meta-step (meta-epoch): DN over the shuffled domain sequence + DR
(sample_num + 1 support/query pass pairs per query domain) at batch 1024, exactly the body of the
epoch loop of /root/reference/model_zoo/mamdr.py:41-143, validation / test excluded.
value = real samples pushed through training mini-batches per second, whole job.

  --impl b200       our path: one process per GPU (torchrun for N > 1), DN replicated, DR query domains
                    sharded, one NCCL all-reduce per meta-step.  Prints ONE JSON line on rank 0.
  --impl reference  the CPU oracle restatement of the reference's TF path (TF 1.12 cannot be installed:
                    BASELINE.md section 2) timed on the host cores, a bounded sample per step.
  --workload W      secondary workloads (Taobao-10-batch / -20 / -20-star / -30, Amazon-6, Amazon-13-{mmoe,ple}[-sharded],
                    Amazon-13-sharded); the default line is config #1.
  --virtual-ranks V opt-in, a SEPARATE line: V DR chains side by side on SM partitions of one GPU (the V-rank sharded
                    schedule, bit-identical to V real ranks); the default line also reports it as `fill_the_machine`.
  --graphs          sharded workloads: the tower part of a step is replayed from a CUDA graph (collectives stay eager).
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD_CONFIG = {"Taobao-10": "config/Taobao-10/deepctr_DN+DR.json", "Taobao-10-batch": "config/Taobao-10/deepctr_DN+DR.json", "Taobao-20": "config/Taobao_20/deepctr_DN+DR.json",
                   "Taobao-20-star": "config/Taobao_20/star_DN+DR.json",
                   "Taobao-30": "config/Taobao_30/deepctr_DN+DR.json", "Amazon-6": "config/Amazon_6/deepctr.json",
                   "Amazon-13-sharded": "config/Amazon_6/deepctr.json", "Amazon-13-mmoe": "config/Amazon_13/mmoe_DN.json",
                   "Amazon-13-ple": "config/Amazon_13/ple_DN.json", "Amazon-13-mmoe-sharded": "config/Amazon_13/mmoe_DN.json",
                   "Amazon-13-ple-sharded": "config/Amazon_13/ple_DN.json"}
METRIC = "MAMDR meta-train samples/sec (Taobao-10 shape)"
_emit = lambda line: print(json.dumps(line), flush=True)   # noqa: E731  (main() re-binds it to the real stdout)


def load_config(workload):
    with open(os.path.join(ROOT, WORKLOAD_CONFIG[workload])) as f:
        c = json.load(f)
    c.setdefault("b200", {})["verbose"] = False
    if workload.endswith("-batch"):     # the `batch` variant of the wrapper (mamdr.py:100-108,182-196): (query, support) PAIRS are the shards
        c["model"]["name"] = "mlp_meta_mamdr_batch"
    c["train"]["result_save_path"] = "/tmp/mamdr_bench/result"
    c["train"]["checkpoint_path"] = "/tmp/mamdr_bench/checkpoint"
    return c


# ---- clocks during the timed region (B200_PROFILING.md) ------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.t0 = self.t1 = None

    def start(self):
        """Start sampling (20 ms period).  Call `mark_begin()` / `mark_end()` around the timed region: the summary uses the samples
        taken inside it.  NVML is queried IN-PROCESS (pynvml: clocks.sm, clocks.max.sm, the clocks-event-reasons bit mask -- the same
        counters `nvidia-smi --query-gpu` prints): eight `nvidia-smi -lms 20` child processes polling the driver during a 70 ms
        multi-GPU timed region measurably slowed the launches they were meant to observe (88 -> 68 M samples/s at 8 GPUs).  Falls
        back to the `nvidia-smi` loop if pynvml is unavailable."""
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(int(self.gpu))
            self.stop_flag = False
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            self.proc = True
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        bits = [("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", "nvmlClocksThrottleReasonSwPowerCap")]
        masks = [(n, getattr(nv, a, None) or getattr(nv, b, 0)) for n, a, b in bits]
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                row = ["", str(sm), str(mx), "", ""] + ["Active" if (r & m) else "Not Active" for _, m in masks]
                self.rows.append((time.time(), row))
            except Exception:
                pass
            time.sleep(0.02)

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.t1 = time.time()
        time.sleep(0.05)
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [r for t, r in self.rows if t0 <= t <= self.t1 + 0.02]
        window = "timed region"
        if not inside:
            inside = [r for t, r in self.rows if self.t1 - 1.0 <= t <= self.t1 + 0.02]
            window = "the loaded second before the end of the timed region (the region is shorter than the sampling period)"
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in inside:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window, "source": "NVML in-process (pynvml)" if self.nvml is not None else "nvidia-smi -lms 20"}


# ---- the reference arm / cpu_baseline: the oracle on the host cores --------------------------------------
def oracle_sample_runner(config, max_batches):
    """Returns (run_once, cores, description).  run_once() executes the first `max_batches`
    mini-batches of a MAMDR meta-step (DN passes first, then DR) with the CPU oracle and returns
    (samples, seconds)."""
    import numpy as np
    import torch
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    from mamdr_b200.schedule import Schedule
    from oracle import build as obuild
    from oracle.mlp import MLPSpec, OracleMLP
    try:
        obuild.build()
    except Exception:
        pass
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = config["dataset"]["synthetic"]
    g = synth.generate(sc["shape"], seed=sc.get("seed", config["dataset"]["seed"]), scale=sc.get("scale", 1.0),
                       signal=sc.get("signal", 1.0))
    mc = config["model"]
    emb = (mc["user_dim"], mc["item_dim"], mc["domain_dim"])
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], emb, mc["hidden_dim"], False)
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], emb, tuple(mc["hidden_dim"]), dropout=mc["dropout"])
    model = OracleMLP(spec, init_mlp_weights(lo, [123, 0]), g["user_emb"], g["item_emb"],
                      lr=config["train"]["learning_rate"])
    bs = config["dataset"]["batch_size"]
    from oracle.meta import OracleMAMDR
    data = {"train": g["train"], "val": g["val"], "test": g["test"]}
    om = OracleMAMDR(model, data, config["train"], bs, Schedule(config["dataset"]["seed"]),
                     {d: init_mlp_weights(lo, [123, d + 1]) for d in range(g["n_domain"])}, name=mc["name"])

    class _Budget(Exception):
        pass

    def run_once():
        """The reference's own schedule (mamdr.py:41-116: DN over the shuffled domains, then the DR chains), stopped after
        max_batches mini-batches -- a bounded sample of one meta-step."""
        state = {"done": 0, "samples": 0}
        inner = model.train_on_batch

        def counted(uid, pid, domain, label, **kw):
            if state["done"] >= max_batches:
                raise _Budget()
            r = inner(uid, pid, domain, label, **kw)
            state["done"] += 1
            state["samples"] += len(uid)
            return r
        model.train_on_batch = counted
        t0 = time.perf_counter()
        try:
            om.train_epoch()
        except _Budget:
            pass
        finally:
            model.train_on_batch = inner
        return state["samples"], time.perf_counter() - t0

    desc = ("the first %d mini-batches (batch %d) of a meta-step in the reference's order -- the DN phase over the shuffled "
            "domains, then DR chains (support pass, query pass, theta_i update) -- on synthetic %s, CPU oracle (numpy + torch-CPU "
            "GEMM + C Philox/AUC)" % (max_batches, bs, sc["shape"]))
    return run_once, cores, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    config = load_config(args.workload)
    per_step = 300
    run_once, cores, desc = oracle_sample_runner(config, per_step)
    for _ in range(args.warmup):
        run_once()
    tot_s, tot_t = 0, 0.0
    for _ in range(args.steps):
        s, t = run_once()
        tot_s += s
        tot_t += t
    v = tot_s / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_desc(config, args.workload, args.gpus), "precision": "fp32 (CPU)",
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def workload_desc(config, workload, n_gpus, model=None):
    """`config` of the JSON line: the workload only (identical for both arms); the arithmetic mode is reported beside it
    (`precision`)."""
    from mamdr_b200 import synth
    tc = config["train"]
    mc = config["model"]
    sc = config["dataset"].get("synthetic", {})
    frozen = not tc.get("emb_trainable", False)
    emb = "%s %d-d embeddings" % ("frozen (pretrained)" if frozen else "trainable", mc["user_dim"])
    _, n_uid, n_pid = synth.SHAPES[sc.get("shape", workload)][:3]
    scale = sc.get("scale", 1.0)
    table_mb = 4.0 * (int(n_uid * scale) * mc["user_dim"] + int(n_pid * scale) * mc["item_dim"]) / 1e6
    return {"workload": "%s synthetic %s: DN + DR(sample_num=%d%s), batch %d, %s %s, %s, one meta-step per bench step"
                        % (mc["name"], workload, tc["sample_num"], "+query" if tc["add_query_domain"] else "",
                           config["dataset"]["batch_size"], "star (PartitionedNorm + StarFCN)" if "star" in mc["name"] else "mlp",
                           "/".join(str(h) for h in mc["hidden_dim"]), emb),
            "parallelism": "dr-shard%d" % n_gpus if n_gpus > 1 else "single",
            "l2": "working set (%.1f MB of %s tables + the parameter arena) is L2-resident by construction; the b200 arm writes a "
                  "256 MiB buffer between timed steps to flush L2" % (table_mb, "frozen" if frozen else "trainable")}


PRECISION_NOTE = ("fp32 storage everywhere; tf32x3 = tcgen05 kind::tf32 MMAs on round-to-nearest hi / lo operand pairs (3 products, "
                  "dropped term 2^-24), 4 TMEM accumulator groups summed with RN adds; tf32 = 1 pass; fp32 = FFMA SIMT (the mode "
                  "the north-star 1e-4 bar is stated for: parity_mode in this line); the reference arm computes in fp32 on the CPU")


# ---- roofline of the dominant kernels (measured live with CUDA events) -----------------------------------
def micro_rooflines(model, peaks, torch):
    """HBM-scale micro-benchmarks of the two HBM-bound kernels named by the north star (gather, fused
    Adam sweep) + the per-step launch mix.  Algorithmic bytes: gather 2*n*dim*4; Adam 28 B/param."""
    import ctypes as C
    from mamdr_b200.engine import _ptr
    ctx = model.ctx
    out = {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    st = model.stream
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    # gather: 4 Mi rows of 128 floats from a 2 Mi-row table (1 GiB) -> 2 GiB written, 2+ GiB read
    rows, dim, n = 1 << 21, 128, 1 << 22
    table = torch.empty(rows, dim, device=model.device).normal_()
    ids = torch.randint(0, rows, (n,), dtype=torch.int32, device=model.device)
    dst = torch.empty(n, dim, device=model.device)
    for _ in range(3):
        ctx.call("mamdr_gather_f32", _ptr(table), rows, dim, _ptr(ids), n, _ptr(dst), dim, st)
    a, b = ev(), ev()
    reps = 10
    a.record()
    for _ in range(reps):
        ctx.call("mamdr_gather_f32", _ptr(table), rows, dim, _ptr(ids), n, _ptr(dst), dim, st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gb = 2.0 * n * dim * 4 / 1e9
    out["gather"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": hbm, "unit": "GB/s",
                     "frac": gb / (ms * 1e-3) / hbm, "rows": n, "dim": dim, "ms": ms}
    del table, ids, dst
    # Adam sweep at the Amazon-6 arena size (79.3 M params)
    P = 79301152
    p = torch.empty(P, device=model.device).normal_()
    m = torch.zeros(P, device=model.device)
    v = torch.zeros(P, device=model.device)
    g = torch.empty(P, device=model.device).normal_()
    state = torch.zeros(ctx.lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device=model.device)
    ctx.call("mamdr_opt_state_init", _ptr(state), 0.9, 0.999, st)
    for _ in range(3):
        ctx.call("mamdr_adam_step", _ptr(p), _ptr(m), _ptr(v), _ptr(g), P, _ptr(state), 1e-3, 0.9, 0.999, 1e-8, st)
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        ctx.call("mamdr_adam_step", _ptr(p), _ptr(m), _ptr(v), _ptr(g), P, _ptr(state), 1e-3, 0.9, 0.999, 1e-8, st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gb = 28.0 * P / 1e9
    out["adam"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": hbm, "unit": "GB/s",
                   "frac": gb / (ms * 1e-3) / hbm, "params": P, "ms": ms}
    del g
    # fused table sweep (sparse rows + l2 + non-lazy Adam) on the Amazon-6 user table: 24 B per element
    rows_t, dim_t = 445789, 128
    n_el = rows_t * dim_t
    tp, tm, tv = p[:n_el], m[:n_el], v[:n_el]
    slot = torch.full((rows_t,), -1, dtype=torch.int32, device=model.device)
    tws = torch.zeros(ctx.lib.mamdr_adam_table_workspace_bytes(), dtype=torch.uint8, device=model.device)
    uids = torch.unique(torch.randint(0, rows_t, (1024,), dtype=torch.int32, device=model.device))
    urows = torch.randn(1024, dim_t, device=model.device)
    ucnt = torch.tensor([uids.numel()], dtype=torch.int32, device=model.device)
    args = (_ptr(tp), _ptr(tm), _ptr(tv), rows_t, dim_t, _ptr(uids), _ptr(urows), _ptr(ucnt), 1024, _ptr(slot), 1e-5,
            _ptr(state), 1e-3, 0.9, 0.999, 1e-8, None, _ptr(tws), tws.numel(), st)
    for _ in range(3):
        ctx.call("mamdr_adam_table_step", *args)
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        ctx.call("mamdr_adam_table_step", *args)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gb = (24.0 * n_el + 4.0 * rows_t) / 1e9
    out["table_adam"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": hbm, "unit": "GB/s",
                         "frac": gb / (ms * 1e-3) / hbm, "rows": rows_t, "dim": dim_t, "ms": ms}
    # the finetune stage's plain SGD over a trainable table (mamdr_sgd_table_step): 8 B per element
    sargs = (_ptr(tp), rows_t, dim_t, _ptr(uids), _ptr(urows), _ptr(ucnt), 1024, _ptr(slot), 1e-5, 1e-3, None, _ptr(tws), tws.numel(), st)
    for _ in range(3):
        ctx.call("mamdr_sgd_table_step", *sargs)
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        ctx.call("mamdr_sgd_table_step", *sargs)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    gb = (8.0 * n_el + 4.0 * rows_t) / 1e9
    out["table_sgd"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": hbm, "unit": "GB/s",
                        "frac": gb / (ms * 1e-3) / hbm, "rows": rows_t, "dim": dim_t, "ms": ms}
    del tp, tm, tv, p, m, v
    # sparse-gradient de-duplication at HBM scale (multi-CTA path): 2 Mi gradient rows x 128 over Zipf ids -> read n x 512 B once,
    # write u x 512 B
    n_s, dim_s = 1 << 21, 128
    zi = torch.from_numpy((__import__("numpy").random.default_rng(5).zipf(1.05, n_s) - 1).clip(0, 445788).astype("int32")).to(model.device)
    grows = torch.randn(n_s, dim_s, device=model.device)
    uo = torch.empty(n_s, dtype=torch.int32, device=model.device)
    ro = torch.empty(n_s, dim_s, device=model.device)
    nu = torch.zeros(1, dtype=torch.int32, device=model.device)
    sws = torch.zeros(ctx.lib.mamdr_scatter_large_workspace_bytes(n_s, dim_s), dtype=torch.uint8, device=model.device)
    sargs = (_ptr(zi), _ptr(grows), dim_s, n_s, dim_s, _ptr(uo), _ptr(ro), _ptr(nu), _ptr(sws), sws.numel(), st)
    for _ in range(3):
        ctx.call("mamdr_scatter_dedup_large_f32", *sargs)
    a, b = ev(), ev()
    a.record()
    for _ in range(reps):
        ctx.call("mamdr_scatter_dedup_large_f32", *sargs)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    u = int(nu.item())
    gb = (n_s + u) * dim_s * 4.0 / 1e9
    out["scatter"] = {"bound": "hbm", "achieved": gb / (ms * 1e-3), "peak": hbm, "unit": "GB/s", "frac": gb / (ms * 1e-3) / hbm,
                      "rows": n_s, "unique": u, "dim": dim_s, "ms": ms,
                      "note": "whole call: 4-pass radix sort of (id, position) + head scan + ONE pass over the gradient rows"}
    return out


def run_amazon(args):
    """BASELINE config #2 (secondary workload, not the driver's default): the joint `mlp` baseline with TRAINABLE
    embedding tables on the synthetic Amazon-6 shape (79.3 M parameters).  A step = `mb` consecutive mini-batches of
    the largest domain's training pass (gather from the arena tables, fp32 tower, sort/segment-sum dedup, fused
    L2 + non-lazy Adam sweep over every table row).  HBM-bound: 24 B per table element per mini-batch."""
    import torch
    import run as runpy
    config = load_config(args.workload)
    # config #2 (the mlp tower): tcgen05 pass kernel per mini-batch by default; the multi-task towers are fp32 SIMT
    prec = (args.precision or "tf32x3") if args.workload == "Amazon-6" else "fp32"
    config["b200"]["precision"] = prec
    base = runpy.build(config)
    base = getattr(base, "base_model", base)   # config #5 wraps the MTL base model in DomainNegotiation
    model = base.model
    model.reset_optimizer()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    idx = max(base.dataset.train_dataset, key=lambda i: base.dataset.train_dataset[i]['n_step'])
    data = base.dataset.train_dataset[idx]['data']
    mb = 100
    for _ in range(max(args.warmup, 3)):
        model.fit_pass(data, mb)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    evs = []
    launches0 = sum(lm.ctx.launches for lm in lane_models)
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        model.fit_pass(data, mb)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    n_table = sum(r * d for _, r, d, _ in model._tables)
    mtl = hasattr(model, "spans")
    # dense part: the whole arena for the mlp; for a multi-task tower only the spans of the sub-model that trains
    n_dense = sum(int(x) for x in model.spans[idx][1]) if mtl else model.params.numel() - n_table
    alg_bytes = 24.0 * n_table + 28.0 * n_dense
    hbm = peaks.get("hbm_gbs", 6650.0)
    step_achieved = alg_bytes * mb / (ms * 1e-3) / 1e9
    shape = config["dataset"]["synthetic"]["shape"]
    # the dominant kernel (fused table sweep), timed live on the model's own tables: CUDA events around 10 launches per table
    import ctypes as C
    from mamdr_b200.engine import _ptr
    k_bytes, k_ms, k_n = 0.0, 0.0, 0
    for off, n_rows, dim, slot in model._tables:
        n_el = n_rows * dim
        targs = (_ptr(model.params[off:off + n_el]), _ptr(model.m[off:off + n_el]), _ptr(model.v[off:off + n_el]), n_rows, dim, None, None,
                 None, 0, _ptr(slot), model.l2_emb, _ptr(model.opt_state), model.lr, model.beta1, model.beta2, model.eps, None,
                 _ptr(model.table_ws), model.table_ws_bytes, model.stream)
        for _ in range(2):
            model.ctx.call("mamdr_adam_table_step", *targs)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            model.ctx.call("mamdr_adam_table_step", *targs)
        b.record()
        torch.cuda.synchronize()
        k_ms += a.elapsed_time(b)
        k_bytes += 10 * (24.0 * n_el + 4.0 * n_rows)
        k_n += 10
    achieved = k_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    if shape == "Amazon-13":
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_v5_table_sweep_ncu.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {"metric": "joint-train samples/sec (%s shape, trainable 128-d tables)" % shape, "value": mb * 1024 / (ms * 1e-3),
            "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s train steps of sub-model %d, trainable user/item tables, synthetic %s (%d + %d rows x 128), "
                                   "%d mini-batches of 1024 per step" % (config["model"]["name"], idx, shape, model.n_uid, model.n_pid, mb),
                       "precision": prec, "l2": "tables + Adam slots (%.0f MB per sweep) are far beyond L2" % (12e-6 * n_table)},
            "gpu_launches": model.ctx.launches - launches0, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json (burst copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                         "kernel": "adam_table_kernel (fused sparse merge + l2 + non-lazy Adam over every table row; 2 launches per mini-batch), "
                                   "timed live on the model's tables",
                         "alg_bytes_per_launch": k_bytes / k_n, "avg_launch_us": 1e3 * k_ms / k_n,
                         "kernel_share_of_step": 2.0 * (k_ms / k_n) / (ms / mb),
                         "step_achieved": step_achieved, "step_frac": step_achieved / hbm,
                         "alg_bytes_per_minibatch": alg_bytes, "us_per_minibatch": 1e3 * ms / mb,
                         "note": "achieved / frac = the table sweep alone (24 B per table element + 4 B per row); step_* = all algorithmic "
                                 "bytes of a mini-batch (tables + 28 B per dense parameter that trains) over the whole step time"}}
    _emit(line)


def run_sharded(args):
    """BASELINE config #5's table layout (secondary workload): trainable 128-d tables of the synthetic Amazon-13 shape
    (502 222 + 215 403 rows) row-sharded over the ranks, joint `mlp` steps data-parallel over a 1024-row batch, NCCL
    all-to-all for ids / rows / gradient rows + one all-reduce of the dense gradients per step (mamdr_b200/sharded.py)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    from mamdr_b200.sharded import ShardedJointTrainer
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29577")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_uid, n_pid, D = 502222, 215403, 13
    lo = mlp_layout(n_uid, n_pid, D, (128, 128, 128), (256, 128, 64), True)
    rng = np.random.Generator(np.random.PCG64(7))
    user0 = (rng.standard_normal((n_uid, 128)) * 1e-4).astype(np.float32)
    item0 = (rng.standard_normal((n_pid, 128)) * 1e-4).astype(np.float32)
    lo_d = mlp_layout(n_uid, n_pid, D, (128, 128, 128), (256, 128, 64), False)
    dense0 = init_mlp_weights(lo_d, [7, 0])
    tower = "mlp"
    if args.workload != "Amazon-13-sharded":      # config #5 proper: the MMOE / PLE sub-models over the sharded tables
        from mamdr_b200.deep_mtl_ctr import MTLTopology, init_mtl_weights
        from mamdr_b200.sharded import ShardedMTLTrainer
        mc = load_config(args.workload)["model"]
        tower = "ple" if "ple" in mc["name"] else "mmoe"
        arch = dict(expert_hidden=tuple(mc["hidden_dim"]), tower_hidden=tuple(mc["tower_hidden_dim"]), gate_hidden=tuple(mc["gate_dnn_hidden_units"]),
                    num_experts=mc.get("num_experts", 0), specific_expert_num=mc.get("specific_expert_num", 0),
                    shared_expert_num=mc.get("shared_expert_num", 0))
        topo_d = MTLTopology(tower, 4, 4, D, (128, 128, 128), emb_trainable=False, **arch)
        t = ShardedMTLTrainer(tower, n_uid, n_pid, D, user0, item0, init_mtl_weights(topo_d.layout, [7, 0]), dropout=mc["dropout"], lr=1e-4,
                              batch_size=1024, device="cuda:%d" % local_rank, use_graphs=args.graphs, **arch)
    else:
        t = ShardedJointTrainer(n_uid, n_pid, D, user0, item0, dense0, dropout=0.5, batch_size=1024, device="cuda:%d" % local_rank,
                                use_graphs=args.graphs)
    g = torch.Generator(device="cuda").manual_seed(11)     # same Zipf-ish id stream on every rank
    mb = 30
    uid = (torch.rand(mb, 1024, device="cuda", generator=g) ** 3 * n_uid).to(torch.int32).clamp_(0, n_uid - 1)
    pid = (torch.rand(mb, 1024, device="cuda", generator=g) ** 3 * n_pid).to(torch.int32).clamp_(0, n_pid - 1)
    lab = (torch.rand(mb, 1024, device="cuda", generator=g) < 0.3).float()

    def step():
        for k in range(mb):
            t.step(uid[k], pid[k], lab[k], k % D)
    for _ in range(max(args.warmup, 1)):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # where a mini-batch goes (rank 0, one extra untimed step): CUDA-event marks between the phases of train_on_batch, and the
    # host time to ENQUEUE a step (no synchronisation inside): host-bound if it is not below the device time
    t.phase_events = []
    t0 = time.perf_counter()
    step()
    host_us = 1e6 * (time.perf_counter() - t0) / mb
    torch.cuda.synchronize()
    marks, t.phase_events = t.phase_events, None
    phases = {}
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        if n1 != "begin":
            phases[n1] = phases.get(n1, 0.0) + 1e3 * e0.elapsed_time(e1) / mb
    if rank == 0:
        msv = float(ms.item())
        alg = 24.0 * (n_uid + n_pid) * 128
        _emit(({"metric": "joint-train samples/sec (Amazon-13 shape, row-sharded trainable tables)", "value": mb * 1024 / (msv * 1e-3),
                          "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": msv,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "%s train steps, trainable tables row-sharded over %d rank(s), NCCL all-to-all, %d mini-batches of 1024 per step%s" % (tower, world, mb, ", tower replayed from a CUDA graph" if args.graphs else "")},
                          "roofline": {"bound": "hbm", "achieved": alg * mb / (msv * 1e-3) / 1e9, "unit": "GB/s",
                                       "note": "aggregate table-sweep bytes (24 B per table element per mini-batch) over all ranks / step time"},
                          "us_per_minibatch": 1e3 * msv / mb, "phase_us_per_minibatch": phases, "host_enqueue_us_per_minibatch": host_us}))
    dist.destroy_process_group()


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import run as runpy
    from mamdr_b200 import dist as mdist

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    rank, world = mdist.init_from_env("nccl")
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE is %d (launch N > 1 with torchrun)" % (args.gpus, world))
    config = load_config(args.workload)
    config["b200"]["device"] = "cuda:%d" % local_rank
    config["b200"]["precision"] = args.precision or ("fp32" if "star" in config["model"]["name"] else "tf32x3")   # STAR: fp32 path only
    if args.virtual_ranks:
        config["b200"]["virtual_ranks"] = args.virtual_ranks
    wrapper = runpy.build(config)
    base = wrapper.base_model
    model = base.model
    wrapper.prepare()
    lane_models = [lm for lm, _ in (getattr(wrapper, "_lane_models", None) or [(model, None)])]
    dev = model.device
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(e2e):
        """One meta-step.  e2e: the step's inputs (every domain's train columns) come from pinned host
        memory inside the timed region and the per-pass losses of the last pass are read back."""
        h2d = 0
        if e2e:
            for d in base.dataset.train_dataset.values():
                dd = d["data"]
                for k in ("uid", "pid", "label"):
                    getattr(dd, k).copy_(pinned[(dd.domain, k)], non_blocking=True)
                    h2d += pinned[(dd.domain, k)].numel() * 4
        before = getattr(base, "h2d_bytes", 0)
        wrapper.train_epoch(0)
        h2d += getattr(base, "h2d_bytes", 0) - before
        d2h = 0
        if e2e:
            host = base.last_pass_losses.cpu()   # D2H read of the step's result (per-batch losses of the last pass)
            d2h = host.numel() * 4
        return h2d, d2h

    pinned = {}
    for d in base.dataset.train_dataset.values():
        dd = d["data"]
        for k in ("uid", "pid", "label"):
            pinned[(dd.domain, k)] = torch.from_numpy(dd.host[k]).pin_memory()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        one_step(False)
    barrier()
    # ---- device-resident timing: K steps, per-step CUDA events, L2 flushed between steps
    base.samples_trained = 0
    launches0 = sum(lm.ctx.launches for lm in lane_models)
    evs = []
    barrier()
    sampler.mark_begin()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        one_step(False)
        b.record()
        evs.append((a, b))
    barrier()
    t_wall = time.perf_counter() - t_wall
    sampler.mark_end()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    launches = sum(lm.ctx.launches for lm in lane_models) - launches0
    my_samples = base.samples_trained
    # whole-job samples: DN passes are replicated (count once), DR passes are sharded (sum over ranks)
    dn_samples = args.steps * sum(d["n_data"] for d in base.dataset.train_dataset.values())
    t = torch.tensor([ms, float(my_samples - dn_samples)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, dr_samples = float(tmax[0]), float(t[1])
    else:
        dr_samples = float(t[1])
    total_samples = dn_samples + dr_samples
    value = total_samples / (ms * 1e-3)

    # ---- end-to-end through the public API with host buffers
    for _ in range(2):
        one_step(True)
    barrier()
    base.samples_trained = 0
    t0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(args.steps):
        x, y = one_step(True)
        h2d, d2h = x, y
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_samples / float(t[0])

    # ---- the dominant kernel, timed live: CUDA events around every launch of the per-pass kernel (pass_kernel in
    # the tcgen05 modes; the captured per-mini-batch graph in fp32 mode) over one extra, untimed-for-value meta-step
    P = sum(model.layout.numels)
    alg_bytes_mb = 1572864 + 12288 + 4096 + 2 * 4 * (P - model.n_domain * 128) + 28 * P   # 6.65 MB / mini-batch
    is_star = "star" in config["model"]["name"]
    if is_star:
        # STAR (config #4): gather + the effective-weight build / forward / backward reads of ONE domain's slices (~6 x 4 B x 141 K)
        # + the gradient-arena memset and the non-lazy Adam over the WHOLE arena (every domain's specific tensors): 32 B / parameter
        alg_bytes_mb = 1572864 + 12288 + 4096 + 6 * 4 * 141057 + 32 * P
    flops_mb = 0.72e9
    if model.pass_kernel:
        for lm in lane_models:
            lm.launch_times = []
        one_step(False)
        torch.cuda.synchronize()
        launches_ev = []
        for lm in lane_models:   # (virtual ranks: the lanes' launches overlap in time; their durations are summed)
            launches_ev += lm.launch_times
            lm.launch_times = None
    else:
        orig_fit = model.fit_pass
        launches_ev = []

        def timed_fit(data, steps=None, order=None):
            n = data.n_step if steps is None else int(steps)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = orig_fit(data, steps, order=order)
            b.record()
            launches_ev.append((a, b, n))
            return out
        model.fit_pass = timed_fit
        one_step(False)
        torch.cuda.synchronize()
        model.fit_pass = orig_fit
    k_ms = sum(a.elapsed_time(b) for a, b, _ in launches_ev)
    k_mb = sum(n for _, _, n in launches_ev)
    k_launches = len(launches_ev)
    k_ms_max = k_ms
    if world > 1:   # the limiting rank: most kernel time inside one meta-step
        t = torch.tensor([k_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_ms_max = float(t[0])

    # ---- the parity mode beside the benched one (N = 1): the same meta-step in fp32 mode (FFMA SIMT tower), the mode whose
    # free-running parameters meet the north-star 1e-4 bar (tests/test_gpu_trajectory.py)
    def side_run(over, steps):
        """The same meta-step under another b200 setting, device-timed (2 warm-up steps, `steps` timed)."""
        c2 = load_config(args.workload)
        c2["b200"]["device"] = "cuda:%d" % local_rank
        c2["b200"]["precision"] = config["b200"]["precision"]
        c2["b200"].update(over)
        w2 = runpy.build(c2)
        w2.prepare()
        for _ in range(2):
            w2.train_epoch(0)
        torch.cuda.synchronize()
        w2.base_model.samples_trained = 0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            flush.fill_(1)
            w2.train_epoch(0)
        b.record()
        torch.cuda.synchronize()
        v = w2.base_model.samples_trained / (a.elapsed_time(b) * 1e-3)
        del w2
        return v

    def guarded(over, steps):
        try:
            return side_run(over, steps)
        except Exception as e:     # a side measurement must never cost the headline line
            sys.stderr.write("side run %r failed: %r\n" % (over, e))
            return None

    parity_mode = fill_mode = None
    if world == 1 and model.pass_kernel and not args.no_micro and not args.virtual_ranks:
        parity_mode = {"precision": "fp32", "value": guarded({"precision": "fp32"}, 2), "unit": "samples/s",
                       "steps": 2, "note": "device-timed, same workload; fp32 FFMA tower (per-pass parity vs the CPU oracle 1e-6 in both modes, tests/test_gpu_trajectory.py)"}
        if "batch" not in config["model"]["name"] and not config["train"].get("finetune_every_epoch"):
            fill_mode = {"virtual_ranks": 2, "value": guarded({"virtual_ranks": 2}, 5), "unit": "samples/s", "steps": 5,
                         "note": "OPT-IN, NOT the headline: b200.virtual_ranks = 2 runs the DR chains of different query domains side by side on two "
                                 "74-SM partitions of this GPU (a row-local chain occupies 64 SMs).  Semantics = the 2-rank sharded schedule "
                                 "(per-lane Adam state during DR, the last owner's state adopted), bit-identical to two real ranks and judged against "
                                 "OracleMAMDR.train_epoch_sharded(2) (tests/test_gpu_virtual_ranks.py); `value` above is the reference's sequential schedule"}

    if rank != 0:
        return
    # ---- roofline + cpu baseline (rank 0)
    steps_per_epoch = total_samples / args.steps / config["dataset"]["batch_size"]
    hbm = peaks.get("hbm_gbs", 6650.0)
    avg_launch_ms = k_ms / max(k_launches, 1)
    alg_bytes_launch = alg_bytes_mb * k_mb / max(k_launches, 1)
    achieved = alg_bytes_launch / (avg_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        if model.pass_kernel:   # DRAM bytes cannot be measured inside a run: the ncu --set full capture of passk::pass_kernel
            for fn in ("r2_pass_kernel_ncu.json", "r1_pass_kernel_ncu.json"):
                fp = os.path.join(ROOT, "profiles", fn)
                if os.path.exists(fp):
                    traffic = json.load(open(fp)).get("dram_bytes_per_launch")
                    traffic_src = "profiles/" + fn + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch)"
                    break
    except Exception:
        pass
    kname = ("passk::pass_kernel (one persistent launch per DN phase / DR chain: passes + meta sweeps in-kernel)"
             if model.pass_kernel else "per-mini-batch SIMT graph")
    roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_source": "MEASURED_PEAKS.json (burst copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "kernel": kname, "launches_per_meta_step": k_launches, "avg_launch_us": 1e3 * avg_launch_ms,
            "minibatches_per_launch": k_mb / max(k_launches, 1), "alg_bytes_per_minibatch": alg_bytes_mb,
            "us_per_minibatch_in_kernel": 1e3 * k_ms / max(k_mb, 1), "kernel_share_of_step": k_ms / (ms / args.steps),
            "limiting_rank_kernel_share_of_step": k_ms_max / (ms / args.steps),
            "note": ("STAR: algorithmic bytes = gather + one domain's weight slices + 32 B per arena parameter (memset + non-lazy Adam over "
                     "every domain's specific tensors) per mini-batch; ~21 launches per mini-batch replayed as one CUDA graph per pass")
            if is_star else "algorithmic bytes = 6.65 MB per mini-batch (SURVEY.md 8(d)) x mini-batches of the launch; at batch "
                    "1024 the whole working set (2.3 MB) is L2-resident and the kernel is bound by the latency of the "
                    "row-local layer chain (every CTA streams all weights through shared memory once per mini-batch: "
                    "shared-memory port) plus 2 grid barriers, not by HBM: see DESIGN.md; HBM-scale rooflines of the "
                    "gather / Adam / scatter sweeps are under 'micro'",
            "tensor_tflops_achieved": flops_mb * k_mb / (k_ms * 1e-3) / 1e12}
    micro = {}
    if not args.no_micro and args.gpus == 1:
        try:
            micro = micro_rooflines(model, peaks, torch)
        except Exception as e:     # the micro-benchmarks must never cost the headline line
            sys.stderr.write("micro rooflines failed: %r\n" % (e,))
            micro = {"error": repr(e)}
    cpu = None
    if args.gpus == 1 and not args.no_cpu:
        run_once, cores, desc = oracle_sample_runner(config, 600)
        run_once()
        s, tsec = run_once()
        s2, tsec2 = run_once()
        cpu = {"value": (s + s2) / (tsec + tsec2), "unit": "samples/s", "cores": cores, "kind": "port", "sample": desc}
    metric = METRIC if args.workload == "Taobao-10" else METRIC.replace("Taobao-10", args.workload)
    line = {"metric": metric, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_desc(config, args.workload, args.gpus, model),
            "precision": config["b200"]["precision"], "precision_note": PRECISION_NOTE,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "micro": micro, "cpu_baseline": cpu,
            "parity_mode": parity_mode, "fill_the_machine": fill_mode,
            "minibatches_per_step": steps_per_epoch, "wall_s": t_wall}
    if args.virtual_ranks:
        line["virtual_ranks"] = args.virtual_ranks
        line["config"]["schedule"] = ("OPT-IN virtual ranks: the %d-rank sharded schedule on ONE GPU (DR chains side by side on SM partitions); "
                                      "not the reference's sequential schedule -- compare with a --gpus %d line, not with the default line"
                                      % (args.virtual_ranks, args.virtual_ranks))
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="Taobao-10", choices=sorted(WORKLOAD_CONFIG))
    ap.add_argument("--precision", default=None, choices=[None, "fp32", "tf32", "tf32x3"],
                    help="tower GEMM mode (default tf32x3: tcgen05 with fp32-equivalent products)")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--virtual-ranks", type=int, default=0, help="opt-in: V DR chains side by side on SM partitions of one GPU "
                    "(the V-rank sharded schedule; a separate bench line, never the headline)")
    ap.add_argument("--graphs", action="store_true", help="sharded workloads: replay the TOWER part of a step from a CUDA graph (no collective inside; "
                    "the collectives stay eager)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything the library prints while building (dataset banners ...) goes to stderr
    # (file-descriptor level: NCCL prints its version banner on fd 1 from C)
    sys.stdout.flush()
    real_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    global _emit
    _emit = lambda line: os.write(real_fd, (json.dumps(line) + "\n").encode())   # noqa: E731
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("Amazon-6", "Amazon-13-mmoe", "Amazon-13-ple"):
        run_amazon(args)
    elif args.workload.endswith("-sharded"):
        run_sharded(args)
    else:
        run_b200(args)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
