"""Row-sharded trainable embedding tables with NCCL all-to-all (BASELINE config #5's table layout, north_star: "embedding
tables that exceed a single GPU are row-sharded with NCCL all-to-all"), driving the joint `mlp` baseline of
``/root/reference/model_zoo/DeepCTR/deepctr.py:63-93`` data-parallel over the mini-batch.

The reference is single-process (``run.py:27-30``) and keeps whole tables in one TF variable; the sharded semantics are
the mathematically identical data-parallel restatement of one Keras train step on a batch of B rows:

  * table row r lives on rank ``r % G`` at local index ``r // G`` together with its Adam slots (m, v) and slot map;
  * every rank takes a contiguous slice of the batch (B_r rows) and runs the tower on it (fp32 per-mini-batch kernels);
  * forward : all-to-all(ids) -> owners gather their rows (``mamdr_gather_f32``) -> all-to-all(rows);
  * backward: ``mamdr_mlp_input_grads`` -> all-to-all(gradient rows, weighted B_r / B) -> owners de-duplicate
    (``mamdr_scatter_dedup_f32``) and run the fused l2 + non-lazy Adam sweep over their shard (``mamdr_adam_table_step``);
  * dense tower gradients: one all-reduce(sum) of the B_r / B weighted arenas, then ``mamdr_adam_step`` on every rank
    (replicas stay bit-identical: same reduced bits, same update).
Index bucketing is ONE kernel per step for both id columns (`mamdr_route_plan`), the gradient rows are packed into the exchange
buffer by `mamdr_route_pack_rows`, rows are picked out of the received blocks by `mamdr_gather_f32`: no tensor-op plumbing.  The
exchange buffers have a FIXED capacity (world x local-batch entries, -1 padded ids that the de-duplication skips), so every
all-to-all has static, equal splits and a step never synchronises with the host.
Dropout masks are indexed by the GLOBAL batch row (``mamdr_batch.row0`` = the slice start), so the sharded step draws exactly
the masks of the unsharded one; parity against the oracle is asserted with dropout 0 and 0.5 (tests/test_gpu_sharded.py).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import DomainData, MLPModel, _ptr


class _Plan(object):
    """Routing of the two id columns of a step with FIXED-capacity exchange buffers: every rank sends each owner a block of
    [cap user entries | cap item entries] (local row index, -1 = padding), so all split sizes are static -- no host
    synchronisation anywhere in a step -- and both tables travel in ONE all-to-all per direction.  The plan itself is one
    kernel (`mamdr_route_plan`: owner, stable position within the owner's block); the buffers are persistent."""

    def __init__(self, world, cap, dim, device):
        self.world, self.cap, self.dim, self.n = int(world), int(cap), int(dim), 0
        self.block = 2 * self.cap
        self.n_recv = self.world * self.block
        i32, f32 = dict(dtype=torch.int32, device=device), dict(dtype=torch.float32, device=device)
        self.slot_u, self.slot_i = torch.zeros(self.cap, **i32), torch.zeros(self.cap, **i32)   # slots of local row i in the exchange buffers
        self.send = torch.full((self.n_recv,), -1, **i32)
        self.recv_idx = torch.empty_like(self.send)                                # block r = the rows rank r asks of me
        self.ids_u, self.ids_i = torch.empty_like(self.send), torch.empty_like(self.send)   # recv_idx split per table (-1 elsewhere)
        self.got, self.back = torch.zeros(self.n_recv, self.dim, **f32), torch.zeros(self.n_recv, self.dim, **f32)
        self.gsend, self.grecv = torch.zeros(self.n_recv, self.dim, **f32), torch.zeros(self.n_recv, self.dim, **f32)
        self.out_u, self.out_i = torch.zeros(self.cap, self.dim, **f32), torch.zeros(self.cap, self.dim, **f32)

    def fetch(self, ctx, users, items, ids_u, ids_i, n, stream):
        """ids -> owners (one all-to-all), owners gather both tables (padding skipped), rows back (one all-to-all), rows picked
        out of the received blocks by slot: returns (user rows [n, dim], item rows [n, dim]) in the original order."""
        self.n = int(n)
        ctx.call("mamdr_route_plan", _ptr(ids_u), _ptr(ids_i), self.n, self.world, self.cap, self.block, _ptr(self.slot_u), _ptr(self.slot_i),
                 _ptr(self.send), C.c_void_p(self.send.data_ptr() + 4 * self.cap), stream)
        dist.all_to_all_single(self.recv_idx, self.send)
        ctx.call("mamdr_route_gather2", _ptr(users.table), _ptr(items.table), _ptr(self.recv_idx), self.world, self.cap, self.dim,
                 _ptr(self.got), _ptr(self.ids_u), _ptr(self.ids_i), stream)
        dist.all_to_all_single(self.back, self.got)
        ctx.launches += 2
        if self.n:
            for slot, out in ((self.slot_u, self.out_u), (self.slot_i, self.out_i)):
                ctx.call("mamdr_gather_f32", _ptr(self.back), self.n_recv, self.dim, _ptr(slot), self.n, _ptr(out), self.dim, stream)
            ctx.launches += 2
        return self.out_u, self.out_i

    def send_grads(self, ctx, dX, du, scale, stream):
        """The gradient rows of both tables (the two column blocks of dX [n, du + di], x scale) to their owners: one all-to-all.
        The padding rows of the exchange buffer are never read (their ids are -1)."""
        if self.n:
            stride = dX.shape[1]
            ctx.call("mamdr_route_pack_rows", _ptr(dX), stride, _ptr(self.slot_u), self.n, self.dim, float(scale), _ptr(self.gsend), stream)
            ctx.call("mamdr_route_pack_rows", C.c_void_p(dX.data_ptr() + 4 * du), stride, _ptr(self.slot_i), self.n, self.dim, float(scale),
                     _ptr(self.gsend), stream)
            ctx.launches += 2
        dist.all_to_all_single(self.grecv, self.gsend)


class ShardedTable(object):
    def __init__(self, ctx, full_init, rank, world, device, l2, cap):
        self.ctx, self.rank, self.world, self.l2 = ctx, rank, world, float(l2)
        local = np.ascontiguousarray(full_init[rank::world], dtype=np.float32)
        self.rows, self.dim = int(local.shape[0]), int(local.shape[1])
        self.table = torch.from_numpy(local).to(device)
        self.m, self.v = torch.zeros_like(self.table), torch.zeros_like(self.table)
        self.slot = torch.full((max(self.rows, 1),), -1, dtype=torch.int32, device=device)
        self.ws_bytes = ctx.lib.mamdr_adam_table_workspace_bytes()
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=device)
        max_recv = int(world) * 2 * int(cap)       # the shared exchange buffer of both tables (_Plan)
        if max_recv > ctx.lib.mamdr_scatter_max_n():
            raise ValueError("world x local batch = %d exceeds the de-duplication limit %d" % (max_recv, ctx.lib.mamdr_scatter_max_n()))
        self.sc_bytes = ctx.lib.mamdr_scatter_workspace_bytes(max_recv)
        self.sc_ws = torch.zeros(max(self.sc_bytes, 16), dtype=torch.uint8, device=device)
        self.uniq_ids = torch.zeros(max_recv, dtype=torch.int32, device=device)
        self.uniq_rows = torch.zeros(max_recv, self.dim, dtype=torch.float32, device=device)
        self.n_uniq = torch.zeros(4, dtype=torch.int32, device=device)
        self.max_recv = max_recv

    def apply(self, plan, ids, opt_state, lr, beta1, beta2, eps, loss_slot, stream):
        """De-duplicate the received gradient rows of this table (`ids`: the received id list with -1 for padding and for the
        other table's entries) and run the fused l2 + Adam sweep over the local shard."""
        self.ctx.call("mamdr_scatter_dedup_f32", _ptr(ids), _ptr(plan.grecv), self.dim, plan.n_recv, self.dim,
                      _ptr(self.uniq_ids), _ptr(self.uniq_rows), _ptr(self.n_uniq), _ptr(self.sc_ws), self.sc_ws.numel(), stream)
        self.ctx.launches += 2
        if self.rows:
            self.ctx.call("mamdr_adam_table_step", _ptr(self.table), _ptr(self.m), _ptr(self.v), self.rows, self.dim,
                          _ptr(self.uniq_ids), _ptr(self.uniq_rows), _ptr(self.n_uniq), plan.n_recv, _ptr(self.slot), self.l2,
                          _ptr(opt_state), lr, beta1, beta2, eps, _ptr(loss_slot), _ptr(self.ws), self.ws_bytes, stream)
            self.ctx.launches += 2

    def full(self):
        """All-gather of the shard into the full table (tests / checkpoints)."""
        world = self.world
        sizes = [torch.zeros(1, dtype=torch.int64, device=self.table.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([self.rows], dtype=torch.int64, device=self.table.device))
        parts = [torch.empty(int(s.item()), self.dim, dtype=torch.float32, device=self.table.device) for s in sizes]
        dist.all_gather(parts, self.table) if len(set(int(s.item()) for s in sizes)) == 1 else _uneven_gather(parts, self.table, sizes)
        n = sum(p.shape[0] for p in parts)
        out = torch.empty(n, self.dim, dtype=torch.float32, device=self.table.device)
        for r, p in enumerate(parts):
            out[r::world] = p
        return out


def _uneven_gather(parts, mine, sizes):
    mx = max(int(s.item()) for s in sizes)
    pad = torch.zeros(mx, mine.shape[1], dtype=mine.dtype, device=mine.device)
    pad[:mine.shape[0]] = mine
    bufs = [torch.empty_like(pad) for _ in parts]
    dist.all_gather(bufs, pad)
    for p, b in zip(parts, bufs):
        p.copy_(b[:p.shape[0]])


class _Steps(object):
    """`step` = one `train_on_batch`.  The collectives of a step are always issued eagerly (round 1's CUDA-graph replay of WHOLE
    steps, NCCL collectives inside the capture, dead-locked for some shapes and was removed).  `use_graphs=True` captures only
    the TOWER part of a step -- the ~15 (mlp) / ~40 (multi-task) kernel launches between the row fetch and the gradient
    exchange, no collective inside -- once per (sub-model, slice) and replays it: at 8 GPUs the eager step is bound by the host's
    launch rate, not by the device."""

    def _init_graphs(self, use_graphs=False):
        self.tower_graphs = bool(use_graphs)
        self._tower_cache = {}

    def _tower(self, key, fn):
        if not self.tower_graphs:
            fn()
            return
        ent = self._tower_cache.get(key)
        ctx = self.model.ctx
        if ent is None:
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            before = ctx.launches
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                fn()
            ent = self._tower_cache[key] = (g, ctx.launches - before)
            ctx.launches = before
        ent[0].replay()
        ctx.launches += ent[1]

    def step(self, uid, pid, label, domain):
        return self.train_on_batch(uid, pid, label, domain)

    phase_events = None   # bench / diagnostics: set to a list to get (phase name, CUDA event) marks of every step

    def _mark(self, name):
        if self.phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(self.device))
            self.phase_events.append((name, e))


class ShardedJointTrainer(_Steps):
    """Joint `mlp` training (``DeepCTR.train``) with row-sharded trainable tables; one instance per rank."""

    def __init__(self, n_uid, n_pid, n_domain, user_init, item_init, dense_init, emb_dim=(128, 128, 128), hidden=(256, 128, 64),
                 dropout=0.0, lr=1e-3, l2_emb=1e-5, batch_size=1024, device="cuda:0", use_graphs=False):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        self.batch_size = int(batch_size)
        self._init_graphs(use_graphs)
        bl = (self.batch_size + self.world - 1) // self.world
        self.max_local = bl
        # the tower runs in frozen-table mode on the rows received for this rank's slice of the batch
        z_u = np.zeros((bl, emb_dim[0]), dtype=np.float32)
        z_i = np.zeros((bl, emb_dim[1]), dtype=np.float32)
        self.model = MLPModel(bl, bl, n_domain, emb_dim=emb_dim, hidden=hidden, dropout=dropout, l2_emb=l2_emb, emb_trainable=False,
                              user_table=z_u, item_table=z_i, init_weights=dense_init, lr=lr, max_batch=bl, precision=_lib.PREC_FP32,
                              device=device, use_graphs=False)
        m = self.model
        m.desc.frozen_reg = 0.0
        ctx = m.ctx
        self.users = ShardedTable(ctx, user_init, self.rank, self.world, self.device, l2_emb, bl)
        self.items = ShardedTable(ctx, item_init, self.rank, self.world, self.device, l2_emb, bl)
        self.arange = torch.arange(bl, dtype=torch.int32, device=self.device)
        if emb_dim[0] != emb_dim[1]:
            raise NotImplementedError("the shared exchange buffer needs user_dim == item_dim (every shipped config: 128 / 128)")
        self.plan = _Plan(self.world, bl, emb_dim[0], self.device)
        self.dX = torch.zeros(bl, emb_dim[0] + emb_dim[1], dtype=torch.float32, device=self.device)
        self.loss_local = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.loss_tab = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.y_buf = torch.zeros(bl, dtype=torch.float32, device=self.device)
        self.comm_bytes = 0

    def _slice(self, n):
        """Contiguous split of a batch of n rows over the ranks (numpy.array_split sizes)."""
        base, rem = divmod(n, self.world)
        sizes = [base + (1 if r < rem else 0) for r in range(self.world)]
        start = sum(sizes[:self.rank])
        return start, sizes[self.rank]

    def _local_data(self, y, domain, rows):
        d = DomainData.__new__(DomainData)
        d.domain, d.n_data, d.batch_size, d.n_step = int(domain), int(rows), self.max_local, 1
        d.uid, d.pid, d.label, d.order, d.device = self.arange, self.arange, y, self.arange, self.device
        return d

    def train_on_batch(self, uid, pid, label, domain):
        """One Keras train step on the GLOBAL batch (device int32 / fp32 columns, identical on every rank)."""
        m = self.model
        st = m.stream
        n = int(uid.numel())
        start, bl = self._slice(n)
        u, p, y = uid[start:start + bl], pid[start:start + bl], label[start:start + bl]    # contiguous views of the columns
        plan = self.plan
        self._mark("begin")
        rows_u, rows_i = plan.fetch(m.ctx, self.users, self.items, u, p, bl, st)
        self._mark("fetch")
        w = float(bl) / float(n)                         # this rank's share of the batch mean
        self.loss_local.zero_()
        self.loss_tab.zero_()
        if bl:
            m.user_table, m.item_table = rows_u, rows_i
            self.y_buf[:bl].copy_(y)

            def tower():
                stt = m.stream
                self.loss_local.zero_()
                data = self._local_data(self.y_buf, domain, bl)
                b = m._batch(data, 0, bl, False)
                b.row0 = start            # dropout masks follow the GLOBAL batch row: the same draw as the unsharded step
                m.ctx.call("mamdr_mlp_train_step", C.byref(m.desc), C.byref(b), _ptr(rows_u), _ptr(rows_i), _ptr(m.params), _ptr(m.grads),
                           _ptr(m.ws), m.ws_bytes, _ptr(m.opt_state), _ptr(self.loss_local), None, _ptr(m.auc_acc), _ptr(m.thresholds),
                           m.num_thresholds, m.precision, stt)
                m.ctx.call("mamdr_mlp_input_grads", C.byref(m.desc), bl, _ptr(m.params), _ptr(m.ws), m.ws_bytes, _ptr(self.dX), stt)
                m.ctx.launches += 14
                m.grads.mul_(w)
                self.loss_local.mul_(w)
            self._tower((int(domain), bl, start, n), tower)
        else:
            m.grads.zero_()
        # tables first (they read the beta powers), then the dense arena (its apply advances them); the gradient rows are
        # the two column blocks of dX [bl, du + di], weighted by this rank's share while they are packed.  The all-reduce of
        # the dense gradients runs on NCCL's stream under the table sweeps
        self._mark("tower")
        plan.send_grads(m.ctx, self.dX, self.users.dim, w, st)
        dense_work = dist.all_reduce(m.grads, async_op=True)     # queued on NCCL's stream BEHIND the gradient-row exchange
        self._mark("send_grads")
        self.users.apply(plan, plan.ids_u, m.opt_state, m.lr, m.beta1, m.beta2, m.eps, self.loss_tab, st)
        self.items.apply(plan, plan.ids_i, m.opt_state, m.lr, m.beta1, m.beta2, m.eps, self.loss_tab, st)
        self._mark("tables")
        dense_work.wait()
        m.ctx.call("mamdr_adam_step", _ptr(m.params), _ptr(m.m), _ptr(m.v), _ptr(m.grads), m.params.numel(), _ptr(m.opt_state), m.lr,
                   m.beta1, m.beta2, m.eps, st)
        m.ctx.launches += 1
        self.comm_bytes += 4 * m.grads.numel() + 4 * plan.n_recv * (1 + 2 * self.users.dim)   # ids + rows out + gradient rows back, fixed-capacity blocks
        both = torch.cat([self.loss_local, self.loss_tab])
        dist.all_reduce(both)
        self._mark("dense")
        return both   # [mean BCE + l2 |E_d|^2, l2 (|E_u|^2 + |E_i|^2)]; their sum is the Keras loss

    def train_pass(self, host_split, domain, order, dev_cols=None):
        """One pass over a domain's training split in the given sample order (numpy), batch by batch."""
        uid = torch.from_numpy(np.ascontiguousarray(host_split['uid'][order], dtype=np.int32)).to(self.device)
        pid = torch.from_numpy(np.ascontiguousarray(host_split['pid'][order], dtype=np.int32)).to(self.device)
        lab = torch.from_numpy(np.ascontiguousarray(host_split['label'][order], dtype=np.float32)).to(self.device)
        losses = []
        for s in range(0, len(order), self.batch_size):
            e = min(len(order), s + self.batch_size)
            losses.append(self.step(uid[s:e], pid[s:e], lab[s:e], domain))
        return losses

    def dense_weights(self):
        return self.model.layout.unpack(self.model.params.cpu().numpy())


class ShardedMTLTrainer(_Steps):
    """BASELINE config #5 end to end: DomainNegotiation (``model_zoo/domain_negotiation.py:18-123``) over an MMOE / PLE /
    SharedBottom tower (``DeepMTLCTR/deep_mtl_ctr.py:21-66``) whose trainable user / item tables are row-sharded over the
    ranks.  One instance per rank.  A Keras train step of sub-model t on a global batch: all-to-all(ids) -> owners gather ->
    all-to-all(rows) -> ``mamdr_mtl_train_step`` on this rank's slice -> ``mamdr_mtl_input_grads`` -> all-to-all(gradient
    rows x B_r / B) -> owners de-duplicate + fused l2 / Adam sweep of their shard; the gradients of sub-model t's two arena
    spans take one all-reduce each and ``mamdr_adam_ranges_step`` runs on every rank (replicas stay bit-identical).
    theta of the DN outer update is kept like the model: dense arena replicated, one theta shard per table shard."""

    def __init__(self, kind, n_uid, n_pid, n_domain, user_init, item_init, dense_init, emb_dim=(128, 128, 128), expert_hidden=(256, 128),
                 tower_hidden=(64,), gate_hidden=(64,), num_experts=0, specific_expert_num=0, shared_expert_num=0, dropout=0.0,
                 lr=1e-4, l2_emb=1e-5, batch_size=1024, device="cuda:0", use_graphs=False):
        from .deep_mtl_ctr import MTLModel, MTLTopology
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        self.batch_size = int(batch_size)
        self._init_graphs(use_graphs)
        bl = (self.batch_size + self.world - 1) // self.world
        self.max_local = bl
        # the tower runs in frozen-table mode on the rows received for this rank's slice of the batch
        topo = MTLTopology(kind, bl, bl, n_domain, emb_dim, expert_hidden, tower_hidden, gate_hidden, num_experts=num_experts,
                           specific_expert_num=specific_expert_num, shared_expert_num=shared_expert_num, emb_trainable=False)
        self.model = MTLModel(topo, dense_init, user_table=np.zeros((bl, emb_dim[0]), dtype=np.float32),
                              item_table=np.zeros((bl, emb_dim[1]), dtype=np.float32), dropout=dropout, l2_emb=l2_emb, lr=lr,
                              max_batch=bl, device=device, use_graphs=False)
        m = self.model
        m.desc.frozen_reg = 0.0
        self.users = ShardedTable(m.ctx, user_init, self.rank, self.world, self.device, l2_emb, bl)
        self.items = ShardedTable(m.ctx, item_init, self.rank, self.world, self.device, l2_emb, bl)
        self.arange = torch.arange(bl, dtype=torch.int32, device=self.device)
        if emb_dim[0] != emb_dim[1]:
            raise NotImplementedError("the shared exchange buffer needs user_dim == item_dim (every shipped config: 128 / 128)")
        self.plan = _Plan(self.world, bl, emb_dim[0], self.device)
        self.dX = torch.zeros(bl, emb_dim[0] + emb_dim[1], dtype=torch.float32, device=self.device)
        self.loss_local = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.loss_tab = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.y_buf = torch.zeros(bl, dtype=torch.float32, device=self.device)
        self.theta = None
        self.comm_bytes = 0

    _slice = ShardedJointTrainer._slice
    _local_data = ShardedJointTrainer._local_data

    def train_on_batch(self, uid, pid, label, domain):
        """One Keras train step of sub-model `domain` on the GLOBAL batch (device columns, identical on every rank)."""
        m = self.model
        st = m.stream
        t = int(domain)
        n = int(uid.numel())
        start, bl = self._slice(n)
        u, p, y = uid[start:start + bl], pid[start:start + bl], label[start:start + bl]    # contiguous views of the columns
        plan = self.plan
        self._mark("begin")
        rows_u, rows_i = plan.fetch(m.ctx, self.users, self.items, u, p, bl, st)
        self._mark("fetch")
        w = float(bl) / float(n)
        self.loss_local.zero_()
        self.loss_tab.zero_()
        begin, length, n_spans = m.spans[t]
        spans = [m.grads[int(begin[q]):int(begin[q]) + int(length[q])] for q in range(n_spans)]
        if bl:
            self.y_buf[:bl].copy_(y)

            def tower():
                stt = m.stream
                self.loss_local.zero_()
                data = self._local_data(self.y_buf, t, bl)
                b = m._batch(data, 0, bl, False)
                b.row0 = start            # dropout masks follow the GLOBAL batch row: the same draw as the unsharded step
                m.ctx.call("mamdr_mtl_train_step", C.byref(m.desc), C.byref(m.domains[t]), C.byref(b), _ptr(rows_u), _ptr(rows_i),
                           _ptr(m.params), _ptr(m.grads), _ptr(m.ws), m.ws_bytes, _ptr(m.opt_state), _ptr(self.loss_local), None,
                           _ptr(m.auc_acc), _ptr(m.thresholds), m.num_thresholds, stt)
                m.ctx.call("mamdr_mtl_input_grads", C.byref(m.desc), C.byref(m.domains[t]), bl, _ptr(m.params), _ptr(m.ws), m.ws_bytes,
                           _ptr(self.dX), stt)
                m.ctx.launches += m.launches_per_train_step()
                for g in spans:
                    g.mul_(w)
                self.loss_local.mul_(w)
            self._tower((t, bl, start, n), tower)
        else:
            for g in spans:
                g.zero_()
        # tables first (they read the beta powers), then sub-model t's spans of the dense arena (that apply advances them).  The
        # all-reduces of the spans run on NCCL's stream under the table sweeps
        self._mark("tower")
        plan.send_grads(m.ctx, self.dX, self.users.dim, w, st)
        dense_work = [dist.all_reduce(g, async_op=True) for g in spans]     # queued on NCCL's stream BEHIND the gradient-row exchange
        self._mark("send_grads")
        self.users.apply(plan, plan.ids_u, m.opt_state, m.lr, m.beta1, m.beta2, m.eps, self.loss_tab, st)
        self.items.apply(plan, plan.ids_i, m.opt_state, m.lr, m.beta1, m.beta2, m.eps, self.loss_tab, st)
        self._mark("tables")
        for wk in dense_work:
            wk.wait()
        m.ctx.call("mamdr_adam_ranges_step", _ptr(m.params), _ptr(m.m), _ptr(m.v), _ptr(m.grads), begin, length, n_spans,
                   _ptr(m.opt_state), m.lr, m.beta1, m.beta2, m.eps, st)
        self.comm_bytes += 4 * sum(g.numel() for g in spans) + 4 * plan.n_recv * (1 + 2 * self.users.dim)
        both = torch.cat([self.loss_local, self.loss_tab])
        dist.all_reduce(both)
        self._mark("dense")
        return both

    train_pass = ShardedJointTrainer.train_pass

    # ---- DomainNegotiation over the sharded model ------------------------------------------------------------------
    def dn_prepare(self):
        """domain_negotiation.py:27-31 -- theta <- the model's weights; the optimizer slots start at zero."""
        m = self.model
        self.theta = {"dense": m.params.clone(), "user": self.users.table.clone(), "item": self.items.table.clone()}
        m.reset_optimizer()
        for tab in (self.users, self.items):
            tab.m.zero_()
            tab.v.zero_()

    def _triples(self):
        m = self.model
        return ((self.theta["dense"], m.params), (self.theta["user"], self.users.table), (self.theta["item"], self.items.table))

    def dn_meta_step(self, splits, sequence, orders, meta_lr, max_steps=0):
        """One DN meta-step (domain_negotiation.py:41-88): model <- theta; one pass of sub-model idx per domain of the
        (already shuffled) sequence, the single Adam threading through; theta += beta (model - theta); model <- theta."""
        m = self.model
        for theta, live in self._triples():
            if live.numel():
                m.copy_(live.view(-1), theta.view(-1))
        losses = []
        for idx, order in zip(sequence, orders):
            if max_steps and max_steps > 0:
                order = order[:max_steps * self.batch_size]
            m.reset_states()
            losses.append(self.train_pass(splits[idx], idx, order))
        for theta, live in self._triples():
            if live.numel():
                m.ctx.call("mamdr_dn_update", _ptr(theta), _ptr(live), float(meta_lr), live.numel(), _ptr(live), m.stream)
                m.ctx.launches += 1
        return losses

    def dense_weights(self):
        return self.model.layout.unpack(self.model.params.cpu().numpy())

    def theta_dense(self):
        return self.model.layout.unpack(self.theta["dense"].cpu().numpy())

    def theta_tables(self):
        """Full theta tables (all-gather of the shards) -- tests / checkpoints."""
        out = []
        for tab, key in ((self.users, "user"), (self.items, "item")):
            live = tab.table
            tab.table = self.theta[key]
            out.append(tab.full())
            tab.table = live
        return out
