"""Device-resident mlp model: the stand-in for the compiled ``tf.keras.Model`` that the reference
builds in ``/root/reference/model_zoo/DeepCTR/deepctr.py:20-61`` (BCE loss, ``AdamOptimizer``,
``AUC(num_thresholds=500)``) and drives with ``train_on_batch`` / ``fit`` / ``evaluate``.

All arithmetic runs in ``libmamdr_b200.so`` through the C-ABI; PyTorch only owns device memory and
streams.  Nothing here falls back to the CPU.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from . import _lib
from .auc import thresholds as auc_thresholds
from .layout import mlp_layout


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DomainData(object):
    """One domain's split as a device-resident column store (``uid, pid, label``), the replacement
    for the per-domain ``tf.data`` CSV pipeline of ``utils/dataset.py:20-38`` (which re-reads and
    re-parses the CSV on every pass).  ``n_step = ceil(n / batch_size)``, ragged tail kept."""

    def __init__(self, uid, pid, label, domain, batch_size, device):
        self.domain = int(domain)
        self.n_data = int(len(uid))
        self.batch_size = int(batch_size)
        self.n_step = int(np.ceil(self.n_data / float(batch_size))) if self.n_data else 0
        self.host = {"uid": np.ascontiguousarray(uid, dtype=np.int32),
                     "pid": np.ascontiguousarray(pid, dtype=np.int32),
                     "label": np.ascontiguousarray(label, dtype=np.float32)}
        self.device = device
        self.uid = self.pid = self.label = self.order = None
        if device is not None:
            self.upload()

    def upload(self, non_blocking=False):
        dev = self.device
        self.uid = torch.from_numpy(self.host["uid"]).to(dev, non_blocking=non_blocking)
        self.pid = torch.from_numpy(self.host["pid"]).to(dev, non_blocking=non_blocking)
        self.label = torch.from_numpy(self.host["label"]).to(dev, non_blocking=non_blocking)
        if self.order is None:
            self.order = torch.arange(max(self.n_data, 1), dtype=torch.int32, device=dev)

    def set_order(self, order):
        """Install the sample order of the next training pass (host numpy or device tensor)."""
        if isinstance(order, np.ndarray):
            order = torch.from_numpy(np.ascontiguousarray(order, dtype=np.int32))
        self.order.copy_(order, non_blocking=True)


class SplitView(object):
    """A sub-dataset of one ``DomainData`` that shares its device columns: the ``dataset.take(n)`` / ``dataset.skip(n)``
    meta-train / meta-val split of ``model_zoo/maml.py:296-318``.  The view covers the samples ``[lo, hi)`` of the split
    (file order); a pass runs over ``n_data`` of them in the order installed by ``set_order`` (a permutation of the
    window, of which the positions ``pick`` are kept -- ``shuffle(...).take(n)`` / ``.skip(n)`` of the non-exclusive mode)."""

    def __init__(self, data, lo, hi, pick=None):
        self.base, self.lo, self.hi = data, int(lo), int(hi)
        self.pick = pick if pick is not None else slice(0, self.hi - self.lo)
        self.domain, self.batch_size, self.device = data.domain, data.batch_size, data.device
        self.n_data = len(range(self.hi - self.lo)[self.pick])
        self.n_step = int(np.ceil(self.n_data / float(self.batch_size))) if self.n_data else 0
        self.uid, self.pid, self.label = data.uid, data.pid, data.label
        self.order = None
        if data.uid is not None:
            self.order = (torch.arange(self.lo, self.hi, dtype=torch.int32)[self.pick]).to(data.uid.device)
            if self.order.numel() == 0:
                self.order = torch.zeros(1, dtype=torch.int32, device=data.uid.device)

    def window_order(self, perm):
        """Sample ids of the next pass from a permutation of ``range(hi - lo)`` (host int32)."""
        return (np.asarray(perm, dtype=np.int32)[self.pick] + np.int32(self.lo)).astype(np.int32)

    def set_order(self, perm):
        if self.n_data and self.order is not None:   # (not uploaded: host-only use)
            self.order[:self.n_data].copy_(torch.from_numpy(np.ascontiguousarray(self.window_order(perm))), non_blocking=True)


class NamedWeight(object):
    """Minimal stand-in for a ``tf.Variable`` in ``model.trainable_weights`` (has ``.name``)."""

    def __init__(self, name, view, offset, numel):
        self.name, self.value, self.offset, self.numel = name, view, offset, numel
        self.shape = tuple(view.shape)

    def __repr__(self):
        return "<NamedWeight %s %s>" % (self.name, self.shape)


class MLPModel(object):
    # variable names as DeepCTR 0.9.0 creates them: used by the substring matching
    # of ``model_zoo/maml.py:160-177`` ("emb" in p.name ...)
    TF_NAMES = {"user_emb": "sparse_emb_user_emb/embeddings:0", "item_emb": "sparse_emb_item_emb/embeddings:0",
                "domain_emb": "sparse_emb_domain_emb/embeddings:0", "dense_kernel": "dense/kernel:0",
                "global_bias": "prediction_layer/global_bias:0"}

    def __init__(self, n_uid, n_pid, n_domain, emb_dim=(128, 128, 128), hidden=(256, 128, 64), dropout=0.5,
                 dropout_seed=1024, l2_emb=1e-5, emb_trainable=False, user_table=None, item_table=None,
                 init_weights=None, lr=1e-3, max_batch=1024, precision=_lib.PREC_FP32, device="cuda:0",
                 use_graphs=True):
        self._ctor = dict(n_uid=n_uid, n_pid=n_pid, n_domain=n_domain, emb_dim=emb_dim, hidden=hidden, dropout=dropout,
                          dropout_seed=dropout_seed, l2_emb=l2_emb, emb_trainable=emb_trainable, user_table=user_table,
                          item_table=item_table, lr=lr, max_batch=max_batch, precision=precision, device=device,
                          use_graphs=use_graphs)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mamdr_b200 runs on CUDA devices only (no CPU fallback)")
        torch.cuda.set_device(self.device)
        self.ctx = _lib.Context(self.device.index or 0)
        self.layout = mlp_layout(n_uid, n_pid, n_domain, emb_dim, hidden, emb_trainable)
        self.n_uid, self.n_pid, self.n_domain = int(n_uid), int(n_pid), int(n_domain)
        self.emb_dim, self.hidden = tuple(emb_dim), tuple(hidden)
        self.emb_trainable = bool(emb_trainable)
        self.lr, self.beta1, self.beta2, self.eps = float(lr), 0.9, 0.999, 1e-8
        self.max_batch = int(max_batch)
        self.precision = int(precision)
        self.use_graphs = bool(use_graphs)
        self.optimizer = "adam"
        self.sgd_lr = 0.0
        dev = self.device
        P = self.layout.total
        f32 = dict(dtype=torch.float32, device=dev)
        self.params = torch.zeros(P, **f32)
        self.grads = torch.zeros(P, **f32)
        self.m = torch.zeros(P, **f32)
        self.v = torch.zeros(P, **f32)
        if init_weights is not None:
            self.params.copy_(torch.from_numpy(self.layout.pack(init_weights)))
        self.frozen_reg = 0.0
        if not emb_trainable:
            if user_table is None or item_table is None:
                raise ValueError("frozen embeddings need user_table and item_table")
            ut = np.ascontiguousarray(user_table, dtype=np.float32)
            it = np.ascontiguousarray(item_table, dtype=np.float32)
            assert ut.shape == (n_uid, emb_dim[0]) and it.shape == (n_pid, emb_dim[1])
            self.frozen_reg = float(l2_emb * (np.sum(ut.astype(np.float64) ** 2) + np.sum(it.astype(np.float64) ** 2)))
            self.user_table = torch.from_numpy(ut).to(dev)
            self.item_table = torch.from_numpy(it).to(dev)
        else:
            # trainable tables live inside the arena (user_emb, item_emb first, like model.trainable_weights)
            self.user_table = self.item_table = None   # (tcgen05 modes: one pass-kernel launch per mini-batch, _train_step_tc_tables)
        # ---- C-ABI descriptor
        d = _lib.MlpDesc()
        d.n_layers = len(hidden)
        for i in range(3):
            d.emb_dim[i] = int(emb_dim[i])
        for i, h in enumerate(hidden):
            d.hidden[i] = int(h)
        d.n_domain, d.emb_trainable = int(n_domain), int(emb_trainable)
        d.n_uid, d.n_pid = int(n_uid), int(n_pid)
        d.dropout_rate, d.dropout_seed, d.l2_emb = float(dropout), int(dropout_seed), float(l2_emb)
        d.frozen_reg = self.frozen_reg
        lo = self.layout
        d.off_user_emb, d.off_item_emb = lo.offset("user_emb"), lo.offset("item_emb")
        d.off_domain_emb = lo.offset("domain_emb")
        for i in range(len(hidden)):
            d.off_kernel[i], d.off_bias[i] = lo.offset("kernel%d" % i), lo.offset("bias%d" % i)
        d.off_dense_kernel, d.off_global_bias = lo.offset("dense_kernel"), lo.offset("global_bias")
        d.arena_floats = P
        self.desc = d
        lib = self.ctx.lib
        ws_bytes = lib.mamdr_mlp_workspace_bytes(C.byref(d), self.max_batch)
        if ws_bytes == 0:
            raise _lib.MamdrError(-1, "mamdr_mlp_workspace_bytes rejected the descriptor")
        self.ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
        self.ws_bytes = ws_bytes
        self.opt_state = torch.zeros(lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device=dev)
        self.num_thresholds = 500
        self.thresholds = torch.from_numpy(auc_thresholds(self.num_thresholds)).to(dev)
        self.auc_acc = torch.zeros(4, self.num_thresholds, **f32)
        self._auc_out = torch.zeros(1, **f32)
        self._auc_zero = torch.zeros(4, self.num_thresholds, **f32)
        self._recording = False
        self.reset_optimizer()
        self._graphs = {}
        self._loss_bufs = {}
        if self.emb_trainable:
            lo = self.layout
            self._tables = []   # (arena offset, rows, dim, slot map)
            for name, n_rows, dim in (("user_emb", self.n_uid, emb_dim[0]), ("item_emb", self.n_pid, emb_dim[1])):
                self._tables.append((lo.offset(name), int(n_rows), int(dim),
                                     torch.full((int(n_rows),), -1, dtype=torch.int32, device=dev)))
            self.table_ws_bytes = lib.mamdr_adam_table_workspace_bytes()
            self.table_ws = torch.zeros(self.table_ws_bytes, dtype=torch.uint8, device=dev)
            self.l2_emb = float(l2_emb)
            self._sq = torch.zeros(1, dtype=torch.float64, device=dev)
            self.dense_off = lo.offset("domain_emb")
            self._opt_prev = torch.zeros_like(self.opt_state)   # the optimizer state before the current step (table sweeps)
        # the tcgen05 modes are served by the persistent pass kernel only (one cooperative launch per domain
        # pass); fp32 is the per-mini-batch SIMT path.  No silent fallback between them.
        self.pass_kernel = False
        if self.precision != _lib.PREC_FP32:
            rc = lib.mamdr_mlp_pass_supported(self.ctx.handle, C.byref(d), self.max_batch)
            if rc != 0:
                raise _lib.MamdrError(rc, (lib.mamdr_last_error(self.ctx.handle) or b"").decode() +
                                      " -- use b200.precision = 'fp32' for this model shape")
            self.pass_ws_bytes = lib.mamdr_mlp_pass_workspace_bytes(C.byref(d), self.max_batch)
            self.pass_ws = torch.zeros(self.pass_ws_bytes, dtype=torch.uint8, device=dev)
            self.pass_kernel = True
            self._prog_buf = torch.zeros(4096 * int(lib.mamdr_program_op_bytes()), dtype=torch.uint8, device=dev)
        self._recording = False
        self.program_ops = 0
        self.launch_times = None

    # ---- Keras-like surface -------------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def trainable_weights(self):
        out = []
        for name, view, off, n in zip(self.layout.names, self.layout.views(self.params), self.layout.offsets,
                                      self.layout.numels):
            tf_name = self.TF_NAMES.get(name)
            if tf_name is None:
                kind, idx = ("kernel", name[6:]) if name.startswith("kernel") else ("bias", name[4:])
                tf_name = "dnn/%s%s:0" % (kind, idx)
            out.append(NamedWeight(tf_name, view, off, n))
        return out

    @property
    def stateful_metric_functions(self):
        return [self]  # the AUC metric lives on the model; exposes reset_states()

    def reset_states(self):
        """``AUC.reset_states`` (utils/auc.py:283-284)."""
        if self._recording:   # keep the reset in program order
            self.ctx.call("mamdr_copy", _ptr(self.auc_acc), _ptr(self._auc_zero), self.auc_acc.numel(), self.stream)
        else:
            self.auc_acc.zero_()

    def reset_optimizer(self):
        """``tf.global_variables_initializer()`` on the optimizer slots
        (model_zoo/mamdr.py:35, model_zoo/domain_negotiation.py:31): m = v = 0, beta powers reset."""
        self.m.zero_()
        self.v.zero_()
        self.ctx.call("mamdr_opt_state_init", _ptr(self.opt_state), self.beta1, self.beta2, self.stream)
        self.ctx.launches += 1

    def compile(self, optimizer="adam", lr=None):
        """``model.compile`` with ``AdamOptimizer`` (DeepCTR/deepctr.py:54-60) or the finetune stage's
        ``GradientDescentOptimizer`` (specific_base_model.py:120, base_model.py:69)."""
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' or 'sgd'")
        self.optimizer = optimizer
        if optimizer == "sgd":
            self.sgd_lr = float(lr)
        elif lr is not None:
            self.lr = float(lr)

    def get_weights(self):
        return self.params.clone()

    def set_weights(self, flat):
        self.copy_(self.params, flat)

    def copy_(self, dst, src):
        self.ctx.call("mamdr_copy", _ptr(dst), _ptr(src), dst.numel(), self.stream)
        self.ctx.launches += 1

    def opt_words(self):
        """(step, b1pow, b2pow) of the device optimizer state as a float32 device tensor [3] (no sync)."""
        st = self.opt_state
        return torch.stack([st[:8].view(torch.int64)[0].to(torch.float32), st[8:12].view(torch.float32)[0],
                            st[12:16].view(torch.float32)[0]])

    def set_opt_words(self, words):
        st = self.opt_state
        st[:8].view(torch.int64)[0] = words[0].to(torch.int64)
        st[8:12].view(torch.float32)[0] = words[1]
        st[12:16].view(torch.float32)[0] = words[2]

    def read_step(self):
        step, b1, b2 = C.c_int64(), C.c_float(), C.c_float()
        self.ctx.call("mamdr_opt_state_read", _ptr(self.opt_state), C.byref(step), C.byref(b1), C.byref(b2),
                      self.stream)
        return step.value, b1.value, b2.value

    # ---- one mini-batch ---------------------------------------------------------------------------------
    def _batch(self, data, offset, rows, use_order):
        b = _lib.Batch()
        b.uid_dev, b.pid_dev, b.label_dev = data.uid.data_ptr(), data.pid.data_ptr(), data.label.data_ptr()
        b.order_dev = data.order.data_ptr() if use_order else None
        b.offset, b.rows, b.domain = int(offset), int(rows), int(data.domain)
        return b

    @contextlib.contextmanager
    def program(self, enabled=True):
        """Deferred execution: every training pass and meta sweep issued inside the block is recorded and the whole
        sequence runs as ONE persistent cooperative launch at the end of the block (``mamdr_program_begin / _end``).
        Only the tcgen05 modes have the in-kernel executor; otherwise (or with ``enabled=False``) this is a no-op and
        calls execute immediately."""
        if not (enabled and self.pass_kernel) or self._recording or self.emb_trainable:   # (trainable tables: the table sweeps run
            yield False                                                                    #  between mini-batches -- not recordable)
            return
        self.ctx.call("mamdr_program_begin")
        self._recording = True
        self._prog_minibatches = 0
        try:
            yield True
        except BaseException:
            self._recording = False
            self.ctx.lib.mamdr_program_abort(self.ctx.handle)
            raise
        self._recording = False
        n_ops = C.c_int32(0)
        ev = self._launch_event()
        self.ctx.call("mamdr_program_end", _ptr(self._prog_buf), self._prog_buf.numel(), C.byref(n_ops), self.stream)
        self._launch_event(ev, self._prog_minibatches)
        self.ctx.launches += 3   # H2D of the op list, barrier memset, the persistent kernel
        self.program_ops = n_ops.value

    def _pass(self, data, steps, use_order, offset=0, rows=None, order=None):
        """Descriptor of a pass over ``data``; with ``rows`` given: the single mini-batch [offset, offset+rows).
        ``order``: an int32 device tensor holding the pass's sample order (default: ``data.order``)."""
        ps = _lib.Pass()
        ps.uid_dev, ps.pid_dev, ps.label_dev = data.uid.data_ptr(), data.pid.data_ptr(), data.label.data_ptr()
        ps.n_data, ps.batch_size, ps.steps, ps.domain = data.n_data, data.batch_size, int(steps), int(data.domain)
        ps.order_dev = (order if order is not None else data.order).data_ptr() if use_order else None
        if rows is not None:
            ps.n_data, ps.batch_size, ps.steps = int(rows), int(rows), 1
            if use_order:
                ps.order_dev = data.order.data_ptr() + 4 * int(offset)
            else:
                ps.uid_dev, ps.pid_dev = data.uid.data_ptr() + 4 * int(offset), data.pid.data_ptr() + 4 * int(offset)
                ps.label_dev = data.label.data_ptr() + 4 * int(offset)
        return ps

    def _launch_event(self, start=None, minibatches=0):
        """bench hook: with ``self.launch_times`` set to a list, every launch of the persistent kernel is bracketed
        by CUDA events on its stream and (start, end, mini-batches) is appended."""
        if self.launch_times is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(self.device))
        if start is not None:
            self.launch_times.append((start, e, minibatches))
        return e

    def _train_pass(self, ps, losses, with_auc=True):
        adam = self.optimizer == "adam"
        ev = None if self._recording else self._launch_event()
        self.ctx.call("mamdr_mlp_train_pass", C.byref(self.desc), C.byref(ps), _ptr(self.user_table),
                      _ptr(self.item_table), _ptr(self.params), _ptr(self.m), _ptr(self.v), _ptr(self.grads),
                      _ptr(self.pass_ws), self.pass_ws_bytes, _ptr(self.opt_state), _ptr(losses),
                      _ptr(self.auc_acc if with_auc else None), _ptr(self.thresholds), self.num_thresholds,
                      0 if adam else 1, self.lr if adam else self.sgd_lr, self.beta1, self.beta2, self.eps,
                      self.precision, self.stream)
        if self._recording:
            self._prog_minibatches += ps.steps
        else:
            self.ctx.launches += 2   # 64-byte barrier memset + the persistent kernel
            self._launch_event(ev, ps.steps)

    def _eval_pass(self, ps, losses, probs=None, with_auc=True):
        self.ctx.call("mamdr_mlp_eval_pass", C.byref(self.desc), C.byref(ps), _ptr(self.user_table),
                      _ptr(self.item_table), _ptr(self.params), _ptr(self.pass_ws), self.pass_ws_bytes,
                      _ptr(self.opt_state), _ptr(losses), _ptr(probs), _ptr(self.auc_acc if with_auc else None),
                      _ptr(self.thresholds), self.num_thresholds, self.precision, self.stream)
        self.ctx.launches += 2

    def _train_step(self, data, offset, rows, loss_slot, probs=None, with_auc=True):
        """One mini-batch (forward + backward + optimizer apply); gradients are left in ``self.grads``."""
        if self.pass_kernel and self.emb_trainable:
            self._train_step_tc_tables(data, offset, rows, loss_slot, with_auc)
            return
        if self.pass_kernel:
            self._train_pass(self._pass(data, 1, True, offset, rows), loss_slot, with_auc)
            return
        b = self._batch(data, offset, rows, True)
        st = self.stream
        if self.emb_trainable:
            self.desc.frozen_reg = 0.0   # training adds the tables' l2 penalty inside the fused table sweep
        self.ctx.call("mamdr_mlp_train_step", C.byref(self.desc), C.byref(b), _ptr(self.user_table),
                      _ptr(self.item_table), _ptr(self.params), _ptr(self.grads), _ptr(self.ws), self.ws_bytes,
                      _ptr(self.opt_state), _ptr(loss_slot), _ptr(probs), _ptr(self.auc_acc if with_auc else None),
                      _ptr(self.thresholds), self.num_thresholds, self.precision, st)
        if self.emb_trainable:
            # K6 + K7 fused per table: de-duplicated sparse rows + dense l2 term + Adam over every row, BEFORE the
            # dense apply advances the beta powers; then the dense part of the arena
            for t, (off, n_rows, dim, slot) in enumerate(self._tables):
                ids, srows, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
                rc = self.ctx.lib.mamdr_mlp_sparse_grads(C.byref(self.desc), int(rows), _ptr(self.ws), t, C.byref(ids),
                                                         C.byref(srows), C.byref(cnt))
                if rc != 0:
                    raise _lib.MamdrError(rc, "mamdr_mlp_sparse_grads")
                self._table_sweep(off, n_rows, dim, slot, ids, srows, cnt, int(rows), self.opt_state, loss_slot, st)
            do = self.dense_off
            if self.optimizer == "adam":
                self.ctx.call("mamdr_adam_step", _ptr(self.params[do:]), _ptr(self.m[do:]), _ptr(self.v[do:]),
                              _ptr(self.grads[do:]), self.params.numel() - do, _ptr(self.opt_state), self.lr, self.beta1,
                              self.beta2, self.eps, st)
            else:
                self.ctx.call("mamdr_sgd_step", _ptr(self.params[do:]), _ptr(self.grads[do:]), self.params.numel() - do,
                              _ptr(self.opt_state), self.sgd_lr, st)
            self.ctx.launches += 1 + 2 + 2 * 2   # dX GEMM, sort + segment-sum (both tables per launch), 2 x (slot scatter, table sweep)
        elif self.optimizer == "adam":
            self.ctx.call("mamdr_adam_step", _ptr(self.params), _ptr(self.m), _ptr(self.v), _ptr(self.grads),
                          self.params.numel(), _ptr(self.opt_state), self.lr, self.beta1, self.beta2, self.eps, st)
        else:
            self.ctx.call("mamdr_sgd_step", _ptr(self.params), _ptr(self.grads), self.params.numel(),
                          _ptr(self.opt_state), self.sgd_lr, st)
        self.ctx.launches += 3 + 3 * len(self.hidden) + 3  # memset, assemble, L fwd, head, L-1 dH, L dW, colsum, dEd, opt

    def _train_step_tc_tables(self, data, offset, rows, loss_slot, with_auc=True):
        """Config #2 in the tcgen05 modes: one mini-batch = the pass kernel (tables gathered from the arena, tower + dX on the
        tensor cores, dense variables applied in-kernel, the two sparse gradients de-duplicated behind it) + the fused
        sparse-merge / l2 / non-lazy Adam sweep of each table, which reads the beta powers of BEFORE the step."""
        st = self.stream
        self.desc.frozen_reg = 0.0   # training adds the tables' l2 penalty inside the fused table sweep
        self._opt_prev.copy_(self.opt_state)
        self._train_pass(self._pass(data, 1, True, offset, rows), loss_slot, with_auc)
        for t, (off, n_rows, dim, slot) in enumerate(self._tables):
            ids, srows, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
            rc = self.ctx.lib.mamdr_mlp_pass_sparse_grads(C.byref(self.desc), int(rows), _ptr(self.pass_ws), t, C.byref(ids),
                                                          C.byref(srows), C.byref(cnt))
            if rc != 0:
                raise _lib.MamdrError(rc, "mamdr_mlp_pass_sparse_grads")
            self._table_sweep(off, n_rows, dim, slot, ids, srows, cnt, int(rows), self._opt_prev, loss_slot, st)
        self.ctx.launches += 2 + 2 * 2   # sort + segment-sum (both tables per launch), 2 x (slot scatter, table sweep)

    def _table_sweep(self, off, n_rows, dim, slot, ids, srows, cnt, max_uniq, opt_state, loss_slot, st):
        """K6 + K7 fused for one trainable table: the de-duplicated sparse rows + the dense l2 term, then the non-lazy Adam over
        every row -- or, in the finetune stage (``compile('sgd')``), plain SGD over every row."""
        n_el = n_rows * dim
        if self.optimizer == "adam":
            self.ctx.call("mamdr_adam_table_step", _ptr(self.params[off:off + n_el]), _ptr(self.m[off:off + n_el]),
                          _ptr(self.v[off:off + n_el]), n_rows, dim, ids, srows, cnt, max_uniq, _ptr(slot),
                          self.l2_emb, _ptr(opt_state), self.lr, self.beta1, self.beta2, self.eps,
                          _ptr(loss_slot), _ptr(self.table_ws), self.table_ws_bytes, st)
        else:
            self.ctx.call("mamdr_sgd_table_step", _ptr(self.params[off:off + n_el]), n_rows, dim, ids, srows, cnt, max_uniq,
                          _ptr(slot), self.l2_emb, self.sgd_lr, _ptr(loss_slot), _ptr(self.table_ws), self.table_ws_bytes, st)

    # ---- virtual ranks: several models side by side on ONE GPU, each pass kernel on its own SM partition -------------
    def set_pass_ctas(self, n):
        """CTAs of this model's persistent pass kernel (0 = one per SM): ``mamdr_ctx_set_pass_ctas``."""
        self.ctx.call("mamdr_ctx_set_pass_ctas", int(n))

    def clone_lane(self):
        """A second model of the same architecture on the same device with its own context, arena, optimizer slots and
        workspaces (a "virtual rank").  State is brought over with ``copy_state_from``."""
        lane = type(self)(init_weights=None, **self._ctor)
        lane.optimizer, lane.sgd_lr, lane.lr = self.optimizer, self.sgd_lr, self.lr
        return lane

    def copy_state_from(self, other):
        """params, Adam slots and beta powers / global step of ``other`` -> this model (device copies on the current stream)."""
        for dst, src in ((self.params, other.params), (self.m, other.m), (self.v, other.v)):
            self.copy_(dst, src)
        self.opt_state.copy_(other.opt_state, non_blocking=True)

    # ---- gradient-only step + a second optimizer: what MAML / MLDG / PCGrad add to the compiled model ----------------
    def grads_on_batch(self, data, offset, rows, loss_slot, with_auc=True):
        """The ``K.function(model._feed_inputs + model._feed_targets, [total_loss] + metrics_tensors, updates=...)`` that
        ``_make_meta_train_function`` builds (model_zoo/maml.py:196-233, mldg.py, pcgrad.py): forward + backward of
        ``model.total_loss`` at the live weights WITHOUT an optimizer apply; the gradients of every variable are left in
        ``self.grads``.  [EXT] the function is built without ``K.learning_phase()`` among its inputs, so the placeholder
        takes its default 0: the forward is the INFERENCE one (no dropout); the stateful AUC is updated like in any step.
        Runs the fp32 tower (the per-mini-batch path that leaves the gradients in memory) whatever ``self.precision`` is."""
        if self.emb_trainable:
            raise NotImplementedError("gradient accumulation over trainable tables (sparse IndexedSlices) is not built: the "
                                      "shipped MAML / MLDG / PCGrad configs train on frozen pretrained tables")
        d = getattr(self, "_desc_nodrop", None)
        if d is None:
            d = self._desc_nodrop = type(self.desc).from_buffer_copy(self.desc)
            d.dropout_rate = 0.0
        b = self._batch(data, offset, rows, True)
        self.ctx.call("mamdr_mlp_train_step", C.byref(d), C.byref(b), _ptr(self.user_table), _ptr(self.item_table),
                      _ptr(self.params), _ptr(self.grads), _ptr(self.ws), self.ws_bytes, _ptr(self.opt_state),
                      _ptr(loss_slot), None, _ptr(self.auc_acc if with_auc else None), _ptr(self.thresholds),
                      self.num_thresholds, _lib.PREC_FP32, self.stream)
        self.ctx.launches += 2 + 3 * len(self.hidden) + 3

    def new_optimizer_slots(self):
        """Slots of a SECOND ``tf.train.AdamOptimizer`` (``self.meta_optimizer``, maml.py:201): (m, v, beta-power state)."""
        m2, v2 = torch.zeros_like(self.params), torch.zeros_like(self.params)
        st = torch.zeros_like(self.opt_state)
        self.ctx.call("mamdr_opt_state_init", _ptr(st), self.beta1, self.beta2, self.stream)
        self.ctx.launches += 1
        return m2, v2, st

    def train_on_batch(self, data, offset, rows):
        """``Model.train_on_batch`` -> (loss, auc) host floats.  Synchronises: debugging / parity only;
        the wrappers use ``fit_pass``."""
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._train_step(data, offset, rows, loss)
        return float(loss.item()), self.auc_result()

    def _pass_plan(self, data, steps):
        bs = data.batch_size
        return [(s * bs, min(bs, data.n_data - s * bs)) for s in range(steps)]

    def fit_pass(self, data, steps=None, order=None):
        """One pass of ``steps`` mini-batches over ``data`` in the order installed by
        ``data.set_order`` (== the ``for step in range(train_step): model.train_on_batch(iter)`` loops
        of mamdr.py:85-97 and domain_negotiation.py:71-72, and ``model.fit(iter, steps_per_epoch)`` of
        mamdr.py:54).  Asynchronous; returns the device tensor of per-batch losses.  The whole pass is
        captured once into a CUDA graph per (domain split, steps, optimizer) and replayed."""
        steps = data.n_step if steps is None else int(steps)
        if steps <= 0:
            return torch.zeros(0, dtype=torch.float32, device=self.device)
        if data.batch_size > self.max_batch:
            raise ValueError("batch_size %d exceeds max_batch %d" % (data.batch_size, self.max_batch))
        key = (id(data), steps, self.optimizer, self.sgd_lr, self.lr)
        losses = self._loss_bufs.get(key)
        if losses is None:
            losses = self._loss_bufs[key] = torch.zeros(steps, dtype=torch.float32, device=self.device)
        if self.pass_kernel and self.emb_trainable:
            # the tables change between mini-batches: one launch (+ the two table sweeps) per mini-batch
            if order is not None:
                data.order.copy_(order, non_blocking=True)
            for s, (off, rows) in enumerate(self._pass_plan(data, steps)):
                self._train_step_tc_tables(data, off, rows, losses[s:s + 1])
            return losses
        if self.pass_kernel:
            self._train_pass(self._pass(data, steps, True, order=order), losses)
            return losses
        if order is not None:
            data.order.copy_(order, non_blocking=True)   # the captured graphs read data.order
        plan = self._pass_plan(data, steps)
        if not self.use_graphs:
            for s, (off, rows) in enumerate(plan):
                self._train_step(data, off, rows, losses[s:s + 1])
            return losses
        g = self._graphs.get(key)
        if g is None:
            # capture happens on a side stream; the launches recorded there are not executed
            before = self.ctx.launches
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(self.device)
            with torch.cuda.graph(g):
                for s, (off, rows) in enumerate(plan):
                    self._train_step(data, off, rows, losses[s:s + 1])
            self._graphs[key] = (g, self.ctx.launches - before)
            self.ctx.launches = before
            g = self._graphs[key]
        g[0].replay()
        self.ctx.launches += g[1]
        return losses

    # ---- evaluate ---------------------------------------------------------------------------------------
    def evaluate(self, data, steps=None):
        """``Model.evaluate(dataset, steps)`` -> (mean of per-batch losses, AUC) as host floats: resets
        the stateful AUC, runs the inference forward (no dropout) batch by batch in file order."""
        steps = data.n_step if steps is None else int(steps)
        self.reset_states()
        losses = torch.zeros(max(steps, 1), dtype=torch.float32, device=self.device)
        st = self.stream
        if data.batch_size > self.max_batch:
            raise ValueError("batch_size %d exceeds max_batch %d" % (data.batch_size, self.max_batch))
        if self.emb_trainable:
            self._refresh_table_reg()
        if self.pass_kernel and steps > 0:
            self._eval_pass(self._pass(data, steps, False), losses)
            auc = self.auc_result()
            return float(losses[:steps].double().mean().item()), auc
        for s, (off, rows) in enumerate(self._pass_plan(data, steps)):
            b = self._batch(data, off, rows, False)
            self.ctx.call("mamdr_mlp_eval_step", C.byref(self.desc), C.byref(b), _ptr(self.user_table),
                          _ptr(self.item_table), _ptr(self.params), _ptr(self.ws), self.ws_bytes,
                          _ptr(losses[s:s + 1]), None, _ptr(self.auc_acc), _ptr(self.thresholds),
                          self.num_thresholds, self.precision, st)
            self.ctx.launches += 2 + len(self.hidden)
        auc = self.auc_result()
        return float(losses[:steps].double().mean().item()) if steps else 0.0, auc

    def _refresh_table_reg(self):
        """Inference loss with trainable tables: the l2 penalty of the two tables enters through ``frozen_reg``
        (training adds it inside the fused table sweep)."""
        tot = 0.0
        for off, n_rows, dim, _ in self._tables:
            self.ctx.call("mamdr_sum_squares_f64", _ptr(self.params[off:off + n_rows * dim]), n_rows * dim, _ptr(self._sq),
                          _ptr(self.table_ws), self.table_ws_bytes, self.stream)
            self.ctx.launches += 1
            tot += float(self._sq.item())
        self.desc.frozen_reg = self.l2_emb * tot

    def predict(self, data, offset, rows, use_order=False):
        """Sigmoid outputs of one inference mini-batch (device tensor) -- test hook."""
        probs = torch.zeros(rows, dtype=torch.float32, device=self.device)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        if self.emb_trainable:
            self._refresh_table_reg()
        if self.pass_kernel:
            self._eval_pass(self._pass(data, 1, use_order, offset, rows), loss, probs, with_auc=False)
            return probs, loss
        b = self._batch(data, offset, rows, use_order)
        self.ctx.call("mamdr_mlp_eval_step", C.byref(self.desc), C.byref(b), _ptr(self.user_table),
                      _ptr(self.item_table), _ptr(self.params), _ptr(self.ws), self.ws_bytes, _ptr(loss),
                      _ptr(probs), None, None, 0, self.precision, self.stream)
        self.ctx.launches += 2 + len(self.hidden)
        return probs, loss

    def auc_result(self):
        self.ctx.call("mamdr_auc_result", _ptr(self.auc_acc), self.num_thresholds, _ptr(self._auc_out), self.stream)
        self.ctx.launches += 1
        return float(self._auc_out.item())

    # ---- persistence.  The reference writes Keras HDF5 (base_model.py:177-181); h5py is not available here, so the weight
    # files are numpy .npz archives keyed by the variables' TF names (WEIGHT_EXT; INTEGRATION.md shows the h5 converter).
    def state_arrays(self, flat=None):
        """{TF variable name: array} of a weight snapshot (default: the live model)."""
        flat = self.params if flat is None else flat
        return {w.name: v.detach().cpu().numpy() for w, v in zip(self.trainable_weights, self.layout.views(flat[:self.params.numel()]))}

    def save_weights(self, path, flat=None):
        import numpy as np
        with open(path, "wb") as f:
            np.savez(f, **self.state_arrays(flat))

    def load_weights(self, path):
        import numpy as np
        blob = np.load(path)
        names = [w.name for w in self.trainable_weights]
        if sorted(names) != sorted(k for k in blob.files if not k.startswith("__")):
            raise ValueError("checkpoint layout mismatch")
        self.params.copy_(torch.from_numpy(self.layout.pack([blob[n] for n in names])))
        return blob


WEIGHT_EXT = ".npz"
