"""``DeepMTLCTR`` base model (MMOE / PLE with num_levels = 1 / SharedBottom, one task per domain) -- mirrors
``/root/reference/model_zoo/DeepMTLCTR/deep_mtl_ctr.py``: ``build_model`` (:21-66: ``deepctr.models.MMOE / PLE /
SharedBottom``, then ONE compiled sub-``Model(inputs, outputs[t])`` per domain, all sharing one ``AdamOptimizer``, :53),
the joint ``train`` loop (:68-98) and ``val_and_test`` through the per-domain sub-models (:185-221), over the
device-resident ``MTLModel`` (C-ABI ``mamdr_mtl_train_step`` / ``mamdr_mtl_eval_step``; fp32 path).

BASELINE config #5 wraps it in ``DomainNegotiation`` (``mmoe_meta_domain_negotiation`` /
``ple_meta_domain_negotiation``).  The reference cannot run that combination (the multi-output
Keras model is never compiled); the defined semantics are: meta parameters = ``model.trainable_weights``, a pass on
domain t runs sub-model t's train function, validation uses sub-model t.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .auc import thresholds as auc_thresholds
from .base_model import BaseModel
from .engine import MLPModel, NamedWeight, _ptr
from .layout import ParamLayout, glorot_normal


class MTLTopology(object):
    """Which experts domain t mixes (gate-column order: its specific experts first, then the shared ones, as deepctr's
    CGC concatenates them) and the physical order of the trainable weights: [user_emb?, item_emb?, domain_emb,
    shared experts..., then per domain t: its specific experts..., gate t DNN, gate t out, tower t DNN, tower t out,
    bias t]; inside a DNN: kernel0.., bias0..  -- so the variables of sub-model t are two contiguous arena spans."""

    def __init__(self, kind, n_uid, n_pid, n_domain, emb_dim, expert_hidden, tower_hidden, gate_hidden, num_experts=0,
                 specific_expert_num=0, shared_expert_num=0, emb_trainable=True):
        self.kind = kind
        self.n_uid, self.n_pid, self.n_domain = int(n_uid), int(n_pid), int(n_domain)
        self.emb_dim = tuple(int(x) for x in emb_dim)
        self.in_dim = sum(self.emb_dim)
        self.expert_hidden = tuple(int(x) for x in expert_hidden)
        self.tower_hidden = tuple(int(x) for x in tower_hidden)
        self.gate_hidden = tuple(int(x) for x in gate_hidden)
        self.emb_trainable = bool(emb_trainable)
        if kind == 'mmoe':
            self.n_shared, self.n_specific, self.has_gate = int(num_experts), 0, True
        elif kind == 'ple':
            self.n_shared, self.n_specific, self.has_gate = int(shared_expert_num), int(specific_expert_num), True
        elif kind == 'shared_bottom':
            self.n_shared, self.n_specific, self.has_gate = 1, 0, False
        else:
            raise ValueError("MTL kind must be mmoe / ple / shared_bottom, not {!r}".format(kind))
        S, Q = self.n_shared, self.n_specific
        self.k = S + Q
        if not 1 <= self.k <= _lib.MTL_MAX_K:
            raise ValueError("a domain mixes %d experts; supported: 1..%d" % (self.k, _lib.MTL_MAX_K))
        self.expert_sets = [[S + t * Q + q for q in range(Q)] + list(range(S)) for t in range(self.n_domain)]
        names, shapes = [], []
        if self.emb_trainable:
            names += ['user_emb', 'item_emb']
            shapes += [(self.n_uid, self.emb_dim[0]), (self.n_pid, self.emb_dim[1])]
        names += ['domain_emb']
        shapes += [(self.n_domain, self.emb_dim[2])]

        def dnn(prefix, dims):
            L = len(dims) - 1
            return (['%s_kernel%d' % (prefix, l) for l in range(L)] + ['%s_bias%d' % (prefix, l) for l in range(L)],
                    [(dims[l], dims[l + 1]) for l in range(L)] + [(dims[l + 1],) for l in range(L)])

        def add(ns):
            names.extend(ns[0])
            shapes.extend(ns[1])

        for e in range(S):
            add(dnn('expert%d' % e, (self.in_dim,) + self.expert_hidden))
        for t in range(self.n_domain):
            for q in range(Q):
                add(dnn('expert%d' % (S + t * Q + q), (self.in_dim,) + self.expert_hidden))
            if self.has_gate:
                add(dnn('gate%d' % t, (self.in_dim,) + self.gate_hidden))
                add((['gate%d_out' % t], [(self.gate_hidden[-1], self.k)]))
            add(dnn('tower%d' % t, (self.expert_hidden[-1],) + self.tower_hidden))
            add((['tower%d_out' % t, 'bias%d' % t], [(self.tower_hidden[-1], 1), (1,)]))
        self.layout = ParamLayout(names, shapes)

    def reachable(self, t):
        """Names of the variables sub-model t trains (everything output t depends on)."""
        lo = self.layout
        out = (['user_emb', 'item_emb'] if self.emb_trainable else []) + ['domain_emb']
        pre = tuple('expert%d_' % e for e in self.expert_sets[t]) + ('gate%d_' % t, 'tower%d_' % t)
        out += [n for n in lo.names if n.startswith(pre)] + ['bias%d' % t]
        return out

    def dense_spans(self, t):
        """Maximal contiguous arena spans [(begin, len)] of sub-model t's variables, tables excluded."""
        lo = self.layout
        spans = []
        for n in self.reachable(t):
            if n in ('user_emb', 'item_emb'):
                continue
            i = lo.index(n)
            b, e = lo.offsets[i], (lo.offsets[i] + lo.numels[i] + 31) // 32 * 32
            if spans and spans[-1][1] == b:
                spans[-1][1] = e
            else:
                spans.append([b, e])
        return [(b, e - b) for b, e in spans]


def init_mtl_weights(layout, seed):
    """deepctr / Keras defaults: embeddings RandomNormal(0, 1e-4); DNN kernels glorot_normal; zero biases; the gate /
    tower output Dense layers glorot_uniform; PredictionLayer bias zero."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for name, shape in zip(layout.names, layout.shapes):
        if name.endswith('_emb'):
            out.append((rng.standard_normal(shape) * 1e-4).astype(np.float32))
        elif name.endswith('_out'):
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            out.append(rng.uniform(-lim, lim, size=shape).astype(np.float32))
        elif '_kernel' in name:
            out.append(glorot_normal(rng, shape))
        else:
            out.append(np.zeros(shape, dtype=np.float32))
    return out


class MTLModel(MLPModel):
    """Device-resident multi-task model with the Keras-like surface of ``MLPModel`` (fit_pass / evaluate / arenas); the
    sub-model that runs is chosen by the domain of the data it is given."""

    def __init__(self, topo, init_weights, user_table=None, item_table=None, dropout=0.5, dropout_seed=1024, l2_emb=1e-5,
                 lr=1e-3, max_batch=1024, device="cuda:0", use_graphs=True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mamdr_b200 runs on CUDA devices only (no CPU fallback)")
        torch.cuda.set_device(self.device)
        self.ctx = _lib.Context(self.device.index or 0)
        lib = self.ctx.lib
        self.topo, self.layout = topo, topo.layout
        lo = self.layout
        self.n_uid, self.n_pid, self.n_domain = topo.n_uid, topo.n_pid, topo.n_domain
        self.emb_dim, self.hidden = topo.emb_dim, topo.expert_hidden
        self.emb_trainable = topo.emb_trainable
        self.lr, self.beta1, self.beta2, self.eps = float(lr), 0.9, 0.999, 1e-8
        self.max_batch, self.precision, self.use_graphs = int(max_batch), _lib.PREC_FP32, bool(use_graphs)
        self.optimizer, self.sgd_lr = "adam", 0.0
        self.l2_emb = float(l2_emb)
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        P_ = lo.total
        self.params, self.grads = torch.zeros(P_, **f32), torch.zeros(P_, **f32)
        self.m, self.v = torch.zeros(P_, **f32), torch.zeros(P_, **f32)
        self.params.copy_(torch.from_numpy(lo.pack(init_weights)))
        self.frozen_reg = 0.0
        if not self.emb_trainable:
            ut = np.ascontiguousarray(user_table, dtype=np.float32)
            it = np.ascontiguousarray(item_table, dtype=np.float32)
            assert ut.shape == (self.n_uid, self.emb_dim[0]) and it.shape == (self.n_pid, self.emb_dim[1])
            self.frozen_reg = float(l2_emb * (np.sum(ut.astype(np.float64) ** 2) + np.sum(it.astype(np.float64) ** 2)))
            self.user_table, self.item_table = torch.from_numpy(ut).to(dev), torch.from_numpy(it).to(dev)
        else:
            self.user_table = self.item_table = None
        d = _lib.MtlDesc()
        for i in range(3):
            d.emb_dim[i] = self.emb_dim[i]
        d.n_domain, d.n_uid, d.n_pid = self.n_domain, self.n_uid, self.n_pid
        d.emb_trainable, d.has_gate, d.k = int(self.emb_trainable), int(topo.has_gate), topo.k
        d.n_expert_layers, d.n_tower_layers = len(topo.expert_hidden), len(topo.tower_hidden)
        d.n_gate_layers = len(topo.gate_hidden) if topo.has_gate else 0
        for l, h in enumerate(topo.expert_hidden):
            d.expert_hidden[l] = h
        for l, h in enumerate(topo.tower_hidden):
            d.tower_hidden[l] = h
        if topo.has_gate:
            for l, h in enumerate(topo.gate_hidden):
                d.gate_hidden[l] = h
        d.dropout_rate, d.dropout_seed, d.l2_emb, d.frozen_reg = float(dropout), int(dropout_seed), float(l2_emb), self.frozen_reg
        d.off_user_emb, d.off_item_emb, d.off_domain_emb = lo.offset('user_emb'), lo.offset('item_emb'), lo.offset('domain_emb')
        d.arena_floats = P_
        self.desc = d
        # one descriptor + one Adam span list per sub-model
        self.domains, self.spans = [], []
        for t in range(self.n_domain):
            dm = _lib.MtlDomain()
            dm.domain = t
            for j, e in enumerate(topo.expert_sets[t]):
                dm.expert_id[j] = e
                for l in range(len(topo.expert_hidden)):
                    dm.off_expert_kernel[j][l] = lo.offset('expert%d_kernel%d' % (e, l))
                    dm.off_expert_bias[j][l] = lo.offset('expert%d_bias%d' % (e, l))
            if topo.has_gate:
                for l in range(len(topo.gate_hidden)):
                    dm.off_gate_kernel[l], dm.off_gate_bias[l] = lo.offset('gate%d_kernel%d' % (t, l)), lo.offset('gate%d_bias%d' % (t, l))
                dm.off_gate_out = lo.offset('gate%d_out' % t)
            for l in range(len(topo.tower_hidden)):
                dm.off_tower_kernel[l], dm.off_tower_bias[l] = lo.offset('tower%d_kernel%d' % (t, l)), lo.offset('tower%d_bias%d' % (t, l))
            dm.off_tower_out, dm.off_bias = lo.offset('tower%d_out' % t), lo.offset('bias%d' % t)
            self.domains.append(dm)
            sp = topo.dense_spans(t)
            self.spans.append(((C.c_int64 * len(sp))(*[b for b, _ in sp]), (C.c_int64 * len(sp))(*[n for _, n in sp]), len(sp)))
        self.ws_bytes = lib.mamdr_mtl_workspace_bytes(C.byref(d), self.max_batch)
        if self.ws_bytes == 0:
            raise _lib.MamdrError(-1, "mamdr_mtl_workspace_bytes rejected the descriptor")
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.opt_state = torch.zeros(lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device=dev)
        self.num_thresholds = 500
        self.thresholds = torch.from_numpy(auc_thresholds(self.num_thresholds)).to(dev)
        self.auc_acc = torch.zeros(4, self.num_thresholds, **f32)
        self._auc_out = torch.zeros(1, **f32)
        self._auc_zero = torch.zeros(4, self.num_thresholds, **f32)
        self._recording, self.program_ops, self.launch_times, self.pass_kernel = False, 0, None, False
        if self.emb_trainable:
            self._tables = []
            for name, n_rows, dim in (("user_emb", self.n_uid, self.emb_dim[0]), ("item_emb", self.n_pid, self.emb_dim[1])):
                self._tables.append((lo.offset(name), int(n_rows), int(dim), torch.full((int(n_rows),), -1, dtype=torch.int32, device=dev)))
            self.table_ws_bytes = lib.mamdr_adam_table_workspace_bytes()
            self.table_ws = torch.zeros(self.table_ws_bytes, dtype=torch.uint8, device=dev)
            self._sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.reset_optimizer()
        self._graphs, self._loss_bufs = {}, {}

    @property
    def trainable_weights(self):
        """Names carry the substrings the reference's meta-parameter selection matches on (maml.py:160-177):
        "emb" for the embedding tables, "expert" / "gate" / "tower" for the DNN blocks."""
        out = []
        for name, view, off, n in zip(self.layout.names, self.layout.views(self.params), self.layout.offsets, self.layout.numels):
            tf_name = "sparse_emb_%s/embeddings:0" % name if name.endswith('_emb') else name.replace('_', '/', 1) + ":0"
            out.append(NamedWeight(tf_name, view, off, n))
        return out

    def launches_per_train_step(self):
        t = self.topo
        Le, Lg, Lt = len(t.expert_hidden), (len(t.gate_hidden) if t.has_gate else 0), len(t.tower_hidden)
        n = 2 + Le + Lg + (1 if t.has_gate else 0) + Lt + 1          # memset, assemble, fwd, gate_mix, head
        n += (Lt - 1) + 1 + (2 if t.has_gate else 0) + (Le - 1) + max(Lg - 1, 0)   # dH chains, dMix, mix bwd, gate_out grad
        n += 1 + Le + Lg + Lt + 2                                    # dX (K-segmented), dW, colsum, domain grad
        n += (2 + 4 if self.emb_trainable else 0) + 1                # sort + segment sum, 2 x (slot scatter, sweep), ranges Adam
        return n

    def _train_step(self, data, offset, rows, loss_slot, probs=None, with_auc=True):
        if self.optimizer != "adam":
            raise NotImplementedError("the multi-task path applies Adam (the sub-models' compile, deep_mtl_ctr.py:53-65)")
        t = int(data.domain)
        b = self._batch(data, offset, rows, True)
        st = self.stream
        if self.emb_trainable:
            self.desc.frozen_reg = 0.0   # training adds the tables' l2 penalty inside the fused table sweep
        self.ctx.call("mamdr_mtl_train_step", C.byref(self.desc), C.byref(self.domains[t]), C.byref(b), _ptr(self.user_table),
                      _ptr(self.item_table), _ptr(self.params), _ptr(self.grads), _ptr(self.ws), self.ws_bytes, _ptr(self.opt_state),
                      _ptr(loss_slot), _ptr(probs), _ptr(self.auc_acc if with_auc else None), _ptr(self.thresholds),
                      self.num_thresholds, st)
        if self.emb_trainable:
            for ti, (off, n_rows, dim, slot) in enumerate(self._tables):
                ids, srows, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
                rc = self.ctx.lib.mamdr_mtl_sparse_grads(C.byref(self.desc), int(rows), _ptr(self.ws), ti, C.byref(ids), C.byref(srows),
                                                         C.byref(cnt))
                if rc != 0:
                    raise _lib.MamdrError(rc, "mamdr_mtl_sparse_grads")
                n_el = n_rows * dim
                self.ctx.call("mamdr_adam_table_step", _ptr(self.params[off:off + n_el]), _ptr(self.m[off:off + n_el]),
                              _ptr(self.v[off:off + n_el]), n_rows, dim, ids, srows, cnt, int(rows), _ptr(slot), self.l2_emb,
                              _ptr(self.opt_state), self.lr, self.beta1, self.beta2, self.eps, _ptr(loss_slot), _ptr(self.table_ws),
                              self.table_ws_bytes, st)
        begin, length, n = self.spans[t]
        self.ctx.call("mamdr_adam_ranges_step", _ptr(self.params), _ptr(self.m), _ptr(self.v), _ptr(self.grads), begin, length, n,
                      _ptr(self.opt_state), self.lr, self.beta1, self.beta2, self.eps, st)
        self.ctx.launches += self.launches_per_train_step()

    def _eval_batch(self, data, off, rows, use_order, loss, probs, with_auc):
        b = self._batch(data, off, rows, use_order)
        self.ctx.call("mamdr_mtl_eval_step", C.byref(self.desc), C.byref(self.domains[int(data.domain)]), C.byref(b),
                      _ptr(self.user_table), _ptr(self.item_table), _ptr(self.params), _ptr(self.ws), self.ws_bytes, _ptr(loss),
                      _ptr(probs), _ptr(self.auc_acc if with_auc else None), _ptr(self.thresholds),
                      self.num_thresholds if with_auc else 0, self.stream)
        t = self.topo
        self.ctx.launches += 2 + len(t.expert_hidden) + (len(t.gate_hidden) + 1 if t.has_gate else 0) + len(t.tower_hidden)

    def evaluate(self, data, steps=None):
        steps = data.n_step if steps is None else int(steps)
        self.reset_states()
        if self.emb_trainable:
            self._refresh_table_reg()
        losses = torch.zeros(max(steps, 1), dtype=torch.float32, device=self.device)
        for s, (off, rows) in enumerate(self._pass_plan(data, steps)):
            self._eval_batch(data, off, rows, False, losses[s:s + 1], None, True)
        auc = self.auc_result()
        return float(losses[:steps].double().mean().item()) if steps else 0.0, auc

    def predict(self, data, offset, rows, use_order=False):
        probs = torch.zeros(rows, dtype=torch.float32, device=self.device)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        if self.emb_trainable:
            self._refresh_table_reg()
        self._eval_batch(data, offset, rows, use_order, loss, probs, False)
        return probs, loss


class DeepMTLCTR(BaseModel):
    def __init__(self, dataset, config):
        super(DeepMTLCTR, self).__init__(dataset, config)

    def build_model(self):
        mc, tc = self.model_config, self.train_config
        name = mc['name']
        # deep_mtl_ctr.py:25-48 -- substring dispatch on the model name
        kind = "shared_bottom" if "shared_bottom" in name else ("mmoe" if "mmoe" in name else ("ple" if "ple" in name else None))
        if kind is None:
            raise ValueError("model: {} is not a multi-task tower".format(name))
        if kind == "ple" and mc.get('num_levels', 1) != 1:
            raise NotImplementedError("PLE is built for num_levels = 1 (every shipped config)")
        if tc['optimizer'] != 'adam' or tc['loss'] != 'binary_crossentropy':
            raise NotImplementedError("only adam + binary_crossentropy are on the hot path")
        if self.b200_config.get('precision', 'fp32') != 'fp32':
            raise ValueError("the multi-task towers run in the fp32 mode (b200.precision = 'fp32')")
        # build_emb (:108-121): emb_trainable is honoured only with load_pretrain_emb; default tables always train
        if tc['load_pretrain_emb']:
            if self.dataset.user_table is None or self.dataset.item_table is None:
                raise AttributeError("dataset has no pretrained user_emb / item_emb")
            emb_trainable = bool(tc['emb_trainable'])
        else:
            emb_trainable = True
        self.emb_trainable = emb_trainable
        emb_dim = (mc['user_dim'], mc['item_dim'], mc['domain_dim'])
        self.topo = MTLTopology(kind, self.n_uid, self.n_pid, self.n_domain, emb_dim, mc['hidden_dim'], mc['tower_hidden_dim'],
                                mc.get('gate_dnn_hidden_units', ()), num_experts=mc.get('num_experts', 0),
                                specific_expert_num=mc.get('specific_expert_num', 0),
                                shared_expert_num=mc.get('shared_expert_num', 0), emb_trainable=emb_trainable)
        self.layout = self.topo.layout
        self._init_draws = 0
        self.init_seed = self.b200_config.get('init_seed', self.dataset.conf['seed'])
        w0 = self.draw_initial_weights()
        if emb_trainable and tc['load_pretrain_emb']:
            w0[self.layout.index('user_emb')] = self.dataset.user_table
            w0[self.layout.index('item_emb')] = self.dataset.item_table
        return MTLModel(self.topo, w0, user_table=None if emb_trainable else self.dataset.user_table,
                        item_table=None if emb_trainable else self.dataset.item_table, dropout=mc.get('dropout', 0.0),
                        dropout_seed=1024, l2_emb=1e-5, lr=tc['learning_rate'], max_batch=self.dataset.batch_size,
                        device=self.b200_config.get('device', self.dataset.device),
                        use_graphs=self.b200_config.get('cuda_graphs', True))

    def draw_initial_weights(self):
        w = init_mtl_weights(self.layout, [self.init_seed, self._init_draws])
        self._init_draws += 1
        return w

    def train(self):
        """deep_mtl_ctr.py:68-98 -- joint training: shuffled domains, one full pass of sub-model idx each, one Adam."""
        self.model.reset_optimizer()
        train_sequence = list(range(self.n_domain))
        for epoch in range(self.train_config['epoch']):
            self.log("Epoch: {}".format(epoch), "-" * 30)
            train_sequence = self.schedule.shuffle_sequence(train_sequence)
            self.stage_epoch_orders(list(train_sequence))
            for idx in train_sequence:
                self.log("Train on: Domain {}".format(idx))
                old_time = time.time()
                self.model.reset_states()
                self.run_train_pass(idx)
                self.log("Training time: ", time.time() - old_time)
            self.log("Val Result: ")
            avg_loss, avg_auc, domain_loss, domain_auc = self.val_and_test("val")
            if self.early_stop_step(avg_auc):
                break
            self.log("Test Result: ")
            self.val_and_test("test")
