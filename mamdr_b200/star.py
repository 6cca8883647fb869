"""``Star`` base model -- mirrors ``/root/reference/model_zoo/Star/star.py``: ``build_model_structure`` (:70-97:
PartitionedNorm -> StarFCN x L -> Dense(1, sigmoid)), compile (:22-33) and the joint ``train`` loop (:35-68), over the
device-resident ``StarModel`` (C-ABI ``mamdr_star_train_step`` / ``mamdr_star_eval_step``; fp32 path).
BASELINE config #4 wraps it in MAMDR (``star_meta_mamdr_finetune``, meta_parms = ["emb", "kernel_shared", "bias_shared"]).
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .auc import thresholds as auc_thresholds
from .base_model import BaseModel
from .engine import MLPModel, NamedWeight, _ptr
from .layout import ParamLayout


def star_layout(n_domain, emb_dim, hidden):
    """Keras creation order of the trainable weights (oracle/star.py:StarSpec)."""
    n = sum(emb_dim)
    dims = (n,) + tuple(hidden)
    names = ['domain_emb', 'gamma_specific', 'beta_specific', 'gamma_shared', 'beta_shared']
    shapes = [(n_domain, emb_dim[2]), (n_domain, n), (n_domain, n), (n,), (n,)]
    for l in range(len(hidden)):
        names += ['kernel_specific%d' % l, 'bias_specific%d' % l, 'kernel_shared%d' % l, 'bias_shared%d' % l]
        shapes += [(n_domain, dims[l], dims[l + 1]), (n_domain, dims[l + 1]), (dims[l], dims[l + 1]), (dims[l + 1],)]
    names += ['out_kernel', 'out_bias']
    shapes += [(dims[-1], 1), (1,)]
    return ParamLayout(names, shapes)


def init_star_weights(layout, seed):
    """Keras defaults: Embedding uniform(+-0.05); glorot_uniform kernels (fan_in / fan_out = the last two dims);
    zero biases; gamma ones; beta zeros."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for name, shape in zip(layout.names, layout.shapes):
        if name == 'domain_emb':
            out.append(rng.uniform(-0.05, 0.05, size=shape).astype(np.float32))
        elif name.startswith('kernel') or name == 'out_kernel':
            lim = np.sqrt(6.0 / (shape[-2] + shape[-1]))
            out.append(rng.uniform(-lim, lim, size=shape).astype(np.float32))
        elif name.startswith('gamma'):
            out.append(np.ones(shape, dtype=np.float32))
        else:
            out.append(np.zeros(shape, dtype=np.float32))
    return out


StarDesc = _lib.StarDesc


class StarModel(MLPModel):
    """Device-resident STAR model with the Keras-like surface of ``MLPModel`` (fit_pass / evaluate / arenas)."""

    def __init__(self, n_uid, n_pid, n_domain, emb_dim, hidden, user_table, item_table, init_weights, lr=1e-3,
                 max_batch=1024, device="cuda:0", use_graphs=True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mamdr_b200 runs on CUDA devices only (no CPU fallback)")
        torch.cuda.set_device(self.device)
        self.ctx = _lib.Context(self.device.index or 0)
        lib = self.ctx.lib
        self.layout = lo = star_layout(n_domain, emb_dim, hidden)
        self.n_uid, self.n_pid, self.n_domain = int(n_uid), int(n_pid), int(n_domain)
        self.emb_dim, self.hidden = tuple(emb_dim), tuple(hidden)
        self.emb_trainable = False
        self.lr, self.beta1, self.beta2, self.eps = float(lr), 0.9, 0.999, 1e-8
        self.max_batch, self.precision, self.use_graphs = int(max_batch), _lib.PREC_FP32, bool(use_graphs)
        self.optimizer, self.sgd_lr = "adam", 0.0
        dev = self.device
        f32 = dict(dtype=torch.float32, device=dev)
        P_ = lo.total
        self.params, self.grads = torch.zeros(P_, **f32), torch.zeros(P_, **f32)
        self.m, self.v = torch.zeros(P_, **f32), torch.zeros(P_, **f32)
        self.params.copy_(torch.from_numpy(lo.pack(init_weights)))
        ut = np.ascontiguousarray(user_table, dtype=np.float32)
        it = np.ascontiguousarray(item_table, dtype=np.float32)
        assert ut.shape == (n_uid, emb_dim[0]) and it.shape == (n_pid, emb_dim[1])
        self.user_table, self.item_table = torch.from_numpy(ut).to(dev), torch.from_numpy(it).to(dev)
        d = StarDesc()
        d.n_layers = len(hidden)
        for i in range(3):
            d.emb_dim[i] = int(emb_dim[i])
        for l, h in enumerate(hidden):
            d.hidden[l] = int(h)
            d.off_ksp[l], d.off_bsp[l] = lo.offset('kernel_specific%d' % l), lo.offset('bias_specific%d' % l)
            d.off_ksh[l], d.off_bsh[l] = lo.offset('kernel_shared%d' % l), lo.offset('bias_shared%d' % l)
        d.n_domain, d.n_uid, d.n_pid = int(n_domain), int(n_uid), int(n_pid)
        d.pn_eps, d.pn_momentum = 1e-3, 0.99
        d.off_domain_emb = lo.offset('domain_emb')
        d.off_gamma_sp, d.off_beta_sp = lo.offset('gamma_specific'), lo.offset('beta_specific')
        d.off_gamma_sh, d.off_beta_sh = lo.offset('gamma_shared'), lo.offset('beta_shared')
        d.off_out_kernel, d.off_out_bias = lo.offset('out_kernel'), lo.offset('out_bias')
        d.arena_floats = P_
        self.desc = d
        self.ws_bytes = lib.mamdr_star_workspace_bytes(C.byref(d), self.max_batch)
        if self.ws_bytes == 0:
            raise _lib.MamdrError(-1, "mamdr_star_workspace_bytes rejected the descriptor")
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=dev)
        # non-trainable PartitionedNorm state: [moving_mean | moving_var | biased_mean | biased_var] [D, n] + steps
        n = sum(emb_dim)
        self.pn_state = torch.zeros(lib.mamdr_star_state_bytes(C.byref(d)), dtype=torch.uint8, device=dev)
        self.pn_state[4 * n_domain * n:8 * n_domain * n].view(torch.float32).fill_(1.0)   # moving_var = 1 (Keras)
        self.opt_state = torch.zeros(lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device=dev)
        self.num_thresholds = 500
        self.thresholds = torch.from_numpy(auc_thresholds(self.num_thresholds)).to(dev)
        self.auc_acc = torch.zeros(4, self.num_thresholds, **f32)
        self._auc_out = torch.zeros(1, **f32)
        self._auc_zero = torch.zeros(4, self.num_thresholds, **f32)
        self._recording, self.program_ops, self.launch_times, self.pass_kernel = False, 0, None, False
        self.reset_optimizer()
        self._graphs, self._loss_bufs = {}, {}

    TF = {'domain_emb': 'domain_emb/embeddings:0', 'gamma_specific': 'partitioned_norm/gamma_specific:0',
          'beta_specific': 'partitioned_norm/beta_specific:0', 'gamma_shared': 'partitioned_norm/gamma_shared:0',
          'beta_shared': 'partitioned_norm/beta_shared:0', 'out_kernel': 'dense/kernel:0', 'out_bias': 'dense/bias:0'}

    @property
    def trainable_weights(self):
        out = []
        for name, view, off, n in zip(self.layout.names, self.layout.views(self.params), self.layout.offsets, self.layout.numels):
            tf_name = self.TF.get(name)
            if tf_name is None:   # kernel_specific2 -> star_fcn_2/kernel_specific:0
                base, l = name.rstrip('0123456789'), name[len(name.rstrip('0123456789')):]
                tf_name = "star_fcn%s/%s:0" % ("" if l == "0" else "_" + l, base)
            out.append(NamedWeight(tf_name, view, off, n))
        return out

    # ---- weight snapshots.  Keras get_weights / set_weights / save_weights / load_weights include the NON-trainable
    # PartitionedNorm moving statistics (Star/partitioned_norm.py:74-99), so the best checkpoint, the finetune stage's
    # per-domain restarts and val_and_test("test") all carry them: a snapshot here is [arena | pn_state words].
    def pn_parts(self):
        """(float32 view of the statistics [4, D, n] flattened, int32 view of the per-domain update counters)."""
        n, D = sum(self.emb_dim), self.n_domain
        return self.pn_state[:16 * D * n].view(torch.float32), self.pn_state[16 * D * n:].view(torch.int32)

    def get_weights(self):
        return torch.cat([self.params, self.pn_state.view(torch.float32)])

    def set_weights(self, flat):
        P_ = self.params.numel()
        self.copy_(self.params, flat[:P_])
        if flat.numel() > P_:
            self.pn_state.view(torch.float32).copy_(flat[P_:])

    def save_weights(self, path, flat=None):
        import numpy as np
        arrays = self.state_arrays(flat)
        pn = self.pn_state if flat is None or flat.numel() == self.params.numel() else flat[self.params.numel():].view(torch.uint8)
        arrays["__pn_state__"] = pn.detach().cpu().numpy().view(np.uint8)
        with open(path, "wb") as f:
            np.savez(f, **arrays)

    def load_weights(self, path):
        blob = super(StarModel, self).load_weights(path)
        if "__pn_state__" in blob.files:
            self.pn_state.copy_(torch.from_numpy(blob["__pn_state__"].copy()))
        return blob

    def moving_stats(self):
        n, D = sum(self.emb_dim), self.n_domain
        f = self.pn_state[:16 * D * n].view(torch.float32).view(4, D, n)
        return f[0], f[1]

    def _train_step(self, data, offset, rows, loss_slot, probs=None, with_auc=True):
        b = self._batch(data, offset, rows, True)
        st = self.stream
        self.ctx.call("mamdr_star_train_step", C.byref(self.desc), C.byref(b), _ptr(self.user_table), _ptr(self.item_table),
                      _ptr(self.params), _ptr(self.grads), _ptr(self.pn_state), _ptr(self.ws), self.ws_bytes, _ptr(loss_slot),
                      _ptr(probs), _ptr(self.auc_acc if with_auc else None), _ptr(self.thresholds), self.num_thresholds, st)
        if self.optimizer == "adam":     # the reference's compile, star.py:24-27
            self.ctx.call("mamdr_adam_step", _ptr(self.params), _ptr(self.m), _ptr(self.v), _ptr(self.grads),
                          self.params.numel(), _ptr(self.opt_state), self.lr, self.beta1, self.beta2, self.eps, st)
        else:                            # the finetune stage's GradientDescentOptimizer (specific_base_model.py:120): the gradient
            #                              arena is fully written (zero outside the batch's domain slices), one sweep over it
            self.ctx.call("mamdr_sgd_step", _ptr(self.params), _ptr(self.grads), self.params.numel(), _ptr(self.opt_state),
                          self.sgd_lr, st)
        L = len(self.hidden)
        self.ctx.launches += 2 + 4 + L + 1 + L + L + 1 + 2 + 1   # memsets, assemble, pn x2, eff, fwd, head, dH/dY, dW, colsum, grads, pn bwd, adam

    def _eval_batch(self, data, off, rows, use_order, loss, probs, with_auc):
        b = self._batch(data, off, rows, use_order)
        self.ctx.call("mamdr_star_eval_step", C.byref(self.desc), C.byref(b), _ptr(self.user_table), _ptr(self.item_table),
                      _ptr(self.params), _ptr(self.pn_state), _ptr(self.ws), self.ws_bytes, _ptr(loss), _ptr(probs),
                      _ptr(self.auc_acc if with_auc else None), _ptr(self.thresholds), self.num_thresholds if with_auc else 0, self.stream)
        self.ctx.launches += 5 + len(self.hidden)

    def evaluate(self, data, steps=None):
        steps = data.n_step if steps is None else int(steps)
        self.reset_states()
        losses = torch.zeros(max(steps, 1), dtype=torch.float32, device=self.device)
        for s, (off, rows) in enumerate(self._pass_plan(data, steps)):
            self._eval_batch(data, off, rows, False, losses[s:s + 1], None, True)
        auc = self.auc_result()
        return float(losses[:steps].double().mean().item()) if steps else 0.0, auc

    def predict(self, data, offset, rows, use_order=False):
        probs = torch.zeros(rows, dtype=torch.float32, device=self.device)
        loss = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._eval_batch(data, offset, rows, use_order, loss, probs, False)
        return probs, loss


class Star(BaseModel):
    def __init__(self, dataset, config):
        super(Star, self).__init__(dataset, config)

    def build_model(self):
        mc, tc = self.model_config, self.train_config
        if mc.get('norm') != "pn" or mc.get('dense') != "star":
            raise NotImplementedError("only norm='pn', dense='star' (the STAR topology) is built")
        if mc.get('auxiliary_net'):
            raise NotImplementedError("auxiliary_net is false in every shipped config (SURVEY.md 2.1 #12)")
        if not (tc['load_pretrain_emb'] and not tc['emb_trainable']):
            raise NotImplementedError("the STAR path covers frozen pretrained tables (Taobao configs)")
        if tc['optimizer'] != 'adam' or tc['loss'] != 'binary_crossentropy':
            raise NotImplementedError("only adam + binary_crossentropy are on the hot path")
        if self.b200_config.get('precision', 'fp32') != 'fp32':
            raise ValueError("the STAR tower runs in the fp32 mode (b200.precision = 'fp32')")
        self.emb_trainable = False
        emb_dim = (mc['user_dim'], mc['item_dim'], mc['domain_dim'])
        self.layout = star_layout(self.n_domain, emb_dim, mc['hidden_dim'])
        self._init_draws = 0
        self.init_seed = self.b200_config.get('init_seed', self.dataset.conf['seed'])
        return StarModel(self.n_uid, self.n_pid, self.n_domain, emb_dim, tuple(mc['hidden_dim']), self.dataset.user_table,
                         self.dataset.item_table, self.draw_initial_weights(), lr=tc['learning_rate'],
                         max_batch=self.dataset.batch_size, device=self.b200_config.get('device', self.dataset.device),
                         use_graphs=self.b200_config.get('cuda_graphs', True))

    def draw_initial_weights(self):
        w = init_star_weights(self.layout, [self.init_seed, self._init_draws])
        self._init_draws += 1
        return w

    def train(self):
        """star.py:35-68 -- joint training: shuffled domains, one full pass each, one Adam."""
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
