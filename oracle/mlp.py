"""MLP tower train / eval step + TF-flavoured Adam, numpy restatement.
TEST INFRASTRUCTURE (see oracle/__init__.py).  parity unpinned ([EXT] TF 1.12 /
DeepCTR 0.9.0 semantics).

Reference call sites restated here:
  * topology  ``model_zoo/DeepCTR/deepctr.py:95-136``  (3 embeddings -> concat ->
    DNN(256,128,64; relu; dropout) -> Dense(1, no bias) -> PredictionLayer)
  * compile   ``model_zoo/DeepCTR/deepctr.py:54-60``   (BCE, AdamOptimizer, AUC)
  * one batch ``Model.train_on_batch`` as driven by ``model_zoo/mamdr.py:85-97`` and
    ``model_zoo/domain_negotiation.py:71-72``.

Weights are a list of arrays ordered like ``model.trainable_weights``
(``model_zoo/maml.py:158-159``):
  [user_emb?, item_emb?, domain_emb, kernel0..kernelL-1, bias0..biasL-1, dense_kernel, global_bias]
"""
import numpy as np

from . import philox
from .auc import AUC

try:  # CPU GEMM backend: torch (MKL/oneDNN) is 10-50x faster than this image's OpenBLAS; numpy otherwise
    import torch as _torch
except Exception:  # pragma: no cover
    _torch = None
MATMUL_BACKEND = "torch" if _torch is not None else "numpy"


def mm(a, b):
    """a @ b.  Both backends are plain fp32 (or fp64) CPU GEMMs; they differ only in summation order."""
    if MATMUL_BACKEND == "torch" and a.dtype == b.dtype and a.dtype in (np.float32, np.float64):
        return _torch.matmul(_torch.from_numpy(a), _torch.from_numpy(b)).numpy()
    return a @ b


CLIP_LO = np.float32(1e-7)
CLIP_HI = np.float32(1.0) - np.float32(1e-7)
# the clip's zero-gradient region expressed on the logit: |s| > ln((1 - 1e-7) / 1e-7).  Exact-arithmetic equivalent of
# `CLIP_LO <= p <= CLIP_HI`, but robust in fp32 where p's spacing next to 1 (6e-8) spans a 0.7-wide band of logits.
LOGIT_CLIP = np.float32(16.118095)


class MLPSpec(object):
    def __init__(self, n_uid, n_pid, n_domain, emb_dim=(128, 128, 128), hidden=(256, 128, 64),
                 dropout=0.5, dropout_seed=1024, l2_emb=1e-5, emb_trainable=False):
        self.n_uid, self.n_pid, self.n_domain = int(n_uid), int(n_pid), int(n_domain)
        self.emb_dim = tuple(int(x) for x in emb_dim)
        self.hidden = tuple(int(x) for x in hidden)
        self.dropout = float(dropout)
        self.dropout_seed = int(dropout_seed)
        self.l2_emb = float(l2_emb)
        self.emb_trainable = bool(emb_trainable)
        self.in_dim = sum(self.emb_dim)

    @property
    def names(self):
        L = len(self.hidden)
        n = (['user_emb', 'item_emb'] if self.emb_trainable else []) + ['domain_emb']
        n += ['kernel%d' % i for i in range(L)] + ['bias%d' % i for i in range(L)]
        return n + ['dense_kernel', 'global_bias']

    @property
    def shapes(self):
        dims = (self.in_dim,) + self.hidden
        L = len(self.hidden)
        s = ([(self.n_uid, self.emb_dim[0]), (self.n_pid, self.emb_dim[1])] if self.emb_trainable else [])
        s += [(self.n_domain, self.emb_dim[2])]
        s += [(dims[i], dims[i + 1]) for i in range(L)] + [(dims[i + 1],) for i in range(L)]
        return s + [(dims[-1], 1), (1,)]


class AdamState(object):
    """``tf.train.AdamOptimizer`` slots.  beta-power accumulators are fp32
    variables multiplied once per apply; created once and never reset by the meta loop
    (``model_zoo/maml.py:181-187`` assigns weights only)."""

    def __init__(self, weights, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr, self.beta1, self.beta2, self.eps = (np.float32(lr), np.float32(beta1),
                                                     np.float32(beta2), np.float32(eps))
        self.m = [np.zeros_like(w) for w in weights]
        self.v = [np.zeros_like(w) for w in weights]
        self.b1pow = np.float32(beta1)
        self.b2pow = np.float32(beta2)
        self.step = 0  # number of applies so far == dropout global_step

    def apply(self, weights, grads):
        """TF ``ApplyAdam`` kernel order of operations (non-Nesterov):
        alpha = lr*sqrt(1-b2pow)/(1-b1pow); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
        var -= (m*alpha)/(sqrt(v)+eps); then b1pow*=b1, b2pow*=b2."""
        one = np.float32(1.0)
        for w, g, m, v in zip(weights, grads, self.m, self.v):
            dt = w.dtype.type
            alpha = dt(self.lr) * np.sqrt(dt(one) - dt(self.b2pow)) / (dt(one) - dt(self.b1pow))
            g = g.astype(w.dtype, copy=False)
            m += (g - m) * (dt(one) - dt(self.beta1))
            v += (g * g - v) * (dt(one) - dt(self.beta2))
            w -= (m * alpha) / (np.sqrt(v) + dt(self.eps))
        self.b1pow = np.float32(self.b1pow * self.beta1)
        self.b2pow = np.float32(self.b2pow * self.beta2)
        self.step += 1


def sgd_apply(weights, grads, lr):
    """``train.GradientDescentOptimizer`` used by the finetune stage
    (``model_zoo/specific_base_model.py:120``, ``model_zoo/base_model.py:69``)."""
    for w, g in zip(weights, grads):
        w -= g.astype(w.dtype, copy=False) * w.dtype.type(np.float32(lr))


class OracleMLP(object):
    """One model instance: current weights + frozen tables + Adam slots + AUC metric."""

    def __init__(self, spec, weights, user_table=None, item_table=None, lr=1e-3, dtype=np.float32):
        self.spec = spec
        self.dtype = np.dtype(dtype)
        assert len(weights) == len(spec.names)
        self.weights = [np.array(w, dtype=self.dtype) for w in weights]
        for w, s in zip(self.weights, spec.shapes):
            assert w.shape == tuple(s), (w.shape, s)
        if not spec.emb_trainable:
            self.user_table = np.asarray(user_table, dtype=self.dtype)
            self.item_table = np.asarray(item_table, dtype=self.dtype)
            # constant part of the L2 penalty (frozen tables still carry the regulariser, A-2)
            self.frozen_reg = float(spec.l2_emb * (np.sum(self.user_table.astype(np.float64) ** 2)
                                                  + np.sum(self.item_table.astype(np.float64) ** 2)))
        else:
            self.user_table = self.item_table = None
            self.frozen_reg = 0.0
        self.adam = AdamState(self.weights, lr=lr)
        self.auc = AUC(500)

    # ---- weight plumbing (maml.py:181-194, utils/tool.py:36-45) ---------------------------
    preact_log = None   # set to a list to record min |pre-activation| per (training forward, layer); see forward()

    def w(self, name):
        return self.weights[self.spec.names.index(name)]

    def get_weights(self):
        return [x.copy() for x in self.weights]

    def set_weights(self, values):
        for dst, src in zip(self.weights, values):
            dst[...] = src

    def _tables(self):
        if self.spec.emb_trainable:
            return self.w('user_emb'), self.w('item_emb')
        return self.user_table, self.item_table

    # ---- forward --------------------------------------------------------------------------
    def forward(self, uid, pid, domain, train, masks=None):
        """SURVEY.md A-10 forward.  ``domain`` is one int (uniform per batch, dataset.py:73-99)."""
        sp, dt = self.spec, self.dtype.type
        Eu, Ei = self._tables()
        Ed = self.w('domain_emb')
        b = len(uid)
        X = np.concatenate([Eu[uid], Ei[pid], np.broadcast_to(Ed[domain], (b, sp.emb_dim[2]))], axis=1)
        H = [X]
        L = len(sp.hidden)
        for l in range(L):
            Z = mm(H[l], self.w('kernel%d' % l)) + self.w('bias%d' % l)
            A = np.maximum(Z, dt(0))
            M = None
            if train and sp.dropout > 0:
                M = masks[l] if masks is not None else philox.dropout_mask(
                    b, sp.hidden[l], sp.dropout_seed + l, self.adam.step, sp.dropout, self.dtype.type)
                A = A * M
            if self.preact_log is not None:
                # test diagnostic: the smallest |pre-activation| among the units that survive dropout -- a ReLU gate whose
                # pre-activation is within rounding distance of zero can open on one implementation and close on another
                az = np.abs(Z) if M is None else np.where(M > 0, np.abs(Z), np.inf)
                self.preact_log.append(float(np.min(az)))
            H.append(A)
        z = mm(H[L], self.w('dense_kernel'))                   # [b,1]
        s = z[:, 0] + self.w('global_bias')[0]
        self._last_logit = s
        p = dt(1) / (dt(1) + np.exp(-s))
        return H, p

    def loss_from_p(self, p, y):
        """Keras binary_crossentropy on probabilities (A-3) + embedding L2 penalties (A-2)."""
        dt = self.dtype.type
        ph = np.clip(p, dt(CLIP_LO), dt(CLIP_HI))
        lg = np.log(ph / (dt(1) - ph))
        bce = np.maximum(lg, dt(0)) - lg * y + np.log1p(np.exp(-np.abs(lg)))
        reg = self.frozen_reg
        reg_names = ['domain_emb'] + (['user_emb', 'item_emb'] if self.spec.emb_trainable else [])
        for n in reg_names:
            reg += self.spec.l2_emb * float(np.sum(self.w(n).astype(np.float64) ** 2))
        return float(np.mean(bce, dtype=np.float64)) + reg

    # ---- one training mini-batch ------------------------------------------------------------
    def gradients(self, uid, pid, domain, label, masks=None, train=True):
        """``train=False``: the gradients of the INFERENCE-mode loss (no dropout) -- what the K.function of
        ``model_zoo/maml.py:196-233`` computes, which is built without the learning-phase placeholder ([EXT] default 0)."""
        sp, dt = self.spec, self.dtype.type
        b = len(uid)
        y = np.asarray(label, dtype=self.dtype).reshape(-1)
        H, p = self.forward(uid, pid, domain, train=train, masks=masks)
        loss = self.loss_from_p(p, y)
        L = len(sp.hidden)
        inv_keep = (np.float32(1.0) / np.float32(1.0 - sp.dropout)).astype(self.dtype) if (sp.dropout > 0 and train) else dt(1)
        ds = (p - y) / dt(b)
        ds = np.where(np.abs(self._last_logit) <= dt(LOGIT_CLIP), ds, dt(0)).astype(self.dtype)
        g = {}
        g['global_bias'] = np.array([np.sum(ds)], dtype=self.dtype)
        g['dense_kernel'] = mm(H[L].T, ds.reshape(-1, 1))
        dH = ds.reshape(-1, 1) * self.w('dense_kernel').reshape(1, -1)
        for l in range(L - 1, -1, -1):
            # H[l+1] = relu(Z) * M  =>  M * 1[Z>0] == inv_keep * 1[H[l+1] > 0]
            dZ = dH * np.where(H[l + 1] > 0, inv_keep, dt(0)).astype(self.dtype)
            g['kernel%d' % l] = mm(H[l].T, dZ)
            g['bias%d' % l] = np.sum(dZ, axis=0)
            dH = mm(dZ, self.w('kernel%d' % l).T)
        du, di = sp.emb_dim[0], sp.emb_dim[1]
        two_l2 = dt(2.0 * sp.l2_emb)
        gEd = two_l2 * self.w('domain_emb')
        gEd[domain] += np.sum(dH[:, du + di:], axis=0)
        g['domain_emb'] = gEd
        if sp.emb_trainable:
            # TF order: the IndexedSlices gradient is de-duplicated first (unsorted_segment_sum:
            # duplicates of one id added in batch order), THEN aggregated with the dense regulariser gradient
            su = np.zeros_like(self.w('user_emb'))
            np.add.at(su, uid, dH[:, :du])
            si = np.zeros_like(self.w('item_emb'))
            np.add.at(si, pid, dH[:, du:du + di])
            g['user_emb'], g['item_emb'] = two_l2 * self.w('user_emb') + su, two_l2 * self.w('item_emb') + si
        return loss, p, [g[n] for n in sp.names]

    def train_on_batch(self, uid, pid, domain, label, masks=None, optimizer='adam', sgd_lr=None):
        loss, p, grads = self.gradients(uid, pid, domain, label, masks)
        if optimizer == 'adam':
            self.adam.apply(self.weights, grads)
        else:
            sgd_apply(self.weights, grads, sgd_lr)
        self.auc.update_state(label, p.astype(np.float32))
        return loss, self.auc.result()

    # ---- evaluate (Keras Model.evaluate, A-3 last paragraph) -------------------------------
    def evaluate(self, uid, pid, domain, label, batch_size=1024):
        self.auc.reset_states()
        n = len(uid)
        losses = []
        for s in range(0, n, batch_size):
            e = min(n, s + batch_size)
            _, p = self.forward(uid[s:e], pid[s:e], domain, train=False)
            y = np.asarray(label[s:e], dtype=self.dtype).reshape(-1)
            losses.append(self.loss_from_p(p, y))
            self.auc.update_state(y, p.astype(np.float32))
        return float(np.mean(losses)), self.auc.result()
