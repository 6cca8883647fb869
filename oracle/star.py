"""CPU restatement of the reference's STAR tower -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``/root/reference/model_zoo/Star/star.py:70-113`` (topology), ``Star/partitioned_norm.py:44-203``
(PartitionedNorm) and ``Star/star_fcn.py:50-139`` (StarFCN); compile / loss / optimizer as in ``star.py:23-33``
(BCE, one AdamOptimizer, AUC(500)).  SURVEY.md Appendix A-8.  The forward of both layers is PINNED to the reference's own `call` methods executed on numpy
(tests/golden/reference_star_layers_v1.npz, tests/test_reference_golden.py); the [EXT] Keras pieces (zero-debiased moving
average, initialisers, BCE, Adam) are defined here and remain unpinned.

  X = [E_u[uid] | E_i[pid] | E_d[dom]]                                   plain Keras Embeddings, NO l2 regulariser
  PartitionedNorm (norm = "pn"), d = domain of the batch:
      gamma = gamma_shared * gamma_specific[d] ; beta = beta_shared + beta_specific[d]
      train: mu, var = batch mean / biased variance over rows ; Y = gamma * (X - mu) / sqrt(var + 1e-3) + beta ;
             moving_mean[d], moving_var[d] <- zero-debiased EMA (momentum 0.99) of mu, var  ([EXT] TF-1.x
             K.moving_average_update -> assign_moving_average(zero_debias=True): biased = 0.99*biased + 0.01*value,
             t += 1, moving = biased / (1 - 0.99^t), per domain)
      eval : Y = gamma * (X - moving_mean[d]) / sqrt(moving_var[d] + 1e-3) + beta
  StarFCN x L (dense = "star"):  W = W_shared * W_specific[d] ; b = b_shared + b_specific[d] ; H = relu(H W + b)
  output Dense(1, sigmoid) with bias ; loss = mean Keras binary_crossentropy.  No dropout anywhere.
Trainable weights in Keras creation order: [domain_emb, gamma_specific, beta_specific, gamma_shared, beta_shared,
(kernel_specific, bias_specific, kernel_shared, bias_shared) x L, out_kernel, out_bias]; moving statistics are
non-trainable state (never part of the meta parameters).

All rows of a batch share one domain, so the E_d columns of X are constant within a batch: PartitionedNorm maps them to
beta exactly and the gradient of E_d is mathematically zero.  Both sides define it as exactly zero (the reference's
value is rounding noise of the batch statistics).
"""
import numpy as np

from . import auc as auc_mod
from .mlp import CLIP_HI, CLIP_LO, LOGIT_CLIP, AdamState, mm, sgd_apply

PN_EPS = 1e-3
PN_MOMENTUM = 0.99


class StarSpec(object):
    def __init__(self, n_uid, n_pid, n_domain, emb_dim=(128, 128, 128), hidden=(256, 128, 64)):
        self.n_uid, self.n_pid, self.n_domain = int(n_uid), int(n_pid), int(n_domain)
        self.emb_dim, self.hidden = tuple(emb_dim), tuple(hidden)
        self.in_dim = sum(self.emb_dim)

    @property
    def names(self):
        n = ['domain_emb', 'gamma_specific', 'beta_specific', 'gamma_shared', 'beta_shared']
        for l in range(len(self.hidden)):
            n += ['kernel_specific%d' % l, 'bias_specific%d' % l, 'kernel_shared%d' % l, 'bias_shared%d' % l]
        return n + ['out_kernel', 'out_bias']

    @property
    def shapes(self):
        D, dims = self.n_domain, (self.in_dim,) + self.hidden
        s = [(D, self.emb_dim[2]), (D, self.in_dim), (D, self.in_dim), (self.in_dim,), (self.in_dim,)]
        for l in range(len(self.hidden)):
            s += [(D, dims[l], dims[l + 1]), (D, dims[l + 1]), (dims[l], dims[l + 1]), (dims[l + 1],)]
        return s + [(dims[-1], 1), (1,)]


def init_star_weights(spec, seed):
    """Keras defaults: Embedding uniform(-0.05, 0.05); glorot_uniform kernels (shared and specific: fan_in / fan_out are
    the last two dims); zeros biases; gamma ones; beta zeros; Dense(1) glorot_uniform + zero bias."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for name, shape in zip(spec.names, spec.shapes):
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


class OracleStar(object):
    def __init__(self, spec, weights, user_table, item_table, lr=1e-3, dtype=np.float32):
        self.spec, self.dtype = spec, np.dtype(dtype)
        self.weights = [np.array(w, dtype=self.dtype) for w in weights]
        self.user_table = np.asarray(user_table, dtype=self.dtype)
        self.item_table = np.asarray(item_table, dtype=self.dtype)
        self.adam = AdamState(self.weights, lr=lr)
        self.auc = auc_mod.AUC(500)
        D, n = spec.n_domain, spec.in_dim
        self.biased_mean = np.zeros((D, n), dtype=self.dtype)
        self.biased_var = np.zeros((D, n), dtype=self.dtype)
        self.pn_steps = np.zeros(D, dtype=np.int64)
        self.moving_mean = np.zeros((D, n), dtype=self.dtype)       # Keras: zeros / ones initialisers
        self.moving_var = np.ones((D, n), dtype=self.dtype)

    def w(self, name):
        return self.weights[self.spec.names.index(name)]

    # ---- forward ---------------------------------------------------------------------------------------------
    def forward(self, uid, pid, domain, train):
        sp, dt = self.spec, self.dtype.type
        b = len(uid)
        X = np.concatenate([self.user_table[uid], self.item_table[pid],
                            np.broadcast_to(self.w('domain_emb')[domain], (b, sp.emb_dim[2]))], axis=1)
        gamma = self.w('gamma_shared') * self.w('gamma_specific')[domain]
        beta = self.w('beta_shared') + self.w('beta_specific')[domain]
        nu = sp.emb_dim[0] + sp.emb_dim[1]
        if train:
            mu = np.mean(X, axis=0, dtype=self.dtype)
            mu[nu:] = X[0, nu:]                                     # constant columns: mean = the value, exactly centred
            xc = X - mu
            xc[:, nu:] = 0
            var = np.mean(xc * xc, axis=0, dtype=self.dtype)
        else:
            mu, var = self.moving_mean[domain], self.moving_var[domain]
            xc = X - mu
        rstd = dt(1) / np.sqrt(var + dt(PN_EPS))
        xhat = xc * rstd
        H = [xhat * gamma + beta]
        cache = {'xhat': xhat, 'rstd': rstd, 'gamma': gamma, 'mu': mu, 'var': var}
        for l in range(len(sp.hidden)):
            W = self.w('kernel_shared%d' % l) * self.w('kernel_specific%d' % l)[domain]
            bias = self.w('bias_shared%d' % l) + self.w('bias_specific%d' % l)[domain]
            H.append(np.maximum(mm(H[l], W) + bias, dt(0)))
        z = mm(H[-1], self.w('out_kernel'))[:, 0] + self.w('out_bias')[0]
        self._last_logit = z
        p = dt(1) / (dt(1) + np.exp(-z))
        return H, p, cache

    def loss_from_p(self, p, y):
        dt = self.dtype.type
        ph = np.clip(p, dt(CLIP_LO), dt(CLIP_HI))
        lg = np.log(ph / (dt(1) - ph))
        bce = np.maximum(lg, dt(0)) - lg * y + np.log1p(np.exp(-np.abs(lg)))
        return float(np.mean(bce, dtype=np.float64))

    def _update_moving(self, domain, mu, var):
        dt = self.dtype.type
        self.pn_steps[domain] += 1
        t = int(self.pn_steps[domain])
        self.biased_mean[domain] = self.biased_mean[domain] * dt(PN_MOMENTUM) + mu * dt(1 - PN_MOMENTUM)
        self.biased_var[domain] = self.biased_var[domain] * dt(PN_MOMENTUM) + var * dt(1 - PN_MOMENTUM)
        corr = dt(1.0 - PN_MOMENTUM ** t)
        self.moving_mean[domain] = self.biased_mean[domain] / corr
        self.moving_var[domain] = self.biased_var[domain] / corr

    # ---- one training mini-batch --------------------------------------------------------------------------------
    def gradients(self, uid, pid, domain, label, update_stats=True):
        sp, dt = self.spec, self.dtype.type
        b = len(uid)
        y = np.asarray(label, dtype=self.dtype).reshape(-1)
        H, p, c = self.forward(uid, pid, domain, train=True)
        if update_stats:
            self._update_moving(domain, c['mu'], c['var'])
        loss = self.loss_from_p(p, y)
        L = len(sp.hidden)
        ds = (p - y) / dt(b)
        ds = np.where(np.abs(self._last_logit) <= dt(LOGIT_CLIP), ds, dt(0)).astype(self.dtype)
        g = {n: np.zeros_like(w) for n, w in zip(sp.names, self.weights)}
        g['out_bias'] = np.array([np.sum(ds)], dtype=self.dtype)
        g['out_kernel'] = mm(H[L].T, ds.reshape(-1, 1))
        dH = ds.reshape(-1, 1) * self.w('out_kernel').reshape(1, -1)
        for l in range(L - 1, -1, -1):
            dZ = dH * (H[l + 1] > 0).astype(self.dtype)
            dW = mm(H[l].T, dZ)
            db = np.sum(dZ, axis=0)
            Wsh, Wsp = self.w('kernel_shared%d' % l), self.w('kernel_specific%d' % l)
            g['kernel_shared%d' % l] = dW * Wsp[domain]
            g['kernel_specific%d' % l][domain] = dW * Wsh
            g['bias_shared%d' % l] = db
            g['bias_specific%d' % l][domain] = db
            dH = mm(dZ, (Wsh * Wsp[domain]).T)
        dgam = np.sum(dH * c['xhat'], axis=0)
        dbet = np.sum(dH, axis=0)
        g['gamma_shared'] = dgam * self.w('gamma_specific')[domain]
        g['gamma_specific'][domain] = dgam * self.w('gamma_shared')
        g['beta_shared'] = dbet
        g['beta_specific'][domain] = dbet
        # domain_emb: exactly zero (module docstring)
        return loss, p, [g[n] for n in sp.names]

    def train_on_batch(self, uid, pid, domain, label, masks=None, optimizer='adam', sgd_lr=None):
        loss, p, grads = self.gradients(uid, pid, domain, label)
        if optimizer == 'adam':
            self.adam.apply(self.weights, grads)
        else:   # the finetune stage's GradientDescentOptimizer (specific_base_model.py:118-122)
            sgd_apply(self.weights, grads, sgd_lr)
        self.auc.update_state(label, p.astype(np.float32))
        return loss, self.auc.result()

    def evaluate(self, uid, pid, domain, label, batch_size=1024):
        self.auc.reset_states()
        n = len(uid)
        losses = []
        for s in range(0, n, batch_size):
            e = min(n, s + batch_size)
            _, p, _ = self.forward(uid[s:e], pid[s:e], domain, train=False)
            y = np.asarray(label[s:e], dtype=self.dtype).reshape(-1)
            losses.append(self.loss_from_p(p, y))
            self.auc.update_state(y, p.astype(np.float32))
        return float(np.mean(losses)), self.auc.result()

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, values):
        for w, v in zip(self.weights, values):
            w[...] = v
