"""CPU restatement of the reference's multi-task towers (MMOE / PLE with num_levels = 1 / SharedBottom) --
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows ``/root/reference/model_zoo/DeepMTLCTR/deep_mtl_ctr.py:21-66``: ``deepctr.models.MMOE / PLE / SharedBottom``
([EXT] deepctr 0.9.0, absent from /root/reference; its published topology is restated below) with one task per domain,
one compiled sub-``Model(inputs, outputs[t])`` per domain, all sharing ONE ``AdamOptimizer`` (:53, :57-65).
SURVEY.md Appendix A-9.  Parity unpinned: the [EXT] semantics are defined here and checked against torch.autograd
(tests/test_oracle_mtl.py).

  X = [E_u[uid] | E_i[pid] | E_d[dom]]                       embeddings carry l2 = 1e-5 (deepctr l2_reg_embedding)
  DNN(h) = (Dense(h_l) -> relu -> dropout) for every l       (deepctr.layers.DNN; dropout after every activation)
  mmoe : experts e = 0..E-1 = DNN(hidden_dim)(X), shared by every domain
         domain t: a = softmax(DNN(gate_dnn_hidden_units)(X) . G_t)   G_t [g_last, E], no bias
                   mix = sum_e a[:, e] * expert_e ; T = DNN(tower_hidden_dim)(mix) ; p = sigmoid(T . w_t + g_t)
  ple  : (num_levels = 1, i.e. CGC, last level => task gates only)  S shared experts + Q specific experts per domain;
         domain t mixes [its Q specific experts..., the S shared experts...] through its gate; tower as above
  shared_bottom : one bottom DNN(hidden_dim), no gate; tower per domain

A step on domain t differentiates sub-model t only: the variables reachable from output t (tables, the experts domain
t mixes, gate t, tower t) are updated; every other variable keeps its value AND its Adam slots (TF creates the slots
per variable; `apply_gradients` touches only the variables it is given) while the beta powers -- one pair per
optimizer -- advance on every step.

Dropout masks follow oracle/philox.py with one stream per DNN layer: seed = dropout_seed + stream,
stream = 8 * e + l (expert e), 4096 + 8 * t + l (gate t), 8192 + 8 * t + l (tower t).

Physical / list order of the trainable weights ([EXT]; ours): [user_emb?, item_emb?, domain_emb, shared experts...,
then per domain t: its specific experts..., gate t DNN, gate t out, tower t DNN, tower t out, bias t]; inside a DNN:
kernel0.., bias0.. (deepctr creates the kernels first).
"""
import numpy as np

from . import philox
from .auc import AUC
from .mlp import CLIP_HI, CLIP_LO, LOGIT_CLIP, AdamState, mm


class MTLSpec(object):
    def __init__(self, n_uid, n_pid, n_domain, kind='mmoe', emb_dim=(128, 128, 128), expert_hidden=(256, 128),
                 tower_hidden=(64,), gate_hidden=(64,), num_experts=5, specific_expert_num=5, shared_expert_num=2,
                 dropout=0.5, dropout_seed=1024, l2_emb=1e-5, emb_trainable=True):
        self.n_uid, self.n_pid, self.n_domain = int(n_uid), int(n_pid), int(n_domain)
        self.kind = kind
        self.emb_dim = tuple(int(x) for x in emb_dim)
        self.expert_hidden = tuple(int(x) for x in expert_hidden)
        self.tower_hidden = tuple(int(x) for x in tower_hidden)
        self.gate_hidden = tuple(int(x) for x in gate_hidden)
        self.dropout, self.dropout_seed, self.l2_emb = float(dropout), int(dropout_seed), float(l2_emb)
        self.emb_trainable = bool(emb_trainable)
        self.in_dim = sum(self.emb_dim)
        D = self.n_domain
        if kind == 'mmoe':
            self.n_shared, self.n_specific, self.has_gate = int(num_experts), 0, True
        elif kind == 'ple':
            self.n_shared, self.n_specific, self.has_gate = int(shared_expert_num), int(specific_expert_num), True
        elif kind == 'shared_bottom':
            self.n_shared, self.n_specific, self.has_gate = 1, 0, False
        else:
            raise ValueError("kind must be mmoe / ple / shared_bottom")
        S, Q = self.n_shared, self.n_specific
        self.n_experts = S + D * Q
        # expert ids mixed by domain t, in gate-column order (deepctr: specific experts first, then the shared ones)
        self.expert_sets = [[S + t * Q + q for q in range(Q)] + list(range(S)) for t in range(D)]
        self.k = S + Q
        assert len(self.expert_hidden) >= 1 and len(self.tower_hidden) >= 1
        assert not self.has_gate or len(self.gate_hidden) >= 1

    # ---- layout ---------------------------------------------------------------------------------------------
    def _dnn(self, prefix, dims):
        L = len(dims) - 1
        names = ['%s_kernel%d' % (prefix, l) for l in range(L)] + ['%s_bias%d' % (prefix, l) for l in range(L)]
        shapes = [(dims[l], dims[l + 1]) for l in range(L)] + [(dims[l + 1],) for l in range(L)]
        return names, shapes

    def _expert(self, e):
        return self._dnn('expert%d' % e, (self.in_dim,) + self.expert_hidden)

    def _build(self):
        names, shapes = [], []
        if self.emb_trainable:
            names += ['user_emb', 'item_emb']
            shapes += [(self.n_uid, self.emb_dim[0]), (self.n_pid, self.emb_dim[1])]
        names += ['domain_emb']
        shapes += [(self.n_domain, self.emb_dim[2])]
        for e in range(self.n_shared):
            n, s = self._expert(e)
            names += n
            shapes += s
        for t in range(self.n_domain):
            for q in range(self.n_specific):
                n, s = self._expert(self.n_shared + t * self.n_specific + q)
                names += n
                shapes += s
            if self.has_gate:
                n, s = self._dnn('gate%d' % t, (self.in_dim,) + self.gate_hidden)
                names += n + ['gate%d_out' % t]
                shapes += s + [(self.gate_hidden[-1], self.k)]
            n, s = self._dnn('tower%d' % t, (self.expert_hidden[-1],) + self.tower_hidden)
            names += n + ['tower%d_out' % t, 'bias%d' % t]
            shapes += s + [(self.tower_hidden[-1], 1), (1,)]
        return names, shapes

    @property
    def names(self):
        return self._build()[0]

    @property
    def shapes(self):
        return self._build()[1]

    def reachable(self, t):
        """Names of the variables sub-model t trains."""
        out = (['user_emb', 'item_emb'] if self.emb_trainable else []) + ['domain_emb']
        for e in self.expert_sets[t]:
            out += self._expert(e)[0]
        pre = ('gate%d_' % t, 'tower%d_' % t)
        out += [n for n in self.names if n.startswith(pre)] + ['bias%d' % t]
        return out

    @staticmethod
    def stream(kind, idx, layer):
        return {'expert': 0, 'gate': 4096, 'tower': 8192}[kind] + 8 * idx + layer


def init_mtl_weights(spec, seed):
    """deepctr / Keras defaults: embeddings RandomNormal(0, 1e-4); DNN kernels glorot_normal; zero biases; the gate /
    tower output Dense layers glorot_uniform; PredictionLayer bias zero."""
    rng = np.random.Generator(np.random.PCG64(seed))

    def glorot_normal(shape):                                # truncated normal (|x| <= 2 sigma), Keras' stddev correction
        out = rng.standard_normal(shape)
        bad = np.abs(out) > 2.0
        while bad.any():
            out[bad] = rng.standard_normal(int(bad.sum()))
            bad = np.abs(out) > 2.0
        return (out * (np.sqrt(2.0 / (shape[0] + shape[1])) / 0.87962566103423978)).astype(np.float32)

    out = []
    for name, shape in zip(spec.names, spec.shapes):
        if name.endswith('_emb'):
            out.append((rng.standard_normal(shape) * 1e-4).astype(np.float32))
        elif name.endswith('_out'):
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            out.append(rng.uniform(-lim, lim, size=shape).astype(np.float32))
        elif '_kernel' in name:
            out.append(glorot_normal(shape))
        else:
            out.append(np.zeros(shape, dtype=np.float32))
    return out


class OracleMTL(object):
    def __init__(self, spec, weights, user_table=None, item_table=None, lr=1e-3, dtype=np.float32):
        self.spec, self.dtype = spec, np.dtype(dtype)
        self.names = spec.names
        self.index = {n: i for i, n in enumerate(self.names)}
        assert len(weights) == len(self.names)
        self.weights = [np.array(w, dtype=self.dtype) for w in weights]
        for w, s in zip(self.weights, spec.shapes):
            assert w.shape == tuple(s), (w.shape, s)
        if not spec.emb_trainable:
            self.user_table = np.asarray(user_table, dtype=self.dtype)
            self.item_table = np.asarray(item_table, dtype=self.dtype)
            self.frozen_reg = float(spec.l2_emb * (np.sum(self.user_table.astype(np.float64) ** 2)
                                                  + np.sum(self.item_table.astype(np.float64) ** 2)))
        else:
            self.user_table = self.item_table = None
            self.frozen_reg = 0.0
        self.adam = AdamState(self.weights, lr=lr)
        self.auc = AUC(500)

    def w(self, name):
        return self.weights[self.index[name]]

    def get_weights(self):
        return [x.copy() for x in self.weights]

    def set_weights(self, values):
        for dst, src in zip(self.weights, values):
            dst[...] = src

    def _tables(self):
        if self.spec.emb_trainable:
            return self.w('user_emb'), self.w('item_emb')
        return self.user_table, self.item_table

    # ---- one DNN: list of post-dropout activations H[0..L] ------------------------------------------------------
    def _dnn_forward(self, prefix, kind, idx, X, widths, train, masks):
        sp, dt = self.spec, self.dtype.type
        H = [X]
        for l, n in enumerate(widths):
            A = np.maximum(mm(H[l], self.w('%s_kernel%d' % (prefix, l))) + self.w('%s_bias%d' % (prefix, l)), dt(0))
            if train and sp.dropout > 0:
                key = (kind, idx, l)
                M = masks[key] if masks is not None else philox.dropout_mask(
                    len(X), n, sp.dropout_seed + sp.stream(kind, idx, l), self.adam.step, sp.dropout, self.dtype.type)
                A = A * M
            H.append(A)
        return H

    def _dnn_backward(self, prefix, H, dOut, g):
        """dOut = gradient w.r.t. the DNN's (post-dropout) output; fills g[kernel / bias]; returns d(input)."""
        sp, dt = self.spec, self.dtype.type
        inv_keep = (np.float32(1.0) / np.float32(1.0 - sp.dropout)).astype(self.dtype) if sp.dropout > 0 else dt(1)
        dH = dOut
        for l in range(len(H) - 2, -1, -1):
            dZ = dH * np.where(H[l + 1] > 0, inv_keep, dt(0)).astype(self.dtype)
            g['%s_kernel%d' % (prefix, l)] = mm(H[l].T, dZ)
            g['%s_bias%d' % (prefix, l)] = np.sum(dZ, axis=0)
            dH = mm(dZ, self.w('%s_kernel%d' % (prefix, l)).T)
        return dH

    def forward(self, uid, pid, domain, train, masks=None):
        sp, dt = self.spec, self.dtype.type
        Eu, Ei = self._tables()
        b, t = len(uid), int(domain)
        X = np.concatenate([Eu[uid], Ei[pid], np.broadcast_to(self.w('domain_emb')[t], (b, sp.emb_dim[2]))], axis=1)
        c = {'X': X, 'experts': []}
        for e in sp.expert_sets[t]:
            c['experts'].append(self._dnn_forward('expert%d' % e, 'expert', e, X, sp.expert_hidden, train, masks))
        if sp.has_gate:
            c['gate'] = self._dnn_forward('gate%d' % t, 'gate', t, X, sp.gate_hidden, train, masks)
            logits = mm(c['gate'][-1], self.w('gate%d_out' % t))
            ex = np.exp(logits - np.max(logits, axis=1, keepdims=True))
            a = ex / np.sum(ex, axis=1, keepdims=True)
            mix = np.zeros_like(c['experts'][0][-1])
            for j in range(sp.k):
                mix = mix + a[:, j:j + 1] * c['experts'][j][-1]
            c['a'] = a
        else:
            mix = c['experts'][0][-1]
        c['tower'] = self._dnn_forward('tower%d' % t, 'tower', t, mix, sp.tower_hidden, train, masks)
        s = mm(c['tower'][-1], self.w('tower%d_out' % t))[:, 0] + self.w('bias%d' % t)[0]
        self._last_logit = s
        p = dt(1) / (dt(1) + np.exp(-s))
        return c, p

    def loss_from_p(self, p, y):
        dt = self.dtype.type
        ph = np.clip(p, dt(CLIP_LO), dt(CLIP_HI))
        lg = np.log(ph / (dt(1) - ph))
        bce = np.maximum(lg, dt(0)) - lg * y + np.log1p(np.exp(-np.abs(lg)))
        reg = self.frozen_reg
        for n in ['domain_emb'] + (['user_emb', 'item_emb'] if self.spec.emb_trainable else []):
            reg += self.spec.l2_emb * float(np.sum(self.w(n).astype(np.float64) ** 2))
        return float(np.mean(bce, dtype=np.float64)) + reg

    def gradients(self, uid, pid, domain, label, masks=None):
        """-> loss, p, {name: gradient} over ``spec.reachable(domain)``."""
        sp, dt = self.spec, self.dtype.type
        b, t = len(uid), int(domain)
        y = np.asarray(label, dtype=self.dtype).reshape(-1)
        c, p = self.forward(uid, pid, t, train=True, masks=masks)
        loss = self.loss_from_p(p, y)
        ds = (p - y) / dt(b)
        ds = np.where(np.abs(self._last_logit) <= dt(LOGIT_CLIP), ds, dt(0)).astype(self.dtype)
        g = {}
        g['bias%d' % t] = np.array([np.sum(ds)], dtype=self.dtype)
        g['tower%d_out' % t] = mm(c['tower'][-1].T, ds.reshape(-1, 1))
        dT = ds.reshape(-1, 1) * self.w('tower%d_out' % t).reshape(1, -1)
        dMix = self._dnn_backward('tower%d' % t, c['tower'], dT, g)
        dX = np.zeros_like(c['X'])
        if sp.has_gate:
            a = c['a']
            da = np.stack([np.sum(dMix * c['experts'][j][-1], axis=1) for j in range(sp.k)], axis=1)
            dlogit = a * (da - np.sum(a * da, axis=1, keepdims=True))
            for j, e in enumerate(sp.expert_sets[t]):
                dX = dX + self._dnn_backward('expert%d' % e, c['experts'][j], a[:, j:j + 1] * dMix, g)
            g['gate%d_out' % t] = mm(c['gate'][-1].T, dlogit)
            dG = mm(dlogit, self.w('gate%d_out' % t).T)
            dX = dX + self._dnn_backward('gate%d' % t, c['gate'], dG, g)
        else:
            dX = dX + self._dnn_backward('expert%d' % sp.expert_sets[t][0], c['experts'][0], dMix, g)
        du, di = sp.emb_dim[0], sp.emb_dim[1]
        two_l2 = dt(2.0 * sp.l2_emb)
        gEd = two_l2 * self.w('domain_emb')
        gEd[t] += np.sum(dX[:, du + di:], axis=0)
        g['domain_emb'] = gEd
        if sp.emb_trainable:
            su = np.zeros_like(self.w('user_emb'))
            np.add.at(su, uid, dX[:, :du])
            si = np.zeros_like(self.w('item_emb'))
            np.add.at(si, pid, dX[:, du:du + di])
            g['user_emb'], g['item_emb'] = two_l2 * self.w('user_emb') + su, two_l2 * self.w('item_emb') + si
        return loss, p, g

    def train_on_batch(self, uid, pid, domain, label, masks=None, optimizer='adam', sgd_lr=None):
        loss, p, g = self.gradients(uid, pid, domain, label, masks)
        idx = [self.index[n] for n in self.spec.reachable(domain)]
        assert sorted(g.keys()) == sorted(self.spec.reachable(domain))
        sub = AdamState.__new__(AdamState)                    # a view of the shared optimizer on sub-model t's variables
        sub.__dict__.update(self.adam.__dict__)
        sub.m, sub.v = [self.adam.m[i] for i in idx], [self.adam.v[i] for i in idx]
        sub.apply([self.weights[i] for i in idx], [g[self.names[i]] for i in idx])
        self.adam.b1pow, self.adam.b2pow, self.adam.step = sub.b1pow, sub.b2pow, sub.step
        self.auc.update_state(label, p.astype(np.float32))
        return loss, self.auc.result()

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
