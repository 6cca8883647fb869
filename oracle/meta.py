"""Domain Negotiation / Domain Regularization outer loops, numpy restatement.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, line by line, the host control flow of the reference:
  * ``model_zoo/domain_negotiation.py:37-116,118-123``   (DN, Alg. 1)
  * ``model_zoo/mamdr.py:41-159,168-196``               (DN on shared + DR on specific)
  * ``model_zoo/specific_base_model.py:44-97,164-172``  (early stop snapshots, val/test, merge)
  * ``model_zoo/base_model.py:111-175,208-224``         (plain val/test, weighted AUC, early stop)
  * ``model_zoo/DeepCTR/deepctr.py:63-93``              (joint 'alternate' training)

Python's ``random`` is unseeded in the reference (``run.py:26`` seeds TF only), so the
domain order, the DR support samples and the per-pass batch order are *injected* through a
``schedule`` object exposing ``shuffle_sequence(seq)``, ``sample_support(candidates, k)`` and
``batch_order(domain, n)``; the product consumes the same object type
(``mamdr_b200/schedule.py``) so both sides see identical draws in identical order.

``data`` is ``{'train'|'val'|'test': {domain: {'uid','pid','label'}}}`` of numpy arrays.

Pinned: the update rules / merge / accumulate / early-stop / weighted-AUC functions below reproduce, bit for bit, vectors
obtained by executing the reference's own methods (tests/golden/reference_meta_v1.npz, tests/test_reference_golden.py).
"""
from copy import deepcopy

import numpy as np


def n_steps(n, batch_size):
    return int(np.ceil(n / float(batch_size)))  # utils/dataset.py:25


def train_pass(model, d, domain, order, batch_size, max_steps=0, optimizer='adam', sgd_lr=None):
    """One full pass (``train_step`` mini-batches, ragged tail kept) over a domain's train set."""
    n = len(order)
    steps = n_steps(n, batch_size)
    if max_steps and max_steps > 0:
        steps = min(steps, max_steps)
    tot, auc = 0.0, 0.0
    for s in range(steps):
        sel = order[s * batch_size:(s + 1) * batch_size]
        loss, auc = model.train_on_batch(d['uid'][sel], d['pid'][sel], domain, d['label'][sel],
                                         optimizer=optimizer, sgd_lr=sgd_lr)
        tot += loss
    return tot / max(steps, 1), auc, steps


def merge_weights(shared, specific, method):  # specific_base_model.py:164-172
    if method == 'plus':
        return [x + y for x, y in zip(shared, specific)]
    if method == 'times':
        return [x * y for x, y in zip(shared, specific)]
    return []


def weighted_auc(data, mode, domain_auc):  # base_model.py:157-175
    num = sum(len(data[mode][k]['uid']) * v for k, v in domain_auc.items())
    den = sum(len(data[mode][k]['uid']) for k in domain_auc)
    return num / den


class EarlyStop(object):  # base_model.py:202-224
    def __init__(self, patience):
        self.patience, self.counter, self.best_metric, self.early_stop = patience, 0, None, False

    def step(self, metric, on_improve):
        if self.best_metric is None:
            self.best_metric = metric
            on_improve()
        elif metric <= self.best_metric:
            self.counter += 1
            if self.counter >= self.patience:
                self.early_stop = True
        else:
            on_improve()
            self.best_metric = metric
            self.counter = 0
        return self.early_stop


class OracleDN(object):
    """``DomainNegotiation.train`` with ``target_domain=-1`` (every shipped config)."""

    def __init__(self, model, data, train_config, batch_size, schedule):
        self.model, self.data, self.tc, self.bs, self.schedule = model, data, train_config, batch_size, schedule
        self.meta_weights = model.get_weights()                     # :29
        self.target = train_config.get('target_domain', -1)
        self.sequence = [k for k in sorted(data['train'].keys()) if not (self.target >= 0 and k == self.target)]   # :135-141
        if isinstance(train_config.get('meta_sequence'), list):    # :142-145
            if len(train_config['meta_sequence']) != len(self.sequence):
                raise ValueError("All the domains must be given in the sequence")
            self.sequence = list(train_config['meta_sequence'])
        self.es = EarlyStop(train_config['patience'])
        self.best_weights = None
        self.log = []

    def train_epoch(self):
        tc = self.tc
        if tc['shuffle_sequence']:                                   # :41-42
            self.sequence = self.schedule.shuffle_sequence(self.sequence)
        self.model.set_weights(self.meta_weights)                    # :50
        train_sequence = list(self.sequence) + ([self.target] if self.target >= 0 else [])   # :44-47
        for idx in train_sequence:                                   # :53-84
            self.model.auc.reset_states()                            # :56-57
            d = self.data['train'][idx]
            order = self.schedule.batch_order(idx, len(d['uid']))
            cap = tc.get('meta_train_step', 0) if idx != self.target else 0   # :67 (the target domain always trains a full pass)
            loss, auc, steps = train_pass(self.model, d, idx, order, self.bs, cap)
            self.log.append((idx, loss, auc, steps))
        new = self.model.get_weights()                               # :118-123
        beta = tc['meta_learning_rate']
        for var in range(len(new)):
            self.meta_weights[var] += (new[var] - self.meta_weights[var]) * beta
        self.model.set_weights(self.meta_weights)                    # :88
        self._fit_target()                                           # :89-93

    def _fit_target(self, steps=0):
        """``model.fit(self.target_iter, steps_per_epoch=...)`` on the target domain (a full pass unless `steps` is given)."""
        if self.target >= 0:
            d = self.data['train'][self.target]
            train_pass(self.model, d, self.target, self.schedule.batch_order(self.target, len(d['uid'])), self.bs, steps)

    def val_and_test(self, mode, weights=None):  # base_model.py:111-144
        if mode not in ('val', 'test'):
            raise ValueError("Mode can be either val or test, not: {}".format(mode))
        if mode == 'test' and self.best_weights is not None:
            self.model.set_weights(self.best_weights)                # load_model(best h5), :121
        domain_loss, domain_auc = {}, {}
        for idx, d in self.data[mode].items():
            l, a = self.model.evaluate(d['uid'], d['pid'], idx, d['label'], self.bs)
            domain_loss[idx], domain_auc[idx] = float(l), float(a)
        all_loss, all_auc = 0, 0          # plain left-to-right adds like the reference (Python >= 3.12's sum() is compensated)
        for k in domain_loss:
            all_loss += domain_loss[k]
            all_auc += domain_auc[k]
        avg_loss = all_loss / len(domain_loss)
        avg_auc = all_auc / len(domain_auc)
        return avg_loss, avg_auc, domain_loss, domain_auc

    def early_stop_step(self, metric):
        def keep():
            self.best_weights = self.model.get_weights()             # save_model(h5)
        return self.es.step(metric, keep)

    def val(self):                                                   # maml.py:343-353
        if self.tc.get('meta_finetune_step', 0) > 0:
            return self.meta_finetune_val()
        return self.val_and_test("val")

    def meta_finetune_val(self):                                     # maml.py:244-287
        domain_loss, domain_auc = {}, {}
        all_loss, all_auc = 0, 0
        weights = self.model.get_weights()
        for idx, d in self.data['train'].items():
            self.model.set_weights(weights)
            for epoch in range(self.tc['meta_finetune_step']):
                train_pass(self.model, d, idx, self.schedule.batch_order(idx, len(d['uid'])), self.bs)
            v = self.data['val'][idx]
            l, a = self.model.evaluate(v['uid'], v['pid'], idx, v['label'], self.bs)
            domain_loss[idx], domain_auc[idx] = float(l), float(a)
            all_loss += domain_loss[idx]
            all_auc += domain_auc[idx]
        self.model.set_weights(weights)
        return all_loss / len(domain_loss), all_auc / len(domain_auc), domain_loss, domain_auc


class OracleReptile(OracleDN):
    """``Reptile.train`` with ``target_domain=-1`` (``model_zoo/reptile.py:45-99,127-142``): the model is reset to theta
    before every domain; theta moves after every domain, or (``batch`` names) once per epoch by the summed deltas."""

    def __init__(self, model, data, train_config, batch_size, schedule, name='mlp_meta_reptile'):
        OracleDN.__init__(self, model, data, train_config, batch_size, schedule)
        self.name = name
        self.sequence = list(range(len(data['train'])))             # :36 (the target domain stays in the list and is skipped, :47)
        self.accum = [np.zeros_like(w) for w in self.meta_weights]  # :31

    def train_epoch(self):
        tc = self.tc
        beta = np.float32(tc['meta_learning_rate'])
        self.sequence = self.schedule.shuffle_sequence(self.sequence)   # :46
        for idx in self.sequence:
            if self.target >= 0 and idx == self.target:              # :47-48
                continue
            self.model.auc.reset_states()                            # :53-54
            self.model.set_weights(self.meta_weights)                # :57
            d = self.data['train'][idx]
            order = self.schedule.batch_order(idx, len(d['uid']))
            loss, auc, steps = train_pass(self.model, d, idx, order, self.bs, tc.get('meta_train_step', 0))
            self.log.append((idx, loss, auc, steps))
            self._fit_target(steps=1)                                # :82-85 one step on the target domain
            new = self.model.get_weights()
            if "batch" in self.name:                                 # :134-137
                for var in range(len(new)):
                    self.accum[var] += new[var] - self.meta_weights[var]
            else:                                                    # :127-132
                for var in range(len(new)):
                    self.meta_weights[var] += (new[var] - self.meta_weights[var]) * beta
        if "batch" in self.name:                                     # :139-142
            for var in range(len(self.accum)):
                self.meta_weights[var] += self.accum[var] * beta
                self.accum[var] = np.zeros_like(self.accum[var])
        self.model.set_weights(self.meta_weights)                    # :99
        self._fit_target()                                           # :98-102


class OracleMAML(OracleDN):
    """First-order MAML, ``model_zoo/maml.py:35-151,196-242,289-341`` with ``target_domain=-1``: per domain the model is reset
    to theta, trained with its own Adam on the meta-train part, the gradients of the meta-val batches at the adapted weights
    are accumulated (inference-mode forward, ``OracleMLP.gradients(train=False)``) and a SECOND Adam (``meta_learning_rate``)
    applies them to theta -- per domain, or once per epoch for ``batch`` names."""

    def __init__(self, model, data, train_config, batch_size, schedule, name='mlp_meta'):
        from .mlp import AdamState
        OracleDN.__init__(self, model, data, train_config, batch_size, schedule)
        self.name = name
        self.sequence = list(range(len(data['train'])))                      # :56
        mode = train_config.get('average_meta_grad', 'none')
        if mode in ('moving_mean', 'drop'):
            raise NotImplementedError(mode)
        self.scale = None
        if mode == 'mean' and train_config['meta_train_step'] > 0:          # :208-210
            self.scale = np.float32(float(len(self.sequence) * train_config['meta_train_step']))
        self.accum = [np.zeros_like(w) for w in self.meta_weights]           # :202
        self.meta_adam = AdamState(self.meta_weights, lr=train_config['meta_learning_rate'])   # :201
        self.split = self.build_meta_data_split()

    def build_meta_data_split(self):                                         # :289-341
        tc = self.tc
        if tc.get('target_domain', -1) >= 0:
            raise NotImplementedError("target_domain >= 0 under MAML / MLDG (meta-val on the target domain) is not restated")
        out = {}
        for idx, d in self.data['train'].items():
            n = len(d['uid'])
            if tc['meta_split'] == "meta-train/val":
                n_train = int(n * tc['meta_split_ratio'])
                tr, mv = (0, n_train, slice(0, n_train)), (n_train, n, slice(0, n - n_train))
            elif tc['meta_split'] == "meta-train/val-no-exclusive":
                n_train = int(n * tc['meta_split_ratio'])
                tr, mv = (0, n, slice(0, n_train)), (0, n, slice(n_train, n))
            else:
                n_train = n
                tr, mv = (0, n, slice(0, n)), (0, n, slice(0, n))
            n_test = n - n_train if n_train != n else n
            out[idx] = {'train': tr, 'meta': mv, 'train_step': n_steps(n_train, self.bs), 'meta_val_step': n_steps(n_test, self.bs)}
        return out

    def _order(self, idx, window):
        lo, hi, pick = window
        return (self.schedule.batch_order(idx, hi - lo)[pick] + np.int32(lo)).astype(np.int32)

    def meta_train_pass(self, idx, order, steps):
        """`steps` calls of the accumulating K.function (:231 update_add)."""
        d = self.data['train'][idx]
        for s in range(steps):
            sel = order[s * self.bs:(s + 1) * self.bs]
            loss, p, grads = self.model.gradients(d['uid'][sel], d['pid'][sel], idx, d['label'][sel], train=False)
            self.model.auc.update_state(d['label'][sel], p.astype(np.float32))
            for a, g in zip(self.accum, grads):
                a += g

    def meta_parms_update_step(self):                                        # :214 (on the LIVE variables)
        grads = self.accum if self.scale is None else [a / self.scale for a in self.accum]
        self.meta_adam.apply(self.model.weights, grads)

    def _meta_train_step(self):                                              # :235-242
        self.meta_parms_update_step()
        for a in self.accum:
            a[...] = 0
        return self.model.get_weights()

    def _inner_loop(self, idx, sp, order_train, order_meta):
        tc = self.tc
        train_step, meta_val_step = sp['train_step'], sp['meta_val_step']
        if tc['meta_train_step'] > 0:
            train_step, meta_val_step = min(train_step, tc['meta_train_step']), min(meta_val_step, tc['meta_train_step'])
        if train_step > 0:
            train_pass(self.model, self.data['train'][idx], idx, order_train, self.bs, train_step)   # :86-93
        self.meta_train_pass(idx, order_meta, meta_val_step)                # :101-104

    def domain_step(self, idx):
        sp = self.split[idx]
        self.model.auc.reset_states()
        self.model.set_weights(self.meta_weights)                            # :76
        order_train = self._order(idx, sp['train'])                          # :84
        order_meta = self._order(idx, sp['meta'])                            # :85
        self._inner_loop(idx, sp, order_train, order_meta)
        if "batch" in self.name:
            return
        self.model.set_weights(self.meta_weights)                            # :115
        self.meta_weights = self._meta_train_step()                          # :116

    def finish_epoch(self):
        if "batch" in self.name:                                             # :119-121
            self.model.set_weights(self.meta_weights)
            self.meta_weights = self._meta_train_step()
        self.model.set_weights(self.meta_weights)                            # :122

    def train_epoch(self):
        self.sequence = self.schedule.shuffle_sequence(self.sequence)       # :66
        for idx in self.sequence:
            if self.target >= 0 and idx == self.target:                      # :68-69
                continue
            self.domain_step(idx)
        self.finish_epoch()


class OracleMLDG(OracleMAML):
    """``model_zoo/mldg.py:35-155``: accumulate at theta over meta-train, one meta-Adam apply (accumulators kept), accumulate
    over meta-val at the moved weights, reset to theta, apply + clear."""

    def _inner_loop(self, idx, sp, order_train, order_meta):
        tc = self.tc
        train_step, meta_val_step = sp['train_step'], sp['meta_val_step']
        if tc['meta_train_step'] > 0:
            train_step, meta_val_step = min(train_step, tc['meta_train_step']), min(meta_val_step, tc['meta_train_step'])
        self.meta_train_pass(idx, order_train, train_step)                   # :92-96
        self.meta_parms_update_step()                                        # :107-108
        self.meta_train_pass(idx, order_meta, meta_val_step)                 # :112-114


def pcgrad_project(final_grads, current_grads, aux_grads):
    """``PCGrad.PCGrad`` (model_zoo/pcgrad.py:152-160), verbatim semantics (the caller passes the SAME list as final and current)."""
    for var in range(len(final_grads)):
        grad_dot = np.sum(current_grads[var] * aux_grads[var], axis=-1)
        aux_grads[var][grad_dot > 0] -= np.expand_dims(
            grad_dot[grad_dot > 0] / np.linalg.norm(current_grads[var][grad_dot > 0], axis=-1), -1) * current_grads[var][grad_dot > 0]
        final_grads[var] += aux_grads[var]


class OraclePCGrad(OracleMAML):
    """``model_zoo/pcgrad.py:35-150``."""

    def build_meta_data_split(self):                                         # :324-330
        return {idx: {'train_step': n_steps(len(d['uid']), self.bs)} for idx, d in self.data['train'].items()}

    def finish_epoch(self):
        pass

    def domain_step(self, idx):
        tc = self.tc
        self.model.auc.reset_states()
        n = len(self.data['train'][idx]['uid'])
        order = self.schedule.batch_order(idx, n)                        # :79
        train_step = self.split[idx]['train_step']
        if tc['meta_train_step'] > 0:
            train_step = min(train_step, tc['meta_train_step'])
        for a in self.accum:                                             # :86
            a[...] = 0
        self.meta_train_pass(idx, order, train_step)                     # :88-91
        current = [a.copy() for a in self.accum]                         # :103
        final = current                                                  # :104 (alias)
        candidates = list(self.sequence)
        candidates.remove(idx)
        for aux_idx in self.schedule.sample_support(candidates, tc['sample_num']):   # :109-111
            na = len(self.data['train'][aux_idx]['uid'])
            aux_order = self.schedule.batch_order(aux_idx, na)           # :115
            for a in self.accum:                                         # :118
                a[...] = 0
            self.meta_train_pass(aux_idx, aux_order, self.split[aux_idx]['train_step'])   # :120-121
            aux = [a.copy() for a in self.accum]                         # :123
            pcgrad_project(final, current, aux)                          # :124
        for a, f in zip(self.accum, final):                              # :127
            a[...] = f
        self._meta_train_step()                                          # :128


class MetaSubset(object):
    """View of a model whose ``get_weights`` / ``set_weights`` only cover the meta parameters selected by
    ``MAML._get_model_meta_parms`` (``model_zoo/maml.py:153-179``: name-substring lists such as STAR's
    ["emb", "kernel_shared", "bias_shared"]); everything else is delegated, so the remaining tensors keep training
    continuously underneath the DN / DR algebra."""

    def __init__(self, model, index):
        self._model, self._index = model, list(index)

    def get_weights(self):
        return [self._model.weights[i].copy() for i in self._index]

    def set_weights(self, values):
        for i, v in zip(self._index, values):
            self._model.weights[i][...] = v

    def __getattr__(self, item):
        return getattr(self._model, item)


class OracleMAMDR(object):
    """``MAMDR.train`` (``model_zoo/mamdr.py:18-166``) with ``target_domain=-1``."""

    def __init__(self, model, data, train_config, batch_size, schedule, domain_weights, name='mlp_meta_mamdr'):
        self.model, self.data, self.tc, self.bs, self.schedule = model, data, train_config, batch_size, schedule
        self.name = name
        self.meta_weights = model.get_weights()                      # :29
        # :30-33  theta_d^0 = an independent re-initialisation of every layer (injected)
        self.domain_weights = {k: [np.array(w, dtype=model.dtype) for w in v] for k, v in domain_weights.items()}
        target = train_config.get('target_domain', -1)               # specific_base_model.py:33-36: the target domain is skipped
        self.sequence = [k for k in sorted(data['train'].keys()) if not (target >= 0 and k == target)]
        if isinstance(train_config.get('meta_sequence'), list):
            if len(train_config['meta_sequence']) != len(self.sequence):
                raise ValueError("All the domains must be given in the sequence")
            self.sequence = list(train_config['meta_sequence'])
        self.es = EarlyStop(train_config['patience'])
        self.best_shared_weights = None
        self.best_domain_weights = None

    def _update_meta_weight(self, update_vars, merged_weights=None, meta_lr=1):  # :173-180
        new_vars = self.model.get_weights()
        old_vars = merged_weights if merged_weights is not None else update_vars
        for var in range(len(new_vars)):
            update_vars[var] += (new_vars[var] - old_vars[var]) * meta_lr

    def _accumulate_grad(self, accum, old_vars, shared):  # :182-191
        new_vars = self.model.get_weights()
        for var in range(len(accum)):
            if self.tc['merged_method'] == 'plus':
                accum[var] += (new_vars[var] - old_vars[var]) / 1
            elif self.tc['merged_method'] == 'times':
                accum[var] += (new_vars[var] - old_vars[var]) * shared[var] / 1

    def _update_meta_weight_by_grads(self, grads, old_vars):  # :193-196
        for var in range(len(old_vars)):
            old_vars[var] += grads[var] / self.tc['sample_num'] * self.tc['meta_learning_rate']
            grads[var] = np.zeros_like(grads[var])

    def train_epoch(self):
        tc, model = self.tc, self.model
        beta = tc['meta_learning_rate']
        if tc['shuffle_sequence']:                                   # :45-46
            self.sequence = self.schedule.shuffle_sequence(self.sequence)
        seq = self.sequence
        # ---- DN on the shared parameters (:48-57)
        model.set_weights(self.meta_weights)
        for idx in seq:
            d = self.data['train'][idx]
            train_pass(model, d, idx, self.schedule.batch_order(idx, len(d['uid'])), self.bs)
        self._update_meta_weight(self.meta_weights, meta_lr=beta)
        # ---- DR on the specific parameters (:59-108)
        for idx in seq:
            d = self.data['train'][idx]
            cands = list(seq)
            cands.remove(idx)
            aux_idxs = self.schedule.sample_support(cands, tc['sample_num'])
            if tc['add_query_domain']:
                aux_idxs = list(aux_idxs) + [idx]
            merged = merge_weights(self.meta_weights, self.domain_weights[idx], tc['merged_method'])
            accum = [np.zeros_like(w) for w in merged]
            for aux_idx in aux_idxs:
                model.set_weights(merged)                            # :78
                aux_d = self.data['train'][aux_idx]
                train_pass(model, aux_d, aux_idx, self.schedule.batch_order(aux_idx, len(aux_d['uid'])), self.bs)
                train_pass(model, d, idx, self.schedule.batch_order(idx, len(d['uid'])), self.bs,
                           tc.get('domain_regulation_step', 0))
                if 'batch' in self.name:                             # :100-105
                    self._accumulate_grad(accum, merged, self.meta_weights)
                else:
                    self._update_meta_weight(self.domain_weights[idx], merged, meta_lr=beta)
                    merged = merge_weights(self.meta_weights, self.domain_weights[idx], tc['merged_method'])
            if 'batch' in self.name:                                 # :107-108
                self._update_meta_weight_by_grads(accum, self.domain_weights[idx])
            if tc.get('finetune_every_epoch'):                       # :110-143
                merged = merge_weights(self.meta_weights, self.domain_weights[idx], tc['merged_method'])
                model.set_weights(merged)
                model.auc.reset_states()
                train_pass(model, d, idx, self.schedule.batch_order(idx, len(d['uid'])), self.bs)
                new_vars = model.get_weights()                       # :168-171
                for var in range(len(new_vars)):
                    self.domain_weights[idx][var] = new_vars[var] - merged[var]

    def train_epoch_sharded(self, world):
        """The multi-GPU semantics DEFINED in SURVEY.md 8(e) (not in the reference, which is
        single-process), simulated rank by rank on the CPU: DN replicated; the DR query domains are
        LPT-sharded by cost sum_j (S_j + S_i); every rank continues from the post-DN Adam state with its
        own chains in sequence order; afterwards all ranks adopt the Adam state of the rank owning the
        last query domain of the sequence.  Sample orders use the global (sequential) pass ids.
        'plus'/'times' merge, no finetune_every_epoch.  'batch' names shard the (query, support) PAIRS instead
        (`_train_epoch_pair_sharded`)."""
        tc, model = self.tc, self.model
        beta = tc['meta_learning_rate']
        if tc['shuffle_sequence']:
            self.sequence = self.schedule.shuffle_sequence(self.sequence)
        seq = self.sequence
        supports = {}
        for idx in seq:
            cands = list(seq)
            cands.remove(idx)
            aux = self.schedule.sample_support(cands, tc['sample_num'])
            supports[idx] = list(aux) + ([idx] if tc['add_query_domain'] else [])
        S = {i: n_steps(len(self.data['train'][i]['uid']), self.bs) for i in seq}
        cost = {i: sum(S[j] + S[i] for j in supports[i]) for i in seq}
        load, owner = [0] * world, {}
        for i in sorted(cost, key=lambda k: (-cost[k], k)):
            r = min(range(world), key=lambda q: (load[q], q))
            owner[i] = r
            load[r] += cost[i]
        if 'batch' in self.name:
            return self._train_epoch_pair_sharded(world, seq, supports, S)
        n_pass = len(seq) + sum(2 * len(supports[i]) for i in seq)
        pid = self.schedule.reserve(n_pass)
        # ---- DN (replicated)
        model.set_weights(self.meta_weights)
        for idx in seq:
            d = self.data['train'][idx]
            train_pass(model, d, idx, self.schedule.batch_order_at(pid, idx, len(d['uid'])), self.bs)
            pid += 1
        self._update_meta_weight(self.meta_weights, meta_lr=beta)
        ids = {}
        for idx in seq:
            ids[idx] = pid
            pid += 2 * len(supports[idx])
        adam = model.adam
        snap = (deepcopy(adam.m), deepcopy(adam.v), adam.b1pow, adam.b2pow, adam.step)
        final = None
        for r in range(world):
            for mm_, s_ in zip(adam.m, snap[0]):
                mm_[...] = s_
            for vv_, s_ in zip(adam.v, snap[1]):
                vv_[...] = s_
            adam.b1pow, adam.b2pow, adam.step = snap[2], snap[3], snap[4]
            for idx in seq:
                if owner[idx] != r:
                    continue
                d = self.data['train'][idx]
                p = ids[idx]
                merged = merge_weights(self.meta_weights, self.domain_weights[idx], tc['merged_method'])
                for aux_idx in supports[idx]:
                    model.set_weights(merged)
                    aux_d = self.data['train'][aux_idx]
                    train_pass(model, aux_d, aux_idx, self.schedule.batch_order_at(p, aux_idx, len(aux_d['uid'])), self.bs)
                    train_pass(model, d, idx, self.schedule.batch_order_at(p + 1, idx, len(d['uid'])), self.bs)
                    p += 2
                    self._update_meta_weight(self.domain_weights[idx], merged, meta_lr=beta)
                    merged = merge_weights(self.meta_weights, self.domain_weights[idx], tc['merged_method'])
            if r == owner[seq[-1]]:
                final = (deepcopy(adam.m), deepcopy(adam.v), adam.b1pow, adam.b2pow, adam.step)
                final_live = self._live_state()
        for mm_, s_ in zip(adam.m, final[0]):
            mm_[...] = s_
        for vv_, s_ in zip(adam.v, final[1]):
            vv_[...] = s_
        adam.b1pow, adam.b2pow, adam.step = final[2], final[3], final[4]
        self._set_live_state(final_live)
        return owner

    def _train_epoch_pair_sharded(self, world, seq, supports, S):
        """'batch' names (mamdr.py:100-108): every (query i, support j) pair starts from theta (+|*) theta_i, so the pairs are
        LPT-sharded by S_j + S_i; each rank accumulates its pairs' deltas per query domain (:182-191) from the post-DN Adam
        state; the accumulators are summed over the ranks (rank order) and every theta_i takes :193-196; the Adam state and
        the live model of the rank that owns the LAST pair of the sequence are adopted by all."""
        tc, model = self.tc, self.model
        beta = tc['meta_learning_rate']
        cost = {(pos, k): S[j] + S[i] for pos, i in enumerate(seq) for k, j in enumerate(supports[i])}
        load, owner = [0] * world, {}
        for key in sorted(cost, key=lambda q: (-cost[q], q)):
            r = min(range(world), key=lambda q: (load[q], q))
            owner[key] = r
            load[r] += cost[key]
        n_pass = len(seq) + sum(2 * len(supports[i]) for i in seq)
        pid = self.schedule.reserve(n_pass)
        model.set_weights(self.meta_weights)
        for idx in seq:
            d = self.data['train'][idx]
            train_pass(model, d, idx, self.schedule.batch_order_at(pid, idx, len(d['uid'])), self.bs)
            pid += 1
        self._update_meta_weight(self.meta_weights, meta_lr=beta)
        ids = {}
        for idx in seq:
            ids[idx] = pid
            pid += 2 * len(supports[idx])
        adam = model.adam
        snap = (deepcopy(adam.m), deepcopy(adam.v), adam.b1pow, adam.b2pow, adam.step)
        merged = {i: merge_weights(self.meta_weights, self.domain_weights[i], tc['merged_method']) for i in seq}
        accum_r = []
        final = final_live = None
        last_key = (len(seq) - 1, len(supports[seq[-1]]) - 1)
        for r in range(world):
            for mm_, s_ in zip(adam.m, snap[0]):
                mm_[...] = s_
            for vv_, s_ in zip(adam.v, snap[1]):
                vv_[...] = s_
            adam.b1pow, adam.b2pow, adam.step = snap[2], snap[3], snap[4]
            acc = {i: [np.zeros_like(w) for w in merged[i]] for i in seq}
            for pos, idx in enumerate(seq):
                d = self.data['train'][idx]
                for k, aux_idx in enumerate(supports[idx]):
                    if owner[(pos, k)] != r:
                        continue
                    p = ids[idx] + 2 * k
                    model.set_weights(merged[idx])
                    aux_d = self.data['train'][aux_idx]
                    train_pass(model, aux_d, aux_idx, self.schedule.batch_order_at(p, aux_idx, len(aux_d['uid'])), self.bs)
                    train_pass(model, d, idx, self.schedule.batch_order_at(p + 1, idx, len(d['uid'])), self.bs)
                    self._accumulate_grad(acc[idx], merged[idx], self.meta_weights)
            accum_r.append(acc)
            if r == owner[last_key]:
                final = (deepcopy(adam.m), deepcopy(adam.v), adam.b1pow, adam.b2pow, adam.step)
                final_live = self._live_state()
        for idx in seq:
            total = [np.zeros_like(w) for w in merged[idx]]
            for r in range(world):
                for t_, a_ in zip(total, accum_r[r][idx]):
                    t_ += a_
            self._update_meta_weight_by_grads(total, self.domain_weights[idx])
        for mm_, s_ in zip(adam.m, final[0]):
            mm_[...] = s_
        for vv_, s_ in zip(adam.v, final[1]):
            vv_[...] = s_
        adam.b1pow, adam.b2pow, adam.step = final[2], final[3], final[4]
        self._set_live_state(final_live)
        return owner

    _LIVE_EXTRA = ("biased_mean", "biased_var", "pn_steps", "moving_mean", "moving_var")   # OracleStar's non-trainable state

    def _live_state(self):
        """Everything of the live model a rank hands over at the end of a sharded meta-step: ALL variables (also the ones
        outside the meta-parameter subset) and the non-trainable normalisation state."""
        base = getattr(self.model, "_model", self.model)
        return ([w.copy() for w in base.weights], {k: np.copy(getattr(base, k)) for k in self._LIVE_EXTRA if hasattr(base, k)})

    def _set_live_state(self, state):
        base = getattr(self.model, "_model", self.model)
        for w, s_ in zip(base.weights, state[0]):
            w[...] = s_
        for k, v in state[1].items():
            getattr(base, k)[...] = v

    def val_and_test(self, mode):  # specific_base_model.py:64-97
        if mode == 'val':
            shared, specific = deepcopy(self.meta_weights), deepcopy(self.domain_weights)
        elif mode == 'test':
            shared, specific = deepcopy(self.best_shared_weights), deepcopy(self.best_domain_weights)
        else:
            raise ValueError("Mode can be either val or test, not: {}".format(mode))
        domain_loss, domain_auc = {}, {}
        for idx, d in self.data[mode].items():
            self.model.set_weights(merge_weights(shared, specific[idx], self.tc['merged_method']))
            l, a = self.model.evaluate(d['uid'], d['pid'], idx, d['label'], self.bs)
            domain_loss[idx], domain_auc[idx] = float(l), float(a)
        all_loss, all_auc = 0, 0          # plain left-to-right adds like the reference (Python >= 3.12's sum() is compensated)
        for k in domain_loss:
            all_loss += domain_loss[k]
            all_auc += domain_auc[k]
        avg_loss = all_loss / len(domain_loss)
        avg_auc = all_auc / len(domain_auc)
        return avg_loss, avg_auc, domain_loss, domain_auc

    def early_stop_step(self, metric):  # specific_base_model.py:44-62
        def keep():
            self.best_shared_weights = deepcopy(self.meta_weights)
            self.best_domain_weights = deepcopy(self.domain_weights)
        return self.es.step(metric, keep)


def joint_train_epoch(model, data, batch_size, schedule, sequence):
    """``DeepCTR.train`` inner part (``model_zoo/DeepCTR/deepctr.py:70-78``): shuffle the
    domains, then one full ``model.fit`` pass per domain with the single Adam."""
    sequence = schedule.shuffle_sequence(sequence)
    for idx in sequence:
        d = data['train'][idx]
        model.auc.reset_states()   # Keras fit() resets stateful metrics at epoch start
        train_pass(model, d, idx, schedule.batch_order(idx, len(d['uid'])), batch_size)
    return sequence
