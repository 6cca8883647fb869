"""Meta-parameter plumbing shared by the DN / MAMDR wrappers -- mirrors the parts of
``/root/reference/model_zoo/maml.py`` that are on the hot path: ``__getattr__`` delegation (:27-33),
``_get_model_meta_parms`` (:153-179), ``_set_model_meta_parms`` (:181-187), ``_get_meta_weights``
(:189-194) and ``val`` (:343-353) -- plus, as SURVEY.md section 8(f) row f4, the first-order MAML loop itself
(``train`` :35-151, ``_make_meta_train_function`` :196-233, ``_meta_train_step`` :235-242, ``build_meta_data_split`` :289-341)
that MLDG and PCGrad (``mldg.py`` / ``pcgrad.py`` here) re-order.

The reference keeps weights as lists of host numpy arrays and crosses PCIe on every get / set; here a
weight set is one device arena (``MetaWeights``) and get / set are device-side multi-tensor copies.
"""
import ctypes as C
import math

import torch

from .engine import SplitView, _ptr


class MetaWeights(object):
    """A full-arena snapshot that behaves like the reference's list of arrays (ordered like
    ``model.trainable_weights``); ``ranges`` are the (offset, numel) spans of the meta parameters."""

    def __init__(self, flat, layout, ranges):
        self.flat, self.layout, self.ranges = flat, layout, ranges

    def views(self):
        return self.layout.views(self.flat)

    def __len__(self):
        return len(self.layout.names)

    def __getitem__(self, i):
        return self.views()[i]

    def __iter__(self):
        return iter(self.views())

    def clone(self):
        return MetaWeights(self.flat.clone(), self.layout, self.ranges)

    def numpy(self):
        return [v.detach().cpu().numpy() for v in self.views()]


class MAML(object):
    def __init__(self, base_model):
        self.base_model = base_model

    def __getattr__(self, item):
        # Delegate the base model (maml.py:27-33)
        return getattr(self.base_model, item)

    # ---- maml.py:153-179
    def _get_model_meta_parms(self):
        tw = self.model.trainable_weights
        if self.train_config['meta_parms'][0] == "all":
            meta = list(tw)
        elif self.train_config['meta_parms'][0] == "all_hidden":
            meta = [p for p in tw if "emb" not in p.name]
        else:
            meta = []
            for name in self.train_config['meta_parms']:
                found = False
                for p in tw:
                    if name in p.name:
                        meta.append(p)
                        found = True
                if not found:
                    raise ValueError("meta parms: {} not found in the model".format(name))
        self.model_meta_parms = meta
        # merge the selected tensors into maximal contiguous arena spans (padding included: it is zero)
        spans = sorted(set((p.offset, p.numel) for p in meta))
        align = 32
        ranges = []
        for off, n in spans:
            end = (off + n + align - 1) // align * align
            if ranges and ranges[-1][1] == off:
                ranges[-1][1] = end
            else:
                ranges.append([off, end])
        self.meta_ranges = [(a, b - a) for a, b in ranges]

    def _ranges(self, *tensors):
        for off, n in self.meta_ranges:
            yield n, [t[off:off + n] if t is not None else None for t in tensors]

    # ---- maml.py:181-187 / utils/tool.py:36-45
    def _set_model_meta_parms(self, meta_weights):
        m = self.model
        for n, (dst, src) in self._ranges(m.params, meta_weights.flat):
            m.ctx.call("mamdr_copy", _ptr(dst), _ptr(src), n, m.stream)
            m.ctx.launches += 1

    # ---- maml.py:189-194
    def _get_meta_weights(self):
        return MetaWeights(self.model.params.clone(), self.model.layout, self.meta_ranges)

    # ---- maml.py:196-233: the gradient-accumulating function and the second (meta) Adam --------------------------------
    def _make_meta_train_function(self):
        """``accum_grads`` (one arena, only the meta spans are used), the ``meta_optimizer`` slots and the function that
        accumulates the gradients of ``model.total_loss`` w.r.t. the meta parameters (inference-mode forward, see
        ``engine.MLPModel.grads_on_batch``)."""
        tc = self.train_config
        mode = tc.get('average_meta_grad', 'none')
        if mode in ("moving_mean", "drop"):   # :220-229 ([EXT] K.moving_average_update / a Dropout layer's unseeded mask on g)
            raise NotImplementedError("average_meta_grad = %r is not built (no shipped config uses it)" % mode)
        m = self.model
        self.accum_grads = torch.zeros_like(m.params)                  # :202
        self._zeros = torch.zeros_like(m.params)
        self._meta_m, self._meta_v, self._meta_opt_state = m.new_optimizer_slots()   # :201 AdamOptimizer(meta_learning_rate)
        self._grad_scale = None
        if mode == "mean" and tc['meta_train_step'] > 0:               # :208-210  ag / float(n_domain * meta_train_step)
            self._grad_scale = float(self.n_domain * tc['meta_train_step'])
            self._scaled = torch.zeros_like(m.params)
            self._scratch = torch.zeros_like(m.params)
        n = len(self.meta_ranges)
        if n > 16:
            raise NotImplementedError("more than 16 disjoint meta-parameter spans")
        self._meta_begin = (C.c_int64 * n)(*[off for off, _ in self.meta_ranges])
        self._meta_len = (C.c_int64 * n)(*[ln for _, ln in self.meta_ranges])
        self._loss_slot = torch.zeros(1, dtype=torch.float32, device=m.params.device)
        return self.meta_train

    def meta_train(self, data, offset, rows):
        """One call of the K.function: gradients of one mini-batch added into ``accum_grads`` (:231  K.update_add(ag, g))."""
        m = self.model
        m.grads_on_batch(data, offset, rows, self._loss_slot)
        for n, (acc, g, zero) in self._ranges(self.accum_grads, m.grads, self._zeros):
            m.ctx.call("mamdr_axpy_diff", _ptr(acc), _ptr(g), _ptr(zero), 1.0, n, m.stream)   # acc += (g - 0) * 1: exact
            m.ctx.launches += 1

    def meta_train_pass(self, data, steps):
        """``for _ in range(steps): meta_train(next(iterator))`` over the order installed on ``data``."""
        bs = data.batch_size
        for s in range(int(steps)):
            self.meta_train(data, s * bs, min(bs, data.n_data - s * bs))

    def clear_grads(self):                                             # :203
        m = self.model
        for n, (acc, zero) in self._ranges(self.accum_grads, self._zeros):
            m.ctx.call("mamdr_copy", _ptr(acc), _ptr(zero), n, m.stream)
            m.ctx.launches += 1

    def meta_parms_update_step(self):
        """``meta_optimizer.apply_gradients(zip(accum_grads, model_meta_parms))`` (:214): one TF ApplyAdam of the SECOND
        optimizer on the live meta parameters; the accumulators are left as they are."""
        m = self.model
        tc = self.train_config
        grads = self.accum_grads
        if self._grad_scale is not None:
            # scaled = accum / scale, exact: theta_i += accum / sample_num * beta with theta_i = 0, beta = 1 (clears ITS input, a copy)
            for n, (scr, acc, out, zero) in self._ranges(self._scratch, self.accum_grads, self._scaled, self._zeros):
                m.ctx.call("mamdr_copy", _ptr(scr), _ptr(acc), n, m.stream)
                m.ctx.call("mamdr_copy", _ptr(out), _ptr(zero), n, m.stream)
                m.ctx.call("mamdr_dr_apply_accum", _ptr(out), _ptr(scr), self._grad_scale, 1.0, n, m.stream)
                m.ctx.launches += 3
            grads = self._scaled
        m.ctx.call("mamdr_adam_ranges_step", _ptr(m.params), _ptr(self._meta_m), _ptr(self._meta_v), _ptr(grads),
                   self._meta_begin, self._meta_len, len(self.meta_ranges), _ptr(self._meta_opt_state),
                   float(tc['meta_learning_rate']), m.beta1, m.beta2, m.eps, m.stream)
        m.ctx.launches += 1

    # ---- maml.py:235-242
    def _meta_train_step(self):
        self.meta_parms_update_step()
        self.clear_grads()
        return self._get_meta_weights()

    # ---- maml.py:289-341
    def build_meta_data_split(self):
        """Per domain: the meta-train and meta-val datasets as views of the device-resident split.  ``meta-train/val``: the
        first ``int(n * ratio)`` samples / the rest, each reshuffled per pass (take-then-shuffle, exclusive);
        ``meta-train/val-no-exclusive``: the first ``n_train`` / the last ``n - n_train`` of two INDEPENDENT shuffles (the two
        iterators of the reference re-shuffle separately); anything else ("train-train"): both over all samples."""
        tc = self.train_config
        if tc['target_domain'] >= 0:
            raise NotImplementedError("target_domain >= 0 is not used by any shipped config")
        split = {}
        for idx, d in self.dataset.train_dataset.items():
            data, n = d['data'], d['n_data']
            if tc['meta_split'] == "meta-train/val":
                n_train = int(n * tc['meta_split_ratio'])
                train_view, meta_view = SplitView(data, 0, n_train), SplitView(data, n_train, n)
            elif tc['meta_split'] == "meta-train/val-no-exclusive":
                n_train = int(n * tc['meta_split_ratio'])
                train_view, meta_view = SplitView(data, 0, n, slice(0, n_train)), SplitView(data, 0, n, slice(n_train, n))
            else:
                n_train = n
                train_view, meta_view = SplitView(data, 0, n), SplitView(data, 0, n)
            n_test = n - n_train if n_train != n else n
            bs = float(self.dataset.batch_size)
            split[idx] = {"train_iter": train_view, "train_step": int(math.ceil(n_train / bs)),
                          "meta_iter": meta_view, "meta_val_step": int(math.ceil(n_test / bs))}
        return split

    def _init_iter(self, view):
        """``K.get_session().run(iterator.initializer)``: the next pass's shuffle, from the injected schedule."""
        view.set_order(self.schedule.batch_order(view.domain, view.hi - view.lo))

    # ---- maml.py:35-151 (first-order MAML: Adam inner steps on meta-train, gradients of meta-val at the adapted weights,
    # applied to theta by the meta Adam -- per domain, or summed over the epoch for `batch` names)
    def prepare(self):
        self._get_model_meta_parms()                               # :48
        self.meta_weights = self._get_meta_weights()               # :50
        self._make_meta_train_function()                           # :51
        self.model.reset_optimizer()                               # :53 global_variables_initializer (both optimizers' slots)
        self.meta_data_split = self.build_meta_data_split()        # :55
        self.train_sequence = list(range(self.n_domain))           # :56

    def _inner_loop(self, idx, d):
        """:86-104 -- `train_step` x model.train_on_batch(train_iter), then `meta_val_step` x meta_train(meta_iter)."""
        tc = self.train_config
        train_step, meta_val_step = d['train_step'], d['meta_val_step']
        if tc['meta_train_step'] > 0:
            train_step = min(train_step, tc['meta_train_step'])
            meta_val_step = min(meta_val_step, tc['meta_train_step'])
        self.run_view_train_pass(d['train_iter'], train_step)
        self.meta_train_pass(d['meta_iter'], meta_val_step)

    def run_view_train_pass(self, view, steps):
        """`steps` x ``model.train_on_batch`` over a split view (the model's own Adam)."""
        if steps > 0:
            self.last_pass_losses = self.model.fit_pass(view, steps)

    def domain_step(self, idx):
        """The body of the per-domain loop (:67-116)."""
        d = self.meta_data_split[idx]
        for metric in self.model.stateful_metric_functions:    # :72-73
            metric.reset_states()
        self._set_model_meta_parms(self.meta_weights)          # :76
        self._init_iter(d['train_iter'])                       # :84
        self._init_iter(d['meta_iter'])                        # :85
        self._inner_loop(idx, d)
        if "batch" in self.model_config['name']:               # :112-113
            return
        self._set_model_meta_parms(self.meta_weights)          # :115
        self.meta_weights = self._meta_train_step()            # :116

    def train_epoch(self, epoch=0):
        self.train_sequence = self.schedule.shuffle_sequence(self.train_sequence)   # :66
        for idx in self.train_sequence:
            if self.train_config['target_domain'] >= 0 and idx == self.train_config['target_domain']:   # :68-69 (PCGrad; MAML / MLDG
                continue                                                                            #  refuse the knob in their split)
            self.domain_step(idx)
        self.finish_epoch()

    def finish_epoch(self):
        if "batch" in self.model_config['name']:               # :119-121
            self._set_model_meta_parms(self.meta_weights)
            self.meta_weights = self._meta_train_step()
        self._set_model_meta_parms(self.meta_weights)          # :122

    def train(self):
        self.log("Start {} training on model: {}".format(type(self).__name__, self.model_config['name']))
        self.prepare()
        for epoch in range(self.train_config['epoch']):
            self.log("Epoch: {}".format(epoch), "-" * 30)
            self.train_epoch(epoch)
            if epoch % self.train_config['val_every_step'] == 0:   # :130-144
                val_avg_loss, val_avg_auc, val_domain_loss, val_domain_auc = self.val()
                val_metric = val_domain_auc[self.train_config['target_domain']] \
                    if self.train_config['target_domain'] >= 0 else val_avg_auc     # :137-138
                if self.early_stop_step(val_metric):
                    break
                self.log("Test Result: ")
                self.val_and_test("test")

    # ---- maml.py:343-353
    def val(self):
        self.log("Val Result: ")
        if self.train_config['meta_finetune_step'] > 0:
            return self.meta_finetune_val()      # finetune a few epochs per domain to evaluate the meta parameters
        return self.val_and_test("val")

    # ---- maml.py:244-287
    def meta_finetune_val(self):
        """Per domain: restart from the current weights, ``meta_finetune_step`` Keras epochs (one shuffled pass each, the model's
        own optimizer -- whose slots are NOT part of get_weights / set_weights and keep advancing, as in the reference), evaluate
        on the domain's validation split; the weights are restored at the end."""
        train_dataset, val_dataset = self.dataset.train_dataset, self.dataset.val_dataset
        train_epoch = self.train_config['meta_finetune_step']
        domain_loss, domain_auc = {}, {}
        all_loss, all_auc = 0, 0
        weights = self.model.get_weights()                       # :258
        for domain_idx, train_d in train_dataset.items():
            self.model.set_weights(weights)                      # :260
            self.log("Finetune on domain: {}".format(domain_idx))
            for epoch in range(train_epoch):                     # :264-267 model.fit(..., epochs=epoch + 1, initial_epoch=epoch)
                self.run_train_pass(domain_idx)
            p_loss, p_auc = self.model.evaluate(val_dataset[domain_idx]['data'], steps=val_dataset[domain_idx]['n_step'])   # :269-271
            domain_loss[domain_idx], domain_auc[domain_idx] = p_loss, p_auc
            all_loss += p_loss
            all_auc += p_auc
        self.model.set_weights(weights)                          # :279
        avg_loss = all_loss / len(domain_loss)
        avg_auc = all_auc / len(domain_auc)
        self.log("Loss: ", domain_loss)
        self.log("AUC: ", domain_auc)
        self.log("Overall val Loss: {}, AUC: {}".format(avg_loss, avg_auc))
        return avg_loss, avg_auc, domain_loss, domain_auc

    # ---- resume (what the reference lacks -- theta, theta_d[], the Adam state and the schedule are
    # only in RAM there, base_model.py:177-181 saves the live Keras weights alone) ------------------------------------
    _STATE_TENSORS = ("params", "m", "v", "opt_state", "pn_state")
    _STATE_SEQS = ("meta_sequence", "train_sequence")

    def save_state(self, path, epoch=0):
        """Everything a meta-training run needs to continue bit-identically after `epoch`: the live model arena, the Adam
        slots and beta powers, theta, every theta_d and best snapshot, the domain sequence, the schedule's RNG state and the
        early-stop bookkeeping."""
        m = self.model
        base = self.base_model
        sched = base.schedule
        nxt = getattr(self, "_next_plan", None)
        rng_state, pass_ctr = (sched._rng.getstate(), sched._pass) if nxt is None else nxt["rng_before"]
        seqs = {k: list(getattr(self, k)) for k in self._STATE_SEQS if hasattr(self, k)}
        blob = {"epoch": int(epoch), "names": m.layout.names,
                "model": {k: getattr(m, k).detach().cpu() for k in self._STATE_TENSORS if getattr(m, k, None) is not None},
                "meta_weights": self.meta_weights.flat.cpu(),
                "schedule": {"seed": sched.seed, "rng": rng_state, "pass": pass_ctr},   # from before a look-ahead plan, if any
                "early_stop": {"counter": base.counter, "best_metric": base.best_metric, "early_stop": base.early_stop}}
        blob.update(seqs)
        for k in ("domain_weights", "best_domain_weights"):
            if getattr(self, k, None):
                blob[k] = {d: w.flat.cpu() for d, w in getattr(self, k).items()}
        if getattr(self, "best_shared_weights", None) is not None:
            blob["best_shared_weights"] = self.best_shared_weights.flat.cpu()
        if getattr(self, "accum_grads", None) is not None:
            blob["accum_grads"] = self.accum_grads.cpu()
        if getattr(self, "_meta_m", None) is not None:     # the second (meta) Adam of MAML / MLDG / PCGrad
            blob["meta_optimizer"] = {"m": self._meta_m.cpu(), "v": self._meta_v.cpu(), "state": self._meta_opt_state.cpu()}
        torch.save(blob, path)
        return path

    def load_state(self, path):
        """Restore a `save_state` blob into a freshly built (and `prepare`d, where the wrapper has one) wrapper.  Returns the
        epoch the state was saved after."""
        blob = torch.load(path, map_location="cpu", weights_only=True)   # tensors, ints, strings, lists / tuples only
        m = self.model
        base = self.base_model
        if blob["names"] != m.layout.names:
            raise ValueError("state layout mismatch")
        if not hasattr(self, "model_meta_parms"):
            self._get_model_meta_parms()
        for k, v in blob["model"].items():
            getattr(m, k).copy_(v)
        if getattr(self, "meta_weights", None) is None:
            self.meta_weights = self._get_meta_weights()
        self.meta_weights.flat.copy_(blob["meta_weights"])
        for k in self._STATE_SEQS:
            if k in blob:
                setattr(self, k, list(blob[k]))
        for k in ("domain_weights", "best_domain_weights"):
            if k in blob:
                cur = getattr(self, k, None) or {}
                for d, w in blob[k].items():
                    if d not in cur:
                        cur[d] = self.meta_weights.clone()
                    cur[d].flat.copy_(w)
                setattr(self, k, cur)
        if "best_shared_weights" in blob:
            self.best_shared_weights = self.meta_weights.clone()
            self.best_shared_weights.flat.copy_(blob["best_shared_weights"])
        if "accum_grads" in blob and getattr(self, "accum_grads", None) is not None:
            self.accum_grads.copy_(blob["accum_grads"])
        if "meta_optimizer" in blob and getattr(self, "_meta_m", None) is not None:
            self._meta_m.copy_(blob["meta_optimizer"]["m"])
            self._meta_v.copy_(blob["meta_optimizer"]["v"])
            self._meta_opt_state.copy_(blob["meta_optimizer"]["state"])
        sched = base.schedule
        sched.seed, sched._pass = blob["schedule"]["seed"], blob["schedule"]["pass"]
        sched._rng.setstate(blob["schedule"]["rng"])
        self._next_plan = None
        es = blob["early_stop"]
        base.counter, base.best_metric, base.early_stop = es["counter"], es["best_metric"], es["early_stop"]
        return blob["epoch"]
