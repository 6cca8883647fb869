"""Meta-parameter plumbing shared by the DN / MAMDR wrappers -- mirrors the parts of
``/root/reference/model_zoo/maml.py`` that are on the hot path: ``__getattr__`` delegation (:27-33),
``_get_model_meta_parms`` (:153-179), ``_set_model_meta_parms`` (:181-187), ``_get_meta_weights``
(:189-194) and ``val`` (:343-353).  (The MAML training loop itself is out of scope, SURVEY.md 2.1 #3.)

The reference keeps weights as lists of host numpy arrays and crosses PCIe on every get / set; here a
weight set is one device arena (``MetaWeights``) and get / set are device-side multi-tensor copies.
"""
import ctypes as C

import torch

from .engine import _ptr


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

    # ---- maml.py:343-353
    def val(self):
        if self.train_config['meta_finetune_step'] > 0:
            raise NotImplementedError("meta_finetune_step > 0 is not used by any shipped DN / MAMDR config")
        self.log("Val Result: ")
        return self.val_and_test("val")

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
        sched = base.schedule
        sched.seed, sched._pass = blob["schedule"]["seed"], blob["schedule"]["pass"]
        sched._rng.setstate(blob["schedule"]["rng"])
        self._next_plan = None
        es = blob["early_stop"]
        base.counter, base.best_metric, base.early_stop = es["counter"], es["best_metric"], es["early_stop"]
        return blob["epoch"]
