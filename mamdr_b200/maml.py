"""Meta-parameter plumbing shared by the DN / MAMDR wrappers -- mirrors the parts of
``/root/reference/model_zoo/maml.py`` that are on the hot path: ``__getattr__`` delegation (:27-33),
``_get_model_meta_parms`` (:153-179), ``_set_model_meta_parms`` (:181-187), ``_get_meta_weights``
(:189-194) and ``val`` (:343-353).  (The MAML training loop itself is out of scope, SURVEY.md 2.1 #3.)

The reference keeps weights as lists of host numpy arrays and crosses PCIe on every get / set; here a
weight set is one device arena (``MetaWeights``) and get / set are device-side multi-tensor copies.
"""
import ctypes as C

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
