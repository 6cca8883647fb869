"""Generates tests/golden/reference_meta_v1.npz by EXECUTING THE REFERENCE'S OWN CODE (read from /root/reference, never copied):
the host-side numpy algebra of the meta-training hot path -- SURVEY.md section 8 rows a2-a5, a8, the early-stop bookkeeping
of row a18 and Reptile's update rules (row f4).

    python tests/golden/make_reference_golden.py          # only where /root/reference exists (not on the GPU box)

TensorFlow 1.12 / deepctr 0.9.0 cannot be installed here, but these methods are plain Python + numpy: the modules are imported
with `tensorflow*` / `deepctr*` replaced by inert stubs (nothing of TF is ever called on this path) and the UNBOUND methods are
run on a minimal stand-in for `self` (`_get_meta_weights`, `train_config`).  The vectors pin the oracle (tests/test_reference_golden.py)
and, through it and directly (tests/test_gpu_golden.py), the CUDA sweeps to the reference's arithmetic bit for bit.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"


class _AnyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any


class _Any(metaclass=_AnyMeta):
    """Inert stand-in for every TF / deepctr symbol: subclassable, callable, attribute access returns itself."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    PREFIXES = ("tensorflow", "deepctr")

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in self.PREFIXES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def import_reference():
    sys.meta_path.insert(0, _StubFinder())
    sys.path.insert(0, REFERENCE)
    import model_zoo.base_model as base_model
    import model_zoo.domain_negotiation as domain_negotiation
    import model_zoo.mamdr as mamdr
    import model_zoo.reptile as reptile
    import model_zoo.specific_base_model as specific_base_model
    return base_model, domain_negotiation, mamdr, reptile, specific_base_model


SHAPES = [(7, 5), (5,), (3, 4), (1,)]


def weights(rng, scale=1.0):
    return [(rng.standard_normal(s) * scale).astype(np.float32) for s in SHAPES]


def flat(ws):
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in ws])


class _Self(object):
    """What the reference's methods touch on `self`."""

    def __init__(self, new_vars, train_config):
        self._new = new_vars
        self.train_config = train_config

    def _get_meta_weights(self):
        return [w.copy() for w in self._new]


def make():
    base_model, dn, mamdr, reptile, sbm = import_reference()
    rng = np.random.default_rng(77)
    g = {}
    tc = {"meta_learning_rate": 0.1, "domain_meta_learning_rate": 0.1, "sample_num": 5, "merged_method": "plus", "patience": 3}
    theta, theta_i, model = weights(rng), weights(rng, 0.1), weights(rng)
    g["theta"], g["theta_i"], g["model"] = flat(theta), flat(theta_i), flat(model)
    # a8  DomainNegotiation._update_meta_weight (domain_negotiation.py:118-123)
    t = [w.copy() for w in theta]
    dn.DomainNegotiation._update_meta_weight(_Self(model, tc), t)
    g["dn_update"] = flat(t)
    # a5  SpecificBase._merge_weights (specific_base_model.py:164-172)
    for method in ("plus", "times"):
        s = _Self(model, dict(tc, merged_method=method))
        merged = sbm.SpecificBase._merge_weights(s, theta, theta_i)
        g["merge_" + method] = flat(merged)
        # a2  MAMDR._update_meta_weight (mamdr.py:173-180): DR form (explicit merged weights) and DN form
        ti = [w.copy() for w in theta_i]
        mamdr.MAMDR._update_meta_weight(s, ti, merged, tc["domain_meta_learning_rate"])
        g["dr_update_" + method] = flat(ti)
        # a3  MAMDR._accumulate_grad / _update_meta_weight_by_grads (mamdr.py:182-196)
        acc = weights(np.random.default_rng(5), 0.01)
        g["accum0"] = flat(acc)
        mamdr.MAMDR._accumulate_grad(s, acc, merged, theta)
        g["accumulate_" + method] = flat(acc)
        ti = [w.copy() for w in theta_i]
        mamdr.MAMDR._update_meta_weight_by_grads(s, acc, ti)
        g["apply_accum_" + method] = flat(ti)
        assert not flat(acc).any()
        # a4  MAMDR._update_domain_weights (mamdr.py:168-171)
        dw = [w.copy() for w in theta_i]
        mamdr.MAMDR._update_domain_weights(s, dw, merged)
        g["update_domain_weights_" + method] = flat(dw)
    t = [w.copy() for w in theta]
    mamdr.MAMDR._update_meta_weight(_Self(model, tc), t, None, tc["meta_learning_rate"])
    g["mamdr_dn_form"] = flat(t)
    # f4  Reptile (reptile.py:127-142)
    t = [w.copy() for w in theta]
    reptile.Reptile._update_meta_weight(_Self(model, tc), t)
    g["reptile_update"] = flat(t)
    acc = [np.zeros_like(w) for w in theta]
    for k in range(3):
        reptile.Reptile._accumulate_grad(_Self([w + np.float32(0.25 * k) for w in model], tc), acc, theta)
    g["reptile_accum3"] = flat(acc)
    t = [w.copy() for w in theta]
    reptile.Reptile._update_meta_weight_by_grads(_Self(model, tc), acc, t)
    g["reptile_apply"] = flat(t)
    # a18  BaseModel.early_stop_step (base_model.py:208-224) and _weighted_auc (:157-175)
    metrics = np.array([0.5, 0.6, 0.6, 0.55, 0.7, 0.69, 0.70, 0.68, 0.67], dtype=np.float64)
    s = types.SimpleNamespace(train_config=tc, checkpoint_path="/dev/null", saved=[])
    s.save_model = lambda path: s.saved.append(len(s.saved))
    base_model.BaseModel._build_early_stop(s)
    trace = []
    for mval in metrics:
        stop = base_model.BaseModel.early_stop_step(s, float(mval))
        trace.append([s.counter, float(s.best_metric), float(bool(stop)), float(len(s.saved))])
    g["early_stop_metrics"], g["early_stop_trace"] = metrics, np.array(trace, dtype=np.float64)
    info = {0: {"n_train": 10, "n_val": 4, "n_test": 6}, 1: {"n_train": 30, "n_val": 9, "n_test": 1}, 2: {"n_train": 5, "n_val": 2, "n_test": 3}}
    s.dataset = types.SimpleNamespace(dataset_info=info)
    auc = {0: 0.61, 1: 0.72, 2: 0.55}
    g["weighted_auc"] = np.array([base_model.BaseModel._weighted_auc(s, mode, auc) for mode in ("train", "val", "test")], dtype=np.float64)
    return g


if __name__ == "__main__":
    if not os.path.isdir(REFERENCE):
        raise SystemExit("the reference tree is not available here; the committed .npz is the artefact")
    out = os.path.join(HERE, "reference_meta_v1.npz")
    np.savez_compressed(out, **make())
    print(out, os.path.getsize(out), "bytes")
