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
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
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


# ---- the reference's TRAINING LOOPS executed over a toy Keras stand-in ------------------------------------------------------
# MAMDR.train (mamdr.py:18-166), DomainNegotiation.train (domain_negotiation.py:18-116) and Reptile.train (reptile.py:17-125) are
# host-side Python: with a stand-in for the compiled Keras model (a train step is a fixed affine map that depends on the domain,
# so the ORDER of every step matters) and Python's `random` seeded, running them records the exact control flow -- which domain
# trains when, from which weights, which deltas are applied, which snapshots are kept.  The oracle replays the same loops with
# `Schedule(seed)` (the same `random.Random` draws in the same order) and must land on the same bits (tests/test_reference_golden.py).
N_STEP = {0: 3, 1: 2, 2: 2, 3: 1}
LOOP_SEED = 321


def toy_step(weights, domain):
    for w in weights:
        w[...] = w * np.float32(0.9) + np.float32(0.01 * (domain + 1))


def toy_eval(weights, domain):
    tot = float(sum(np.asarray(w, dtype=np.float64).sum() for w in weights))
    import math
    return tot, 0.5 + 0.4 * math.tanh(tot * (domain + 1) * 0.01)


def toy_init(k):
    """Weights after the k-th (re-)initialisation of the layers: k = 0 is the model build."""
    rng = np.random.default_rng(1000 + k)
    return [rng.standard_normal((3, 2)).astype(np.float32), rng.standard_normal((4,)).astype(np.float32)]


class _ToyData(object):
    def __init__(self, domain):
        self.domain = domain

    def make_initializable_iterator(self):
        return types.SimpleNamespace(domain=self.domain, initializer=None)

    def repeat(self):
        return self


class _ToyKeras(object):
    def __init__(self):
        self.weights = toy_init(0)
        self.steps = []                    # the domain of every train step, in execution order
        self.stateful_metric_functions = []
        self.layers = []

    def fit(self, it, steps_per_epoch=1, callbacks=None, validation_data=None, validation_steps=None, epochs=1, initial_epoch=0, **kw):
        """One Keras epoch of the training loops; with `callbacks` (the finetune stage) the Keras epoch loop with validation."""
        if not callbacks:
            for _ in range(steps_per_epoch):
                self.train_on_batch(it)
            return
        self.stop_training = False
        for cb in callbacks:
            cb.model = self
            cb.on_train_begin()
        for epoch in range(initial_epoch, epochs):
            for _ in range(steps_per_epoch):
                self.train_on_batch(it)
            val_loss, val_auc = toy_eval(self.weights, validation_data.domain)
            for cb in callbacks:
                cb.on_epoch_end(epoch, {"val_loss": val_loss, "val_AUC": val_auc})
            if self.stop_training:
                break

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, ws):
        for a, b in zip(self.weights, ws):
            a[...] = b

    def compile(self, **kw):
        self.compiles = getattr(self, "compiles", 0) + 1

    def save_weights(self, path, overwrite=True):
        self.files = getattr(self, "files", {})
        self.files[os.path.basename(path)] = [w.copy() for w in self.weights]

    def load_weights(self, path):
        self.set_weights(self.files[os.path.basename(path)])

    def train_on_batch(self, it):
        toy_step(self.weights, it.domain)
        self.steps.append(it.domain)
        return 0.0, 0.5

    def evaluate(self, data, steps=1, verbose=0):
        return toy_eval(self.weights, data.domain)


def _toy_wrapper(cls, base_model_mod, train_config, name, meta_idx=None):
    import collections
    model = _ToyKeras()
    mk = lambda: collections.OrderedDict((d, {"data": _ToyData(d), "n_step": N_STEP[d]}) for d in sorted(N_STEP))   # noqa: E731
    info = {d: {"n_train": 4 * N_STEP[d], "n_val": 2 + d, "n_test": 3 + d} for d in N_STEP}
    dataset = types.SimpleNamespace(train_dataset=mk(), val_dataset=mk(), test_dataset=mk(), dataset_info=info)
    base = types.SimpleNamespace(train_config=train_config, model_config={"name": name}, dataset=dataset, model=model,
                                 n_domain=len(N_STEP), checkpoint_path="/tmp/unused/model.h5", saved=None)
    B = base_model_mod.BaseModel
    for meth in ("val_and_test", "early_stop_step", "_weighted_auc", "_format_print_domain_metric", "_build_early_stop"):
        setattr(base, meth, types.MethodType(getattr(B, meth), base))
    base.save_model = lambda path: setattr(base, "saved", [w.copy() for w in model.weights])
    base.load_model = lambda path: [a.__setitem__(Ellipsis, b) for a, b in zip(model.weights, base.saved)]
    base._build_early_stop()
    obj = cls.__new__(cls)
    obj.base_model = base
    inits = [0]
    sel = list(range(len(model.weights))) if meta_idx is None else list(meta_idx)   # the meta parameters (maml.py:153-179)
    obj._get_model_meta_parms = lambda: setattr(obj, "model_meta_parms", [types.SimpleNamespace(shape=model.weights[i].shape) for i in sel])
    obj._get_meta_weights = lambda: [model.weights[i].copy() for i in sel]            # K.batch_get_value(model_meta_parms)
    base.last_set = None

    def set_parms(ws):      # MAML._set_model_meta_parms (maml.py:181-187); the last call of a run carries the final theta
        base.last_set = [np.array(w, dtype=np.float32) for w in ws]
        for i, b in zip(sel, ws):
            model.weights[i][...] = b
    obj._set_model_meta_parms = set_parms

    def init_layer(m):
        inits[0] += 1
        for a, b in zip(model.weights, toy_init(inits[0])):
            a[...] = b
    obj.init_layer = init_layer
    return obj, model, base


LOOP_TC = {"epoch": 2, "shuffle_sequence": True, "sample_num": 2, "add_query_domain": True, "merged_method": "plus",
           "meta_learning_rate": 0.1, "domain_meta_learning_rate": 0.1, "finetune_every_epoch": False, "domain_regulation_step": 0,
           "meta_train_step": 0, "meta_finetune_step": 0, "val_every_step": 1, "target_domain": -1, "patience": 3, "histogram_freq": 0,
           "meta_sequence": "random"}
# meta parameters = a SUBSET of the variables (config #4: meta_parms = ["emb", "kernel_shared", "bias_shared"]): the other
# variables are re-initialised by every init_layer call, never reloaded from theta, and evolve freely through all passes
SUBSET_CASES = [("mamdr", "mlp_meta_mamdr", "plus"), ("dn", "mlp_meta_domain_negotiation", "plus"), ("reptile", "mlp_meta_reptile", "plus")]
SUBSET_META_IDX = [0]
# the config knobs of the loops: each variant overrides LOOP_TC
VARIANT_CASES = [("mamdr", "mlp_meta_mamdr", {"finetune_every_epoch": True}),
                 ("mamdr", "mlp_meta_mamdr", {"domain_regulation_step": 1}),
                 ("mamdr", "mlp_meta_mamdr", {"add_query_domain": False, "sample_num": 3}),
                 ("mamdr", "mlp_meta_mamdr_batch", {"add_query_domain": False, "merged_method": "times"}),
                 ("mamdr", "mlp_meta_mamdr", {"shuffle_sequence": False, "meta_sequence": [2, 0, 3, 1]}),
                 ("dn", "mlp_meta_domain_negotiation", {"meta_train_step": 1}),
                 ("dn", "mlp_meta_domain_negotiation", {"shuffle_sequence": False, "meta_sequence": [2, 0, 3, 1]}),
                 ("dn", "mlp_meta_domain_negotiation", {"val_every_step": 2, "epoch": 3}),
                 ("reptile", "mlp_meta_reptile", {"meta_train_step": 2}),
                 ("reptile", "mlp_meta_reptile_batch", {"epoch": 3, "meta_learning_rate": 0.5}),
                 # target_domain >= 0: the target is left out of the meta sequence, trained after the outer update (DN, Reptile; one
                 # step per inner domain in Reptile), only evaluated by MAMDR; the early-stop metric is ITS validation AUC
                 ("dn", "mlp_meta_domain_negotiation", {"target_domain": 2}),
                 ("dn", "mlp_meta_domain_negotiation", {"target_domain": 0, "meta_train_step": 1, "epoch": 3}),
                 ("mamdr", "mlp_meta_mamdr", {"target_domain": 3}),
                 ("reptile", "mlp_meta_reptile", {"target_domain": 0}),
                 ("reptile", "mlp_meta_reptile_batch", {"target_domain": 2, "meta_train_step": 1})]
LOOP_CASES = [("mamdr", "mlp_meta_mamdr", "plus"), ("mamdr", "mlp_meta_mamdr_batch", "plus"), ("mamdr", "mlp_meta_mamdr", "times"),
              ("dn", "mlp_meta_domain_negotiation", "plus"), ("reptile", "mlp_meta_reptile", "plus"),
              ("reptile", "mlp_meta_reptile_batch", "plus")]


def make_loops():
    import contextlib
    import io
    import random
    base_model, dn, mamdr, reptile, sbm = import_reference()
    import numpy
    g = {}
    # mamdr.py:75 / reptile.py:31 size the `batch` accumulators with K.int_shape / K.dtype of the Keras variables (TF): stand-ins only
    for mod_K in (mamdr.K, reptile.K):
        mod_K.int_shape = staticmethod(lambda p: p.shape)
        mod_K.dtype = staticmethod(lambda p: "float32")
    runs = [(k, n, m, None, {}, "%s|%s|" % (n, m)) for k, n, m in LOOP_CASES]
    runs += [(k, n, m, SUBSET_META_IDX, {}, "%s|%s|subset|" % (n, m)) for k, n, m in SUBSET_CASES]
    runs += [(k, n, over.get("merged_method", "plus"), None, over, "variant%d|" % i) for i, (k, n, over) in enumerate(VARIANT_CASES)]
    for kind, name, method, meta_idx, over, key in runs:
        cls = {"mamdr": mamdr.MAMDR, "dn": dn.DomainNegotiation, "reptile": reptile.Reptile}[kind]
        tc = dict(LOOP_TC, merged_method=method)
        tc.update(over)
        obj, model, base = _toy_wrapper(cls, base_model, tc, name, meta_idx)
        random.seed(LOOP_SEED)
        with contextlib.redirect_stdout(io.StringIO()):
            obj.train()
        g[key + "live"] = flat_any(model.weights)               # the live model when train() returns (non-meta variables included)
        g[key + "steps"] = numpy.array(model.steps, dtype=numpy.int32)
        if kind == "mamdr":
            g[key + "theta"] = flat_any(obj.meta_weights)
            for d in sorted(N_STEP):
                g[key + "theta_%d" % d] = flat_any(obj.domain_weights[d])
                g[key + "best_theta_%d" % d] = flat_any(obj.best_domain_weights[d])
            g[key + "best_theta"] = flat_any(obj.best_shared_weights)
            g[key + "es"] = numpy.array([obj.counter, obj.best_metric], dtype=numpy.float64)
        else:
            g[key + "theta"] = flat_any(base.last_set)      # theta is a local of these loops; their last _set_model_meta_parms
            #                                                  (domain_negotiation.py:88, reptile.py:99) loads it into the model
            g[key + "best"] = flat_any(base.saved)
            g[key + "es"] = numpy.array([base.counter, base.best_metric], dtype=numpy.float64)
    return g


# ---- the reference's CLI dispatch (run.py:22-87) and meta-parameter selection (maml.py:153-179) -------------------------------
DISPATCH_NAMES = ["mlp", "mlp_meta_mamdr_finetune", "mlp_meta_mamdr_batch", "mlp_meta_domain_negotiation_finetune",
                  "mlp_meta_domain_negotiation", "star", "star_meta_mamdr_finetune", "mmoe", "ple", "shared_bottom",
                  "mmoe_meta_domain_negotiation", "ple_meta_domain_negotiation", "mlp_meta_reptile_finetune", "mlp_meta_reptile_batch",
                  "mlp_separate", "wdl", "mlp_meta_mldg", "mlp_meta", "mlp_pcgrad", "mlp_uncertainty_weight", "nothing"]
VAR_NAMES = {
    "mlp_frozen": ["sparse_emb_domain_emb/embeddings:0", "dnn/kernel0:0", "dnn/kernel1:0", "dnn/kernel2:0", "dnn/bias0:0", "dnn/bias1:0",
                   "dnn/bias2:0", "dense/kernel:0", "prediction_layer/global_bias:0"],
    "star": ["domain_emb/embeddings:0", "partitioned_norm/gamma_specific:0", "partitioned_norm/beta_specific:0",
             "partitioned_norm/gamma_shared:0", "partitioned_norm/beta_shared:0", "star_fcn/kernel_specific:0", "star_fcn/bias_specific:0",
             "star_fcn/kernel_shared:0", "star_fcn/bias_shared:0", "star_fcn_1/kernel_specific:0", "star_fcn_1/bias_specific:0",
             "star_fcn_1/kernel_shared:0", "star_fcn_1/bias_shared:0", "dense/kernel:0", "dense/bias:0"]}
META_PARMS_CASES = [("mlp_frozen", ["all"]), ("mlp_frozen", ["all_hidden"]), ("mlp_frozen", ["dnn"]), ("mlp_frozen", ["bias", "dense"]),
                    ("mlp_frozen", ["emb", "nope"]), ("star", ["emb", "kernel_shared", "bias_shared"]), ("star", ["all_hidden"]),
                    ("star", ["specific"])]


class _Recorder(object):
    """Stand-in for every class run.py instantiates: records construction and the calls main() makes on the final object."""
    trace = None

    def __init__(self, *args, **kwargs):
        type(self).trace.append(type(self).__name__)
        self.train_config = {"meta_finetune_step": 0}
        self.checkpoint_path = "ckpt"

    def train(self):
        self.trace.append("train")

    def val_and_test(self, mode):
        self.trace.append("val_and_test:" + mode)
        return 0.0, 0.5, {}, {}

    def separate_train_val_test(self, init_parms=True):
        self.trace.append("separate_train_val_test:init_parms=%s" % init_parms)
        return 0.0, 0.5, {}, {}

    def load_model(self, path):
        self.trace.append("load_model")

    def save_result(self, *a):
        self.trace.append("save_result")


def recorder_classes(names, trace):
    return {n: type(n, (_Recorder,), {"trace": trace}) for n in names}


def make_dispatch():
    import contextlib
    import io
    import_reference()
    import run as ref_run
    assert ref_run.__file__.startswith(REFERENCE)
    out = {"dispatch": {}, "meta_parms": []}
    classes = ["MultiDomainDataset", "Star", "DeepCTR", "DeepMTLCTR", "UncertaintyWeight", "PCGrad", "DomainNegotiation", "MAMDR", "Reptile",
               "MLDG", "MAML"]
    for name in DISPATCH_NAMES:
        trace = []
        for k, v in recorder_classes(classes, trace).items():
            setattr(ref_run, k, v)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                ref_run.main({"model": {"name": name}, "dataset": {"seed": 1}})
        except Exception as e:        # run.py:47 only prints for an unknown model, then crashes on None
            trace.append("raises:" + type(e).__name__)
        out["dispatch"][name] = trace
    import model_zoo.maml as maml
    maml.tool.SetVarOp = lambda parms: None          # utils/tool.py:16-45 builds TF assign ops; not part of the selection logic
    for model_kind, meta_parms in META_PARMS_CASES:
        tw = [types.SimpleNamespace(name=n) for n in VAR_NAMES[model_kind]]
        s = types.SimpleNamespace(train_config={"meta_parms": meta_parms}, model=types.SimpleNamespace(trainable_weights=tw))
        try:
            maml.MAML._get_model_meta_parms(s)
            res = [p.name for p in s.model_meta_parms]
        except ValueError as e:
            res = "ValueError: " + str(e)
        out["meta_parms"].append({"model": model_kind, "meta_parms": meta_parms, "selected": res})
    return out


# ---- the reference's on-disk ingest (utils/dataset.py:41-131) and pretrained-embedding parse (DeepCTR/deepctr.py:105-110) -----
ONDISK = dict(shape="Taobao-10", seed=5, scale=0.004, batch_size=64)


def write_ondisk(root):
    """The tiny on-disk dataset both sides read (written in the reference's layout by tests/test_host.py's writer)."""
    for extra in (os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))):   # tests/ and the repo root
        if extra not in sys.path:
            sys.path.append(extra)
    from mamdr_b200 import synth
    from test_host import _write_reference_layout
    g = synth.generate(ONDISK["shape"], seed=ONDISK["seed"], scale=ONDISK["scale"])
    _write_reference_layout(root, g)
    return {"name": "Taobao", "dataset_path": root, "domain_split_path": "split_by_theme_x", "batch_size": ONDISK["batch_size"],
            "shuffle_buffer_size": 10000, "num_parallel_reads": 8, "seed": 123}


def make_dataset():
    import contextlib
    import io
    import tempfile
    import_reference()
    import utils.dataset as ref_dataset
    import model_zoo.DeepCTR.deepctr as ref_deepctr
    root = tempfile.mkdtemp(prefix="mamdr_ref_ondisk_")
    conf = write_ondisk(root)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ds = ref_dataset.MultiDomainDataset(conf)
    info = ds.dataset_info
    out = {"n_uid": ds.n_uid, "n_pid": ds.n_pid, "n_domain": ds.n_domain,
           "dataset_info": {str(k): v for k, v in info.items()},
           "n_step": {split: {str(k): v["n_step"] for k, v in getattr(ds, split + "_dataset").items()} for split in ("train", "val", "test")}}
    # build_emb: capture the matrix handed to tf.keras.initializers.Constant
    captured = {}
    ref_deepctr.tf = types.SimpleNamespace(keras=types.SimpleNamespace(initializers=types.SimpleNamespace(
        Constant=lambda m: captured.setdefault("m", np.array(m)))))
    s = types.SimpleNamespace(train_config={"load_pretrain_emb": True, "emb_trainable": False}, dataset=ds)
    mats = {}
    for emb_name, n in (("user_emb", ds.n_uid), ("item_emb", ds.n_pid)):
        captured.clear()
        ref_deepctr.DeepCTR.build_emb(s, "x", emb_name, n, 128, trainable=False)
        mats[emb_name] = captured["m"]
    return out, mats


# ---- [EXT] tf.keras 1.12 callbacks used by the finetune stage, restated (tensorflow/python/keras/callbacks.py) -------------------
class KerasEarlyStopping(object):
    def __init__(self, monitor='val_loss', min_delta=0, patience=0, verbose=0, mode='auto', baseline=None):
        assert mode == 'max'
        self.monitor, self.patience, self.min_delta = monitor, patience, abs(min_delta)

    def on_train_begin(self, logs=None):
        self.wait, self.best = 0, -np.inf

    def on_epoch_end(self, epoch, logs=None):
        current = logs[self.monitor]
        if np.greater(current - self.min_delta, self.best):
            self.best, self.wait = current, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.model.stop_training = True


class KerasModelCheckpoint(object):
    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, save_weights_only=False, mode='auto', period=1):
        assert mode == 'max' and save_best_only and save_weights_only
        self.filepath, self.monitor, self.best = filepath, monitor, -np.inf

    def on_train_begin(self, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        current = logs[self.monitor]
        if np.greater(current, self.best):
            self.best = current
            self.model.save_weights(self.filepath, overwrite=True)


FINETUNE_CASES = [("mamdr", "mlp_meta_mamdr_finetune", "plus"), ("dn", "mlp_meta_domain_negotiation_finetune", "plus"),
                  ("reptile", "mlp_meta_reptile_finetune", "plus")]
FINETUNE_EPOCHS = 40


def make_finetune():
    """run.py:66-85 for `*_finetune` names: train -> val_and_test("test") -> load_model(best) -> separate_train_val_test(False)
    (specific_base_model.py:99-162 for MAMDR, base_model.py:41-109 for DN / Reptile), over the toy stand-in with the two Keras
    callbacks restated above."""
    import contextlib
    import io
    import random
    base_model, dn, mamdr, reptile, sbm = import_reference()
    cbs = types.SimpleNamespace(EarlyStopping=KerasEarlyStopping, ModelCheckpoint=KerasModelCheckpoint, TensorBoard=_Any)
    sbm.callbacks = cbs
    base_model.callbacks = cbs
    sbm.AUC = base_model.AUC = lambda **kw: "AUC-metric"      # utils/auc.py builds TF variables; only handed to model.compile here
    g = {}
    for kind, name, method in FINETUNE_CASES:
        cls = {"mamdr": mamdr.MAMDR, "dn": dn.DomainNegotiation, "reptile": reptile.Reptile}[kind]
        tc = dict(LOOP_TC, merged_method=method, loss="binary_crossentropy", learning_rate=0.001)
        obj, model, base = _toy_wrapper(cls, base_model, tc, name)
        if kind != "mamdr":
            base.separate_train_val_test = types.MethodType(base_model.BaseModel.separate_train_val_test, base)
            base.save_model = lambda path, m=model: m.save_weights(path)         # base_model.py:177-181
            base.load_model = lambda path, m=model: m.load_weights(path)
        else:
            base.save_model = lambda path, m=model: m.save_weights(path)
            base.load_model = lambda path, m=model: m.load_weights(path)
        random.seed(LOOP_SEED)
        with contextlib.redirect_stdout(io.StringIO()):
            obj.train()
            obj.val_and_test("test")
            n_train_steps = len(model.steps)
            obj.load_model(obj.checkpoint_path)
            tc["epoch"] = FINETUNE_EPOCHS
            avg_loss, avg_auc, domain_loss, domain_auc = obj.separate_train_val_test(init_parms=False)
        key = "finetune|%s|" % name
        g[key + "steps"] = np.array(model.steps[n_train_steps:], dtype=np.int32)
        g[key + "result"] = np.array([avg_loss, avg_auc] + [domain_loss[d] for d in sorted(N_STEP)] + [domain_auc[d] for d in sorted(N_STEP)],
                                     dtype=np.float64)
        g[key + "live"] = flat_any(model.weights)
        for d in sorted(N_STEP):
            g[key + "ckpt_%d" % d] = flat_any(model.files["domain_%d.h5" % d])
    return g


def make_separate():
    """run.py:67-68 for `<name>_separate`: `BaseModel.separate_train_val_test()` with init_parms=True (base_model.py:41-109):
    tf.global_variables_initializer() (here: the toy model's next initialisation), then per domain restart from those
    weights, train with the compiled optimizer under the two Keras callbacks, load the best checkpoint, test."""
    import contextlib
    import io
    base_model, dn, mamdr, reptile, sbm = import_reference()
    cbs = types.SimpleNamespace(EarlyStopping=KerasEarlyStopping, ModelCheckpoint=KerasModelCheckpoint, TensorBoard=_Any)
    base_model.callbacks = cbs
    base_model.AUC = lambda **kw: "AUC-metric"
    tc = dict(LOOP_TC, loss="binary_crossentropy", learning_rate=0.001, epoch=SEPARATE_EPOCHS, patience=2)
    obj, model, base = _toy_wrapper(dn.DomainNegotiation, base_model, tc, "mlp_separate")
    inits = [0]

    def reinit(op=None):
        inits[0] += 1
        for a, b in zip(model.weights, toy_init(inits[0])):
            a[...] = b
    saved_backend = base_model.backend
    base_model.backend = types.SimpleNamespace(get_session=lambda: types.SimpleNamespace(run=reinit))
    base.separate_train_val_test = types.MethodType(base_model.BaseModel.separate_train_val_test, base)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            avg_loss, avg_auc, domain_loss, domain_auc = base.separate_train_val_test()
    finally:
        base_model.backend = saved_backend
    g = {"separate|steps": np.array(model.steps, dtype=np.int32),
         "separate|result": np.array([avg_loss, avg_auc] + [domain_loss[d] for d in sorted(N_STEP)] + [domain_auc[d] for d in sorted(N_STEP)],
                                     dtype=np.float64),
         "separate|live": flat_any(model.weights), "separate|compiles": np.array([getattr(model, "compiles", 0)], dtype=np.int32)}
    for d in sorted(N_STEP):
        g["separate|ckpt_%d" % d] = flat_any(model.files["domain_%d.h5" % d])
    return g


SEPARATE_EPOCHS = 12


# ---- the joint ('alternate') training loops of the base models: DeepCTR.train (DeepCTR/deepctr.py:63-93), Star.train
# (Star/star.py:35-68), DeepMTLCTR.train (DeepMTLCTR/deep_mtl_ctr.py:68-98) -- note the reference's quirk that every
# `val_and_test("test")` reloads the best checkpoint, so the next epoch continues from the BEST weights, not the latest
JOINT_CASES = ["deepctr", "star", "mtl"]
JOINT_EPOCHS = 5


def make_joint():
    import collections
    import contextlib
    import io
    import random
    base_model, dn, mamdr, reptile, sbm = import_reference()
    import importlib
    # model_zoo/__init__.py binds the CLASS `Star` over the sub-package name: go through sys.modules, not attribute traversal
    ref_deepctr = importlib.import_module("model_zoo.DeepCTR.deepctr")
    ref_mtl = importlib.import_module("model_zoo.DeepMTLCTR.deep_mtl_ctr")
    ref_star = importlib.import_module("model_zoo.Star.star")
    g = {}
    for kind in JOINT_CASES:
        cls = {"deepctr": ref_deepctr.DeepCTR, "star": ref_star.Star, "mtl": ref_mtl.DeepMTLCTR}[kind]
        model = _ToyKeras()
        mk = lambda: collections.OrderedDict((d, {"data": _ToyData(d), "n_step": N_STEP[d]}) for d in sorted(N_STEP))   # noqa: E731
        info = {d: {"n_train": 4 * N_STEP[d], "n_val": 2 + d, "n_test": 3 + d} for d in N_STEP}
        obj = cls.__new__(cls)
        obj.model = model
        obj.domain_model_dict = {d: model for d in N_STEP}        # DeepMTLCTR: one compiled sub-model per domain over shared variables
        obj.dataset = types.SimpleNamespace(train_dataset=mk(), val_dataset=mk(), test_dataset=mk(), dataset_info=info)
        obj.n_domain = len(N_STEP)
        obj.train_config = dict(LOOP_TC, epoch=JOINT_EPOCHS, patience=2)
        obj.model_config = {"name": kind}
        obj.checkpoint_path = "/tmp/unused/model.h5"
        saved = {}
        obj.save_model = lambda path, m=model, s=saved: s.__setitem__("w", [w.copy() for w in m.weights])
        obj.load_model = lambda path, m=model, s=saved: [a.__setitem__(Ellipsis, b) for a, b in zip(m.weights, s["w"])]
        obj._build_early_stop()
        random.seed(LOOP_SEED)
        with contextlib.redirect_stdout(io.StringIO()):
            obj.train()
        key = "joint|%s|" % kind
        g[key + "steps"] = np.array(model.steps, dtype=np.int32)
        g[key + "live"] = flat_any(model.weights)
        g[key + "best"] = flat_any(saved["w"])
        g[key + "es"] = np.array([obj.counter, obj.best_metric, float(obj.early_stop)], dtype=np.float64)
    return g


# ---- output layout: BaseModel.__init__ paths (base_model.py:23-28) and save_result (base_model.py:183-200) --------------------
RESULT_CONFIG = {"model": {"name": "mlp_meta_mamdr_finetune", "hidden_dim": [256, 128, 64]},
                 "train": {"checkpoint_path": "checkpoint", "result_save_path": "result", "patience": 3},
                 "dataset": {"name": "Taobao", "domain_split_path": "split_by_theme_10", "seed": 123}}
RESULT_ARGS = (0.512345, 0.734567, {0: 0.5, 1: 0.52}, {0: 0.71, 1: 0.76})
RESULT_INFO = {"n_user": 7, "n_item": 5, 0: {"n_train": 3, "n_val": 1, "n_test": 2, "ctr_ratio": 0.3}, "total_train": 3}


def result_layout(base_model_module, base_cls, root, extra_config=None):
    """Build a BaseModel subclass instance whose model just records `save_weights`, run save_result under `root` with the
    clock frozen, and describe what was written."""
    import copy
    import json
    cfg = copy.deepcopy(RESULT_CONFIG)
    cfg["train"]["checkpoint_path"] = os.path.join(root, "checkpoint")
    cfg["train"]["result_save_path"] = os.path.join(root, "result")
    if extra_config:
        cfg.update(extra_config)
    written = []
    toy = types.SimpleNamespace(save_weights=lambda path: (written.append(path), open(path, "wb").close()), load_weights=lambda path: None)
    sub = type("Toy", (base_cls,), {"build_model": lambda self: toy})
    dataset = types.SimpleNamespace(n_uid=7, n_pid=5, n_domain=1, conf=cfg["dataset"], dataset_info=RESULT_INFO, device=None)
    real = base_model_module.time.strftime
    base_model_module.time.strftime = lambda *a: "Mon-Jan-01-00-00-00"
    try:
        obj = sub(dataset, cfg)
        obj.save_result(*RESULT_ARGS)
    finally:
        base_model_module.time.strftime = real
    rel = lambda p: os.path.relpath(p, root)   # noqa: E731
    files = {}
    for dirpath, _, names in os.walk(os.path.join(root, "result")):
        for n in sorted(names):
            full = os.path.join(dirpath, n)
            files[rel(full)] = json.load(open(full)) if n.endswith((".json", ".example")) else "<binary>"
    for k in list(files):                       # the config echo contains the temp root: normalise
        if k.endswith("config.json.example"):
            files[k]["train"]["checkpoint_path"] = rel(files[k]["train"]["checkpoint_path"])
            files[k]["train"]["result_save_path"] = rel(files[k]["train"]["result_save_path"])
    return {"checkpoint_path": rel(obj.checkpoint_path), "result_path": rel(obj.result_path), "files": files,
            "save_weights": [rel(p) for p in written]}


def make_result_layout():
    import tempfile
    base_model, dn, mamdr, reptile, sbm = import_reference()
    return result_layout(base_model, base_model.BaseModel, tempfile.mkdtemp(prefix="mamdr_ref_result_"))


# ---- the reference's OWN Keras layers of the STAR tower (Star/star_fcn.py:105-139, Star/partitioned_norm.py:102-203) ------------
# Unlike the mlp / MMOE towers (third-party deepctr), these layers are reference code.  Their `call` methods are executed on numpy
# arrays with the handful of TF / Keras-backend ops they use replaced by numpy equivalents ([EXT] tf.nn.moments,
# tf.nn.batch_normalization; K.moving_average_update only RECORDS which moving statistic receives which value).
STAR_SHAPE = dict(n_domain=3, b=9, emb_dim=(4, 4, 2), n_in=10, hidden=(6, 4), domain=2)


class _Arr(np.ndarray):
    def get_shape(self):
        return types.SimpleNamespace(as_list=lambda: list(self.shape))


def star_problem():
    rng = np.random.default_rng(2025)
    S = STAR_SHAPE
    D, n = S["n_domain"], S["n_in"]
    w = {"gamma_sp": 1 + 0.3 * rng.standard_normal((D, n)), "beta_sp": 0.2 * rng.standard_normal((D, n)),
         "gamma_sh": 1 + 0.3 * rng.standard_normal(n), "beta_sh": 0.2 * rng.standard_normal(n),
         "moving_mean": 0.1 * rng.standard_normal((D, n)), "moving_var": 0.5 + rng.random((D, n)),
         "user_table": rng.standard_normal((S["b"], S["emb_dim"][0])), "item_table": rng.standard_normal((S["b"], S["emb_dim"][1])),
         "domain_emb": rng.standard_normal((D, S["emb_dim"][2]))}
    # the tower's input: [E_u[uid] | E_i[pid] | E_d[dom]] with uid = pid = 0..b-1 (the domain block is constant within a batch)
    w["X"] = np.concatenate([w["user_table"], w["item_table"], np.broadcast_to(w["domain_emb"][S["domain"]], (S["b"], S["emb_dim"][2]))], axis=1)
    dims = (n,) + S["hidden"]
    for l in range(len(S["hidden"])):
        w["k_sp%d" % l] = 0.4 * rng.standard_normal((D, dims[l], dims[l + 1]))
        w["b_sp%d" % l] = 0.1 * rng.standard_normal((D, dims[l + 1]))
        w["k_sh%d" % l] = 0.4 * rng.standard_normal((dims[l], dims[l + 1]))
        w["b_sh%d" % l] = 0.1 * rng.standard_normal(dims[l + 1])
    return w


def make_star_layers():
    import importlib
    import_reference()
    sf = importlib.import_module("model_zoo.Star.star_fcn")
    pn = importlib.import_module("model_zoo.Star.partitioned_norm")
    S, w = STAR_SHAPE, star_problem()
    dom = S["domain"]
    ind = np.full((S["b"], 1), dom, dtype=np.int32)
    ns = types.SimpleNamespace
    np_tf = ns(cast=lambda x, t: int(x), multiply=np.multiply, add=np.add, int32=None, equal=lambda a, b: a == b,
               case=lambda pairs, exclusive=False, name=None: next(fn for pred, fn in pairs if pred)())
    np_nn = ns(embedding_lookup=lambda table, idx: table[idx], bias_add=lambda x, b: x + b)
    # ---- PartitionedNorm.call
    updates = []

    def batch_normalization(x, mean, var, beta, gamma, epsilon):        # [EXT] tf.nn.batch_normalization
        inv = gamma / np.sqrt(var + epsilon)
        return x * inv + (beta - mean * inv)

    def normalize_batch_in_training(x, gamma, beta, reduction_axes, epsilon):   # [EXT] tf.nn.moments (biased variance)
        mean = np.mean(x, axis=tuple(reduction_axes))
        var = np.mean((x - mean) ** 2, axis=tuple(reduction_axes))
        return batch_normalization(x, mean, var, beta, gamma, epsilon), mean, var

    pn.tf, pn.nn = np_tf, np_nn
    pn.K = ns(learning_phase=lambda: 1, reshape=np.reshape, batch_normalization=batch_normalization,
              normalize_batch_in_training=normalize_batch_in_training,
              moving_average_update=lambda var, value, momentum: updates.append((var["name"], np.array(value), momentum)),
              in_train_phase=lambda a, b, training=None: a if training else b())
    layer = pn.PartitionedNorm.__new__(pn.PartitionedNorm)
    layer.n_domain, layer.axis, layer.momentum, layer.epsilon = S["n_domain"], -1, 0.99, 1e-3
    layer.PN_Gamma, layer.PN_Beta, layer.Shared_Gamma, layer.Shared_Beta = w["gamma_sp"], w["beta_sp"], w["gamma_sh"], w["beta_sh"]
    layer.PN_Mean = [{"name": "mean_%d" % d, "value": w["moving_mean"][d]} for d in range(S["n_domain"])]
    layer.PN_Var = [{"name": "var_%d" % d, "value": w["moving_var"][d]} for d in range(S["n_domain"])]
    layer.add_update = lambda *a, **k: None
    X = w["X"].view(_Arr)
    h0_train = np.asarray(layer.call([X, ind], training=1))
    g = {"star|h0_train": h0_train}
    assert [u[0] for u in updates] == ["mean_%d" % dom, "var_%d" % dom] and all(u[2] == 0.99 for u in updates)
    g["star|batch_mean"], g["star|batch_var"] = updates[0][1], updates[1][1]
    # inference: the K.batch_normalization branch reads the batch's domain's moving statistics through tf.case
    pn.K.batch_normalization = lambda x, mean, var, beta, gamma, epsilon: batch_normalization(
        x, mean["value"] if isinstance(mean, dict) else mean, var["value"] if isinstance(var, dict) else var, beta, gamma, epsilon)
    g["star|h0_eval"] = np.asarray(layer.call([X, ind], training=0))
    # ---- StarFCN.call, layer by layer on the training-mode output
    sf.tf, sf.nn = np_tf, np_nn
    sf.ops = ns(convert_to_tensor=lambda x, dtype=None: np.asarray(x))
    sf.common_shapes = ns(rank=lambda x: x.ndim)
    sf.gen_math_ops = ns(mat_mul=lambda a, b: a @ b)
    h = h0_train
    for l in range(len(S["hidden"])):
        fcn = sf.StarFCN.__new__(sf.StarFCN)
        fcn.PN_Kernel, fcn.PN_Bias, fcn.Shred_Kernel, fcn.Shared_Bias = w["k_sp%d" % l], w["b_sp%d" % l], w["k_sh%d" % l], w["b_sh%d" % l]
        fcn.use_bias, fcn.activation, fcn.dtype = True, (lambda x: np.maximum(x, 0)), "float64"
        h = np.asarray(fcn.call([h, ind]))
        g["star|h%d" % (l + 1)] = h
    return g


# ---- the reference's OWN streaming AUC (utils/auc.py:110-157,159-177,248-281 + utils/metrics_utils.py:194-354) -------------------
# These two files are part of the reference tree (a back-port of tf.keras.metrics.AUC).  `AUC.__init__`, `update_state` ->
# `update_confusion_matrix_variables` and `result` are executed with the ~20 TF ops they use replaced by their numpy analogues
# (float32 like the graph); the two shape-normalising helpers of metrics_utils are bypassed (dense [b] inputs).
AUC_STREAM = [(1024, 0), (977, 1), (1, 2), (300, 3)]        # (rows, seed) of the batches of one streaming evaluation


def auc_batch(rows, seed):
    rng = np.random.default_rng(500 + seed)
    y = (rng.random(rows) < 0.3).astype(np.float32)
    p = np.clip(0.25 * y + rng.random(rows) * 0.75, 0, 1).astype(np.float32)
    thr = [(i + 1) * 1.0 / 499 for i in range(498)]
    p[: min(rows, 32)] = np.asarray(thr, dtype=np.float32)[rng.integers(0, 498, min(rows, 32))]   # predictions exactly on thresholds
    if rows > 2:
        p[-1], p[-2] = 0.0, 1.0
    return y, p


class _Var(np.ndarray):
    """A tf.Variable stand-in: a float32 array with assign_add."""

    def assign_add(self, delta):
        self += np.asarray(delta, dtype=np.float32)
        return self


def _numpy_tf_for_metrics():
    import contextlib
    ns = types.SimpleNamespace
    f32 = np.float32

    def div_no_nan(a, b, name=None):
        a, b = np.asarray(a, dtype=f32), np.asarray(b, dtype=f32)
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.where(b == 0, f32(0), a / b).astype(f32)
    math_ops = ns(cast=lambda x, dtype=None: np.asarray(getattr(x, "arr", x)).astype(dtype), greater=np.greater, logical_and=np.logical_and,
                  logical_not=np.logical_not, reduce_sum=lambda x, axis=None, name=None: np.sum(np.asarray(x, dtype=f32), axis=axis, dtype=f32),
                  div_no_nan=div_no_nan, multiply=lambda a, b: (np.asarray(a, f32) * np.asarray(b, f32)).astype(f32),
                  maximum=np.maximum, minimum=np.minimum, log=np.log)
    array_ops = ns(size=lambda x: int(np.asarray(x).size), reshape=lambda x, shape: np.reshape(x, shape), tile=lambda x, m: np.tile(x, [int(k) for k in m]),
                   expand_dims=lambda x, axis: np.expand_dims(x, axis), constant=lambda x: np.asarray(x, dtype=f32), stack=lambda xs: list(xs),
                   where=np.where, ones_like=np.ones_like)
    dtypes = ns(float32=np.float32, bool=np.bool_)
    ops = ns(control_dependencies=lambda deps: contextlib.nullcontext())
    check_ops = ns(assert_greater_equal=lambda *a, **k: None, assert_less_equal=lambda *a, **k: None)
    control_flow_ops = ns(group=lambda ops_: None)
    return math_ops, array_ops, dtypes, ops, check_ops, control_flow_ops


def make_auc():
    import importlib
    import_reference()
    mu = importlib.import_module("utils.metrics_utils")
    au = importlib.import_module("utils.auc")
    math_ops, array_ops, dtypes, ops, check_ops, control_flow_ops = _numpy_tf_for_metrics()
    mu.math_ops, mu.array_ops, mu.dtypes, mu.ops, mu.check_ops, mu.control_flow_ops = math_ops, array_ops, dtypes, ops, check_ops, control_flow_ops
    mu.to_list = lambda x: list(x) if isinstance(x, (list, tuple)) else [x]
    compat = types.SimpleNamespace(assert_is_compatible_with=lambda other: None)
    proxy = lambda a: types.SimpleNamespace(arr=np.asarray(a), shape=compat, dtype=np.float32)   # noqa: E731
    mu.ragged_assert_compatible_and_get_flat_values = lambda values, mask=None: ([proxy(v) for v in values], mask)
    mu.squeeze_or_expand_dimensions = lambda y_pred, y_true=None, sample_weight=None: (y_pred.arr, y_true.arr)
    au.math_ops, au.array_ops = math_ops, array_ops
    au.K = types.SimpleNamespace(epsilon=lambda: 1e-7)
    au.tf = types.SimpleNamespace(assign=lambda v, x: None, zeros_like=np.zeros_like)
    au.AUC.variables = []
    g = {}
    for T in (3, 500):
        a = au.AUC(num_thresholds=T, name="AUC")                       # the real __init__: threshold table, curve / summation enums
        for nm in ("true_positives", "true_negatives", "false_positives", "false_negatives"):
            setattr(a, nm, np.zeros(T, dtype=np.float32).view(_Var))
        g["auc|T%d|thresholds" % T] = np.asarray(a.thresholds, dtype=np.float64)
        if T == 3:
            a.update_state(np.float32([0, 0, 1, 1]), np.float32([0, 0.5, 0.3, 0.9]))     # the doc-string example, utils/auc.py:44-56
            g["auc|T3|acc"] = np.stack([np.asarray(a.true_positives), np.asarray(a.false_positives), np.asarray(a.false_negatives),
                                        np.asarray(a.true_negatives)])
            g["auc|T3|result"] = np.float32(a.result())
            continue
        for k, (rows, seed) in enumerate(AUC_STREAM):
            y, p = auc_batch(rows, seed)
            a.update_state(y, p)
            g["auc|T500|acc_after_%d" % k] = np.stack([np.asarray(a.true_positives), np.asarray(a.false_positives),
                                                       np.asarray(a.false_negatives), np.asarray(a.true_negatives)])
            g["auc|T500|result_after_%d" % k] = np.float32(a.result())
    return g


def flat_any(ws):
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in ws])


if __name__ == "__main__":
    if not os.path.isdir(REFERENCE):
        raise SystemExit("the reference tree is not available here; the committed .npz is the artefact")
    out = os.path.join(HERE, "reference_meta_v1.npz")
    np.savez_compressed(out, **make())
    print(out, os.path.getsize(out), "bytes")
    out = os.path.join(HERE, "reference_loops_v1.npz")
    loops = make_loops()
    loops.update(make_joint())
    loops.update(make_finetune())
    loops.update(make_separate())
    np.savez_compressed(out, **loops)
    print(out, os.path.getsize(out), "bytes")
    import json
    out = os.path.join(HERE, "reference_dispatch_v1.json")
    with open(out, "w") as f:
        json.dump(make_dispatch(), f, indent=1)
    print(out, os.path.getsize(out), "bytes")
    out = os.path.join(HERE, "reference_result_layout_v1.json")
    with open(out, "w") as f:
        json.dump(make_result_layout(), f, indent=1)
    print(out, os.path.getsize(out), "bytes")
    out = os.path.join(HERE, "reference_star_layers_v1.npz")
    np.savez_compressed(out, **make_star_layers())
    print(out, os.path.getsize(out), "bytes")
    out = os.path.join(HERE, "reference_auc_v1.npz")
    np.savez_compressed(out, **make_auc())
    print(out, os.path.getsize(out), "bytes")
    info, mats = make_dataset()
    out = os.path.join(HERE, "reference_dataset_v1.json")
    with open(out, "w") as f:
        json.dump(info, f, indent=1)
    print(out, os.path.getsize(out), "bytes")
    out = os.path.join(HERE, "reference_dataset_emb_v1.npz")
    np.savez_compressed(out, **{k: v.astype(np.float32) for k, v in mats.items()})
    for k, v in mats.items():
        assert np.array_equal(v.astype(np.float32).astype(v.dtype), v), "the float64 matrix holds float32 values"
    print(out, os.path.getsize(out), "bytes")
