"""CPU: the oracle's meta algebra against vectors produced by EXECUTING THE REFERENCE'S OWN CODE
(tests/golden/reference_meta_v1.npz, written by tests/golden/make_reference_golden.py from /root/reference with TF / deepctr
stubbed out): SURVEY.md section 8 rows a2-a5, a8, a18 and Reptile's rules (f4) -- bit-exact.  /root/reference is NOT read here."""
import os
import types

import numpy as np

from conftest import ROOT
from oracle import meta as ometa

REF = np.load(os.path.join(ROOT, "tests", "golden", "reference_meta_v1.npz"))
SHAPES = [(7, 5), (5,), (3, 4), (1,)]
TC = {"meta_learning_rate": 0.1, "domain_meta_learning_rate": 0.1, "sample_num": 5, "merged_method": "plus", "patience": 3}


def _split(flat):
    out, o = [], 0
    for s in SHAPES:
        n = int(np.prod(s))
        out.append(np.array(flat[o:o + n], dtype=np.float32).reshape(s))
        o += n
    return out


def _flat(ws):
    return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in ws])


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _same(got, key):
    np.testing.assert_array_equal(_bits(_flat(got)), _bits(REF[key]), err_msg=key)


class _Model(object):
    dtype = np.dtype(np.float32)

    def __init__(self, ws):
        self.ws = ws
        self.auc = types.SimpleNamespace(reset_states=lambda: None)

    def get_weights(self):
        return [w.copy() for w in self.ws]

    def set_weights(self, ws):
        pass


def _mamdr(method):
    o = ometa.OracleMAMDR.__new__(ometa.OracleMAMDR)
    o.model, o.tc = _Model(_split(REF["model"])), dict(TC, merged_method=method)
    return o


def test_merge_and_dr_updates_match_the_reference():
    theta, theta_i = _split(REF["theta"]), _split(REF["theta_i"])
    for method in ("plus", "times"):
        merged = ometa.merge_weights(theta, theta_i, method)                      # a5
        _same(merged, "merge_" + method)
        o = _mamdr(method)
        ti = [w.copy() for w in theta_i]
        o._update_meta_weight(ti, merged, TC["domain_meta_learning_rate"])        # a2 (DR form)
        _same(ti, "dr_update_" + method)
        acc = _split(REF["accum0"])
        o._accumulate_grad(acc, merged, theta)                                    # a3
        _same(acc, "accumulate_" + method)
        ti = [w.copy() for w in theta_i]
        o._update_meta_weight_by_grads(acc, ti)
        _same(ti, "apply_accum_" + method)
        assert not _flat(acc).any()
        new = o.model.get_weights()                                               # a4: theta_i = model - merged
        _same([n - m for n, m in zip(new, merged)], "update_domain_weights_" + method)
    o = _mamdr("plus")
    t = [w.copy() for w in theta]
    o._update_meta_weight(t, None, TC["meta_learning_rate"])                      # a2 (DN form inside MAMDR)
    _same(t, "mamdr_dn_form")


class _PassSetsWeights(object):
    """A stand-in model whose training pass lands on fixed weights: isolates the outer update of the DN / Reptile loops."""
    dtype = np.dtype(np.float32)

    def __init__(self, start, after_pass):
        self.weights = [w.copy() for w in start]
        self.after = after_pass
        self.calls = 0
        self.auc = types.SimpleNamespace(reset_states=lambda: None)

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, ws):
        for a, b in zip(self.weights, ws):
            a[...] = b

    def train_on_batch(self, uid, pid, domain, label, optimizer='adam', sgd_lr=None):
        tgt = self.after(self.calls) if callable(self.after) else self.after
        self.calls += 1
        for a, b in zip(self.weights, tgt):
            a[...] = b
        return 0.0, 0.5


def _data(n_domain):
    return {'train': {d: {'uid': np.zeros(3, np.int32), 'pid': np.zeros(3, np.int32), 'label': np.zeros(3, np.float32)} for d in range(n_domain)}}


def test_dn_and_reptile_outer_updates_match_the_reference():
    from mamdr_b200.schedule import Schedule
    theta, model = _split(REF["theta"]), _split(REF["model"])
    tc = dict(TC, shuffle_sequence=False, meta_train_step=0)
    dn = ometa.OracleDN(_PassSetsWeights(theta, model), _data(1), tc, 8, Schedule(1))      # a8
    dn.train_epoch()
    _same(dn.meta_weights, "dn_update")
    rp = ometa.OracleReptile(_PassSetsWeights(theta, model), _data(1), tc, 8, Schedule(1))  # f4, per-domain rule
    rp.train_epoch()
    _same(rp.meta_weights, "reptile_update")
    # f4, batch rule: three domains whose passes land on model + 0.25 k, deltas summed against the SAME theta, applied once
    shifted = lambda k: [w + np.float32(0.25 * k) for w in model]   # noqa: E731
    rb = ometa.OracleReptile(_PassSetsWeights(theta, shifted), _data(3), tc, 8, Schedule(1), name="mlp_meta_reptile_batch")
    orig = rb.schedule.shuffle_sequence
    rb.schedule.shuffle_sequence = lambda seq: list(seq)              # keep 0, 1, 2: the reference vector used k = 0, 1, 2
    rb.train_epoch()
    rb.schedule.shuffle_sequence = orig
    _same(rb.meta_weights, "reptile_apply")


def test_early_stop_and_weighted_auc_match_the_reference():
    es = ometa.EarlyStop(TC["patience"])
    saved = []
    trace = []
    for mval in REF["early_stop_metrics"]:
        stop = es.step(float(mval), lambda: saved.append(1))
        trace.append([es.counter, float(es.best_metric), float(bool(stop)), float(len(saved))])
    np.testing.assert_array_equal(np.array(trace), REF["early_stop_trace"])
    sizes = {0: (10, 4, 6), 1: (30, 9, 1), 2: (5, 2, 3)}
    data = {m: {k: {'uid': np.zeros(v[i])} for k, v in sizes.items()} for i, m in enumerate(("train", "val", "test"))}
    auc = {0: 0.61, 1: 0.72, 2: 0.55}
    got = np.array([ometa.weighted_auc(data, m, auc) for m in ("train", "val", "test")])
    np.testing.assert_allclose(got, REF["weighted_auc"], rtol=1e-15)


def test_product_early_stop_matches_the_reference():
    """mamdr_b200/base_model.py:early_stop_step / _weighted_auc (host code, no GPU) on the same trace."""
    from mamdr_b200.base_model import BaseModel
    s = types.SimpleNamespace(train_config=TC, checkpoint_path="/dev/null", saved=[], log=lambda *a: None)
    s.save_model = lambda path: s.saved.append(1)
    BaseModel._build_early_stop(s)
    trace = []
    for mval in REF["early_stop_metrics"]:
        stop = BaseModel.early_stop_step(s, float(mval))
        trace.append([s.counter, float(s.best_metric), float(bool(stop)), float(len(s.saved))])
    np.testing.assert_array_equal(np.array(trace), REF["early_stop_trace"])
    info = {0: {"n_train": 10, "n_val": 4, "n_test": 6}, 1: {"n_train": 30, "n_val": 9, "n_test": 1}, 2: {"n_train": 5, "n_val": 2, "n_test": 3}}
    s.dataset = types.SimpleNamespace(dataset_info=info)
    auc = {0: 0.61, 1: 0.72, 2: 0.55}
    got = np.array([BaseModel._weighted_auc(s, m, auc) for m in ("train", "val", "test")])
    np.testing.assert_allclose(got, REF["weighted_auc"], rtol=1e-15)


# ---- the reference's TRAINING LOOPS (executed over a toy Keras stand-in) vs the oracle's loops ---------------------------------
import sys  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_reference_golden as mrg  # noqa: E402  (toy definitions only: /root/reference is not touched by importing it)

LOOPS = np.load(os.path.join(ROOT, "tests", "golden", "reference_loops_v1.npz"))


class _ToyOracleModel(object):
    dtype = np.dtype(np.float32)

    def __init__(self):
        self.weights = mrg.toy_init(0)
        self.steps = []
        self.auc = types.SimpleNamespace(reset_states=lambda: None)

    def get_weights(self):
        return [w.copy() for w in self.weights]

    def set_weights(self, ws):
        for a, b in zip(self.weights, ws):
            a[...] = b

    def train_on_batch(self, uid, pid, domain, label, optimizer='adam', sgd_lr=None):
        mrg.toy_step(self.weights, domain)
        self.steps.append(domain)
        return 0.0, 0.5

    def evaluate(self, uid, pid, domain, label, batch_size):
        return mrg.toy_eval(self.weights, domain)


def _toy_data(bs):
    col = lambda n: {'uid': np.zeros(n, np.int32), 'pid': np.zeros(n, np.int32), 'label': np.zeros(n, np.float32)}   # noqa: E731
    return {'train': {d: col(bs * s) for d, s in sorted(mrg.N_STEP.items())},
            'val': {d: col(2 + d) for d in sorted(mrg.N_STEP)}, 'test': {d: col(3 + d) for d in sorted(mrg.N_STEP)}}


import pytest  # noqa: E402


@pytest.mark.parametrize("kind,name,method", mrg.SUBSET_CASES)
def test_oracle_loops_with_a_meta_parameter_subset(kind, name, method):
    """meta_parms selecting a SUBSET of the variables (config #4's ["emb", "kernel_shared", "bias_shared"]): in the reference the
    other variables are re-initialised by every `init_layer` call (so they start from the LAST draw), are never reloaded from
    theta and evolve through every pass; the oracle's `MetaSubset` view must reproduce that -- steps, theta, theta_d and the
    full live model when train() returns, bit for bit."""
    from mamdr_b200.schedule import Schedule
    bs = 4
    tc = dict(mrg.LOOP_TC, merged_method=method)
    model = _ToyOracleModel()
    sel = mrg.SUBSET_META_IDX
    key = "%s|%s|subset|" % (name, method)
    if kind == "mamdr":
        last = mrg.toy_init(len(mrg.N_STEP))                   # init_layer ran once per domain over ALL layers (mamdr.py:31-33)
        for i, w in enumerate(model.weights):
            if i not in sel:
                w[...] = last[i]
        om = ometa.OracleMAMDR(ometa.MetaSubset(model, sel), _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED),
                               {d: [mrg.toy_init(d + 1)[i] for i in sel] for d in mrg.N_STEP}, name=name)
    elif kind == "dn":
        om = ometa.OracleDN(ometa.MetaSubset(model, sel), _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED))
    else:
        om = ometa.OracleReptile(ometa.MetaSubset(model, sel), _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED), name=name)
    for epoch in range(tc["epoch"]):
        om.train_epoch()
        _, val_auc, _, _ = om.val_and_test("val")
        if om.early_stop_step(val_auc):
            break
        om.val_and_test("test")
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(mrg.flat_any(om.meta_weights)), _bits(LOOPS[key + "theta"]))
    if kind == "mamdr":
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(mrg.flat_any(om.domain_weights[d])), _bits(LOOPS[key + "theta_%d" % d]))
        np.testing.assert_array_equal(_bits(mrg.flat_any(model.weights)), _bits(LOOPS[key + "live"]))
    else:
        # DN / Reptile: `val_and_test("test")` reloads the best h5 (ALL variables); the oracle's snapshot is the meta view, so only
        # the meta part of the live model is comparable here -- the product harness below compares the full arena
        k = sum(int(np.prod(model.weights[i].shape)) for i in sel)
        np.testing.assert_array_equal(_bits(mrg.flat_any([model.weights[i] for i in sel])), _bits(LOOPS[key + "live"][:k]))


@pytest.mark.parametrize("kind,name,method", mrg.LOOP_CASES)
def test_oracle_loops_replay_the_reference_loops(kind, name, method):
    """MAMDR.train / DomainNegotiation.train / Reptile.train of the reference, run for two epochs over a toy model whose train
    step depends on the domain (so every re-ordering shows), against the oracle's loops with `Schedule(seed)`: the same
    sequence of train steps, theta, every theta_d, the kept best snapshots and the early-stop counters -- bit for bit."""
    from mamdr_b200.schedule import Schedule
    bs = 4
    tc = dict(mrg.LOOP_TC, merged_method=method)
    model = _ToyOracleModel()
    key = "%s|%s|" % (name, method)
    if kind == "mamdr":
        om = ometa.OracleMAMDR(model, _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED), {d: mrg.toy_init(d + 1) for d in mrg.N_STEP}, name=name)
    elif kind == "dn":
        om = ometa.OracleDN(model, _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED))
    else:
        om = ometa.OracleReptile(model, _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED), name=name)
    for epoch in range(tc["epoch"]):
        om.train_epoch()
        _, val_auc, _, _ = om.val_and_test("val")
        if om.early_stop_step(val_auc):
            break
        om.val_and_test("test")
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(mrg.flat_any(om.meta_weights)), _bits(LOOPS[key + "theta"]))
    np.testing.assert_array_equal(np.array([om.es.counter, om.es.best_metric], dtype=np.float64), LOOPS[key + "es"])
    if kind == "mamdr":
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(mrg.flat_any(om.domain_weights[d])), _bits(LOOPS[key + "theta_%d" % d]), err_msg="theta_%d" % d)
            np.testing.assert_array_equal(_bits(mrg.flat_any(om.best_domain_weights[d])), _bits(LOOPS[key + "best_theta_%d" % d]))
        np.testing.assert_array_equal(_bits(mrg.flat_any(om.best_shared_weights)), _bits(LOOPS[key + "best_theta"]))
    else:
        np.testing.assert_array_equal(_bits(mrg.flat_any(om.best_weights)), _bits(LOOPS[key + "best"]))


# ---- the reference's CLI dispatch (run.py:22-87) and meta-parameter selection (maml.py:153-179), executed -> product -----------
import json  # noqa: E402

DISPATCH = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_dispatch_v1.json")))
OUT_OF_SCOPE = {"mlp_uncertainty_weight": NotImplementedError, "nothing": ValueError}


@pytest.mark.parametrize("name", sorted(DISPATCH["dispatch"]))
def test_product_cli_dispatch_matches_the_reference(monkeypatch, name):
    """Which base model / wrapper is built for a model name and which calls `main` then makes (train, test, the finetune stage
    after `load_model`, `separate`, `save_result`) -- the trace recorded by running the reference's run.main over recording
    stand-ins, against this repo's run.main over the same stand-ins.  Names outside the hot-path scope raise instead."""
    import run as prod_run
    import mamdr_b200.dataset as p_dataset
    import mamdr_b200.deep_mtl_ctr as p_mtl
    import mamdr_b200.deepctr as p_ctr
    import mamdr_b200.domain_negotiation as p_dn
    import mamdr_b200.maml as p_maml
    import mamdr_b200.mamdr as p_mamdr
    import mamdr_b200.mldg as p_mldg
    import mamdr_b200.pcgrad as p_pcgrad
    import mamdr_b200.reptile as p_rep
    import mamdr_b200.star as p_star
    trace = []
    rec = mrg.recorder_classes(["MultiDomainDataset", "Star", "DeepCTR", "DeepMTLCTR", "DomainNegotiation", "MAMDR", "Reptile", "MAML",
                                "MLDG", "PCGrad"], trace)
    for mod, cls in ((p_dataset, "MultiDomainDataset"), (p_star, "Star"), (p_ctr, "DeepCTR"), (p_mtl, "DeepMTLCTR"),
                     (p_dn, "DomainNegotiation"), (p_mamdr, "MAMDR"), (p_rep, "Reptile"), (p_maml, "MAML"), (p_mldg, "MLDG"),
                     (p_pcgrad, "PCGrad")):
        monkeypatch.setattr(mod, cls, rec[cls])
    config = {"model": {"name": name}, "dataset": {"seed": 1}}
    if name in OUT_OF_SCOPE:
        with pytest.raises(OUT_OF_SCOPE[name]):
            prod_run.main(config)
        return
    prod_run.main(config)
    assert trace == DISPATCH["dispatch"][name]


@pytest.mark.parametrize("case", DISPATCH["meta_parms"], ids=lambda c: "%s-%s" % (c["model"], "+".join(c["meta_parms"])))
def test_product_meta_parameter_selection_matches_the_reference(case):
    """maml.py:153-179 executed on the variable names of the mlp / STAR models vs mamdr_b200/maml.py on the same names: the
    same variables in the same order ("all", "all_hidden", substring lists), the same ValueError for an unmatched name."""
    from mamdr_b200.maml import MAML
    names = mrg.VAR_NAMES[case["model"]]
    tw = [types.SimpleNamespace(name=n, offset=32 * i, numel=8) for i, n in enumerate(names)]
    w = MAML(types.SimpleNamespace(train_config={"meta_parms": case["meta_parms"]}, model=types.SimpleNamespace(trainable_weights=tw)))
    if isinstance(case["selected"], str):
        with pytest.raises(ValueError) as e:
            w._get_model_meta_parms()
        assert "ValueError: " + str(e.value) == case["selected"]
        return
    w._get_model_meta_parms()
    assert [p.name for p in w.model_meta_parms] == case["selected"]
    # the arena spans cover exactly the selected variables
    covered = set()
    for off, n in w.meta_ranges:
        covered |= set(range(off // 32, (off + n) // 32))
    assert covered == {names.index(n) for n in case["selected"]}


def test_product_variable_names_are_the_ones_the_selection_was_pinned_on():
    """The names the product gives its variables (engine.MLPModel.TF_NAMES / star naming) are the ones used above."""
    from mamdr_b200.engine import MLPModel
    from mamdr_b200.layout import mlp_layout
    lo = mlp_layout(10, 10, 3, (128, 128, 128), (256, 128, 64), False)
    got = []
    for name in lo.names:
        tf_name = MLPModel.TF_NAMES.get(name)
        if tf_name is None:
            kind, idx = ("kernel", name[6:]) if name.startswith("kernel") else ("bias", name[4:])
            tf_name = "dnn/%s%s:0" % (kind, idx)
        got.append(tf_name)
    assert got == mrg.VAR_NAMES["mlp_frozen"]


# ---- the reference's on-disk ingest and pretrained-embedding parse, executed -> product ---------------------------------------
def test_product_ondisk_ingest_matches_the_reference(tmp_path):
    """`utils.dataset.MultiDomainDataset` (utils/dataset.py:41-131: id counts, domain discovery,
    `wc -l` sizes, n_step = ceil(n / batch), ctr_ratio, dataset_info) and `DeepCTR.build_emb`'s parse of the `"f f f ..."`
    embedding strings (DeepCTR/deepctr.py:105-110) were EXECUTED on a tiny on-disk dataset; this repo's reader must agree."""
    from mamdr_b200.dataset import MultiDomainDataset
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_dataset_v1.json")))
    emb = np.load(os.path.join(ROOT, "tests", "golden", "reference_dataset_emb_v1.npz"))
    conf = mrg.write_ondisk(str(tmp_path))
    ds = MultiDomainDataset(conf, device=None)
    assert (ds.n_uid, ds.n_pid, ds.n_domain) == (ref["n_uid"], ref["n_pid"], ref["n_domain"])
    info = {str(k): v for k, v in ds.dataset_info.items()}
    assert info == ref["dataset_info"]
    for split in ("train", "val", "test"):
        got = {str(k): v["n_step"] for k, v in getattr(ds, split + "_dataset").items()}
        assert got == ref["n_step"][split]
        assert list(getattr(ds, split + "_dataset").keys()) == sorted(getattr(ds, split + "_dataset").keys())   # domains in numeric order
    np.testing.assert_array_equal(_bits(ds.user_table), _bits(emb["user_emb"]))
    np.testing.assert_array_equal(_bits(ds.item_table), _bits(emb["item_emb"]))


def test_product_output_layout_matches_the_reference(tmp_path):
    """`BaseModel.__init__`'s checkpoint / result paths (base_model.py:23-28) and `save_result`
    (base_model.py:183-200: folder name, dataset_info.json, config.json.example, result.json, model_parameters.h5) EXECUTED
    with the clock frozen vs mamdr_b200/base_model.py run the same way: the same relative paths, file set and JSON contents."""
    import mamdr_b200.base_model as p_base
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_result_layout_v1.json")))
    got = mrg.result_layout(p_base, p_base.BaseModel, str(tmp_path))
    # the ONE deviation: the weight files are numpy archives keyed by the TF variable names (no h5py here), and they say so
    # in their extension: model_parameters.h5 -> model_parameters.npz (INTEGRATION.md shows the 5-line h5 converter)
    npz = lambda s: s[:-3] + ".npz" if s.endswith(".h5") else s   # noqa: E731
    assert got["checkpoint_path"] == npz(ref["checkpoint_path"]) and got["result_path"] == ref["result_path"]
    assert sorted(got["files"]) == sorted(npz(k) for k in ref["files"])
    for k in ref["files"]:
        assert got["files"][npz(k)] == ref["files"][k], k
    assert got["save_weights"] == [npz(k) for k in ref["save_weights"]]


def test_oracle_star_layers_match_the_reference_layers():
    """The STAR tower's layers are the reference's OWN code (Star/partitioned_norm.py:102-203, Star/star_fcn.py:105-139), not a
    third-party dependency: their `call` methods were executed on numpy arrays (the few TF / Keras-backend ops they use replaced
    by numpy equivalents) and oracle/star.py must reproduce the outputs -- gamma = shared * specific[d], beta = shared +
    specific[d] with d read from the first row's indicator, batch statistics in training and the batch's DOMAIN's moving
    statistics at inference, which moving statistic receives the update, W = shared * specific[d], b = shared + specific[d]."""
    from oracle.star import OracleStar, StarSpec
    ref = np.load(os.path.join(ROOT, "tests", "golden", "reference_star_layers_v1.npz"))
    S, w = mrg.STAR_SHAPE, mrg.star_problem()
    spec = StarSpec(S["b"], S["b"], S["n_domain"], emb_dim=S["emb_dim"], hidden=S["hidden"])
    named = {"domain_emb": w["domain_emb"], "gamma_specific": w["gamma_sp"], "beta_specific": w["beta_sp"], "gamma_shared": w["gamma_sh"],
             "beta_shared": w["beta_sh"], "out_kernel": np.zeros((S["hidden"][-1], 1)), "out_bias": np.zeros(1)}
    for l in range(len(S["hidden"])):
        named.update({"kernel_specific%d" % l: w["k_sp%d" % l], "bias_specific%d" % l: w["b_sp%d" % l],
                      "kernel_shared%d" % l: w["k_sh%d" % l], "bias_shared%d" % l: w["b_sh%d" % l]})
    o = OracleStar(spec, [named[n] for n in spec.names], w["user_table"], w["item_table"], dtype=np.float64)
    o.moving_mean[...], o.moving_var[...] = w["moving_mean"], w["moving_var"]
    ids = np.arange(S["b"])
    H, _, cache = o.forward(ids, ids, S["domain"], train=True)
    np.testing.assert_allclose(H[0], ref["star|h0_train"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(cache["mu"], ref["star|batch_mean"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(cache["var"], ref["star|batch_var"], rtol=1e-10, atol=1e-14)
    for l in range(len(S["hidden"])):
        np.testing.assert_allclose(H[l + 1], ref["star|h%d" % (l + 1)], rtol=1e-10, atol=1e-12)
    H_eval, _, _ = o.forward(ids, ids, S["domain"], train=False)
    np.testing.assert_allclose(H_eval[0], ref["star|h0_eval"], rtol=1e-10, atol=1e-12)


def test_oracle_auc_matches_the_reference_metric_code():
    """utils/auc.py + utils/metrics_utils.py are files of the reference tree; `AUC.__init__` / `update_state` / `result` were
    executed (numpy analogues of the TF ops) on the doc-string example and on a four-batch stream with predictions sitting
    exactly on thresholds, 0 and 1.  oracle/auc.py and the product's threshold table reproduce the threshold table and every
    accumulator exactly, the interpolated ROC-AUC within float32 summation order."""
    from mamdr_b200.auc import thresholds as product_thresholds
    from oracle import auc as oauc
    ref = np.load(os.path.join(ROOT, "tests", "golden", "reference_auc_v1.npz"))
    for T in (3, 500):
        want = ref["auc|T%d|thresholds" % T].astype(np.float32)
        np.testing.assert_array_equal(np.asarray(oauc.thresholds(T), dtype=np.float32), want)
        np.testing.assert_array_equal(np.asarray(product_thresholds(T), dtype=np.float32), want)
    a = oauc.AUC(3)
    a.update_state(np.float32([0, 0, 1, 1]), np.float32([0, 0.5, 0.3, 0.9]))
    np.testing.assert_array_equal(np.asarray(a.acc, dtype=np.float32), ref["auc|T3|acc"])
    assert abs(a.result() - float(ref["auc|T3|result"])) < 1e-7
    a = oauc.AUC(500)
    for k, (rows, seed) in enumerate(mrg.AUC_STREAM):
        y, p = mrg.auc_batch(rows, seed)
        a.update_state(y, p)
        np.testing.assert_array_equal(np.asarray(a.acc, dtype=np.float32), ref["auc|T500|acc_after_%d" % k], err_msg="batch %d" % k)
        assert abs(a.result() - float(ref["auc|T500|result_after_%d" % k])) < 2e-6


@pytest.mark.parametrize("i", range(len(mrg.VARIANT_CASES)))
def test_oracle_loops_config_knobs(i):
    """The loops' config keys (finetune_every_epoch, domain_regulation_step, add_query_domain, sample_num, merged_method,
    an explicit meta_sequence without shuffling, meta_train_step, val_every_step, epoch, meta_learning_rate, target_domain >= 0)
    as the reference's executed loops handle them vs the oracle."""
    from mamdr_b200.schedule import Schedule
    kind, name, over = mrg.VARIANT_CASES[i]
    bs = 4
    tc = dict(mrg.LOOP_TC, merged_method=over.get("merged_method", "plus"))
    tc.update(over)
    model = _ToyOracleModel()
    key = "variant%d|" % i
    if kind == "mamdr":
        om = ometa.OracleMAMDR(model, _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED), {d: mrg.toy_init(d + 1) for d in mrg.N_STEP}, name=name)
    elif kind == "dn":
        om = ometa.OracleDN(model, _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED))
    else:
        om = ometa.OracleReptile(model, _toy_data(bs), tc, bs, Schedule(mrg.LOOP_SEED), name=name)
    for epoch in range(tc["epoch"]):
        om.train_epoch()
        if epoch % tc["val_every_step"] == 0:
            _, val_auc, _, val_domain_auc = om.val_and_test("val")
            metric = val_domain_auc[tc["target_domain"]] if tc["target_domain"] >= 0 else val_auc   # the target domain's own AUC
            if om.early_stop_step(metric):
                break
            om.val_and_test("test")
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(mrg.flat_any(om.meta_weights)), _bits(LOOPS[key + "theta"]))
    if kind == "mamdr":
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(mrg.flat_any(om.domain_weights[d])), _bits(LOOPS[key + "theta_%d" % d]), err_msg="theta_%d" % d)
