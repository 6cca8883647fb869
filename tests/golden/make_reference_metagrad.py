"""Generates tests/golden/reference_metagrad_v1.npz by EXECUTING THE REFERENCE'S OWN MAML / MLDG / PCGrad training loops
(`model_zoo/maml.py:35-151`, `model_zoo/mldg.py:35-155`, `model_zoo/pcgrad.py:35-160`, read from /root/reference, never copied)
over the toy Keras stand-in of make_reference_golden.py -- SURVEY.md section 8(f) row f4.

    python tests/golden/make_reference_metagrad.py        # only where /root/reference exists (not on the GPU box)

What runs is the reference's code: `train`, `_meta_train_step`, `val`, `PCGrad.PCGrad`, PCGrad's `build_meta_data_split`, the
early-stop / val_and_test of `BaseModel`.  What is a stand-in (TF graph construction cannot run here): `_make_meta_train_function`
(the accumulating K.function becomes `accum += toy_grad(live weights, domain)`, the meta AdamOptimizer becomes the oracle's TF-ordered
numpy Adam), the session (`K.get_session().run(op)` runs `op` when it is callable), `K.batch_get_value`, and MAML's tf.data
`build_meta_data_split` (only its step arithmetic `int(n * ratio)` / `ceil(n / batch)` is restated).  The oracle
(`oracle/meta.py: OracleMAML / OracleMLDG / OraclePCGrad`) and the product's wrappers (`mamdr_b200/{maml,mldg,pcgrad}.py`, driven
on the CPU harness of tests/test_product_metagrad_vs_reference.py) must land on the same bits; the projection vectors pin
`oracle.meta.pcgrad_project` and the CUDA kernel `mamdr_pcgrad_project`.
"""
import contextlib
import io
import math
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_reference_golden as mrg  # noqa: E402

BATCH = 4
N_DATA = {d: 4 * s for d, s in mrg.N_STEP.items()}
TC = dict(mrg.LOOP_TC, meta_learning_rate=0.05, meta_split="meta-train/val", meta_split_ratio=0.5, average_meta_grad="none",
          sample_num=2)
CASES = [("maml", "mlp_meta", {}), ("maml", "mlp_meta_batch", {}), ("maml", "mlp_meta", {"meta_split": "train-train", "meta_train_step": 1}),
         ("mldg", "mlp_meta_mldg", {}), ("mldg", "mlp_meta_mldg_batch", {"epoch": 3}),
         ("pcgrad", "mlp_pcgrad", {}), ("pcgrad", "mlp_pcgrad", {"meta_train_step": 1, "sample_num": 3}),
         # meta_finetune_step > 0: `val()` finetunes a few epochs per domain before evaluating (maml.py:244-287, 343-353)
         ("maml", "mlp_meta", {"meta_finetune_step": 2}), ("mldg", "mlp_meta_mldg_batch", {"meta_finetune_step": 1}),
         # target_domain >= 0 under PCGrad: the target is skipped as a query domain (it can still be drawn as a support domain) and
         # its validation AUC drives the early stop (pcgrad.py:68-69,137-138)
         ("pcgrad", "mlp_pcgrad", {"target_domain": 1, "epoch": 3})]


def toy_grad(weights, domain):
    return [w * np.float32(0.1) + np.float32(0.01 * (domain + 1)) for w in weights]


def split_steps(tc, n):
    """maml.py:300-327 -- the step arithmetic of the meta split."""
    if tc["meta_split"] in ("meta-train/val", "meta-train/val-no-exclusive"):
        n_train = int(n * tc["meta_split_ratio"])
        n_test = n - n_train
    else:
        n_train = n_test = n
    return int(math.ceil(n_train / float(BATCH))), int(math.ceil(n_test / float(BATCH)))


def import_modules():
    mrg.import_reference()
    import model_zoo.maml as maml
    import model_zoo.mldg as mldg
    import model_zoo.pcgrad as pcgrad
    return maml, mldg, pcgrad


class _Session(object):
    def run(self, op):
        if isinstance(op, (list, tuple)):
            return [self.run(o) for o in op]
        if callable(op) and not isinstance(op, type) and type(op).__name__ != "_Any":
            return op()
        return None


def run_case(kind, name, over):
    from oracle.mlp import AdamState
    maml, mldg, pcgrad = import_modules()
    base_model = sys.modules["model_zoo.base_model"]
    mod, cls = {"maml": (maml, maml.MAML), "mldg": (mldg, mldg.MLDG), "pcgrad": (pcgrad, pcgrad.PCGrad)}[kind]
    tc = dict(TC)
    tc.update(over)
    obj, model, base = mrg._toy_wrapper(cls, base_model, tc, name)
    model._standardize_user_data = lambda it: ([it], [], None)
    model.grad_calls = []
    sel = list(range(len(model.weights)))
    sess = _Session()
    mod.K = types.SimpleNamespace(get_session=lambda: sess, batch_get_value=lambda vs: [np.array(v) for v in vs])
    if kind == "pcgrad":
        # PCGrad.build_meta_data_split (:324-330) is plain Python over d['data'].make_initializable_iterator(): it RUNS
        pass
    else:
        def build_meta_data_split():
            out = {}
            for idx in sorted(mrg.N_STEP):
                ts, ms = split_steps(tc, N_DATA[idx])
                out[idx] = {"train_iter": types.SimpleNamespace(domain=idx, initializer=None), "train_step": ts,
                            "meta_iter": types.SimpleNamespace(domain=idx, initializer=None), "meta_val_step": ms}
            return out
        obj.build_meta_data_split = build_meta_data_split

    def make_meta_train_function():
        obj.accum_grads = [np.zeros_like(model.weights[i]) for i in sel]
        adam = AdamState([model.weights[i] for i in sel], lr=tc["meta_learning_rate"])
        obj.meta_parms_update_step = lambda: adam.apply([model.weights[i] for i in sel], obj.accum_grads)
        obj.clear_grads = lambda: [a.__setitem__(Ellipsis, 0) for a in obj.accum_grads]

        def set_accum(values):
            for a, v in zip(obj.accum_grads, values):
                a[...] = v
        obj.set_accum_grads = set_accum

        def meta_train(inputs):
            it = inputs[0]
            for a, g in zip(obj.accum_grads, toy_grad([model.weights[i] for i in sel], it.domain)):
                a += g
            model.grad_calls.append(it.domain)
            return 0.0, 0.5
        return meta_train
    obj._make_meta_train_function = make_meta_train_function
    random.seed(mrg.LOOP_SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        obj.train()
    return {"live": mrg.flat_any(model.weights), "steps": np.array(model.steps, dtype=np.int32),
            "grad_calls": np.array(model.grad_calls, dtype=np.int32), "best": mrg.flat_any(base.saved),
            "es": np.array([base.counter, base.best_metric], dtype=np.float64)}


def make_projection():
    """`PCGrad.PCGrad` executed on random gradient lists (2-D, 1-D, [n, 1] and [1] variables), two support domains in a row."""
    _, _, pcgrad = import_modules()
    rng = np.random.default_rng(77)
    shapes = [(9, 6), (5,), (64, 1), (1,), (33, 40)]
    g = {"shapes": np.array([s + (0,) * (2 - len(s)) for s in shapes], dtype=np.int32)}
    cur = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    g["current"] = mrg.flat_any(cur)
    final = cur
    for k in range(2):
        aux = [rng.standard_normal(s).astype(np.float32) for s in shapes]
        g["aux%d" % k] = mrg.flat_any(aux)
        pcgrad.PCGrad.PCGrad(None, final, cur, aux)
        g["final%d" % k] = mrg.flat_any(final)
    return g


def make():
    g = {}
    for i, (kind, name, over) in enumerate(CASES):
        for k, v in run_case(kind, name, over).items():
            g["case%d|%s" % (i, k)] = v
    for k, v in make_projection().items():
        g["proj|" + k] = v
    return g


if __name__ == "__main__":
    if not os.path.isdir(mrg.REFERENCE):
        raise SystemExit("the reference tree is not available here; the committed .npz is the artefact")
    out = os.path.join(HERE, "reference_metagrad_v1.npz")
    np.savez_compressed(out, **make())
    print(out, os.path.getsize(out), "bytes")
