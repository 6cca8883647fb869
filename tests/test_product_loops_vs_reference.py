"""CPU: the PRODUCT's wrapper control flow (mamdr_b200/{mamdr,domain_negotiation,reptile,specific_base_model,maml}.py -- the real
code, unmodified) against the reference's training loops EXECUTED in this repo's build container
(tests/golden/reference_loops_v1.npz, make_reference_golden.py): the same sequence of train steps, theta, every theta_d, best
snapshots and early-stop state, bit for bit.

The wrappers drive the device through two seams only: `model.ctx.call(<C-ABI entry point>, pointers...)` and the base model's
`run_train_pass`.  Here both are replaced by TEST stand-ins -- a numpy interpreter of the eight K9 / K10 meta sweeps acting on
the pointed-to (CPU) arenas with the kernels' formulas, and the toy train step of the reference run -- so that the host logic
runs without a GPU.  This is a test harness, not a product path: the product has no CPU fallback (tests/test_abi.py)."""
import contextlib
import ctypes as C
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import ROOT
from mamdr_b200.base_model import BaseModel
from mamdr_b200.engine import NamedWeight
from mamdr_b200.layout import ParamLayout
from mamdr_b200.schedule import Schedule

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_reference_golden as mrg  # noqa: E402

LOOPS = np.load(os.path.join(ROOT, "tests", "golden", "reference_loops_v1.npz"))


def _view(p, n):
    return np.ctypeslib.as_array((C.c_float * int(n)).from_address(p.value))


def _merge(a, b, method):
    return a + b if method == 0 else a * b


class _NumpyMetaOps(object):
    """The K9 / K10 entry points of include/mamdr_b200.h on host arenas, with the formulas of csrc/meta_ops.cuh."""

    def __init__(self):
        self.launches = 0
        self.calls = []

    def call(self, name, *a):
        self.calls.append(name)
        f32 = np.float32
        if name == "mamdr_copy":
            dst, src, n = a[0], a[1], a[2]
            _view(dst, n)[...] = _view(src, n)
        elif name == "mamdr_merge":
            out, th, ti, n, method = a[:5]
            _view(out, n)[...] = _merge(_view(th, n), _view(ti, n), method)
        elif name == "mamdr_dn_update":
            theta, model, beta, n, model_out = a[:5]
            t = _view(theta, n)
            t += (_view(model, n) - t) * f32(beta)
            if model_out is not None and model_out.value:
                _view(model_out, n)[...] = t
        elif name == "mamdr_dr_update":
            ti, th, model, beta, n, method, model_out = a[:7]
            x, t = _view(ti, n), _view(th, n)
            x += (_view(model, n) - _merge(t, x, method)) * f32(beta)
            if model_out is not None and model_out.value:
                _view(model_out, n)[...] = _merge(t, x, method)
        elif name == "mamdr_dr_accumulate":
            acc, model, th, ti, n, method = a[:6]
            t = _view(th, n)
            d = _view(model, n) - _merge(t, _view(ti, n), method)
            _view(acc, n)[...] += d if method == 0 else d * t
        elif name == "mamdr_dr_apply_accum":
            ti, acc, sample_num, beta, n = a[:5]
            g = _view(acc, n)
            _view(ti, n)[...] += g / f32(sample_num) * f32(beta)
            g[...] = 0
        elif name == "mamdr_sub":
            out, x, y, n = a[:4]
            _view(out, n)[...] = _view(x, n) - _view(y, n)
        elif name == "mamdr_axpy_diff":
            out, x, y, alpha, n = a[:5]
            _view(out, n)[...] += (_view(x, n) - _view(y, n)) * f32(alpha)
        else:
            raise AssertionError("the wrappers called an entry point this harness does not interpret: " + name)


class _ToyDeviceModel(object):
    """What the wrappers touch on `model`: one flat arena in trainable_weights order + the toy train / eval step."""

    def __init__(self):
        w0 = mrg.toy_init(0)
        self.layout = ParamLayout(["kernel0", "bias0"], [w.shape for w in w0])
        self.params = torch.from_numpy(self.layout.pack(w0))
        self.m, self.v = torch.zeros_like(self.params), torch.zeros_like(self.params)
        self.ctx = _NumpyMetaOps()
        self.stream = None
        self.steps = []
        self.stateful_metric_functions = [self]

    @property
    def trainable_weights(self):
        lo = self.layout
        return [NamedWeight("dnn/%s:0" % n, v, o, k) for n, v, o, k in zip(lo.names, lo.views(self.params), lo.offsets, lo.numels)]

    def views(self):
        return [v.numpy() for v in self.layout.views(self.params)]

    def reset_states(self):
        pass

    def reset_optimizer(self):
        pass

    @contextlib.contextmanager
    def program(self, enabled=True):
        yield False

    def evaluate(self, data, steps=None):
        return mrg.toy_eval(self.views(), data.domain)

    # the sharded meta-step's view of the optimizer state (engine.MLPModel: m / v arenas, opt_words / set_opt_words)
    def opt_words(self):
        return torch.zeros(3)

    def set_opt_words(self, words):
        pass

    # the finetune stage's Keras-like surface (engine.MLPModel: compile / get_weights / set_weights)
    def compile(self, optimizer="adam", lr=None):
        self.optimizer = optimizer

    def get_weights(self):
        return self.params.clone()

    def set_weights(self, flat):
        self.params.copy_(flat)

    def save_weights(self, path, flat=None):   # engine.MLPModel.save_weights: an .npz keyed by the variables' TF names
        flat = self.params if flat is None else flat
        with open(path, "wb") as f:
            np.savez(f, **{w.name: v.numpy() for w, v in zip(self.trainable_weights, self.layout.views(flat))})


def _base(name, method, meta_parms=("all",)):
    model = _ToyDeviceModel()
    mk = lambda: {d: {"data": types.SimpleNamespace(domain=d), "n_step": mrg.N_STEP[d], "n_data": 4 * mrg.N_STEP[d]} for d in sorted(mrg.N_STEP)}   # noqa: E731
    info = {d: {"n_train": 4 * mrg.N_STEP[d], "n_val": 2 + d, "n_test": 3 + d} for d in mrg.N_STEP}
    base = types.SimpleNamespace(
        model=model, train_config=dict(mrg.LOOP_TC, merged_method=method, meta_parms=list(meta_parms)), model_config={"name": name}, b200_config={},
        dataset=types.SimpleNamespace(train_dataset=mk(), val_dataset=mk(), test_dataset=mk(), dataset_info=info),
        n_domain=len(mrg.N_STEP), schedule=Schedule(mrg.LOOP_SEED), checkpoint_path="unused", log=lambda *a: None, saved=None, inits=[0])
    for meth in ("val_and_test", "early_stop_step", "_weighted_auc", "_format_print_domain_metric", "_build_early_stop"):
        setattr(base, meth, types.MethodType(getattr(BaseModel, meth), base))
    base._build_early_stop()
    base.save_model = lambda path: setattr(base, "saved", model.params.clone())
    base.load_model = lambda path: model.params.copy_(base.saved)
    base.stage_epoch_orders = lambda passes, mine=None: None

    def run_train_pass(idx, steps=None):
        n = base.dataset.train_dataset[idx]["n_step"] if steps is None else steps
        for _ in range(n):
            mrg.toy_step(model.views(), idx)
            model.steps.append(idx)
    base.run_train_pass = run_train_pass

    def draw_initial_weights():
        base.inits[0] += 1
        return mrg.toy_init(base.inits[0])
    base.draw_initial_weights = draw_initial_weights
    return base, model


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _load_ckpt(model, path):
    blob = np.load(path)
    return torch.from_numpy(model.layout.pack([blob[w.name] for w in model.trainable_weights]))


def _flat(model, flat_tensor):
    return mrg.flat_any([v.numpy() for v in model.layout.views(flat_tensor)])


@pytest.mark.parametrize("kind,name,method", mrg.LOOP_CASES)
def test_product_wrappers_replay_the_reference_loops(kind, name, method):
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR
    from mamdr_b200.reptile import Reptile
    base, model = _base(name, method)
    wrapper = {"mamdr": MAMDR, "dn": DomainNegotiation, "reptile": Reptile}[kind](base)
    wrapper.train()                                             # the product's own train(): epochs, val, early stop, test
    key = "%s|%s|" % (name, method)
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(_flat(model, wrapper.meta_weights.flat)), _bits(LOOPS[key + "theta"]))
    np.testing.assert_array_equal(np.array([base.counter, base.best_metric], dtype=np.float64), LOOPS[key + "es"])
    if kind == "mamdr":
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(_flat(model, wrapper.domain_weights[d].flat)), _bits(LOOPS[key + "theta_%d" % d]),
                                          err_msg="theta_%d" % d)
            np.testing.assert_array_equal(_bits(_flat(model, wrapper.best_domain_weights[d].flat)), _bits(LOOPS[key + "best_theta_%d" % d]))
        np.testing.assert_array_equal(_bits(_flat(model, wrapper.best_shared_weights.flat)), _bits(LOOPS[key + "best_theta"]))
        used = set(model.ctx.calls)
        assert "mamdr_merge" in used and ("mamdr_dr_accumulate" in used) == ("batch" in name)
    else:
        np.testing.assert_array_equal(_bits(_flat(model, base.saved)), _bits(LOOPS[key + "best"]))


@pytest.mark.parametrize("kind,name,method", mrg.SUBSET_CASES)
def test_product_wrappers_with_a_meta_parameter_subset(kind, name, method):
    """meta_parms = ["kernel0"]: only the first variable is a meta parameter (as config #4's ["emb", "kernel_shared",
    "bias_shared"] leaves the specific tensors out).  The product keeps full-arena snapshots and restricts every get / set /
    update to the meta spans; the non-meta variable must be re-initialised by every init_layer, never reloaded from theta and
    train through every pass exactly as in the reference run -- steps, theta, theta_d and the FULL live arena, bit for bit."""
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR
    from mamdr_b200.reptile import Reptile
    base, model = _base(name, method, meta_parms=("kernel0",))
    wrapper = {"mamdr": MAMDR, "dn": DomainNegotiation, "reptile": Reptile}[kind](base)
    wrapper.train()
    key = "%s|%s|subset|" % (name, method)
    lo = model.layout
    k = lo.numels[0]

    def meta_part(flat_tensor):
        return lo.views(flat_tensor)[0].numpy().reshape(-1)

    assert [p.name for p in wrapper.model_meta_parms] == ["dnn/kernel0:0"] and wrapper.meta_ranges == [(0, 32)]
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(meta_part(wrapper.meta_weights.flat)), _bits(LOOPS[key + "theta"]))
    if kind == "mamdr":
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(meta_part(wrapper.domain_weights[d].flat)), _bits(LOOPS[key + "theta_%d" % d]))
            np.testing.assert_array_equal(_bits(meta_part(wrapper.best_domain_weights[d].flat)), _bits(LOOPS[key + "best_theta_%d" % d]))
    np.testing.assert_array_equal(_bits(_flat(model, model.params)), _bits(LOOPS[key + "live"]))
    assert k == LOOPS[key + "theta"].size


@pytest.mark.parametrize("kind", mrg.JOINT_CASES)
def test_product_joint_training_loops_replay_the_reference(kind):
    """`DeepCTR.train` (DeepCTR/deepctr.py:63-93), `Star.train` (Star/star.py:35-68), `DeepMTLCTR.train`
    (DeepMTLCTR/deep_mtl_ctr.py:68-98) executed over the toy stand-in vs the product's `train` methods (real code, run on a
    stand-in `self`): shuffled domains, one full pass each, val -> early stop -> test, INCLUDING the reference's quirk that
    `val_and_test("test")` reloads the best checkpoint so the next epoch continues from the best weights -- steps, the live
    model, the kept checkpoint and the early-stop state, bit for bit."""
    from mamdr_b200.deep_mtl_ctr import DeepMTLCTR
    from mamdr_b200.deepctr import DeepCTR
    from mamdr_b200.star import Star
    cls = {"deepctr": DeepCTR, "star": Star, "mtl": DeepMTLCTR}[kind]
    base, model = _base(kind, "plus")
    obj = cls.__new__(cls)                       # no __init__: building the real model needs the GPU
    for k, v in vars(base).items():
        setattr(obj, k, v)
    del obj.val_and_test, obj.early_stop_step, obj._weighted_auc, obj._format_print_domain_metric, obj._build_early_stop   # use the class's own
    obj.train_config = dict(base.train_config, epoch=mrg.JOINT_EPOCHS, patience=2)
    obj._build_early_stop()
    obj.save_model = lambda path: setattr(obj, "saved", model.params.clone())
    obj.load_model = lambda path: model.params.copy_(obj.saved)
    obj.train()
    key = "joint|%s|" % kind
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(_flat(model, model.params)), _bits(LOOPS[key + "live"]))
    np.testing.assert_array_equal(_bits(_flat(model, obj.saved)), _bits(LOOPS[key + "best"]))
    np.testing.assert_array_equal(np.array([obj.counter, obj.best_metric, float(obj.early_stop)], dtype=np.float64), LOOPS[key + "es"])


@pytest.mark.parametrize("kind,name,method", mrg.FINETUNE_CASES)
def test_product_finetune_stage_replays_the_reference(tmp_path, kind, name, method):
    """run.py:66-85 for `*_finetune` names -- train, test, reload the best checkpoint, then `separate_train_val_test(False)`
    (specific_base_model.py:99-162 for MAMDR: restart every domain from best theta (+) best theta_d; base_model.py:41-109 for DN /
    Reptile: restart from the reloaded weights) -- EXECUTED over the toy stand-in with the two Keras callbacks restated
    ([EXT] tf.keras 1.12 EarlyStopping(min_delta=1e-4) / ModelCheckpoint(save_best_only)), vs the product's own code: the same
    finetune steps, per-domain checkpoints, returned losses / AUCs and the restored live model, bit for bit.  40 epochs of a
    contracting toy map make the val_AUC gains fall below min_delta, which is where the EarlyStopping bookkeeping matters."""
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR
    from mamdr_b200.reptile import Reptile
    base, model = _base(name, method)
    base.checkpoint_path = str(tmp_path / "ckpt" / "model_parameters.npz")
    base.train_config.update(loss="binary_crossentropy", learning_rate=0.001)
    base.separate_train_val_test = types.MethodType(BaseModel.separate_train_val_test, base)
    wrapper = {"mamdr": MAMDR, "dn": DomainNegotiation, "reptile": Reptile}[kind](base)
    wrapper.train()
    wrapper.val_and_test("test")
    n_train_steps = len(model.steps)
    wrapper.load_model(wrapper.checkpoint_path)
    base.train_config["epoch"] = mrg.FINETUNE_EPOCHS
    avg_loss, avg_auc, domain_loss, domain_auc = wrapper.separate_train_val_test(init_parms=False)
    key = "finetune|%s|" % name
    np.testing.assert_array_equal(np.array(model.steps[n_train_steps:], dtype=np.int32), LOOPS[key + "steps"])
    got = np.array([avg_loss, avg_auc] + [domain_loss[d] for d in sorted(mrg.N_STEP)] + [domain_auc[d] for d in sorted(mrg.N_STEP)])
    np.testing.assert_array_equal(got, LOOPS[key + "result"])
    np.testing.assert_array_equal(_bits(_flat(model, model.params)), _bits(LOOPS[key + "live"]))
    for d in sorted(mrg.N_STEP):
        ck = _load_ckpt(model, str(tmp_path / "ckpt" / ("domain_%d.npz" % d)))
        np.testing.assert_array_equal(_bits(_flat(model, ck)), _bits(LOOPS[key + "ckpt_%d" % d]), err_msg="checkpoint of domain %d" % d)
    assert model.optimizer == "adam"        # the stage hands the model back compiled with the training optimizer


@pytest.mark.parametrize("i", range(len(mrg.VARIANT_CASES)))
def test_product_wrappers_config_knobs(i):
    """The loops' config keys as the reference's executed loops handle them vs the product's wrapper code (see
    tests/test_reference_golden.py::test_oracle_loops_config_knobs for the list)."""
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR
    from mamdr_b200.reptile import Reptile
    kind, name, over = mrg.VARIANT_CASES[i]
    base, model = _base(name, over.get("merged_method", "plus"))
    base.train_config.update(over)
    wrapper = {"mamdr": MAMDR, "dn": DomainNegotiation, "reptile": Reptile}[kind](base)
    wrapper.train()
    key = "variant%d|" % i
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS[key + "steps"])
    np.testing.assert_array_equal(_bits(_flat(model, wrapper.meta_weights.flat)), _bits(LOOPS[key + "theta"]))
    np.testing.assert_array_equal(_bits(_flat(model, model.params)), _bits(LOOPS[key + "live"]))
    if kind == "mamdr":
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(_flat(model, wrapper.domain_weights[d].flat)), _bits(LOOPS[key + "theta_%d" % d]),
                                          err_msg="theta_%d" % d)


# ---- two ranks (gloo): DR query domains sharded over the ranks (mamdr_b200/dist.py), one all-reduce per meta-step -------------------
def _sharded_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo")
    from mamdr_b200.mamdr import MAMDR
    base, model = _base("mlp_meta_mamdr", "plus")
    wrapper = MAMDR(base)
    wrapper.train()
    owned = sorted(d for d, r in wrapper.dr_owner.items() if r == rank)
    torch.save({"theta": wrapper.meta_weights.flat.clone(), "theta_d": {d: w.flat.clone() for d, w in wrapper.domain_weights.items()},
                "best_theta_d": {d: w.flat.clone() for d, w in wrapper.best_domain_weights.items()}, "steps": list(model.steps),
                "owned": owned, "es": [base.counter, base.best_metric]}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_product_meta_steps_equal_the_reference_sequential_run(tmp_path):
    """With a train step that carries no optimizer state, DR chains are independent given theta, so sharding the query domains
    over two ranks (replicated DN phase, LPT-assigned chains, ONE all-reduce per meta-step) must reproduce the reference's
    SEQUENTIAL run bit for bit on every rank: theta, every theta_d, the best snapshots and the early-stop state; each rank
    executes the full DN phase and only its own chains."""
    import socket
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    key = "mlp_meta_mamdr|plus|"
    lo = ParamLayout(["kernel0", "bias0"], [w.shape for w in mrg.toy_init(0)])
    flat = lambda t: mrg.flat_any([v.numpy() for v in lo.views(t)])   # noqa: E731
    blobs = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r), weights_only=False) for r in range(2)]
    n_seq = len(LOOPS[key + "steps"])
    assert sorted(blobs[0]["owned"] + blobs[1]["owned"]) == sorted(mrg.N_STEP) and blobs[0]["owned"] and blobs[1]["owned"]
    dn_steps_per_epoch = sum(mrg.N_STEP.values())
    for b in blobs:
        np.testing.assert_array_equal(_bits(flat(b["theta"])), _bits(LOOPS[key + "theta"]))
        for d in sorted(mrg.N_STEP):
            np.testing.assert_array_equal(_bits(flat(b["theta_d"][d])), _bits(LOOPS[key + "theta_%d" % d]), err_msg="theta_%d" % d)
            np.testing.assert_array_equal(_bits(flat(b["best_theta_d"][d])), _bits(LOOPS[key + "best_theta_%d" % d]))
        np.testing.assert_array_equal(np.array(b["es"], dtype=np.float64), LOOPS[key + "es"])
        assert len(b["steps"]) < n_seq                                # fewer train steps than the sequential run ...
    # ... and together exactly the sequential run's steps, the replicated DN phase counted once per epoch
    total = len(blobs[0]["steps"]) + len(blobs[1]["steps"]) - mrg.LOOP_TC["epoch"] * dn_steps_per_epoch
    assert total == n_seq


def _pair_sharded_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo")
    from mamdr_b200.mamdr import MAMDR
    base, model = _base("mlp_meta_mamdr_batch", "plus")
    wrapper = MAMDR(base)
    wrapper.train()
    mine = sorted(k for k, r in wrapper.dr_pair_owner.items() if r == rank)
    torch.save({"theta": wrapper.meta_weights.flat.clone(), "theta_d": {d: w.flat.clone() for d, w in wrapper.domain_weights.items()},
                "steps": list(model.steps), "pairs": mine, "n_pairs": len(wrapper.dr_pair_owner), "live": model.params.clone()},
               os.path.join(out_dir, "pair%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharded_batch_meta_steps_equal_the_reference_sequential_run(tmp_path):
    """'batch' names (mamdr.py:100-108,182-196): every (query, support) pair starts from theta (+) theta_i, so with a train step
    that carries no optimizer state the PAIRS are independent.  Two gloo ranks each run their LPT share of the pairs, ONE
    all-reduce sums the accumulated deltas, every rank applies them: theta equals the reference's executed SEQUENTIAL run bit
    for bit, every theta_d to the last ulps (the two ranks' partial sums are added in a different association), the replicas
    are bit-identical, and together the ranks execute exactly the sequential run's train steps."""
    import socket
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_pair_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    key = "mlp_meta_mamdr_batch|plus|"
    lo = ParamLayout(["kernel0", "bias0"], [w.shape for w in mrg.toy_init(0)])
    flat = lambda t: mrg.flat_any([v.numpy() for v in lo.views(t)])   # noqa: E731
    a, b = [torch.load(os.path.join(str(tmp_path), "pair%d.pt" % r), weights_only=False) for r in range(2)]
    assert a["pairs"] and b["pairs"] and not set(a["pairs"]) & set(b["pairs"]) and len(a["pairs"]) + len(b["pairs"]) == a["n_pairs"]
    assert torch.equal(a["theta"], b["theta"]) and torch.equal(a["live"], b["live"])
    np.testing.assert_array_equal(_bits(flat(a["theta"])), _bits(LOOPS[key + "theta"]))
    for d in sorted(mrg.N_STEP):
        assert torch.equal(a["theta_d"][d], b["theta_d"][d])
        np.testing.assert_allclose(flat(a["theta_d"][d]), LOOPS[key + "theta_%d" % d], rtol=2e-6, atol=1e-7, err_msg="theta_%d" % d)
    dn_steps_per_epoch = sum(mrg.N_STEP.values())
    total = len(a["steps"]) + len(b["steps"]) - mrg.LOOP_TC["epoch"] * dn_steps_per_epoch
    assert total == len(LOOPS[key + "steps"])


def _sharded_subset_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo")
    from mamdr_b200.mamdr import MAMDR
    base, model = _base("mlp_meta_mamdr", "plus", meta_parms=("kernel",))   # bias0 is NOT a meta parameter
    wrapper = MAMDR(base)
    wrapper.train()
    torch.save({"theta": wrapper.meta_weights.flat.clone(), "theta_d": {d: w.flat.clone() for d, w in wrapper.domain_weights.items()},
                "live": model.params.clone(), "owner": dict(wrapper.dr_owner), "last": wrapper.train_sequence[-1]},
               os.path.join(out_dir, "sub%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_meta_steps_with_a_subset_of_meta_parameters_keep_the_replicas_identical(tmp_path):
    """meta_parms = a subset (config #4 ships ['emb', 'kernel_shared', 'bias_shared']): the other variables train through every
    pass and are never reloaded from theta, so each rank's copy follows its own chains.  The one all-reduce of the meta-step
    therefore also carries the live model of the rank that owns the last chain; afterwards (and hence at the start of the next
    replicated DN phase) every rank holds the same live arena, theta and theta_d -- bit for bit."""
    import socket
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_sharded_subset_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = [torch.load(os.path.join(str(tmp_path), "sub%d.pt" % r), weights_only=False) for r in range(2)]
    assert a["owner"] == b["owner"] and set(a["owner"].values()) == {0, 1}
    assert torch.equal(a["live"], b["live"])
    assert torch.equal(a["theta"], b["theta"])
    for d in a["theta_d"]:
        assert torch.equal(a["theta_d"][d], b["theta_d"][d]), d


@pytest.mark.parametrize("kind,name", [("mamdr", "mlp_meta_mamdr"), ("dn", "mlp_meta_domain_negotiation"), ("reptile", "mlp_meta_reptile_batch")])
def test_save_state_resume_on_the_cpu_harness(tmp_path, kind, name):
    """`save_state` after the first meta-step, `load_state` into a fresh wrapper, second meta-step: the same theta / theta_d /
    live arena / step sequence as the uninterrupted run (which itself equals the reference's executed loop)."""
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR
    from mamdr_b200.reptile import Reptile
    cls = {"mamdr": MAMDR, "dn": DomainNegotiation, "reptile": Reptile}[kind]

    def fresh():
        base, model = _base(name, "plus")
        w = cls(base)
        if kind == "dn":
            w._get_model_meta_parms()
            w.meta_weights = w._get_meta_weights()
            w.meta_sequence = w.build_meta_data_split()
        else:
            w.prepare()
        return w, base, model

    a, base_a, model_a = fresh()
    a.train_epoch(0)
    path = a.save_state(str(tmp_path / "state.pt"), epoch=0)
    first = len(model_a.steps)
    a.train_epoch(1)
    b, base_b, model_b = fresh()
    assert b.load_state(path) == 0
    b.train_epoch(1)
    assert model_b.steps == model_a.steps[first:]
    assert torch.equal(a.meta_weights.flat, b.meta_weights.flat) and torch.equal(model_a.params, model_b.params)
    if kind == "mamdr":
        for d in a.domain_weights:
            assert torch.equal(a.domain_weights[d].flat, b.domain_weights[d].flat), d
    # and the uninterrupted two meta-steps are the reference's (no val / early stop in between: theta only)
    key = "%s|plus|" % name
    np.testing.assert_array_equal(_bits(_flat(model_a, a.meta_weights.flat)), _bits(LOOPS[key + "theta"]))


def test_product_separate_training_replays_the_reference(tmp_path):
    """run.py:67-68 for `<name>_separate`: `BaseModel.separate_train_val_test()` with init_parms=True (base_model.py:41-109)
    EXECUTED over the toy stand-in (global_variables_initializer = the toy model's next initialisation, the two Keras callbacks
    restated) vs the product's own code: one model per domain from the fresh initialisation, the compiled optimizer kept (no SGD
    re-compile), best-val_AUC checkpoints, test -- steps, results, per-domain checkpoints and the restored model, bit for bit."""
    base, model = _base("mlp_separate", "plus")
    base.checkpoint_path = str(tmp_path / "ckpt" / "model_parameters.npz")
    base.train_config.update(loss="binary_crossentropy", learning_rate=0.001, epoch=mrg.SEPARATE_EPOCHS, patience=2)
    base.separate_train_val_test = types.MethodType(BaseModel.separate_train_val_test, base)
    model.optimizer = "adam"
    avg_loss, avg_auc, domain_loss, domain_auc = base.separate_train_val_test()
    np.testing.assert_array_equal(np.array(model.steps, dtype=np.int32), LOOPS["separate|steps"])
    got = np.array([avg_loss, avg_auc] + [domain_loss[d] for d in sorted(mrg.N_STEP)] + [domain_auc[d] for d in sorted(mrg.N_STEP)])
    np.testing.assert_array_equal(got, LOOPS["separate|result"])
    np.testing.assert_array_equal(_bits(_flat(model, model.params)), _bits(LOOPS["separate|live"]))
    for d in sorted(mrg.N_STEP):
        ck = _load_ckpt(model, str(tmp_path / "ckpt" / ("domain_%d.npz" % d)))
        np.testing.assert_array_equal(_bits(_flat(model, ck)), _bits(LOOPS["separate|ckpt_%d" % d]), err_msg="checkpoint of domain %d" % d)
    assert model.optimizer == "adam" and int(LOOPS["separate|compiles"][0]) == 0      # the compiled optimizer is never replaced
