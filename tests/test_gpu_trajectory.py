"""-m gpu: parity of the benched mode (tf32x3 pass kernel) and of the fp32 mode over WHOLE meta-steps.

Free-running training is chaotic on both sides: a ReLU gate whose pre-activation lies within rounding distance of zero
opens on one implementation and closes on the other, which changes one Adam-normalised update discretely, and the two
trajectories then separate exponentially.  These tests therefore split the claim in two:

* teacher-forced: before EVERY pass of a meta-step the oracle's state (weights, Adam slots, beta powers, global step)
  is loaded into the device model; after the pass the device state must agree with the oracle's to a per-pass bound.
  A pass that exceeds the clean bound must coincide with a recorded near-zero pre-activation in the oracle's own
  forward passes (the counted gate diagnostic), and must still stay within a loose bound.
* free-running at the FULL Taobao-10 size (scale 1.0, sample_num 5 + query, ~1 160 mini-batches per meta-step):
  theta / theta_d / AUC against OracleMAMDR after one whole meta-step.

Reference loop: /root/reference/model_zoo/mamdr.py:41-116.
"""
import numpy as np
import pytest
import torch

from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule
import oracle.meta as ometa
from oracle.meta import OracleMAMDR
from oracle.mlp import MLPSpec, OracleMLP

pytestmark = pytest.mark.gpu

GATE_EPS = 5e-7        # |pre-activation| below this can round to either side of the ReLU (fp32 / 3xTF32 sums of 256 terms of O(0.1))
CLEAN_TOL = {"tf32x3": {"kernel": 1e-5, "small": 1e-4}, "fp32": {"kernel": 1e-5, "small": 1e-4}}
LOOSE_TOL = 5e-3       # a pass in which a gate flipped


def _build(config):
    import run
    return run.build(config)


def _oracle(base, weights, dtype=np.float32):
    mc = base.model_config
    spec = MLPSpec(base.n_uid, base.n_pid, base.n_domain, (mc['user_dim'], mc['item_dim'], mc['domain_dim']),
                   tuple(mc['hidden_dim']), dropout=mc['dropout'])
    return OracleMLP(spec, weights, base.dataset.user_table, base.dataset.item_table, lr=base.train_config['learning_rate'],
                     dtype=dtype)


def _record_oracle_meta_step(om):
    """Runs one oracle meta-step and records, per training pass: the state before, the state after, the order and the
    smallest surviving |pre-activation| seen inside the pass."""
    passes = []
    orig = ometa.train_pass
    model = om.model

    def tp(mdl, d, domain, order, batch_size, max_steps=0, optimizer='adam', sgd_lr=None):
        ad = mdl.adam
        rec = {"domain": domain, "order": np.asarray(order).copy(), "max_steps": max_steps,
               "w0": [x.copy() for x in mdl.weights], "m0": [x.copy() for x in ad.m], "v0": [x.copy() for x in ad.v],
               "opt0": (ad.step, float(ad.b1pow), float(ad.b2pow))}
        model.preact_log = []
        r = orig(mdl, d, domain, order, batch_size, max_steps, optimizer, sgd_lr)
        rec["min_preact"] = min(model.preact_log) if model.preact_log else np.inf
        rec["near_zero"] = int(np.sum(np.asarray(model.preact_log) < GATE_EPS))
        model.preact_log = None
        rec.update({"w1": [x.copy() for x in mdl.weights], "m1": [x.copy() for x in ad.m], "v1": [x.copy() for x in ad.v],
                    "opt1": (ad.step, float(ad.b1pow), float(ad.b2pow)), "steps": r[2]})
        passes.append(rec)
        return r
    ometa.train_pass = tp
    try:
        om.train_epoch()
    finally:
        ometa.train_pass = orig
    return passes


@pytest.mark.parametrize("prec", ["tf32x3", "fp32"])
def test_teacher_forced_meta_step_tracks_the_oracle_pass_by_pass(prec):
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 0.25, "b200.precision": prec,
                       "train.sample_num": 2})
    w = _build(c)
    w.prepare()
    base, m = w.base_model, w.base_model.model
    lo = m.layout
    om = OracleMAMDR(_oracle(base, w.meta_weights.numpy()), base.dataset.host_splits(), c['train'], base.dataset.batch_size,
                     Schedule(123), {k: v.numpy() for k, v in w.domain_weights.items()}, name=c['model']['name'])
    passes = _record_oracle_meta_step(om)
    assert len(passes) >= 10 + 10 * 3 * 2
    tol = CLEAN_TOL[prec]
    worst_clean, flagged, n_steps = 0.0, [], 0
    for k, rec in enumerate(passes):
        data = base.dataset.train_dataset[rec["domain"]]['data']
        # teacher forcing: the oracle's state before the pass
        m.params.copy_(torch.from_numpy(lo.pack(rec["w0"])))
        m.m.copy_(torch.from_numpy(lo.pack(rec["m0"])))
        m.v.copy_(torch.from_numpy(lo.pack(rec["v0"])))
        m.set_opt_words(torch.tensor(rec["opt0"], dtype=torch.float64))
        data.set_order(rec["order"])
        steps = rec["steps"]
        m.fit_pass(data, steps)
        torch.cuda.synchronize()
        n_steps += steps
        assert m.read_step()[0] == rec["opt1"][0]
        got = lo.unpack(m.params.cpu().numpy())
        errs = {}
        for n_, a, b in zip(lo.names, got, rec["w1"]):
            errs[n_] = rel_err(a, b)
        bad = [n_ for n_, e in errs.items() if e > (tol["kernel"] if n_.startswith("kernel") else tol["small"])]
        if bad:
            # the counted gate diagnostic: a divergence event must coincide with a near-zero pre-activation in the oracle
            assert rec["near_zero"] > 0, ("pass %d diverged without a near-zero pre-activation" % k, errs, rec["min_preact"])
            assert max(errs.values()) < LOOSE_TOL, (k, errs)
            flagged.append((k, max(errs.values()), rec["min_preact"]))
        else:
            worst_clean = max(worst_clean, max(e for n_, e in errs.items() if n_.startswith("kernel")))
    n_exposed = sum(1 for r in passes if r["near_zero"] > 0)
    print("%s: %d passes / %d mini-batches teacher-forced; worst clean kernel error %.2e; %d passes saw a pre-activation below %.0e; "
          "gate events (pass, error, min |pre-activation|) %s" % (prec, len(passes), n_steps, worst_clean, n_exposed, GATE_EPS, flagged))
    # gate events are rare: the stated per-pass bound holds for (almost) every pass
    assert len(flagged) <= max(2, len(passes) // 10), flagged


@pytest.mark.parametrize("prec", ["tf32x3", "fp32"])
def test_full_size_taobao10_meta_step_against_the_oracle(prec):
    """ONE meta-step of config #1 at its full size (scale 1.0, batch 1024, sample_num 5 + the query domain, ~1 160
    mini-batches): the free-running device run against OracleMAMDR on the same schedule.

    Free-running over ~1 160 mini-batches BOTH modes leave the oracle's trajectory at ReLU-gate events, and so does the oracle
    against ITSELF: the numpy / torch-CPU oracle run with 1 and with 4 BLAS threads (same fp32 arithmetic, another summation
    order) ends a scale-0.25 meta-step 3e-3 apart on the theta_d kernels and 6e-2 apart on domain_emb
    (tests/test_oracle_chaos.py).  With 16 BLAS threads the oracle happened to stay on the fp32 SIMT kernels' side of every gate
    of the DN phase (theta within 1.1e-6); with the 4 threads the suite pins it does not (5e-3) -- neither is "the" answer.
    Per pass both modes agree with the oracle to ~1e-6 (teacher-forced test above).  Stated bars for both modes: kernels 2e-2
    (theta) / 5e-2 (theta_d), AUC within 1e-3 average / 2e-3 per domain (north_star: AUC within 1e-3)."""
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 1.0, "b200.precision": prec,
                       "train.sample_num": 5})
    w = _build(c)
    w.prepare()
    base, m = w.base_model, w.base_model.model
    lo = m.layout
    om = OracleMAMDR(_oracle(base, w.meta_weights.numpy()), base.dataset.host_splits(), c['train'], base.dataset.batch_size,
                     Schedule(123), {k: v.numpy() for k, v in w.domain_weights.items()}, name=c['model']['name'])
    base.schedule = Schedule(123)
    w.train_epoch(0)
    om.train_epoch()
    torch.cuda.synchronize()
    theta = w.meta_weights.numpy()
    errs = {n_: rel_err(a, b) for n_, a, b in zip(lo.names, theta, om.meta_weights)}
    derr = {}
    for d in range(base.n_domain):
        for n_, a, b in zip(lo.names, w.domain_weights[d].numpy(), om.domain_weights[d]):
            derr[n_] = max(derr.get(n_, 0.0), rel_err(a, b))
    _, g_auc, _, g_dom = w.val_and_test("val")
    _, o_auc, _, o_dom = om.val_and_test("val")
    dauc = max(abs(g_dom[d] - o_dom[d]) for d in g_dom)
    print(prec, "full-size meta-step: theta", {k: "%.1e" % v for k, v in errs.items()}, "theta_d (max over domains)",
          {k: "%.1e" % v for k, v in derr.items()}, "avg AUC %.6f vs %.6f, max per-domain AUC difference %.1e" % (g_auc, o_auc, dauc))
    for n_, e in errs.items():
        if n_.startswith("kernel") or n_ == "dense_kernel":
            assert e < 2e-2, (n_, e)
    for n_, e in derr.items():
        if n_.startswith("kernel") or n_ == "dense_kernel":
            assert e < 5e-2, ("theta_d", n_, e)
    assert abs(g_auc - o_auc) < 1e-3, (g_auc, o_auc)
    assert dauc < 2e-3, dauc
