"""DN / DR algebra of the oracle on a toy problem (closed forms from SURVEY.md A-7) and the schedule
call order of the reference loops."""
import copy

import numpy as np

from conftest import BASE_CONFIG
from mamdr_b200 import synth
from mamdr_b200.layout import init_mlp_weights, mlp_layout
from mamdr_b200.schedule import Schedule
from oracle.meta import OracleDN, OracleMAMDR, merge_weights, n_steps
from oracle.mlp import MLPSpec, OracleMLP


def _toy(D_scale=0.004, emb=8, hidden=(16, 8), dropout=0.5, seed=3):
    g = synth.generate("Taobao-10", seed=seed, scale=D_scale, emb_dim=emb)
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (emb, emb, emb), hidden, dropout=dropout)
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (emb, emb, emb), hidden, False)
    data = {k: g[k] for k in ("train", "val", "test")}
    return g, spec, lo, data


def test_n_steps_and_merge():
    assert n_steps(1024, 1024) == 1 and n_steps(1025, 1024) == 2 and n_steps(1, 1024) == 1
    a, b = [np.float32([1, 2])], [np.float32([3, 4])]
    np.testing.assert_array_equal(merge_weights(a, b, 'plus')[0], [4, 6])
    np.testing.assert_array_equal(merge_weights(a, b, 'times')[0], [3, 8])


def test_dn_outer_update_is_interpolation():
    g, spec, lo, data = _toy()
    tc = copy.deepcopy(BASE_CONFIG["train"])
    w0 = init_mlp_weights(lo, 1)
    m = OracleMLP(spec, w0, g["user_emb"], g["item_emb"])
    dn = OracleDN(m, data, tc, 64, Schedule(1))
    theta0 = [w.copy() for w in dn.meta_weights]
    dn.train_epoch()
    # replay the inner loop by hand with an identical schedule: theta1 = theta0 + beta (theta_K - theta0)
    m2 = OracleMLP(spec, w0, g["user_emb"], g["item_emb"])
    s2 = Schedule(1)
    seq = s2.shuffle_sequence(sorted(data["train"]))
    from oracle.meta import train_pass
    for idx in seq:
        d = data["train"][idx]
        train_pass(m2, d, idx, s2.batch_order(idx, len(d["uid"])), 64)
    for t0, t1, tk in zip(theta0, dn.meta_weights, m2.weights):
        np.testing.assert_array_equal(t1, t0 + (tk - t0) * np.float32(0.1))
    # the model was reloaded with theta (domain_negotiation.py:88) and Adam state kept running
    for a, b in zip(m.weights, dn.meta_weights):
        np.testing.assert_array_equal(a, b)
    assert m.adam.step == sum(n_steps(len(data["train"][d]["uid"]), 64) for d in data["train"])


def test_mamdr_epoch_structure_and_adam_never_reset():
    g, spec, lo, data = _toy()
    tc = copy.deepcopy(BASE_CONFIG["train"])
    tc["sample_num"] = 2
    w0 = init_mlp_weights(lo, 1)
    dw0 = {d: init_mlp_weights(lo, 10 + d) for d in range(10)}
    m = OracleMLP(spec, w0, g["user_emb"], g["item_emb"])
    mm = OracleMAMDR(m, data, tc, 64, Schedule(2), dw0)
    mm.train_epoch()
    S = {d: n_steps(len(data["train"][d]["uid"]), 64) for d in range(10)}
    # replicate the schedule to count the Adam steps: DN sum S_d + DR sum_i sum_{j in J_i} (S_j + S_i)
    s = Schedule(2)
    seq = s.shuffle_sequence(sorted(data["train"]))
    total = 0
    for idx in seq:
        total += S[idx]
        s.batch_order(idx, 1)
    for idx in seq:
        cands = list(seq)
        cands.remove(idx)
        aux = s.sample_support(cands, 2) + [idx]
        assert len(aux) == 3 and idx not in aux[:2]
        for j in aux:
            total += S[j] + S[idx]
            s.batch_order(j, 1)
            s.batch_order(idx, 1)
    assert m.adam.step == total            # one shared optimizer threads through every pass
    for d in range(10):                    # every theta_d moved away from its initialisation
        assert any(not np.array_equal(a, b) for a, b in zip(mm.domain_weights[d], dw0[d]))
    l, a, dl, da = mm.val_and_test("val")
    assert set(dl) == set(range(10)) and 0 <= a <= 1
    assert mm.early_stop_step(a) is False and mm.best_shared_weights is not None
    assert mm.early_stop_step(a) is False and mm.es.counter == 1     # "<=" counts as no improvement
    mm.early_stop_step(a)
    assert mm.early_stop_step(a) is True                             # patience 3


def test_dr_update_closed_form_plus_and_batch():
    g, spec, lo, data = _toy(dropout=0.0)
    tc = copy.deepcopy(BASE_CONFIG["train"])
    tc.update(sample_num=1, add_query_domain=False, shuffle_sequence=False)
    w0 = init_mlp_weights(lo, 1)
    dw0 = {d: [np.zeros_like(w) for w in w0] for d in range(10)}     # theta_d = 0  ->  merged = theta
    for name in ("mlp_meta_mamdr", "mlp_meta_mamdr_batch"):
        m = OracleMLP(spec, w0, g["user_emb"], g["item_emb"])
        mm = OracleMAMDR(m, data, tc, 64, Schedule(4), dw0, name=name)
        mm.train_epoch()
        # with one support pass per query domain both variants give theta_i = beta * (theta_tilde - merged)
        # (batch: accum / sample_num * beta with sample_num = 1); just check they agree with each other
        if name == "mlp_meta_mamdr":
            ref = copy.deepcopy(mm.domain_weights)
        else:
            for d in range(10):
                for a, b in zip(mm.domain_weights[d], ref[d]):
                    np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)


def test_reptile_oracle_algebra():
    """`Reptile.train` (reptile.py:45-99,127-142): the model restarts from theta for every domain; with one domain the
    `batch` variant equals the per-domain one; with beta = 1 and one domain theta becomes the trained weights."""
    import copy
    from oracle.meta import OracleReptile

    class Toy(object):   # a "model" whose pass adds (domain + 1) to every weight: makes the meta algebra visible
        def __init__(self):
            self.weights = [np.zeros(3, dtype=np.float32), np.ones((2, 2), dtype=np.float32)]

            class A(object):
                def reset_states(self):
                    pass
            self.auc = A()

        def get_weights(self):
            return [w.copy() for w in self.weights]

        def set_weights(self, ws):
            for a, b in zip(self.weights, ws):
                a[...] = b

        def train_on_batch(self, uid, pid, domain, label, optimizer='adam', sgd_lr=None):
            for w in self.weights:
                w += np.float32(domain + 1)
            return 0.0, 0.5

    data = {'train': {d: {'uid': np.zeros(5, np.int32), 'pid': np.zeros(5, np.int32), 'label': np.zeros(5, np.float32)} for d in range(3)}}
    tc = {'meta_learning_rate': 0.5, 'patience': 3, 'meta_train_step': 0}
    from mamdr_b200.schedule import Schedule
    r = OracleReptile(Toy(), copy.deepcopy(data), tc, 8, Schedule(1), name='mlp_meta_reptile')
    r.train_epoch()
    # sequential: theta_k = theta_{k-1} + 0.5 * ((theta_{k-1} + (d_k + 1)) - theta_{k-1}) = theta_{k-1} + 0.5 (d_k + 1), any order
    assert np.allclose(r.meta_weights[0], 0.5 * (1 + 2 + 3)) and np.allclose(r.meta_weights[1], 1 + 0.5 * 6)
    assert np.array_equal(r.model.weights[0], r.meta_weights[0])          # model <- theta at the end (:99)
    rb = OracleReptile(Toy(), copy.deepcopy(data), tc, 8, Schedule(1), name='mlp_meta_reptile_batch')
    rb.train_epoch()
    assert np.allclose(rb.meta_weights[0], 0.5 * 6) and all(not a.any() for a in rb.accum)   # summed deltas, cleared
    one = {'train': {0: data['train'][0]}}
    r1 = OracleReptile(Toy(), one, dict(tc, meta_learning_rate=1.0), 8, Schedule(1))
    r1.train_epoch()
    assert np.allclose(r1.meta_weights[0], 1.0)
