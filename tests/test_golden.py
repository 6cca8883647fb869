"""CPU: the committed golden vectors (tests/golden/kernels_v1.npz, written by tests/golden/make_golden.py) are reproduced by
the oracle -- exactly for the integer / element-wise primitives, to 1e-10 for the float64 train-step records -- and the fp32
oracle agrees with the float64 records within the single-batch tolerance."""
import os
import sys

import numpy as np

from conftest import ROOT, rel_err

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "kernels_v1.npz"))


def test_oracle_reproduces_the_golden_vectors():
    fresh = mg.make()
    assert sorted(fresh) == sorted(GOLD.files)
    for k in GOLD.files:
        a, b = np.asarray(fresh[k]), GOLD[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if a.dtype == np.float64:
            assert rel_err(a, b) < 1e-10, (k, rel_err(a, b))
        else:
            np.testing.assert_array_equal(a.view(np.uint32) if a.dtype == np.float32 else a,
                                          b.view(np.uint32) if b.dtype == np.float32 else b, err_msg=k)


def test_reference_auc_docstring_example_is_in_the_goldens():
    """utils/auc.py:44-56: the only known-answer vector the reference ships."""
    from oracle import auc as oauc
    a = oauc.AUC(3)
    a.update_state(GOLD['auc_kat_y'], GOLD['auc_kat_p'])
    np.testing.assert_array_equal(np.asarray(a.acc, dtype=np.float32), GOLD['auc_kat_acc'])
    assert abs(a.result() - 0.75) < 1e-7 and GOLD['auc_kat_result'] == np.float32(0.75)


def test_c_philox_matches_the_golden_mask():
    from oracle import philox
    np.testing.assert_array_equal(philox.dropout_mask(8, 16, 1030, 3, 0.5, use_c=True), GOLD['philox_mask'])
    assert set(np.unique(GOLD['philox_mask'])) <= {0.0, 2.0}


def test_fp32_oracle_within_tolerance_of_the_float64_records():
    from oracle.mlp import MLPSpec, OracleMLP
    from oracle.mtl import MTLSpec, OracleMTL
    lo, w, ut, it, uid, pid, y = mg.mlp_problem()
    M = mg.MLP
    o = OracleMLP(MLPSpec(M['n_uid'], M['n_pid'], M['n_domain'], M['emb_dim'], M['hidden'], dropout=0.5), w, ut, it, lr=1e-3)
    loss, p, grads = o.gradients(uid, pid, M['domain'], y)
    assert abs(loss - GOLD['mlp_loss']) < 2e-6 * abs(GOLD['mlp_loss'])
    for n, a, b in zip(lo.names, grads, lo.unpack(GOLD['mlp_grads'])):
        assert rel_err(a, b) < 2e-5, (n, rel_err(a, b))
    topo, w, uid, pid, y = mg.mtl_problem()
    T = mg.MTL
    spec = MTLSpec(T['n_uid'], T['n_pid'], T['n_domain'], kind='mmoe', emb_dim=T['emb_dim'], expert_hidden=T['expert_hidden'],
                   tower_hidden=T['tower_hidden'], gate_hidden=T['gate_hidden'], num_experts=T['num_experts'], dropout=0.5, emb_trainable=True)
    o = OracleMTL(spec, w, None, None, lr=1e-3)
    loss, p, gd = o.gradients(uid, pid, T['domain'], y)
    assert abs(loss - GOLD['mtl_loss']) < 2e-6 * abs(GOLD['mtl_loss'])
    gold = dict(zip(topo.layout.names, topo.layout.unpack(GOLD['mtl_grads'])))
    for n in spec.reachable(T['domain']):
        assert rel_err(gd[n], gold[n]) < 3e-5, (n, rel_err(gd[n], gold[n]))
