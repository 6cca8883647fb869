"""-m gpu: the CUDA path against the COMMITTED golden vectors (tests/golden/kernels_v1.npz; generator:
tests/golden/make_golden.py) through the C-ABI -- bit-exact for gather, de-duplication, Adam, the DN / DR sweeps and the AUC
counts; loss 2e-5 and every gradient tensor 3e-5 vs the float64 records for one mlp and one MMOE train step."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu

from gpu_util import bits, ctx, dev, ptr, stream  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as mg  # noqa: E402  (problem definitions only; the expected values come from the committed file)

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "kernels_v1.npz"))


def test_golden_gather_dedup_adam_meta_auc_bit_exact():
    c, lib = ctx(), ctx().lib
    # K1
    t, i = dev(GOLD['gather_table']), dev(GOLD['gather_ids'])
    out = torch.zeros(37, 8, device="cuda")
    c.call("mamdr_gather_f32", ptr(t), 50, 8, ptr(i), 37, ptr(out), 8, stream())
    np.testing.assert_array_equal(bits(out.cpu().numpy()), bits(GOLD['gather_out']))
    # K6
    n = 200
    ws = torch.zeros(lib.mamdr_scatter_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    uo, ro, nu = torch.full((n,), -1, dtype=torch.int32, device="cuda"), torch.zeros(n, 8, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
    d_ids, d_rows = dev(GOLD['dedup_ids']), dev(GOLD['dedup_rows'])
    c.call("mamdr_scatter_dedup_f32", ptr(d_ids), ptr(d_rows), 8, n, 8, ptr(uo), ptr(ro), ptr(nu), ptr(ws), ws.numel(), stream())
    k = int(nu.item())
    assert k == len(GOLD['dedup_uniq'])
    np.testing.assert_array_equal(uo[:k].cpu().numpy(), GOLD['dedup_uniq'])
    np.testing.assert_array_equal(bits(ro[:k].cpu().numpy()), bits(GOLD['dedup_sums']))
    # K7
    p, m, v = dev(GOLD['adam_p0']), torch.zeros(260, device="cuda"), torch.zeros(260, device="cuda")
    state = torch.zeros(lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device="cuda")
    c.call("mamdr_opt_state_init", ptr(state), 0.9, 0.999, stream())
    for tstep in range(5):
        d_g = dev(GOLD['adam_g'][tstep])
        c.call("mamdr_adam_step", ptr(p), ptr(m), ptr(v), ptr(d_g), 260, ptr(state), 1e-3, 0.9, 0.999, 1e-8, stream())
    for got, key in ((p, 'adam_p'), (m, 'adam_m'), (v, 'adam_v')):
        np.testing.assert_array_equal(bits(got.cpu().numpy()), bits(GOLD[key]), err_msg=key)
    step, b1, b2 = C.c_int64(), C.c_float(), C.c_float()
    c.call("mamdr_opt_state_read", ptr(state), C.byref(step), C.byref(b1), C.byref(b2), stream())
    assert step.value == 5 and np.float32(b1.value) == GOLD['adam_pows'][0] and np.float32(b2.value) == GOLD['adam_pows'][1]
    # K9 / K10
    d_th = dev(GOLD['meta_th'])
    th, mo = dev(GOLD['meta_th']), dev(GOLD['meta_mo'])
    c.call("mamdr_dn_update", ptr(th), ptr(mo), 0.1, 256, ptr(mo), stream())
    np.testing.assert_array_equal(bits(th.cpu().numpy()), bits(GOLD['meta_dn']))
    np.testing.assert_array_equal(bits(mo.cpu().numpy()), bits(GOLD['meta_dn']))
    for method, key in ((0, 'meta_dr_plus'), (1, 'meta_dr_times')):
        ti, mo = dev(GOLD['meta_ti']), dev(GOLD['meta_mo'])
        c.call("mamdr_dr_update", ptr(ti), ptr(d_th), ptr(mo), 0.1, 256, method, None, stream())
        np.testing.assert_array_equal(bits(ti.cpu().numpy()), bits(GOLD[key]), err_msg=key)
    # K8
    from mamdr_b200.auc import thresholds
    acc = torch.zeros(4, 3, device="cuda")
    d_thr, d_p, d_y = dev(thresholds(3)), dev(GOLD['auc_kat_p']), dev(GOLD['auc_kat_y'])
    c.call("mamdr_auc_update", ptr(d_p), ptr(d_y), 4, ptr(acc), ptr(d_thr), 3, stream())
    np.testing.assert_array_equal(acc.cpu().numpy(), GOLD['auc_kat_acc'])
    acc = torch.zeros(4, 500, device="cuda")
    res = torch.zeros(1, device="cuda")
    d_thr, d_p, d_y = dev(thresholds(500)), dev(GOLD['auc_p']), dev(GOLD['auc_y'])
    c.call("mamdr_auc_update", ptr(d_p), ptr(d_y), 300, ptr(acc), ptr(d_thr), 500, stream())
    np.testing.assert_array_equal(acc.cpu().numpy(), GOLD['auc_acc'])
    c.call("mamdr_auc_result", ptr(acc), 500, ptr(res), stream())
    assert abs(res.item() - float(GOLD['auc_result'])) < 2e-6


def test_golden_mlp_train_step():
    from mamdr_b200 import _lib
    from mamdr_b200.engine import DomainData, MLPModel
    lo, w, ut, it, uid, pid, y = mg.mlp_problem()
    M = mg.MLP
    m = MLPModel(M['n_uid'], M['n_pid'], M['n_domain'], emb_dim=M['emb_dim'], hidden=M['hidden'], dropout=0.5, dropout_seed=1024,
                 l2_emb=1e-5, emb_trainable=False, user_table=ut, item_table=it, init_weights=w, lr=1e-3, max_batch=64,
                 precision=_lib.PREC_FP32, use_graphs=False)
    assert m.layout.names == lo.names
    data = DomainData(uid, pid, y, M['domain'], 64, m.device)
    loss = torch.zeros(1, device="cuda")
    probs = torch.zeros(M['rows'], device="cuda")
    m._train_step(data, 0, M['rows'], loss, probs=probs)
    torch.cuda.synchronize()
    assert abs(loss.item() - GOLD['mlp_loss']) < 2e-5 * abs(GOLD['mlp_loss'])
    np.testing.assert_allclose(probs.cpu().numpy(), GOLD['mlp_p'], rtol=2e-5, atol=1e-7)
    for n, a, b in zip(lo.names, lo.unpack(m.grads.cpu().numpy()), lo.unpack(GOLD['mlp_grads'])):
        assert rel_err(a, b) < 3e-5, (n, rel_err(a, b))


def test_golden_mmoe_train_step():
    from mamdr_b200.deep_mtl_ctr import MTLModel
    from mamdr_b200.engine import DomainData
    topo, w, uid, pid, y = mg.mtl_problem()
    T = mg.MTL
    m = MTLModel(topo, w, dropout=0.5, dropout_seed=1024, l2_emb=1e-5, lr=1e-3, max_batch=64, use_graphs=False)
    data = DomainData(uid, pid, y, T['domain'], 64, m.device)
    loss = torch.zeros(1, device="cuda")
    probs = torch.zeros(T['rows'], device="cuda")
    m._train_step(data, 0, T['rows'], loss, probs=probs)
    torch.cuda.synchronize()
    # the loss slot also holds the tables' l2 term, added by the fused table sweeps, exactly as the record does
    assert abs(loss.item() - GOLD['mtl_loss']) < 2e-5 * abs(GOLD['mtl_loss'])
    np.testing.assert_allclose(probs.cpu().numpy(), GOLD['mtl_p'], rtol=2e-5, atol=1e-7)
    lo = topo.layout
    gold = dict(zip(lo.names, lo.unpack(GOLD['mtl_grads'])))
    got = dict(zip(lo.names, lo.unpack(m.grads.cpu().numpy())))
    for n in topo.reachable(T['domain']):
        if n in ('user_emb', 'item_emb'):
            continue        # sparse: applied by the table sweep, checked through the ids below and in tests/test_gpu_mtl.py
        assert rel_err(got[n], gold[n]) < 3e-5, (n, rel_err(got[n], gold[n]))
    for ti, col in enumerate((uid, pid)):
        ids, srows, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        assert m.ctx.lib.mamdr_mtl_sparse_grads(C.byref(m.desc), T['rows'], C.c_void_p(m.ws.data_ptr()), ti, C.byref(ids), C.byref(srows),
                                                C.byref(cnt)) == 0
        off, noff = ids.value - m.ws.data_ptr(), cnt.value - m.ws.data_ptr()
        n_u = int(m.ws[noff:noff + 4].view(torch.int32).item())
        np.testing.assert_array_equal(m.ws[off:off + 4 * n_u].view(torch.int32).cpu().numpy(), np.unique(col))


def test_cuda_meta_sweeps_match_the_reference_executed_vectors():
    """tests/golden/reference_meta_v1.npz was produced by EXECUTING the reference's own numpy algebra (mamdr.py:168-196,
    domain_negotiation.py:118-123, specific_base_model.py:164-172, reptile.py:127-142; generator: make_reference_golden.py).
    The K9 / K10 sweeps reproduce it bit for bit."""
    ref = np.load(os.path.join(ROOT, "tests", "golden", "reference_meta_v1.npz"))
    c = ctx()
    n_real = ref["theta"].size
    n = (n_real + 3) // 4 * 4

    def pad(a):
        out = np.zeros(n, dtype=np.float32)
        out[:a.size] = a
        return out

    def same(t, key):
        np.testing.assert_array_equal(bits(t.cpu().numpy()[:n_real]), bits(ref[key]), err_msg=key)

    th, ti, mo = pad(ref["theta"]), pad(ref["theta_i"]), pad(ref["model"])
    d_th, d_ti, d_mo = dev(th), dev(ti), dev(mo)
    t = dev(th)
    c.call("mamdr_dn_update", ptr(t), ptr(d_mo), 0.1, n, None, stream())
    same(t, "dn_update")
    same(t, "mamdr_dn_form")
    same(t, "reptile_update")
    for method, name in ((0, "plus"), (1, "times")):
        merged = torch.zeros(n, device="cuda")
        c.call("mamdr_merge", ptr(merged), ptr(d_th), ptr(d_ti), n, method, stream())
        same(merged, "merge_" + name)
        x = dev(ti)
        c.call("mamdr_dr_update", ptr(x), ptr(d_th), ptr(d_mo), 0.1, n, method, None, stream())
        same(x, "dr_update_" + name)
        acc = dev(pad(ref["accum0"]))
        c.call("mamdr_dr_accumulate", ptr(acc), ptr(d_mo), ptr(d_th), ptr(d_ti), n, method, stream())
        same(acc, "accumulate_" + name)
        x = dev(ti)
        c.call("mamdr_dr_apply_accum", ptr(x), ptr(acc), 5.0, 0.1, n, stream())
        same(x, "apply_accum_" + name)
        assert float(acc.abs().max()) == 0.0
        out = torch.zeros(n, device="cuda")
        c.call("mamdr_sub", ptr(out), ptr(d_mo), ptr(merged), n, stream())
        same(out, "update_domain_weights_" + name)
    # Reptile, batch names: three deltas against the same theta, applied once
    acc, zeros = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for k in range(3):
        shifted = dev(pad(ref["model"] + np.float32(0.25 * k)))
        c.call("mamdr_axpy_diff", ptr(acc), ptr(shifted), ptr(d_th), 1.0, n, stream())
        torch.cuda.synchronize()
    same(acc, "reptile_accum3")
    t = dev(th)
    c.call("mamdr_axpy_diff", ptr(t), ptr(acc), ptr(zeros), 0.1, n, stream())
    same(t, "reptile_apply")


def test_cuda_auc_matches_the_reference_executed_metric():
    """tests/golden/reference_auc_v1.npz: the reference's own utils/auc.py + utils/metrics_utils.py executed on a four-batch
    stream (generator: make_reference_golden.py).  `mamdr_auc_update` reproduces every accumulator exactly, `mamdr_auc_result`
    the interpolated ROC-AUC within float32 summation order."""
    from mamdr_b200.auc import thresholds
    ref = np.load(os.path.join(ROOT, "tests", "golden", "reference_auc_v1.npz"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_reference_golden as mrg
    c = ctx()
    acc = torch.zeros(4, 500, device="cuda")
    res = torch.zeros(1, device="cuda")
    d_thr = dev(thresholds(500))
    for k, (rows, seed) in enumerate(mrg.AUC_STREAM):
        y, p = mrg.auc_batch(rows, seed)
        d_p, d_y = dev(p), dev(y)
        c.call("mamdr_auc_update", ptr(d_p), ptr(d_y), rows, ptr(acc), ptr(d_thr), 500, stream())
        np.testing.assert_array_equal(acc.cpu().numpy(), ref["auc|T500|acc_after_%d" % k], err_msg="batch %d" % k)
        c.call("mamdr_auc_result", ptr(acc), 500, ptr(res), stream())
        assert abs(res.item() - float(ref["auc|T500|result_after_%d" % k])) < 2e-6
