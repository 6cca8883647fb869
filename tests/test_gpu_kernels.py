"""-m gpu: bit-exact parity of the integer / element-wise kernels against the oracle, through the C-ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import auc as oauc
from oracle.mlp import AdamState

pytestmark = pytest.mark.gpu

from gpu_util import bits, ctx, dev, ptr, stream  # noqa: E402


# ---- K1 gather -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,dim,n", [(6932, 128, 1024), (1000, 128, 977), (50, 4, 1), (300, 36, 4097),
                                        (23778, 128, 100000)])
def test_gather_bit_exact(rows, dim, n):
    rng = np.random.default_rng(rows + n)
    table = rng.standard_normal((rows, dim)).astype(np.float32)
    table[rng.integers(0, rows, 8)] = np.float32(np.nan)  # payload bits must survive
    ids = rng.integers(0, rows, n).astype(np.int32)
    ids[: min(n, 16)] = ids[0]  # duplicates
    t, i = dev(table), dev(ids)
    out = torch.full((n, dim + 4), -7.0, device="cuda")
    ctx().call("mamdr_gather_f32", ptr(t), rows, dim, ptr(i), n, ptr(out), dim + 4, stream())
    got = out.cpu().numpy()
    np.testing.assert_array_equal(bits(got[:, :dim]), bits(table[ids]))
    assert np.all(got[:, dim:] == -7.0)  # the stride padding is untouched


def test_gather_empty_and_errors():
    from mamdr_b200._lib import MamdrError
    t = torch.zeros(4, 8, device="cuda")
    i = torch.zeros(1, dtype=torch.int32, device="cuda")
    o = torch.zeros(1, 8, device="cuda")
    ctx().call("mamdr_gather_f32", ptr(t), 4, 8, ptr(i), 0, ptr(o), 8, stream())  # n == 0 is a no-op
    with pytest.raises(MamdrError):
        ctx().call("mamdr_gather_f32", ptr(t), 4, 6, ptr(i), 1, ptr(o), 8, stream())  # dim % 4
    with pytest.raises(MamdrError):
        ctx().call("mamdr_gather_f32", None, 4, 8, ptr(i), 1, ptr(o), 8, stream())
    with pytest.raises(MamdrError):
        ctx().call("mamdr_gather_f32", ptr(t), 4, 8, ptr(i), 1, ptr(o), 4, stream())  # stride < dim


# ---- K6 dedup ----------------------------------------------------------------------------------------------
def _oracle_dedup(ids, rows):
    uniq = np.unique(ids)
    out = np.zeros((len(uniq), rows.shape[1]), dtype=np.float32)
    first = np.ones(len(uniq), bool)
    pos = {int(u): k for k, u in enumerate(uniq)}
    for i, r in zip(ids, rows):      # sequential adds in batch order == np.add.at order
        k = pos[int(i)]
        out[k] = r if first[k] else (out[k] + r).astype(np.float32)
        first[k] = False
    return uniq.astype(np.int32), out


@pytest.mark.parametrize("n,n_ids,dim", [(1024, 200, 128), (977, 100000, 128), (1, 5, 8), (4096, 3, 64),
                                         (8192, 5000, 128)])
def test_scatter_dedup_bit_exact(n, n_ids, dim):
    rng = np.random.default_rng(n)
    ids = rng.integers(0, n_ids, n).astype(np.int32)
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    lib = ctx().lib
    ws = torch.zeros(lib.mamdr_scatter_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    uo = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    ro = torch.zeros(n, dim, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_ids, d_rows = dev(ids), dev(rows)   # keep the device inputs alive across the async call
    ctx().call("mamdr_scatter_dedup_f32", ptr(d_ids), ptr(d_rows), dim, n, dim, ptr(uo), ptr(ro), ptr(nu),
               ptr(ws), ws.numel(), stream())
    k = int(nu.item())
    eu, er = _oracle_dedup(ids, rows)
    assert k == len(eu)
    np.testing.assert_array_equal(uo[:k].cpu().numpy(), eu)           # sorted unique ids, bit-exact
    np.testing.assert_array_equal(bits(ro[:k].cpu().numpy()), bits(er))


def _oracle_dedup_windowed(ids, rows, win=256, group=32):
    """numpy restatement of the large-n summation order (include/mamdr_b200.h: mamdr_scatter_dedup_large_f32): stable sort by
    id; windows of `win` consecutive sorted positions; inside a window the rows of an id are added sequentially in batch order;
    the window partials of an id are added in window order inside groups of `group` windows, then the group sums in order.
    Negative ids are padding."""
    keep = np.nonzero(ids >= 0)[0]
    order = keep[np.argsort(ids[keep], kind="stable")]
    sid = ids[order]
    uniq, start = np.unique(sid, return_index=True)
    end = np.append(start[1:], len(sid))
    out = np.zeros((len(uniq), rows.shape[1]), dtype=np.float32)
    for k, (s0, e0) in enumerate(zip(start, end)):
        pieces = []
        w = s0 // win
        while w * win < e0:
            a, b = max(s0, w * win), min(e0, (w + 1) * win)
            part = rows[order[a]].copy()
            for p in order[a + 1:b]:
                part = (part + rows[p]).astype(np.float32)
            pieces.append(part)
            w += 1
        groups = []
        for g0 in range(0, len(pieces), group):
            acc = pieces[g0]
            for part in pieces[g0 + 1:g0 + group]:
                acc = (acc + part).astype(np.float32)
            groups.append(acc)
        acc = groups[0]
        for gsum in groups[1:]:
            acc = (acc + gsum).astype(np.float32)
        out[k] = acc
    return uniq.astype(np.int32), out


@pytest.mark.parametrize("n,n_ids,dim,zipf", [(20000, 3000, 128, True), (8193, 50, 64, False), (70001, 100000, 32, True), (300000, 5, 8, False)])
def test_scatter_dedup_large_bit_exact(n, n_ids, dim, zipf):
    """The multi-CTA path (n beyond the single-CTA limit): unique ids bit-exact, sums bit-exact against the numpy restatement of
    its windowed order; hot ids (Zipf draws / 5 ids over 300 000 rows) span hundreds of windows; padding ids are skipped."""
    rng = np.random.default_rng(n)
    if zipf:
        ids = np.minimum(rng.zipf(1.2, n) - 1, n_ids - 1).astype(np.int32)
    else:
        ids = rng.integers(0, n_ids, n).astype(np.int32)
    ids[rng.integers(0, n, n // 50)] = -1            # padding entries
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    lib = ctx().lib
    assert n > lib.mamdr_scatter_max_n()
    ws = torch.zeros(lib.mamdr_scatter_large_workspace_bytes(n, dim), dtype=torch.uint8, device="cuda")
    uo = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    ro = torch.zeros(n, dim, device="cuda")
    nu = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_ids, d_rows = dev(ids), dev(rows)
    ctx().call("mamdr_scatter_dedup_large_f32", ptr(d_ids), ptr(d_rows), dim, n, dim, ptr(uo), ptr(ro), ptr(nu),
               ptr(ws), ws.numel(), stream())
    k = int(nu.item())
    eu, er = _oracle_dedup_windowed(ids, rows)
    assert k == len(eu)
    np.testing.assert_array_equal(uo[:k].cpu().numpy(), eu)
    np.testing.assert_array_equal(bits(ro[:k].cpu().numpy()), bits(er))
    # linearity / conservation: the de-duplicated rows sum to the sum of all real rows (fp64 check, size independent)
    tot = rows[ids >= 0].astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(ro[:k].cpu().numpy().astype(np.float64).sum(axis=0), tot, rtol=1e-4, atol=1e-2)


# ---- K7 Adam / SGD --------------------------------------------------------------------------------------------
def test_adam_bit_exact_over_steps():
    rng = np.random.default_rng(3)
    n = 4096 + 32
    w = [rng.standard_normal(n).astype(np.float32)]
    st = AdamState(w, lr=1e-3)
    lib = ctx().lib
    p, m, v = dev(w[0]), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    state = torch.zeros(lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device="cuda")
    ctx().call("mamdr_opt_state_init", ptr(state), 0.9, 0.999, stream())
    for t in range(25):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 1)).astype(np.float32)
        g[:7] = 0.0
        st.apply(w, [g])
        d_g = dev(g)
        ctx().call("mamdr_adam_step", ptr(p), ptr(m), ptr(v), ptr(d_g), n, ptr(state), 1e-3, 0.9, 0.999, 1e-8,
                   stream())
        np.testing.assert_array_equal(bits(p.cpu().numpy()), bits(w[0]), err_msg="step %d" % t)
    np.testing.assert_array_equal(bits(m.cpu().numpy()), bits(st.m[0]))
    np.testing.assert_array_equal(bits(v.cpu().numpy()), bits(st.v[0]))
    step, b1, b2 = C.c_int64(), C.c_float(), C.c_float()
    ctx().call("mamdr_opt_state_read", ptr(state), C.byref(step), C.byref(b1), C.byref(b2), stream())
    assert step.value == 25 and np.float32(b1.value) == st.b1pow and np.float32(b2.value) == st.b2pow


def test_sgd_bit_exact():
    rng = np.random.default_rng(4)
    n = 1024
    w = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal(n).astype(np.float32)
    lib = ctx().lib
    state = torch.zeros(lib.mamdr_opt_state_bytes(), dtype=torch.uint8, device="cuda")
    ctx().call("mamdr_opt_state_init", ptr(state), 0.9, 0.999, stream())
    p, d_g = dev(w), dev(g)
    ctx().call("mamdr_sgd_step", ptr(p), ptr(d_g), n, ptr(state), 0.001, stream())
    np.testing.assert_array_equal(bits(p.cpu().numpy()), bits(w - g * np.float32(0.001)))


# ---- K9 / K10 meta ops ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", [0, 1])
def test_meta_ops_bit_exact(method):
    rng = np.random.default_rng(5)
    n = 141088
    th, ti, mo = (rng.standard_normal(n).astype(np.float32) for _ in range(3))
    beta = np.float32(0.1)
    mg = (lambda a, b: a + b) if method == 0 else (lambda a, b: a * b)
    c = ctx()
    d_th, d_ti, d_mo = dev(th), dev(ti), dev(mo)   # read-only device copies kept alive for the whole test
    # merge
    out = torch.zeros(n, device="cuda")
    c.call("mamdr_merge", ptr(out), ptr(d_th), ptr(d_ti), n, method, stream())
    np.testing.assert_array_equal(bits(out.cpu().numpy()), bits(mg(th, ti)))
    # DN: theta += (model - theta) * beta ; model <- theta
    t, m_ = dev(th), dev(mo)
    c.call("mamdr_dn_update", ptr(t), ptr(m_), 0.1, n, ptr(m_), stream())
    exp = th + (mo - th) * beta
    np.testing.assert_array_equal(bits(t.cpu().numpy()), bits(exp))
    np.testing.assert_array_equal(bits(m_.cpu().numpy()), bits(exp))
    # DR: theta_i += (model - merged) * beta ; model <- theta (+|*) theta_i
    tii, m_ = dev(ti), dev(mo)
    c.call("mamdr_dr_update", ptr(tii), ptr(d_th), ptr(m_), 0.1, n, method, ptr(m_), stream())
    nti = ti + (mo - mg(th, ti)) * beta
    np.testing.assert_array_equal(bits(tii.cpu().numpy()), bits(nti))
    np.testing.assert_array_equal(bits(m_.cpu().numpy()), bits(mg(th, nti)))
    # batch variant: accumulate + apply
    acc0 = rng.standard_normal(n).astype(np.float32)
    acc = dev(acc0)
    c.call("mamdr_dr_accumulate", ptr(acc), ptr(d_mo), ptr(d_th), ptr(d_ti), n, method, stream())
    d = mo - mg(th, ti)
    eacc = acc0 + (d if method == 0 else d * th)
    np.testing.assert_array_equal(bits(acc.cpu().numpy()), bits(eacc))
    tii = dev(ti)
    c.call("mamdr_dr_apply_accum", ptr(tii), ptr(acc), 5.0, 0.1, n, stream())
    np.testing.assert_array_equal(bits(tii.cpu().numpy()), bits(ti + eacc / 5 * beta))
    assert float(acc.abs().max()) == 0.0
    # sub, axpy_diff, copy
    out = torch.zeros(n, device="cuda")
    c.call("mamdr_sub", ptr(out), ptr(d_mo), ptr(d_th), n, stream())
    np.testing.assert_array_equal(bits(out.cpu().numpy()), bits(mo - th))
    u = dev(ti)
    c.call("mamdr_axpy_diff", ptr(u), ptr(d_mo), ptr(d_th), 0.1, n, stream())
    np.testing.assert_array_equal(bits(u.cpu().numpy()), bits(ti + (mo - th) * beta))
    c.call("mamdr_copy", ptr(out), ptr(u), n, stream())
    assert torch.equal(out, u)


def test_meta_idempotence_and_linearity_full_size():
    """size-independent properties at Amazon-6 arena size (79.3 M floats)."""
    n = 79301152
    g = torch.Generator(device="cuda").manual_seed(1)
    th = torch.randn(n, device="cuda", generator=g)
    mo = torch.randn(n, device="cuda", generator=g)
    c = ctx()
    t = th.clone()
    c.call("mamdr_dn_update", ptr(t), ptr(th), 0.1, n, None, stream())   # model == theta -> fixed point
    assert torch.equal(t, th)
    t = th.clone()
    c.call("mamdr_dn_update", ptr(t), ptr(mo), 1.0, n, None, stream())   # beta = 1 -> theta + (model - theta)
    assert torch.equal(t, th + (mo - th))
    z = torch.zeros(n, device="cuda")
    zero = torch.zeros(n, device="cuda")
    c.call("mamdr_merge", ptr(z), ptr(th), ptr(zero), n, 0, stream())
    assert torch.equal(z, th)


# ---- K8 AUC -----------------------------------------------------------------------------------------------------
def test_auc_kat_and_counts_bit_exact():
    c = ctx()
    # the reference's doc-string example (utils/auc.py:44-56)
    thr3 = dev(oauc.thresholds(3))
    acc = torch.zeros(4, 3, device="cuda")
    d_p, d_y = dev(np.float32([0, 0.5, 0.3, 0.9])), dev(np.float32([0, 0, 1, 1]))
    c.call("mamdr_auc_update", ptr(d_p), ptr(d_y), 4, ptr(acc), ptr(thr3), 3, stream())
    np.testing.assert_array_equal(acc.cpu().numpy(), [[2, 1, 0], [2, 0, 0], [0, 1, 2], [0, 2, 2]])
    out = torch.zeros(1, device="cuda")
    c.call("mamdr_auc_result", ptr(acc), 3, ptr(out), stream())
    assert abs(out.item() - 0.75) < 1e-7
    # 500 thresholds, streaming, predictions sitting exactly on thresholds and on 0 / 1
    rng = np.random.default_rng(9)
    thr = oauc.thresholds(500)
    o = oauc.AUC(500)
    acc = torch.zeros(4, 500, device="cuda")
    d_thr = dev(thr)
    for n in (1024, 1024, 977, 1, 5000):
        y = (rng.random(n) < 0.3).astype(np.float32)
        p = np.clip(0.25 * y + rng.random(n) * 0.75, 0, 1).astype(np.float32)
        p[: min(n, 64)] = thr[rng.integers(1, 499, min(n, 64))]
        if n > 2:
            p[-1], p[-2] = 0.0, 1.0
        o.update_state(y, p)
        d_p, d_y = dev(p), dev(y)
        c.call("mamdr_auc_update", ptr(d_p), ptr(d_y), n, ptr(acc), ptr(d_thr), 500, stream())
    np.testing.assert_array_equal(acc.cpu().numpy(), o.acc)
    c.call("mamdr_auc_result", ptr(acc), 500, ptr(out), stream())
    assert abs(out.item() - o.result()) < 2e-6


# ---- routing of the row-sharded tables (route.cu) -----------------------------------------------------------
@pytest.mark.parametrize("n,world,cap,n_rows", [(512, 2, 512, 718_000), (128, 8, 128, 50_000), (1024, 1, 1024, 300), (977, 3, 1000, 5000),
                                                (0, 4, 16, 100), (2500, 64, 2500, 9000)])
def test_route_plan_and_pack_rows_bit_exact(n, world, cap, n_rows):
    """`mamdr_route_plan` (both id columns in one launch) and `mamdr_route_pack_rows` vs the numpy restatement of their
    contract (oracle/route.py)."""
    from oracle.route import pack_rows, route_plan
    rng = np.random.default_rng(n + world)
    ids_a = (rng.zipf(1.2, n) % n_rows).astype(np.int32)      # hot ids: long runs of one owner
    ids_b = rng.integers(0, n_rows, n).astype(np.int32)
    c = ctx()
    da, db = dev(ids_a) if n else torch.zeros(1, dtype=torch.int32, device="cuda"), dev(ids_b) if n else torch.zeros(1, dtype=torch.int32, device="cuda")
    slot_a, slot_b = torch.full((cap,), -5, dtype=torch.int32, device="cuda"), torch.full((cap,), -5, dtype=torch.int32, device="cuda")
    send_a, send_b = torch.full((world * cap,), 7, dtype=torch.int32, device="cuda"), torch.full((world * cap,), 7, dtype=torch.int32, device="cuda")
    c.call("mamdr_route_plan", ptr(da), ptr(db), n, world, cap, cap, ptr(slot_a), ptr(slot_b), ptr(send_a), ptr(send_b), stream())
    for ids, slot, send in ((ids_a, slot_a, send_a), (ids_b, slot_b, send_b)):
        want_slot, want_send = route_plan(ids, world, cap)
        np.testing.assert_array_equal(slot.cpu().numpy()[:n], want_slot)
        np.testing.assert_array_equal(send.cpu().numpy(), want_send)
    if n == 0:
        return
    dim = 36
    src = rng.standard_normal((n, dim + 12)).astype(np.float32)
    dst = torch.full((world * cap, dim), -3.0, device="cuda")
    want_slot, want_send = route_plan(ids_a, world, cap)
    c.call("mamdr_route_pack_rows", ptr(dev(src)), dim + 12, ptr(slot_a), n, dim, 0.37, ptr(dst), stream())
    want = pack_rows(src[:, :dim], want_slot, world, cap, 0.37)
    got = dst.cpu().numpy()
    np.testing.assert_array_equal(bits(got[want_slot]), bits(want[want_slot]))
    untouched = np.ones(world * cap, bool)
    untouched[want_slot] = False
    assert np.all(got[untouched] == -3.0)


@pytest.mark.parametrize("n,world,cap", [(512, 2, 512), (100, 8, 128), (977, 3, 1000), (0, 4, 16)])
def test_route_shared_buffer_plan_and_two_table_gather_bit_exact(n, world, cap):
    """Both id columns in ONE exchange buffer (block = 2 * cap, `mamdr_route_plan`) and the owners' two-table gather with the
    per-table split of the received ids (`mamdr_route_gather2`) vs oracle/route.py."""
    from oracle.route import route_plan_shared
    rng = np.random.default_rng(n * 7 + world)
    rows_a, rows_b, dim = 4001, 977, 128
    ids_a = (rng.zipf(1.3, n) % (rows_a * world)).astype(np.int32)
    ids_b = rng.integers(0, rows_b * world, n).astype(np.int32)
    c = ctx()
    z = torch.zeros(1, dtype=torch.int32, device="cuda")
    da, db = (dev(ids_a), dev(ids_b)) if n else (z, z)
    slot_a, slot_b = torch.zeros(cap, dtype=torch.int32, device="cuda"), torch.zeros(cap, dtype=torch.int32, device="cuda")
    send = torch.full((world * 2 * cap,), 7, dtype=torch.int32, device="cuda")
    c.call("mamdr_route_plan", ptr(da), ptr(db), n, world, cap, 2 * cap, ptr(slot_a), ptr(slot_b), ptr(send), C.c_void_p(send.data_ptr() + 4 * cap), stream())
    wa, wb, wsend = route_plan_shared(ids_a, ids_b, world, cap)
    np.testing.assert_array_equal(send.cpu().numpy(), wsend)
    np.testing.assert_array_equal(slot_a.cpu().numpy()[:n], wa)
    np.testing.assert_array_equal(slot_b.cpu().numpy()[:n], wb)
    # the owner's side: treat `send` as what arrived
    ta, tb = rng.standard_normal((rows_a, dim)).astype(np.float32), rng.standard_normal((rows_b, dim)).astype(np.float32)
    out = torch.full((world * 2 * cap, dim), -9.0, device="cuda")
    ia, ib = torch.zeros_like(send), torch.zeros_like(send)
    c.call("mamdr_route_gather2", ptr(dev(ta)), ptr(dev(tb)), ptr(send), world, cap, dim, ptr(out), ptr(ia), ptr(ib), stream())
    got = out.cpu().numpy().reshape(world, 2, cap, dim)
    w3 = wsend.reshape(world, 2, cap)
    for col, table in ((0, ta), (1, tb)):
        valid = w3[:, col, :] >= 0
        np.testing.assert_array_equal(bits(got[:, col][valid]), bits(table[w3[:, col, :][valid]]))
        assert np.all(got[:, col][~valid] == -9.0)
    want_a, want_b = w3.copy(), w3.copy()
    want_a[:, 1, :] = -1
    want_b[:, 0, :] = -1
    np.testing.assert_array_equal(ia.cpu().numpy(), want_a.reshape(-1))
    np.testing.assert_array_equal(ib.cpu().numpy(), want_b.reshape(-1))
