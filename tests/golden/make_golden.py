"""Generates tests/golden/kernels_v1.npz -- committed input / output vectors of the hot-path primitives.

    python tests/golden/make_golden.py

The reference (TF 1.12 + deepctr 0.9.0) cannot be imported in this environment and ships no fixtures besides the AUC
doc-string example (utils/auc.py:44-56, included here), so these vectors are produced by the CPU ORACLE (oracle/, the
restatement of the reference's arithmetic): they pin the oracle against regressions and let the `-m gpu` tests compare the
CUDA path with fixed numbers that do not depend on the host's BLAS threading.  Element-wise / integer results are exact
fp32 / int32 vectors; the two train-step records hold the float64 oracle's loss and gradients.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mamdr_b200.deep_mtl_ctr import MTLTopology, init_mtl_weights  # noqa: E402  (layout / initialiser only: host code)
from mamdr_b200.layout import init_mlp_weights, mlp_layout        # noqa: E402
from oracle import auc as oauc                                     # noqa: E402
from oracle import philox                                          # noqa: E402
from oracle.mlp import AdamState, MLPSpec, OracleMLP               # noqa: E402
from oracle.mtl import MTLSpec, OracleMTL                          # noqa: E402

MLP = dict(n_uid=60, n_pid=50, n_domain=3, emb_dim=(16, 16, 8), hidden=(32, 16, 8), rows=45, domain=1)
MTL = dict(n_uid=60, n_pid=50, n_domain=3, emb_dim=(16, 16, 8), expert_hidden=(32, 16), tower_hidden=(8,), gate_hidden=(8,),
           num_experts=3, rows=45, domain=2)


def mlp_problem():
    rng = np.random.default_rng(11)
    lo = mlp_layout(MLP['n_uid'], MLP['n_pid'], MLP['n_domain'], MLP['emb_dim'], MLP['hidden'], False)
    w = init_mlp_weights(lo, [11, 0])
    for i, n in enumerate(lo.names):
        if n.startswith('bias') or n in ('global_bias', 'domain_emb'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    ut = (rng.standard_normal((MLP['n_uid'], MLP['emb_dim'][0])) * 0.05).astype(np.float32)
    it = (rng.standard_normal((MLP['n_pid'], MLP['emb_dim'][1])) * 0.05).astype(np.float32)
    uid = rng.integers(0, MLP['n_uid'], MLP['rows']).astype(np.int32)
    pid = rng.integers(0, MLP['n_pid'], MLP['rows']).astype(np.int32)
    y = (rng.random(MLP['rows']) < 0.35).astype(np.float32)
    return lo, w, ut, it, uid, pid, y


def mtl_problem():
    rng = np.random.default_rng(12)
    topo = MTLTopology('mmoe', MTL['n_uid'], MTL['n_pid'], MTL['n_domain'], MTL['emb_dim'], MTL['expert_hidden'], MTL['tower_hidden'],
                       MTL['gate_hidden'], num_experts=MTL['num_experts'], emb_trainable=True)
    w = init_mtl_weights(topo.layout, [12, 0])
    for i, n in enumerate(topo.layout.names):
        if n.endswith('_emb') or 'bias' in n:
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    uid = rng.integers(0, MTL['n_uid'], MTL['rows']).astype(np.int32)
    pid = rng.integers(0, MTL['n_pid'], MTL['rows']).astype(np.int32)
    uid[:6] = uid[0]
    y = (rng.random(MTL['rows']) < 0.35).astype(np.float32)
    return topo, w, uid, pid, y


def make():
    g = {}
    rng = np.random.default_rng(2024)
    f32 = np.float32
    # K1 gather
    g['gather_table'] = rng.standard_normal((50, 8)).astype(f32)
    g['gather_ids'] = rng.integers(0, 50, 37).astype(np.int32)
    g['gather_out'] = g['gather_table'][g['gather_ids']]
    # K6 dedup: sorted unique ids, rows of one id added sequentially in batch order (numpy add.at order)
    ids = rng.integers(0, 40, 200).astype(np.int32)
    rows = rng.standard_normal((200, 8)).astype(f32)
    su = np.zeros((40, 8), dtype=f32)
    np.add.at(su, ids, rows)
    g['dedup_ids'], g['dedup_rows'] = ids, rows
    g['dedup_uniq'] = np.unique(ids).astype(np.int32)
    g['dedup_sums'] = su[g['dedup_uniq']]
    # K7 Adam (TF ApplyAdam order), 5 steps
    p = [rng.standard_normal(260).astype(f32)]
    g['adam_p0'] = p[0].copy()
    st = AdamState(p, lr=1e-3)
    gs = (rng.standard_normal((5, 260)) * 10.0 ** rng.integers(-5, 1, (5, 1))).astype(f32)
    for t in range(5):
        st.apply(p, [gs[t]])
    g['adam_g'], g['adam_p'], g['adam_m'], g['adam_v'] = gs, p[0], st.m[0], st.v[0]
    g['adam_pows'] = np.array([st.b1pow, st.b2pow], dtype=f32)
    # K9 / K10 meta ops
    th, ti, mo = (rng.standard_normal(256).astype(f32) for _ in range(3))
    beta = f32(0.1)
    g['meta_th'], g['meta_ti'], g['meta_mo'] = th, ti, mo
    g['meta_dn'] = th + (mo - th) * beta
    g['meta_dr_plus'] = ti + (mo - (th + ti)) * beta
    g['meta_dr_times'] = ti + (mo - (th * ti)) * beta
    # K8 AUC: the reference's doc-string example + a streaming 500-threshold case
    g['auc_kat_y'], g['auc_kat_p'] = f32([0, 0, 1, 1]), f32([0, 0.5, 0.3, 0.9])
    g['auc_kat_acc'] = f32([[2, 1, 0], [2, 0, 0], [0, 1, 2], [0, 2, 2]])
    g['auc_kat_result'] = f32(0.75)
    y = (rng.random(300) < 0.3).astype(f32)
    pr = np.clip(0.25 * y + rng.random(300) * 0.75, 0, 1).astype(f32)
    a = oauc.AUC(500)
    a.update_state(y, pr)
    g['auc_y'], g['auc_p'], g['auc_acc'], g['auc_result'] = y, pr, np.asarray(a.acc, dtype=f32), f32(a.result())
    # dropout mask convention (Philox4x32-10)
    g['philox_mask'] = philox.dropout_mask(8, 16, 1030, 3, 0.5, use_c=False).astype(f32)
    # one mlp train step (frozen tables) -- float64 oracle
    lo, w, ut, it, uid, pid, yy = mlp_problem()
    spec = MLPSpec(MLP['n_uid'], MLP['n_pid'], MLP['n_domain'], MLP['emb_dim'], MLP['hidden'], dropout=0.5)
    o = OracleMLP(spec, w, ut, it, lr=1e-3, dtype=np.float64)
    loss, pp, grads = o.gradients(uid, pid, MLP['domain'], yy)
    g['mlp_loss'], g['mlp_p'], g['mlp_grads'] = np.float64(loss), np.asarray(pp, dtype=np.float64), lo.pack(grads, dtype=np.float64)
    # one MMOE sub-model train step (trainable tables) -- float64 oracle
    topo, w, uid, pid, yy = mtl_problem()
    spec = MTLSpec(MTL['n_uid'], MTL['n_pid'], MTL['n_domain'], kind='mmoe', emb_dim=MTL['emb_dim'], expert_hidden=MTL['expert_hidden'],
                   tower_hidden=MTL['tower_hidden'], gate_hidden=MTL['gate_hidden'], num_experts=MTL['num_experts'], dropout=0.5,
                   emb_trainable=True)
    o = OracleMTL(spec, w, None, None, lr=1e-3, dtype=np.float64)
    loss, pp, gd = o.gradients(uid, pid, MTL['domain'], yy)
    full = [gd.get(n, np.zeros(s)) for n, s in zip(topo.layout.names, topo.layout.shapes)]
    g['mtl_loss'], g['mtl_p'], g['mtl_grads'] = np.float64(loss), np.asarray(pp, dtype=np.float64), topo.layout.pack(full, dtype=np.float64)
    return g


if __name__ == "__main__":
    out = os.path.join(HERE, "kernels_v1.npz")
    np.savez_compressed(out, **make())
    print(out, os.path.getsize(out), "bytes")
