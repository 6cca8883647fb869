"""CPU, world_size 2, gloo: the host side of the sharded MAMDR meta-step (mamdr_b200/dist.py) -- LPT assignment
is identical on every rank, and the ONE collective of a meta-step (theta_i from their owners + the Adam slots of the
rank that owns the last query domain) is bit-exact.  The kernels themselves need a GPU (tests -m gpu)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mamdr_b200 import dist as mdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    r, w = mdist.init_from_env("gloo")
    assert (r, w) == (rank, world) and mdist.world() == (rank, world)
    # a Taobao-10-like long tail of pass lengths
    n_step = {i: max(1, 31 // (i + 1)) for i in range(10)}
    seq = [3, 7, 0, 9, 1, 4, 8, 2, 6, 5]
    supports = {i: [j for j in seq if j != i][:5] + [i] for i in seq}
    owner = mdist.lpt_assign(mdist.dr_chain_costs(seq, supports, n_step), world)
    P = 1024
    g = torch.Generator().manual_seed(1234)           # same stream on both ranks: the "true" values
    truth = {i: torch.randn(P, generator=g) for i in range(10)}
    m_true, v_true = torch.randn(P, generator=g), torch.rand(P, generator=g)
    words_true = torch.tensor([154.0, 0.9 ** 155, 0.999 ** 155])
    last_owner = owner[seq[-1]]
    # every rank only holds valid data for what it owns; the rest is garbage that the exchange must overwrite
    flats = {i: (truth[i].clone() if owner[i] == rank else torch.full((P,), float(rank + 7))) for i in range(10)}
    m = m_true.clone() if rank == last_owner else torch.zeros(P) - 3
    v = v_true.clone() if rank == last_owner else torch.zeros(P) - 5
    words = words_true.clone() if rank == last_owner else torch.zeros(3)
    nbytes = mdist.exchange(owner, rank, flats, m, v, words, last_owner)
    ok = all(torch.equal(flats[i], truth[i]) for i in range(10)) and torch.equal(m, m_true) and torch.equal(v, v_true) \
        and torch.equal(words, words_true) and nbytes == (12 * P + 4) * 4
    torch.save({"owner": owner, "ok": bool(ok)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_and_assignment_world2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"))
    assert a["ok"] and b["ok"]
    assert a["owner"] == b["owner"]
    loads = [0, 0]
    n_step = {i: max(1, 31 // (i + 1)) for i in range(10)}
    for i, r in a["owner"].items():
        loads[r] += n_step[i]
    assert set(a["owner"].values()) == {0, 1}


def test_lpt_is_balanced_and_deterministic():
    costs = {i: c for i, c in enumerate([310, 150, 100, 75, 60, 50, 44, 38, 34, 31])}
    for world in (1, 2, 4, 8):
        o1, o2 = mdist.lpt_assign(costs, world), mdist.lpt_assign(dict(reversed(list(costs.items()))), world)
        assert o1 == o2
        load = [sum(c for k, c in costs.items() if o1[k] == r) for r in range(world)]
        assert max(load) <= max(max(costs.values()), 1.34 * sum(costs.values()) / world)


def test_sharded_oracle_equals_sequential_oracle_at_world1():
    """With one rank the defined sharded semantics degenerate to the reference schedule."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from conftest import make_config
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    from mamdr_b200.schedule import Schedule
    from oracle.meta import OracleMAMDR
    from oracle.mlp import MLPSpec, OracleMLP
    c = make_config(**{"dataset.synthetic.scale": 0.01, "train.sample_num": 2})
    g = synth.generate("Taobao-10", seed=123, scale=0.01)
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), False)
    data = {"train": g["train"], "val": g["val"], "test": g["test"]}
    outs = []
    for sharded in (False, True):
        spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), dropout=0.5)
        o = OracleMLP(spec, init_mlp_weights(lo, [1, 0]), g["user_emb"], g["item_emb"], lr=1e-3)
        om = OracleMAMDR(o, data, c['train'], 1024, Schedule(9), {d: init_mlp_weights(lo, [1, d + 1]) for d in range(10)})
        if sharded:
            om.train_epoch_sharded(1)
        else:
            om.train_epoch()
        outs.append(om)
    for a, b in zip(outs[0].meta_weights, outs[1].meta_weights):
        assert np.array_equal(a, b)
    for d in outs[0].domain_weights:
        for a, b in zip(outs[0].domain_weights[d], outs[1].domain_weights[d]):
            assert np.array_equal(a, b)


# ---- row-sharded tables: the fixed-capacity routing plan of mamdr_b200/sharded.py (host logic; the kernels need a GPU) ----
def _routing_worker(rank, world, port, out_dir):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    dist.init_process_group("gloo")
    from oracle.route import pack_rows, route_plan
    n_rows, dim, cap = 37, 4, 16
    g = torch.Generator().manual_seed(7)
    full = torch.randn(n_rows, dim, generator=g)                     # the "true" table, same on both ranks
    local = full[rank::world].contiguous()                           # row r lives on rank r % world at index r // world
    gi = torch.Generator().manual_seed(100 + rank)
    n = cap - 3 * rank                                               # ragged slices: ranks hold different row counts
    ids = torch.randint(0, n_rows, (n,), generator=gi, dtype=torch.int32)
    ids[: n // 3] = ids[0]                                           # a hot id (duplicates) like the Zipf batches
    # the plan (mamdr_route_plan's contract, numpy restatement) and the id all-to-all with static, equal splits
    slot_np, send_np = route_plan(ids.numpy(), world, cap)
    slot, send_idx = torch.from_numpy(slot_np).long(), torch.from_numpy(send_np)
    recv_idx = torch.empty_like(send_idx)
    dist.all_to_all_single(recv_idx, send_idx)
    assert recv_idx.numel() == world * cap
    valid = recv_idx >= 0
    assert int(valid.sum()) <= world * cap and bool((recv_idx[valid] < local.shape[0]).all())
    # fetch: owners gather (padding entries skipped), equal-split all-to-all back, rows picked out by slot
    got = torch.zeros(world * cap, dim)
    got[valid] = local[recv_idx[valid].long()]
    back = torch.empty_like(got)
    dist.all_to_all_single(back, got)
    out = back[slot]
    assert torch.equal(out, full[ids.long()])
    # apply: gradient rows travel to their owners aligned with recv_idx; padding rows carry id -1 and are never read
    grads = torch.randn(n, dim, generator=gi)
    send = torch.from_numpy(pack_rows(grads.numpy(), slot_np, world, cap))
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    mine = torch.zeros_like(local)
    mine.index_add_(0, recv_idx[valid].long(), recv[valid])
    # the truth: scatter-add of every rank's gradient rows into the full table, restricted to my rows
    all_ids = [torch.empty(cap, dtype=torch.int32) for _ in range(world)]
    all_g = [torch.empty(cap, dim) for _ in range(world)]
    pad_ids = torch.full((cap,), -1, dtype=torch.int32)
    pad_ids[:n] = ids
    pad_g = torch.zeros(cap, dim)
    pad_g[:n] = grads
    dist.all_gather(all_ids, pad_ids)
    dist.all_gather(all_g, pad_g)
    truth = torch.zeros(n_rows, dim)
    for i_, g_ in zip(all_ids, all_g):
        ok = i_ >= 0
        truth.index_add_(0, i_[ok].long(), g_[ok])
    assert torch.allclose(mine, truth[rank::world], atol=1e-6)
    torch.save({"ok": True}, os.path.join(out_dir, "routing%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_fixed_capacity_routing_plan_world2_gloo(tmp_path):
    """The routing contract of `mamdr_route_plan` / `mamdr_route_pack_rows` (oracle/route.py restates it; the kernels are checked
    against that restatement bit for bit in tests/test_gpu_kernels.py): ids -> owners with static, equal all-to-all splits (-1
    padded), rows back in the original order, gradient rows to their owners; against the unsharded table on two gloo ranks."""
    port = _free_port()
    mp.spawn(_routing_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "routing0.pt")) and os.path.exists(os.path.join(str(tmp_path), "routing1.pt"))
