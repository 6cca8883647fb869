"""-m gpu (needs >= 2 GPUs, skipped otherwise): row-sharded trainable embedding tables with NCCL all-to-all
(mamdr_b200/sharded.py) -- a data-parallel joint `mlp` pass on two ranks against the single-process oracle on the same
global batches, with dropout 0 and 0.5 (the masks are indexed by the GLOBAL batch row, mamdr_batch.row0, so the sharded step
draws exactly the unsharded masks); rel 1e-4 after a ragged multi-batch pass; both ranks hold bit-identical dense replicas."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem():
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    g = synth.generate("Amazon-6", seed=5, scale=0.001)
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), True)
    w = init_mlp_weights(lo, [5, 0])
    rng = np.random.default_rng(1)
    for i, n in enumerate(lo.names):          # lift tables / biases off their near-zero init
        if n.endswith('_emb'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
        if n.startswith('bias'):
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    return g, lo, w


def _worker(rank, world, port, out_dir, dropout=0.0, graphs=False):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch.distributed as dist
    from mamdr_b200.schedule import Schedule
    from mamdr_b200.sharded import ShardedJointTrainer
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl")
    g, lo, w = _problem()
    t = ShardedJointTrainer(g["n_uid"], g["n_pid"], g["n_domain"], w[0], w[1], w[2:], dropout=dropout, batch_size=1024,
                            device="cuda:%d" % rank, use_graphs=graphs)
    d = 0
    split = g["train"][d]
    order = Schedule(3).batch_order(d, len(split["uid"]))
    losses = t.train_pass(split, d, order)
    torch.cuda.synchronize()
    fu, fi = t.users.full(), t.items.full()
    blob = {"dense": t.model.params.cpu(), "step": t.model.read_step(), "loss": [x.cpu() for x in losses]}
    if rank == 0:
        blob["user"], blob["item"] = fu.cpu(), fi.cpu()
    torch.save(blob, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="row-sharded tables need 2 GPUs (NCCL all-to-all)")
@pytest.mark.parametrize("dropout,graphs", [(0.0, False), (0.5, False), (0.5, True)])
def test_row_sharded_tables_two_ranks_match_oracle(tmp_path, dropout, graphs):
    import torch.multiprocessing as mp
    from conftest import rel_err
    from mamdr_b200.schedule import Schedule
    from oracle.meta import train_pass
    from oracle.mlp import MLPSpec, OracleMLP
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), dropout, graphs), nprocs=2, join=True)   # graphs: the tower part of a step replayed from a CUDA graph
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"), weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"), weights_only=False)
    assert torch.equal(a["dense"], b["dense"]) and a["step"] == b["step"]
    g, lo, w = _problem()
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), dropout=dropout, emb_trainable=True)
    o = OracleMLP(spec, w, None, None, lr=1e-3)
    split = g["train"][0]
    order = Schedule(3).batch_order(0, len(split["uid"]))
    assert len(order) > 1024 and len(order) % 1024 != 0
    o_loss, _, steps = train_pass(o, split, 0, order, 1024)
    assert steps == len(a["loss"]) and a["step"][0] == o.adam.step
    got_loss = float(np.mean([float(x.sum()) for x in a["loss"]]))
    assert abs(got_loss - o_loss) < 2e-5 * abs(o_loss)
    assert rel_err(a["user"].numpy(), o.w('user_emb')) < 1e-4 and rel_err(a["item"].numpy(), o.w('item_emb')) < 1e-4
    from mamdr_b200.layout import mlp_layout
    lo_d = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), False)
    for n_, x in zip(lo_d.names, lo_d.unpack(a["dense"].numpy())):
        assert rel_err(x, o.w(n_)) < 1e-4, (n_, rel_err(x, o.w(n_)))


# ---- BASELINE config #5 end to end: mmoe / ple + DomainNegotiation, trainable tables row-sharded over two ranks --------
MTL_ARCH = {"mmoe": dict(expert_hidden=(256, 128), tower_hidden=(64,), gate_hidden=(64,), num_experts=5),
            "ple": dict(expert_hidden=(64, 32), tower_hidden=(32, 16), gate_hidden=(16,), specific_expert_num=2, shared_expert_num=1)}


def _mtl_problem(kind):
    from mamdr_b200 import synth
    from mamdr_b200.deep_mtl_ctr import MTLTopology, init_mtl_weights
    g = synth.generate("Amazon-13", seed=5, scale=0.0005)
    topo = MTLTopology(kind, g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), emb_trainable=True, **MTL_ARCH[kind])
    w = init_mtl_weights(topo.layout, [5, 0])
    rng = np.random.default_rng(1)
    for i, n in enumerate(topo.layout.names):          # lift tables / biases off their near-zero init
        if n.endswith('_emb') or 'bias' in n:
            w[i] = (rng.standard_normal(w[i].shape) * 0.05).astype(np.float32)
    return g, topo, w


def _mtl_worker(rank, world, port, out_dir, kind, dropout=0.0, graphs=False):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch.distributed as dist
    from mamdr_b200.schedule import Schedule
    from mamdr_b200.sharded import ShardedMTLTrainer
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl")
    g, topo, w = _mtl_problem(kind)
    t = ShardedMTLTrainer(kind, g["n_uid"], g["n_pid"], g["n_domain"], w[0], w[1], w[2:], dropout=dropout, lr=1e-4, batch_size=1024,
                          device="cuda:%d" % rank, use_graphs=graphs, **MTL_ARCH[kind])
    t.dn_prepare()
    sched = Schedule(7)
    seq = list(range(g["n_domain"]))
    for _ in range(2):
        seq = sched.shuffle_sequence(seq)
        orders = [sched.batch_order(idx, len(g["train"][idx]["uid"])) for idx in seq]
        t.dn_meta_step(g["train"], seq, orders, 0.1)
    torch.cuda.synchronize()
    tu, ti = t.theta_tables()
    blob = {"dense": t.theta["dense"].cpu(), "live": t.model.params.cpu(), "step": t.model.read_step()}
    if rank == 0:
        blob["user"], blob["item"] = tu.cpu(), ti.cpu()
    torch.save(blob, os.path.join(out_dir, "mtl_rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(240)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="row-sharded tables need 2 GPUs (NCCL all-to-all)")
@pytest.mark.parametrize("kind,dropout,graphs", [("mmoe", 0.0, False), ("ple", 0.0, False), ("mmoe", 0.5, False), ("mmoe", 0.5, True)])
def test_sharded_mtl_domain_negotiation_two_ranks_match_oracle(tmp_path, kind, dropout, graphs):
    """Two DN meta-steps of `<kind>_meta_domain_negotiation` with the tables row-sharded over 2 ranks vs the single-process
    oracle (OracleDN over OracleMTL) on the same global batches: theta (dense + both tables) rel 1e-4, replicas bit-identical;
    with dropout 0.5 the masks follow the global batch row (mamdr_batch.row0), i.e. the sharded run draws the oracle's masks."""
    import torch.multiprocessing as mp
    from conftest import BASE_CONFIG, rel_err
    from mamdr_b200.deep_mtl_ctr import MTLTopology
    from mamdr_b200.schedule import Schedule
    from oracle.meta import OracleDN
    from oracle.mtl import MTLSpec, OracleMTL
    port = _free_port()
    mp.spawn(_mtl_worker, args=(2, port, str(tmp_path), kind, dropout, graphs), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "mtl_rank0.pt"), weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "mtl_rank1.pt"), weights_only=False)
    assert torch.equal(a["dense"], b["dense"]) and torch.equal(a["live"], b["live"]) and a["step"] == b["step"]
    g, topo, w = _mtl_problem(kind)
    arch = MTL_ARCH[kind]
    spec = MTLSpec(g["n_uid"], g["n_pid"], g["n_domain"], kind=kind, expert_hidden=arch["expert_hidden"], tower_hidden=arch["tower_hidden"],
                   gate_hidden=arch["gate_hidden"], num_experts=arch.get("num_experts", 0),
                   specific_expert_num=arch.get("specific_expert_num", 0), shared_expert_num=arch.get("shared_expert_num", 0),
                   dropout=dropout, emb_trainable=True)
    assert spec.names == topo.layout.names
    o = OracleMTL(spec, w, None, None, lr=1e-4)
    tc = dict(BASE_CONFIG["train"], meta_learning_rate=0.1, shuffle_sequence=True, meta_train_step=0)
    od = OracleDN(o, {"train": g["train"], "val": g["val"], "test": g["test"]}, tc, 1024, Schedule(7))
    for _ in range(2):
        od.train_epoch()
    assert a["step"][0] == o.adam.step
    theta = dict(zip(spec.names, od.meta_weights))
    assert rel_err(a["user"].numpy(), theta['user_emb']) < 1e-4 and rel_err(a["item"].numpy(), theta['item_emb']) < 1e-4
    lo_d = MTLTopology(kind, 4, 4, g["n_domain"], (128, 128, 128), emb_trainable=False, **arch).layout
    for n_, x in zip(lo_d.names, lo_d.unpack(a["dense"].numpy())):
        assert rel_err(x, theta[n_]) < 1e-4, (n_, rel_err(x, theta[n_]))
