"""-m gpu (needs >= 2 GPUs, skipped otherwise): row-sharded trainable embedding tables with NCCL all-to-all
(mamdr_b200/sharded.py) -- a data-parallel joint `mlp` pass on two ranks against the single-process oracle on the same
global batches.  Dropout is off (the masks are indexed by the local row in the sharded mode); rel 1e-5 after one batch,
1e-4 after a ragged multi-batch pass; both ranks hold bit-identical dense replicas."""
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


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch.distributed as dist
    from mamdr_b200.schedule import Schedule
    from mamdr_b200.sharded import ShardedJointTrainer
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl")
    g, lo, w = _problem()
    t = ShardedJointTrainer(g["n_uid"], g["n_pid"], g["n_domain"], w[0], w[1], w[2:], dropout=0.0, batch_size=1024,
                            device="cuda:%d" % rank)
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


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="row-sharded tables need 2 GPUs (NCCL all-to-all)")
def test_row_sharded_tables_two_ranks_match_oracle(tmp_path):
    import torch.multiprocessing as mp
    from conftest import rel_err
    from mamdr_b200.schedule import Schedule
    from oracle.meta import train_pass
    from oracle.mlp import MLPSpec, OracleMLP
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"), weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"), weights_only=False)
    assert torch.equal(a["dense"], b["dense"]) and a["step"] == b["step"]
    g, lo, w = _problem()
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), dropout=0.0, emb_trainable=True)
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
