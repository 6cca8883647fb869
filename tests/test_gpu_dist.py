"""-m gpu: the SHARDED MAMDR meta-step on a real device -- two ranks (gloo transport, both on cuda:0
so that it runs on the single-GPU test box; the bench exercises NCCL on 2/4/8 GPUs) against the CPU oracle executing the
same sharded schedule (oracle/meta.py:train_epoch_sharded).  fp32 mode, rel 1e-4; both ranks must end bit-identical."""
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


def _worker(rank, world, port, out_dir, prec, name="mlp_meta_mamdr_finetune"):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    from conftest import make_config
    import run
    from mamdr_b200 import dist as mdist
    from mamdr_b200.schedule import Schedule
    mdist.init_from_env("gloo")
    c = make_config(**{"model.name": name, "dataset.synthetic.scale": 0.03, "b200.precision": prec})
    wrapper = run.build(c)
    wrapper.prepare()
    wrapper.base_model.schedule = Schedule(77)
    for e in range(2):
        wrapper.train_epoch(e)
    torch.cuda.synchronize()
    m = wrapper.model
    blob = {"theta": wrapper.meta_weights.flat.cpu(), "domain": {d: v.flat.cpu() for d, v in wrapper.domain_weights.items()},
            "m": m.m.cpu(), "v": m.v.cpu(), "step": m.read_step(), "owner": wrapper.dr_owner,
            "pair_owner": wrapper.dr_pair_owner}
    torch.save(blob, os.path.join(out_dir, "rank%d.pt" % rank))
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("tf32x3", 1e-2)])
def test_two_rank_sharded_meta_steps_match_sharded_oracle(tmp_path, prec, tol):
    import torch.multiprocessing as mp
    from conftest import make_config, rel_err
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    from mamdr_b200.schedule import Schedule
    from oracle.meta import OracleMAMDR
    from oracle.mlp import MLPSpec, OracleMLP
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), prec), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"), weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"), weights_only=False)
    # after the one collective of the meta-step both ranks hold the same bits
    assert torch.equal(a["theta"], b["theta"]) and torch.equal(a["m"], b["m"]) and torch.equal(a["v"], b["v"]) and a["step"] == b["step"]
    for d in a["domain"]:
        assert torch.equal(a["domain"][d], b["domain"][d])
    assert a["owner"] == b["owner"] and set(a["owner"].values()) == {0, 1}
    # the oracle executes the same sharded schedule on the CPU (same seeds -> same initial draws as the wrappers)
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 0.03})
    g = synth.generate("Taobao-10", seed=123, scale=0.03)
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), False)
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), dropout=0.5)
    o = OracleMLP(spec, init_mlp_weights(lo, [123, 0]), g["user_emb"], g["item_emb"], lr=1e-3)
    om = OracleMAMDR(o, {"train": g["train"], "val": g["val"], "test": g["test"]}, c['train'], 1024, Schedule(77),
                     {d: init_mlp_weights(lo, [123, d + 1]) for d in range(10)}, name=c['model']['name'])
    for e in range(2):
        om.train_epoch_sharded(2)
    theta = lo.unpack(a["theta"].numpy())
    for n_, x, y in zip(lo.names, theta, om.meta_weights):
        assert rel_err(x, y) < tol, ("theta", n_, rel_err(x, y))
    for d in om.domain_weights:
        for n_, x, y in zip(lo.names, lo.unpack(a["domain"][d].numpy()), om.domain_weights[d]):
            assert rel_err(x, y) < (tol if prec == "fp32" else 5e-2), ("theta_%d" % d, n_, rel_err(x, y))
    assert a["step"][0] == o.adam.step


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("tf32x3", 1e-2)])
def test_two_rank_pair_sharded_batch_meta_steps_match_sharded_oracle(tmp_path, prec, tol):
    """'batch' names: the (query, support) PAIRS are sharded (60 units at Taobao-10 instead of 10 chains), one all-reduce of the
    accumulated deltas per meta-step (mamdr.py:100-108,182-196) -- two ranks vs the oracle executing the same pair-sharded
    schedule (oracle/meta.py:_train_epoch_pair_sharded)."""
    import torch.multiprocessing as mp
    from conftest import make_config, rel_err
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    from mamdr_b200.schedule import Schedule
    from oracle.meta import OracleMAMDR
    from oracle.mlp import MLPSpec, OracleMLP
    name = "mlp_meta_mamdr_batch"
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), prec, name), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"), weights_only=False)
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"), weights_only=False)
    assert torch.equal(a["theta"], b["theta"]) and torch.equal(a["m"], b["m"]) and torch.equal(a["v"], b["v"]) and a["step"] == b["step"]
    for d in a["domain"]:
        assert torch.equal(a["domain"][d], b["domain"][d])
    assert a["pair_owner"] == b["pair_owner"] and set(a["pair_owner"].values()) == {0, 1} and len(a["pair_owner"]) == 10 * 3
    c = make_config(**{"model.name": name, "dataset.synthetic.scale": 0.03})
    g = synth.generate("Taobao-10", seed=123, scale=0.03)
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), False)
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), dropout=0.5)
    o = OracleMLP(spec, init_mlp_weights(lo, [123, 0]), g["user_emb"], g["item_emb"], lr=1e-3)
    om = OracleMAMDR(o, {"train": g["train"], "val": g["val"], "test": g["test"]}, c['train'], 1024, Schedule(77),
                     {d: init_mlp_weights(lo, [123, d + 1]) for d in range(10)}, name=name)
    for e in range(2):
        owner = om.train_epoch_sharded(2)
    assert owner == a["pair_owner"]
    for n_, x, y in zip(lo.names, lo.unpack(a["theta"].numpy()), om.meta_weights):
        assert rel_err(x, y) < tol, ("theta", n_, rel_err(x, y))
    for d in om.domain_weights:
        for n_, x, y in zip(lo.names, lo.unpack(a["domain"][d].numpy()), om.domain_weights[d]):
            assert rel_err(x, y) < (tol if prec == "fp32" else 5e-2), ("theta_%d" % d, n_, rel_err(x, y))
    assert a["step"][0] == o.adam.step
