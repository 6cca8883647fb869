"""-m gpu: the opt-in "virtual ranks" mode (b200.virtual_ranks = 2: the DR chains of different query domains side by side on
two SM partitions of ONE GPU, mamdr_b200/mamdr.py:_dr_chains_on_lanes) has exactly the semantics of the two-rank sharded
schedule: it must reproduce, BIT FOR BIT, what two real ranks compute (tests/test_gpu_dist.py's worker, gloo transport, both
on cuda:0), which in turn is judged against `OracleMAMDR.train_epoch_sharded(2)`; and a pass kernel launched with fewer CTAs
than SMs must give the bits of the full-grid launch (every reduction has a fixed order that does not depend on the grid)."""
import os

import numpy as np
import pytest
import torch

from conftest import make_config, rel_err
from mamdr_b200.schedule import Schedule

pytestmark = pytest.mark.gpu


def _run(virtual_ranks, prec="tf32x3", epochs=2, scale=0.03):
    import run
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": scale, "b200.precision": prec,
                       "b200.virtual_ranks": virtual_ranks})
    wrapper = run.build(c)
    wrapper.prepare()
    wrapper.base_model.schedule = Schedule(77)
    for e in range(epochs):
        wrapper.train_epoch(e)
    torch.cuda.synchronize()
    return wrapper


@pytest.mark.parametrize("prec", ["tf32x3", "tf32"])
def test_virtual_ranks_equal_two_real_ranks_bit_for_bit(tmp_path, prec):
    import torch.multiprocessing as mp
    from test_gpu_dist import _free_port, _worker
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), prec), nprocs=2, join=True)
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"), weights_only=False)
    w = _run(2, prec)
    assert w.dr_owner == a["owner"] and set(w.dr_owner.values()) == {0, 1}
    m = w.model
    assert torch.equal(w.meta_weights.flat.cpu(), a["theta"])
    for d in a["domain"]:
        assert torch.equal(w.domain_weights[d].flat.cpu(), a["domain"][d]), d
    assert torch.equal(m.m.cpu(), a["m"]) and torch.equal(m.v.cpu(), a["v"]) and m.read_step() == a["step"]


def test_virtual_ranks_match_the_sharded_oracle():
    from mamdr_b200 import synth
    from mamdr_b200.layout import init_mlp_weights, mlp_layout
    from oracle.meta import OracleMAMDR
    from oracle.mlp import MLPSpec, OracleMLP
    w = _run(2, "tf32x3")
    c = make_config(**{"model.name": "mlp_meta_mamdr_finetune", "dataset.synthetic.scale": 0.03})
    g = synth.generate("Taobao-10", seed=123, scale=0.03)
    lo = mlp_layout(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), False)
    spec = MLPSpec(g["n_uid"], g["n_pid"], g["n_domain"], (128, 128, 128), (256, 128, 64), dropout=0.5)
    o = OracleMLP(spec, init_mlp_weights(lo, [123, 0]), g["user_emb"], g["item_emb"], lr=1e-3)
    om = OracleMAMDR(o, {"train": g["train"], "val": g["val"], "test": g["test"]}, c['train'], 1024, Schedule(77),
                     {d: init_mlp_weights(lo, [123, d + 1]) for d in range(10)}, name=c['model']['name'])
    for e in range(2):
        om.train_epoch_sharded(2)
    for n_, x, y in zip(lo.names, w.meta_weights.numpy(), om.meta_weights):
        assert rel_err(x, y) < 1e-2, ("theta", n_, rel_err(x, y))     # the free-running tf32x3 bar of tests/test_gpu_dist.py
    for d in om.domain_weights:
        for n_, x, y in zip(lo.names, w.domain_weights[d].numpy(), om.domain_weights[d]):
            assert rel_err(x, y) < 5e-2, ("theta_%d" % d, n_, rel_err(x, y))
    assert w.model.read_step()[0] == o.adam.step
    _, a, _, _ = w.val_and_test("val")
    _, oa, _, _ = om.val_and_test("val")
    assert abs(a - oa) < 1e-3


@pytest.mark.parametrize("ctas", [74, 37, 32])
def test_pass_kernel_bits_do_not_depend_on_the_grid(ctas):
    """One DN meta-step with the pass kernel on `ctas` CTAs vs one CTA per SM."""
    import run
    out = []
    for n in (0, ctas):
        c = make_config(**{"model.name": "mlp_meta_domain_negotiation", "dataset.synthetic.scale": 0.05, "b200.precision": "tf32x3"})
        w = run.build(c)
        w._get_model_meta_parms()
        w.meta_weights = w._get_meta_weights()
        w.model.reset_optimizer()
        w.meta_sequence = w.build_meta_data_split()
        w.model.set_pass_ctas(n)
        w.base_model.schedule = Schedule(5)
        w.train_epoch(0)
        torch.cuda.synchronize()
        out.append((w.meta_weights.flat.clone(), w.model.m.clone(), w.model.v.clone(), w.model.auc_acc.clone()))
    for x, y in zip(*out):
        assert torch.equal(x, y)
