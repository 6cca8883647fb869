"""Host-side logic that needs no GPU: layout, schedule, synthetic data, on-disk reader, dispatch."""
import json
import os

import numpy as np
import pytest

from conftest import make_config
from mamdr_b200 import synth
from mamdr_b200.layout import init_mlp_weights, mlp_layout
from mamdr_b200.schedule import Schedule


def test_layout_matches_trainable_weights_order_and_sizes():
    lo = mlp_layout(23778, 6932, 10, (128, 128, 128), (256, 128, 64), emb_trainable=False)
    assert lo.names == ['domain_emb', 'kernel0', 'kernel1', 'kernel2', 'bias0', 'bias1', 'bias2',
                        'dense_kernel', 'global_bias']
    assert sum(lo.numels) == 141057          # P_frozen for Taobao-10
    assert all(o % 32 == 0 for o in lo.offsets) and lo.total % 32 == 0
    lo2 = mlp_layout(445789, 172653, 6, (128, 128, 128), (256, 128, 64), emb_trainable=True)
    assert lo2.names[:3] == ['user_emb', 'item_emb', 'domain_emb']
    assert sum(lo2.numels) == 79301121       # Amazon-6 trainable P
    w = init_mlp_weights(lo, 1)
    flat = lo.pack(w)
    back = lo.unpack(flat)
    for a, b in zip(w, back):
        np.testing.assert_array_equal(a, b)
    pad = np.ones(lo.total, bool)
    for o, n in zip(lo.offsets, lo.numels):
        pad[o:o + n] = False
    assert np.all(flat[pad] == 0)


def test_initialisers():
    lo = mlp_layout(10, 10, 3, (128, 128, 128), (256, 128, 64), emb_trainable=False)
    w = dict(zip(lo.names, init_mlp_weights(lo, [123, 0])))
    assert np.all(w['bias0'] == 0) and np.all(w['global_bias'] == 0)
    k0 = w['kernel0']
    std = np.sqrt(2.0 / (384 + 256))
    assert abs(k0.std() - std) < 0.1 * std and np.abs(k0).max() <= 2 * std / 0.8796 + 1e-6
    assert abs(w['domain_emb'].std() - 1e-4) < 3e-5
    w2 = dict(zip(lo.names, init_mlp_weights(lo, [123, 1])))
    assert not np.array_equal(w2['kernel0'], k0)          # independent re-initialisation draw
    w3 = dict(zip(lo.names, init_mlp_weights(lo, [123, 0])))
    np.testing.assert_array_equal(w3['kernel0'], k0)       # deterministic


def test_schedule_is_deterministic_and_ordered():
    a, b = Schedule(7), Schedule(7)
    seq = list(range(10))
    for _ in range(3):
        sa, sb = a.shuffle_sequence(seq), b.shuffle_sequence(seq)
        assert sa == sb and sorted(sa) == seq
        assert a.sample_support([1, 2, 3, 4, 5], 3) == b.sample_support([1, 2, 3, 4, 5], 3)
        oa, ob = a.batch_order(2, 100), b.batch_order(2, 100)
        np.testing.assert_array_equal(oa, ob)
        assert sorted(oa.tolist()) == list(range(100)) and oa.dtype == np.int32
    assert not np.array_equal(a.batch_order(2, 100), oa)   # fresh permutation per pass
    with pytest.raises(ValueError):
        a.sample_support([1, 2], 5)                         # sample_num must be <= D-1 (mamdr.py:68)


def test_synthetic_shapes_match_table_I():
    g = synth.generate("Taobao-10", seed=123)
    assert (g["n_domain"], g["n_uid"], g["n_pid"]) == (10, 23778, 6932)
    assert sum(len(g["train"][d]["uid"]) for d in range(10)) == 92137
    assert sum(len(g["val"][d]["uid"]) for d in range(10)) == 37645
    assert sum(len(g["test"][d]["uid"]) for d in range(10)) == 43502
    sizes = [len(g["train"][d]["uid"]) for d in range(10)]
    assert sizes == sorted(sizes, reverse=True)            # long tail (d+1)^-1
    d0 = g["train"][0]
    assert d0["uid"].dtype == np.int32 and d0["label"].dtype == np.float32
    assert d0["uid"].min() >= 0 and d0["uid"].max() < 23778 and d0["pid"].max() < 6932
    assert set(np.unique(d0["label"])) <= {0.0, 1.0}
    assert g["user_emb"].shape == (23778, 128) and g["item_emb"].shape == (6932, 128)
    g2 = synth.generate("Taobao-10", seed=123)
    np.testing.assert_array_equal(g2["train"][3]["pid"], g["train"][3]["pid"])
    # Zipf: the most frequent id is much more frequent than the median one
    cnt = np.bincount(np.concatenate([g["train"][d]["uid"] for d in range(10)]), minlength=23778)
    assert cnt.max() > 50 * max(1, int(np.median(cnt)))
    a = synth.generate("Amazon-6", seed=1, scale=0.001)
    assert a["user_emb"] is None and a["n_domain"] == 6


def _write_reference_layout(root, g):
    split = os.path.join(root, "split_by_theme_x")
    os.makedirs(os.path.join(split, "processed_data"))
    json.dump({"id": g["n_uid"], "raw_id2id": {}}, open(os.path.join(split, "processed_data/uid2id.json"), "w"))
    json.dump({"id": g["n_pid"], "raw_id2id": {}}, open(os.path.join(split, "processed_data/pid2id.json"), "w"))
    json.dump({str(i): " ".join("%r" % float(x) for x in g["user_emb"][i]) for i in range(g["n_uid"])},
              open(os.path.join(split, "processed_data/user_emb.json"), "w"))
    json.dump({str(i): " ".join("%r" % float(x) for x in g["item_emb"][i]) for i in range(g["n_pid"])},
              open(os.path.join(split, "processed_data/item_emb.json"), "w"))
    for d in range(g["n_domain"]):
        dp = os.path.join(split, "domain_%d" % d)
        os.makedirs(dp)
        for name in ("train", "val", "test"):
            s = g[name][d]
            with open(os.path.join(dp, name + ".csv"), "w") as f:
                f.write("uid,pid,domain,label\n")
                for u, p, y in zip(s["uid"], s["pid"], s["label"]):
                    f.write("%d,%d,%d,%d\n" % (u, p, d, int(y)))
        json.dump({"ctr_ratio": g["ctr_ratio"][d]}, open(os.path.join(dp, "domain_property.json"), "w"))


def test_on_disk_reference_format_round_trip(tmp_path):
    from mamdr_b200.dataset import MultiDomainDataset
    g = synth.generate("Taobao-10", seed=5, scale=0.004)
    _write_reference_layout(str(tmp_path), g)
    conf = {"name": "Taobao", "dataset_path": str(tmp_path), "domain_split_path": "split_by_theme_x",
            "batch_size": 64, "shuffle_buffer_size": 10000, "num_parallel_reads": 8, "seed": 123}
    ds = MultiDomainDataset(conf, device=None)
    assert (ds.n_uid, ds.n_pid, ds.n_domain) == (g["n_uid"], g["n_pid"], 10)
    np.testing.assert_array_equal(ds.user_table, g["user_emb"])
    for d in range(10):
        np.testing.assert_array_equal(ds.train_dataset[d]["data"].host["uid"], g["train"][d]["uid"])
        np.testing.assert_array_equal(ds.test_dataset[d]["data"].host["label"], g["test"][d]["label"])
        n = len(g["train"][d]["uid"])
        assert ds.train_dataset[d]["n_data"] == n and ds.train_dataset[d]["n_step"] == int(np.ceil(n / 64.0))
    info = ds.dataset_info
    assert info["total_train"] == sum(len(g["train"][d]["uid"]) for d in range(10))
    assert info[0]["ctr_ratio"] == g["ctr_ratio"][0]


def test_dispatch_rules(monkeypatch):
    import run
    from mamdr_b200.dataset import MultiDomainDataset
    c = make_config(**{"dataset.synthetic.scale": 0.002})
    ds = MultiDomainDataset(c["dataset"], device=None)
    for name, exc in [("nothing", ValueError), ("wdl", NotImplementedError)]:
        c["model"]["name"] = name
        with pytest.raises(exc):
            run.build(c, dataset=ds)
    # the wrappers selected for the in-scope names (run.py:55-65) -- checked without building a model
    import mamdr_b200.deepctr as dc
    built = []
    monkeypatch.setattr(dc.DeepCTR, "__init__", lambda self, dataset, config: built.append(config['model']['name']))
    from mamdr_b200.domain_negotiation import DomainNegotiation
    from mamdr_b200.mamdr import MAMDR
    c["model"]["name"] = "mlp_meta_domain_negotiation_finetune_"
    assert type(run.build(c, dataset=ds)) is DomainNegotiation
    c["model"]["name"] = "mlp_meta_mamdr_finetune"
    assert type(run.build(c, dataset=ds)) is MAMDR
    c["model"]["name"] = "mlp"
    assert type(run.build(c, dataset=ds)) is dc.DeepCTR
    # 'star' is tested first (run.py:40): star_* names build the Star base model, then the same wrappers
    import mamdr_b200.star as st
    monkeypatch.setattr(st.Star, "__init__", lambda self, dataset, config: built.append(config['model']['name']))
    c["model"]["name"] = "star"
    assert type(run.build(c, dataset=ds)) is st.Star
    c["model"]["name"] = "star_meta_mamdr_finetune"
    assert type(run.build(c, dataset=ds)) is MAMDR
    # multi-task towers (run.py:38,44-45): mmoe / ple / shared_bottom -> DeepMTLCTR, then the same wrappers
    import mamdr_b200.deep_mtl_ctr as mt
    monkeypatch.setattr(mt.DeepMTLCTR, "__init__", lambda self, dataset, config: built.append(config['model']['name']))
    for name in ("mmoe", "ple", "shared_bottom"):
        c["model"]["name"] = name
        assert type(run.build(c, dataset=ds)) is mt.DeepMTLCTR
    c["model"]["name"] = "ple_meta_domain_negotiation"
    assert type(run.build(c, dataset=ds)) is DomainNegotiation
    from mamdr_b200.reptile import Reptile
    for name in ("mlp_meta_reptile_finetune", "mlp_meta_reptile_batch"):
        c["model"]["name"] = name
        assert type(run.build(c, dataset=ds)) is Reptile
    from mamdr_b200.maml import MAML
    from mamdr_b200.mldg import MLDG
    from mamdr_b200.pcgrad import PCGrad
    for name, cls in (("mlp_meta_mldg", MLDG), ("mlp_meta_maml_finetune", MAML), ("mlp_pcgrad", PCGrad)):   # row f4
        c["model"]["name"] = name
        assert type(run.build(c, dataset=ds)) is cls
    c["model"]["name"] = "mlp_uncertainty_weight"
    with pytest.raises(NotImplementedError):
        run.build(c, dataset=ds)


def test_mtl_topology_matches_oracle_layout():
    """Host logic of BASELINE config #5 (no GPU): the product's weight order / sub-model reachability equal the oracle's,
    and the variables of sub-model t form at most two contiguous arena spans (shared block, domain block)."""
    from mamdr_b200.deep_mtl_ctr import MTLTopology, init_mtl_weights
    from oracle.mtl import MTLSpec
    for kind in ("mmoe", "ple", "shared_bottom"):
        for trainable in (True, False):
            topo = MTLTopology(kind, 50, 40, 4, (16, 16, 8), (24, 12), (8,), (4,), num_experts=3, specific_expert_num=2,
                               shared_expert_num=2, emb_trainable=trainable)
            spec = MTLSpec(50, 40, 4, kind=kind, emb_dim=(16, 16, 8), expert_hidden=(24, 12), tower_hidden=(8,), gate_hidden=(4,),
                           num_experts=3, specific_expert_num=2, shared_expert_num=2, emb_trainable=trainable)
            assert topo.layout.names == spec.names and topo.layout.shapes == [tuple(s) for s in spec.shapes]
            assert topo.expert_sets == spec.expert_sets and topo.k == spec.k
            lo = topo.layout
            for t in range(4):
                assert sorted(topo.reachable(t)) == sorted(spec.reachable(t))
                spans = topo.dense_spans(t)
                assert 1 <= len(spans) <= 2
                covered = np.zeros(lo.total, dtype=bool)
                for b, n in spans:
                    assert b % 4 == 0 and n % 4 == 0
                    covered[b:b + n] = True
                for name, off, numel in zip(lo.names, lo.offsets, lo.numels):
                    if name in ('user_emb', 'item_emb'):
                        assert not covered[off:off + numel].any()
                    else:
                        assert covered[off:off + numel].all() == (name in topo.reachable(t))
                        assert covered[off:off + numel].any() == (name in topo.reachable(t))
            w = init_mtl_weights(lo, [1, 0])
            assert [x.shape for x in w] == lo.shapes and all(x.dtype == np.float32 for x in w)
    with pytest.raises(ValueError):
        MTLTopology("cgc", 5, 5, 2, (8, 8, 8), (8,), (8,), (4,))


def test_package_exports_the_reference_class_names():
    """run.py:6-15 of the reference imports these names; `from mamdr_b200 import <name>` must resolve each of them."""
    import mamdr_b200
    from mamdr_b200 import MAMDR, MAML, BaseModel, DeepCTR, DeepMTLCTR, DomainNegotiation, MultiDomainDataset, Reptile, Star
    assert issubclass(MAMDR, MAML) and issubclass(DomainNegotiation, MAML) and issubclass(Reptile, MAML)
    assert issubclass(DeepCTR, BaseModel) and issubclass(Star, BaseModel) and issubclass(DeepMTLCTR, BaseModel)
    assert MultiDomainDataset.__module__ == "mamdr_b200.dataset"
    assert sorted(mamdr_b200.__all__) == sorted(["MAML", "DomainNegotiation", "MAMDR", "Reptile", "MLDG", "PCGrad", "BaseModel", "DeepCTR", "Star",
                                                 "DeepMTLCTR", "MultiDomainDataset"])
    assert issubclass(mamdr_b200.PCGrad, MAML) and issubclass(mamdr_b200.MLDG, MAML)
    with pytest.raises(AttributeError):
        mamdr_b200.UncertaintyWeight
