import copy
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The oracle's fp32 GEMMs run on torch-CPU: their summation order -- and with it which near-zero ReLU gates open in a
    # free-running trajectory -- depends on the BLAS thread count, i.e. on the host.  Pin it so the checker is the same
    # function on every box.
    try:
        import torch
        torch.set_num_threads(4)
    except Exception:
        pass


BASE_CONFIG = {
    "model": {"name": "mlp_meta_mamdr_finetune", "norm": "none", "dense": "dense", "auxiliary_net": False,
              "user_dim": 128, "item_dim": 128, "domain_dim": 128, "auxiliary_dim": 128,
              "hidden_dim": [256, 128, 64], "dropout": 0.5},
    "train": {"load_pretrain_emb": True, "emb_trainable": False, "epoch": 2, "learning_rate": 0.001,
              "meta_learning_rate": 0.1, "domain_meta_learning_rate": 0.1, "merged_method": "plus",
              "sample_num": 2, "add_query_domain": True, "finetune_every_epoch": False,
              "shuffle_sequence": True, "meta_sequence": "random", "target_domain": -1,
              "domain_regulation_step": 0, "meta_train_step": 0, "meta_finetune_step": 0,
              "meta_split": "train-train", "meta_split_ratio": 0.8, "average_meta_grad": "none",
              "meta_parms": ["all"], "result_save_path": "result", "checkpoint_path": "checkpoint",
              "loss": "binary_crossentropy", "optimizer": "adam", "patience": 3, "val_every_step": 1,
              "histogram_freq": 0, "shuffle_buff_size": 10000},
    "dataset": {"name": "Taobao", "dataset_path": "dataset/Taobao", "domain_split_path": "split_by_theme_10",
                "batch_size": 1024, "shuffle_buffer_size": 10000, "num_parallel_reads": 8, "seed": 123,
                "synthetic": {"shape": "Taobao-10", "scale": 0.05, "signal": 1.0}},
    "b200": {"precision": "fp32", "device": "cuda:0", "cuda_graphs": True, "verbose": False},
}


def make_config(tmp_path=None, **over):
    """over: dotted keys, e.g. make_config(**{'model.name': 'mlp', 'dataset.synthetic.scale': 0.1})"""
    c = copy.deepcopy(BASE_CONFIG)
    if tmp_path is not None:
        c["train"]["result_save_path"] = str(tmp_path / "result")
        c["train"]["checkpoint_path"] = str(tmp_path / "checkpoint")
    for k, v in over.items():
        node = c
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return c


@pytest.fixture
def cfg(tmp_path):
    return lambda **over: make_config(tmp_path, **over)


def rel_err(a, b):
    """max |a-b| / max |b| (tensor-level relative error used for the 1e-4 parameter bar)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = max(float(np.max(np.abs(b))), 1e-30)
    return float(np.max(np.abs(a - b))) / den
