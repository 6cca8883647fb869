"""``MultiDomainDataset`` -- same attribute surface as ``/root/reference/utils/dataset.py:41-130``
(``n_uid n_pid n_domain train_dataset[idx]={'data','n_step','n_data'} val_dataset test_dataset
item_emb user_emb ctr_ratio dataset_info``), but each split is parsed ONCE into a device-resident
column store (``engine.DomainData``) instead of a tf.data CSV pipeline re-read on every pass.

Two sources:
  * on-disk, the reference's format (``<dataset_path>/<domain_split_path>/processed_data/{uid2id,
    pid2id}.json``, Taobao ``{item,user}_emb.json``, ``domain_<k>/{train,val,test}.csv`` with header
    ``uid,pid,domain,label``, ``domain_<k>/domain_property.json``);
  * synthetic (``conf['synthetic'] = {'shape': 'Taobao-10', 'scale': 1.0, 'signal': 1.0}``), used when
    no dataset can be downloaded (``mamdr_b200/synth.py``).
"""
import collections
import glob
import json
import os.path as osp

import numpy as np

from . import synth
from .engine import DomainData


def _read_csv(path):
    with open(path, "r") as f:
        header = f.readline().strip().split(",")
    cols = {name: i for i, name in enumerate(header)}
    for need in ("uid", "pid", "label"):
        if need not in cols:
            raise ValueError("%s: missing column %r (header %r)" % (path, need, header))
    try:
        arr = np.loadtxt(path, delimiter=",", skiprows=1, ndmin=2, dtype=np.float64)
    except ValueError:
        arr = np.zeros((0, len(header)))
    if arr.size == 0:
        arr = np.zeros((0, len(header)))
    return (arr[:, cols["uid"]].astype(np.int32), arr[:, cols["pid"]].astype(np.int32),
            arr[:, cols["label"]].astype(np.float32))


def _emb_table(emb_dict, n, dim):
    """``DeepCTR.build_emb`` parsing (model_zoo/DeepCTR/deepctr.py:105-110): missing ids stay zero."""
    table = np.zeros((n, dim), dtype=np.float32)
    for key in sorted(emb_dict.keys()):
        table[int(key)] = np.asarray(emb_dict[key].split(" "), dtype="float32")
    return table


class MultiDomainDataset(object):
    def __init__(self, conf, device="cuda:0"):
        self.conf = conf
        self.dataset_path = conf['dataset_path']
        self.domain_split_path = osp.join(self.dataset_path, conf['domain_split_path'])
        self.seed = conf['seed']
        self.batch_size = conf['batch_size']
        self.shuffle_buffer_size = conf['shuffle_buffer_size']
        self.device = device
        self.train_dataset = collections.OrderedDict()
        self.val_dataset = collections.OrderedDict()
        self.test_dataset = collections.OrderedDict()
        self.ctr_ratio = collections.OrderedDict()
        self.user_table = self.item_table = None
        if conf.get('synthetic'):
            self._from_synthetic(conf['synthetic'])
        else:
            self._from_disk()

    def _add_domain(self, idx, splits, ctr_ratio):
        self.ctr_ratio[idx] = ctr_ratio
        for name, store in (("train", self.train_dataset), ("val", self.val_dataset), ("test", self.test_dataset)):
            s = splits[name]
            dd = DomainData(s["uid"], s["pid"], s["label"], idx, self.batch_size, self.device)
            store[idx] = {"data": dd, "n_step": dd.n_step, "n_data": dd.n_data}

    def _from_synthetic(self, sconf):
        g = synth.generate(sconf.get('shape', 'Taobao-10'), seed=sconf.get('seed', self.seed),
                           scale=sconf.get('scale', 1.0), signal=sconf.get('signal', 1.0))
        self.synthetic = g
        self.n_uid, self.n_pid, self.n_domain = g["n_uid"], g["n_pid"], g["n_domain"]
        self.user_table, self.item_table = g["user_emb"], g["item_emb"]
        print("Found {} domain, in: synthetic {}".format(self.n_domain, g["name"]))
        for d in range(self.n_domain):
            self._add_domain(d, {k: g[k][d] for k in ("train", "val", "test")}, g["ctr_ratio"][d])

    def _from_disk(self):
        with open(osp.join(self.domain_split_path, "processed_data/uid2id.json"), "r") as f:
            self.n_uid = json.load(f)['id']
        with open(osp.join(self.domain_split_path, "processed_data/pid2id.json"), "r") as f:
            self.n_pid = json.load(f)['id']
        if self.conf['name'] == "Taobao":
            with open(osp.join(self.domain_split_path, "processed_data/item_emb.json"), "r") as f:
                self.item_emb = json.load(f)
            with open(osp.join(self.domain_split_path, "processed_data/user_emb.json"), "r") as f:
                self.user_emb = json.load(f)
            dim_u = len(next(iter(self.user_emb.values())).split(" "))
            dim_i = len(next(iter(self.item_emb.values())).split(" "))
            self.user_table = _emb_table(self.user_emb, self.n_uid, dim_u)
            self.item_table = _emb_table(self.item_emb, self.n_pid, dim_i)
        domains_list = glob.glob(osp.join(self.domain_split_path, "domain_*"))
        domains_list.sort(key=lambda x: int(x.split("_")[-1]))
        self.n_domain = len(domains_list)
        print("Found {} domain, in: {}".format(self.n_domain, self.domain_split_path))
        for d_path in domains_list:
            domain_idx = int(osp.split(d_path)[-1].split("_")[-1])
            splits = {}
            for name in ("train", "val", "test"):
                uid, pid, label = _read_csv(osp.join(d_path, name + ".csv"))
                splits[name] = {"uid": uid, "pid": pid, "label": label}
            with open(osp.join(d_path, "domain_property.json")) as f:
                ctr = json.load(f)['ctr_ratio']
            self._add_domain(domain_idx, splits, ctr)

    def get_train_dataset(self, domain_idx):
        return self.train_dataset[domain_idx]

    def get_val_dataset(self, domain_idx):
        return self.val_dataset[domain_idx]

    def get_test_dataset(self, domain_idx):
        return self.test_dataset[domain_idx]

    def host_splits(self):
        """{'train'|'val'|'test': {domain: {'uid','pid','label'}}} host numpy views (tests / bench)."""
        out = {}
        for name, store in (("train", self.train_dataset), ("val", self.val_dataset), ("test", self.test_dataset)):
            out[name] = collections.OrderedDict((k, v["data"].host) for k, v in store.items())
        return out

    @property
    def dataset_info(self):
        total_train, total_val, total_test = 0, 0, 0
        info = {'n_user': self.n_uid, 'n_item': self.n_pid}
        for i in self.train_dataset:
            info[i] = {"n_train": self.train_dataset[i]['n_data'], "n_val": self.val_dataset[i]['n_data'],
                       "n_test": self.test_dataset[i]['n_data'], "ctr_ratio": self.ctr_ratio[i]}
            total_train += self.train_dataset[i]['n_data']
            total_val += self.val_dataset[i]['n_data']
            total_test += self.test_dataset[i]['n_data']
        info["total_train"], info['total_val'], info['total_test'] = total_train, total_val, total_test
        return info
