"""Synthetic multi-domain CTR data in the reference's dataset contract (no network => no datasets).

Reproduces the *shape* of what ``/root/reference/dataset/{Taobao,Amazon}/*.py`` write and
``utils/dataset.py:41-99`` reads: per-domain ``train/val/test`` tables with columns
``uid, pid, domain, label`` (``dataset/Taobao/split.py:21``), ``n_uid`` / ``n_pid`` counts, and for
Taobao the frozen 128-d user / item embedding tables.  Sizes are the published Table-I numbers
(ICDE'23 slides p.24).  Generator: SURVEY.md section 8(d).
"""
import numpy as np

# name -> (n_domain, n_uid, n_pid, train, val, test, pretrained_tables)
SHAPES = {
    "Taobao-10": (10, 23778, 6932, 92137, 37645, 43502, True),
    "Taobao-20": (20, 58190, 16319, 243592, 96591, 106500, True),
    "Taobao-30": (30, 99143, 29945, 394805, 151369, 179252, True),
    "Amazon-6": (6, 445789, 172653, 9968333, 3372666, 3585877, False),
    "Amazon-13": (13, 502222, 215403, 11999607, 4100756, 4339523, False),
}


def _split_sizes(total, n_domain):
    """n_d proportional to (d+1)^-1 (long tail), largest-remainder rounding, each >= 1."""
    w = 1.0 / np.arange(1, n_domain + 1)
    w /= w.sum()
    raw = w * total
    n = np.maximum(1, np.floor(raw).astype(np.int64))
    rem = int(total - n.sum())
    if rem > 0:
        order = np.argsort(-(raw - np.floor(raw)))
        n[order[:rem]] += 1
    return n


def _zipf_sampler(rng, n_items, s=1.05):
    ranks = np.arange(1, n_items + 1, dtype=np.float64)
    cdf = np.cumsum(ranks ** (-s))
    cdf /= cdf[-1]
    perm = rng.permutation(n_items).astype(np.int32)

    def draw(n):
        r = np.searchsorted(cdf, rng.random(n), side="left")
        return perm[np.minimum(r, n_items - 1)]
    return draw


def generate(shape="Taobao-10", seed=123, scale=1.0, signal=1.0, emb_dim=128):
    """Returns a dict: n_uid, n_pid, n_domain, train/val/test {d: {uid,pid,label}}, ctr_ratio {d},
    user_emb/item_emb (float32 tables or None), name."""
    D, n_uid, n_pid, n_tr, n_va, n_te, pretrained = SHAPES[shape]
    if scale != 1.0:
        n_uid, n_pid = max(8, int(n_uid * scale)), max(8, int(n_pid * scale))
        n_tr, n_va, n_te = (max(D, int(x * scale)) for x in (n_tr, n_va, n_te))
    rng = np.random.Generator(np.random.PCG64(seed))
    draw_u = _zipf_sampler(rng, n_uid)
    draw_p = _zipf_sampler(rng, n_pid)
    # latent factors: the frozen "pretrained" tables N(0, 0.05^2) (stand-in for Taobao's
    # user_embedding.csv / item_embedding.csv); for Amazon they only drive the labels
    fu = (rng.standard_normal((n_uid, emb_dim)) * 0.05).astype(np.float32)
    fi = (rng.standard_normal((n_pid, emb_dim)) * 0.05).astype(np.float32)
    ctr = np.round(rng.uniform(0.2, 0.5, size=D), 2)                 # dataset/Taobao/split.py:110-112
    dom_sign = np.where(rng.random(D) < 0.3, -1.0, 1.0)               # some domains conflict
    out = {"name": shape, "n_uid": n_uid, "n_pid": n_pid, "n_domain": D, "ctr_ratio": {},
           "train": {}, "val": {}, "test": {},
           "user_emb": fu if pretrained else None, "item_emb": fi if pretrained else None}
    sizes = {"train": _split_sizes(n_tr, D), "val": _split_sizes(n_va, D), "test": _split_sizes(n_te, D)}
    for d in range(D):
        out["ctr_ratio"][d] = float(ctr[d])
        base = np.log(ctr[d])  # logit of c/(1+c)
        for split in ("train", "val", "test"):
            n = int(sizes[split][d])
            uid, pid = draw_u(n), draw_p(n)
            logit = base + signal * dom_sign[d] * 35.0 * np.einsum("ij,ij->i", fu[uid], fi[pid])
            label = (rng.random(n) < 1.0 / (1.0 + np.exp(-logit))).astype(np.float32)
            out[split][d] = {"uid": uid.astype(np.int32), "pid": pid.astype(np.int32), "label": label}
    return out
