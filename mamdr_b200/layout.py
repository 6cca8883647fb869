"""Flat parameter arena layout + weight initialisers for the mlp tower.

Every trainable tensor of the model lives in ONE fp32 device buffer, ordered like the reference's
``model.trainable_weights`` (``/root/reference/model_zoo/maml.py:158-159``; for the mlp built in
``model_zoo/DeepCTR/deepctr.py:95-136``:
``[user_emb?, item_emb?, domain_emb, kernel0.., bias0.., dense_kernel, global_bias]``), each tensor
start aligned to 32 floats (128 B), padding kept zero.  theta, every theta_d, the live model, Adam
m / v and the best snapshots share this layout, so the meta ops of ``mamdr.py:168-196`` and the
``SetVarOp`` round trips (``utils/tool.py:36-45``) become single coalesced sweeps.
"""
import numpy as np

ALIGN = 32  # floats


class ParamLayout(object):
    def __init__(self, names, shapes):
        self.names = list(names)
        self.shapes = [tuple(int(x) for x in s) for s in shapes]
        self.numels = [int(np.prod(s)) for s in self.shapes]
        self.offsets = []
        off = 0
        for n in self.numels:
            self.offsets.append(off)
            off = (off + n + ALIGN - 1) // ALIGN * ALIGN
        self.total = off

    def index(self, name):
        return self.names.index(name)

    def offset(self, name):
        return self.offsets[self.index(name)] if name in self.names else -1

    def views(self, flat):
        """List of tensor views (torch or numpy) into a flat arena, in trainable_weights order."""
        return [flat[o:o + n].reshape(s) for o, n, s in zip(self.offsets, self.numels, self.shapes)]

    def pack(self, arrays, dtype=np.float32):
        flat = np.zeros(self.total, dtype=dtype)
        assert len(arrays) == len(self.names)
        for a, o, n, s in zip(arrays, self.offsets, self.numels, self.shapes):
            a = np.asarray(a, dtype=dtype)
            assert a.shape == s, (a.shape, s)
            flat[o:o + n] = a.reshape(-1)
        return flat

    def unpack(self, flat):
        flat = np.asarray(flat)
        return [flat[o:o + n].reshape(s).copy() for o, n, s in zip(self.offsets, self.numels, self.shapes)]


def mlp_layout(n_uid, n_pid, n_domain, emb_dim, hidden, emb_trainable):
    dims = (sum(emb_dim),) + tuple(hidden)
    L = len(hidden)
    names, shapes = [], []
    if emb_trainable:
        names += ['user_emb', 'item_emb']
        shapes += [(n_uid, emb_dim[0]), (n_pid, emb_dim[1])]
    names += ['domain_emb']
    shapes += [(n_domain, emb_dim[2])]
    names += ['kernel%d' % i for i in range(L)] + ['bias%d' % i for i in range(L)]
    shapes += [(dims[i], dims[i + 1]) for i in range(L)] + [(dims[i + 1],) for i in range(L)]
    names += ['dense_kernel', 'global_bias']
    shapes += [(dims[-1], 1), (1,)]
    return ParamLayout(names, shapes)


# ---- initialisers ([EXT] Keras / DeepCTR defaults) -----------------------------------
def _truncated_normal(rng, shape, stddev):
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def glorot_normal(rng, shape):
    fan_in, fan_out = shape[0], shape[1]
    stddev = np.sqrt(2.0 / (fan_in + fan_out)) / 0.87962566103423978
    return _truncated_normal(rng, shape, stddev)


def init_mlp_weights(layout, seed):
    """One draw of every layer's initialiser: Glorot-normal kernels, zero biases,
    RandomNormal(0, 1e-4) embeddings.  ``seed`` selects the draw, so the first build
    (theta) and each ``init_layer`` re-initialisation (theta_d^0,
    ``model_zoo/specific_base_model.py:174-178``) are independent samples."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for name, shape in zip(layout.names, layout.shapes):
        if name.endswith('_emb'):
            out.append((rng.standard_normal(shape) * 1e-4).astype(np.float32))
        elif name.startswith('kernel') or name == 'dense_kernel':
            out.append(glorot_normal(rng, shape))
        else:
            out.append(np.zeros(shape, dtype=np.float32))
    return out
