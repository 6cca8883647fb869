"""Injected meta-training schedule.

The reference draws the domain order (``random.shuffle``, ``/root/reference/model_zoo/mamdr.py:45-46``,
``model_zoo/domain_negotiation.py:41-42``) and the DR support domains (``random.sample``,
``mamdr.py:68``) from Python's *unseeded* global RNG, and reshuffles each training pass through an
unseeded tf.data shuffle buffer (``utils/dataset.py:78-82``) -- runs are not reproducible.  Here all
three come from one seeded object that is consumed in the reference's call order, so the GPU path,
the CPU oracle and every rank of a multi-GPU run see identical draws.
"""
import random

import numpy as np


class Schedule(object):
    def __init__(self, seed=123, shuffle_batches=True):
        self.seed = int(seed)
        self._rng = random.Random(self.seed)
        self._pass = 0
        self.shuffle_batches = shuffle_batches

    def shuffle_sequence(self, seq):
        seq = list(seq)
        self._rng.shuffle(seq)
        return seq

    def sample_support(self, candidates, k):
        return self._rng.sample(list(candidates), k=k)

    def reserve(self, k):
        """Reserve the ids of the next ``k`` training passes (the passes of one meta-step, numbered in
        the reference's sequential execution order) and return the first id.  A rank that runs only a
        shard of the passes still uses the global ids, so a pass's sample order does not depend on the
        sharding."""
        first = self._pass + 1
        self._pass += int(k)
        return first

    def batch_order(self, domain, n):
        """Sample order of the next training pass over ``domain`` (a fresh permutation per pass)."""
        return self.batch_order_at(self.reserve(1), domain, n)

    def batch_order_at(self, pass_id, domain, n):
        if not self.shuffle_batches:
            return np.arange(n, dtype=np.int32)
        g = np.random.Generator(np.random.PCG64([self.seed, int(pass_id), int(domain)]))
        return g.permutation(n).astype(np.int32)
