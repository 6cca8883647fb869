"""Multi-GPU sharding of the MAMDR meta-step (one process per GPU, ``torch.distributed``; NCCL on
the B200 box, gloo in the CPU tests).  The reference is single-process (``/root/reference/run.py:27-30``);
SURVEY.md section 8(e) defines the sharded semantics:

  * DN on the shared theta is strictly sequential (Alg. 1) -> every rank runs it redundantly from the
    same theta and the same schedule.  The kernels use no float atomics, so the replicas stay
    bit-identical with zero communication.
  * DR chains (one per query domain) only read theta and read/write their own theta_i -> the query
    domains are sharded across ranks, LPT-balanced by their cost sum_j (S_j + S_i).
  * ONE collective per meta-step: an all-reduce(sum) over a packed buffer holding, per domain, theta_i
    from its owner and zeros from everyone else (x + 0 is exact, so the exchange is bit-exact), plus the
    Adam slots (m, v, step, beta powers) AND the live model state (the whole parameter arena, i.e. also the variables
    that are not meta parameters when ``meta_parms`` is a subset -- they keep training through every pass and are never
    reloaded from theta -- and STAR's PartitionedNorm moving statistics) of the rank that owns the LAST query domain of
    the epoch's sequence, which all ranks adopt: every rank then starts the next replicated DN phase from the same
    state.  With world_size 1 this is exactly the reference schedule.
"""
import os

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's env (RANK / WORLD_SIZE / MASTER_*)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend)
    return world()


def lpt_assign(costs, n_ranks):
    """Longest-processing-time-first assignment.  ``costs``: {key: cost}.  Deterministic: ties go to
    the lower key / lower rank.  Returns {key: rank}."""
    load = [0] * n_ranks
    owner = {}
    for key in sorted(costs, key=lambda k: (-costs[k], k)):
        r = min(range(n_ranks), key=lambda i: (load[i], i))
        owner[key] = r
        load[r] += costs[key]
    return owner


def dr_chain_costs(sequence, supports, n_step, domain_regulation_step=0):
    """cost_i = sum_{j in J_i} (S_j + S_i') mini-batches (mamdr.py:75-97)."""
    costs = {}
    for idx in sequence:
        s_i = n_step[idx]
        if domain_regulation_step and domain_regulation_step > 0:
            s_i = min(s_i, domain_regulation_step)
        costs[idx] = sum(n_step[j] + s_i for j in supports[idx])
    return costs


def dr_pair_costs(sequence, supports, n_step, domain_regulation_step=0):
    """'batch' names: cost of the (query at sequence position pos, k-th support) pair = S_j + S_i' mini-batches
    (mamdr.py:85-101); keys (pos, k)."""
    costs = {}
    for pos, idx in enumerate(sequence):
        s_i = n_step[idx]
        if domain_regulation_step and domain_regulation_step > 0:
            s_i = min(s_i, domain_regulation_step)
        for k, j in enumerate(supports[idx]):
            costs[(pos, k)] = n_step[j] + s_i
    return costs


def exchange_sum(accum_all, m, v, opt_words, rank, last_owner, extra=()):
    """The one collective of a pair-sharded ('batch') meta-step: the per-query-domain delta accumulators [n_domain, P] are
    SUMMED over the ranks; the Adam slots / beta powers / ``extra`` tensors are adopted from ``last_owner`` (everyone else
    contributes zeros).  In place on every rank."""
    P = m.numel()
    n_acc = accum_all.numel()
    n_extra = sum(int(t.numel()) for t in extra)
    buf = torch.zeros(n_acc + 2 * P + 4 + n_extra, dtype=torch.float32, device=m.device)
    buf[:n_acc].copy_(accum_all.reshape(-1))
    if rank == last_owner:
        buf[n_acc:n_acc + P].copy_(m)
        buf[n_acc + P:n_acc + 2 * P].copy_(v)
        buf[n_acc + 2 * P:n_acc + 2 * P + 3].copy_(opt_words)
        o = n_acc + 2 * P + 4
        for t in extra:
            buf[o:o + t.numel()].copy_(t.reshape(-1))
            o += t.numel()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    accum_all.copy_(buf[:n_acc].reshape(accum_all.shape))
    m.copy_(buf[n_acc:n_acc + P])
    v.copy_(buf[n_acc + P:n_acc + 2 * P])
    opt_words.copy_(buf[n_acc + 2 * P:n_acc + 2 * P + 3])
    o = n_acc + 2 * P + 4
    for t in extra:
        t.copy_(buf[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    return buf.numel() * 4


def exchange(owner, rank, domain_flats, m, v, opt_words, last_owner, extra=()):
    """The one collective of a sharded meta-step (see module docstring).  ``domain_flats``: {idx: flat
    theta_i tensor}; ``opt_words``: float32 tensor [3] = (step, b1pow, b2pow) of this rank; ``extra``: further float32
    tensors (live parameter arena, normalisation state) adopted from ``last_owner``.  All tensors are updated in place
    on every rank."""
    keys = sorted(domain_flats)
    P = m.numel()
    n_extra = sum(int(t.numel()) for t in extra)
    buf = torch.zeros((len(keys) + 2) * P + 4 + n_extra, dtype=torch.float32, device=m.device)
    for k, idx in enumerate(keys):
        if owner[idx] == rank:
            buf[k * P:(k + 1) * P].copy_(domain_flats[idx])
    base = len(keys) * P
    if rank == last_owner:
        buf[base:base + P].copy_(m)
        buf[base + P:base + 2 * P].copy_(v)
        buf[base + 2 * P:base + 2 * P + 3].copy_(opt_words)
        o = base + 2 * P + 4
        for t in extra:
            buf[o:o + t.numel()].copy_(t.reshape(-1))
            o += t.numel()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    for k, idx in enumerate(keys):
        domain_flats[idx].copy_(buf[k * P:(k + 1) * P])
    m.copy_(buf[base:base + P])
    v.copy_(buf[base + P:base + 2 * P])
    opt_words.copy_(buf[base + 2 * P:base + 2 * P + 3])
    o = base + 2 * P + 4
    for t in extra:
        t.copy_(buf[o:o + t.numel()].reshape(t.shape))
        o += t.numel()
    return buf.numel() * 4
