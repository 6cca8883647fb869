"""Routing plan of the row-sharded tables, numpy restatement of the `mamdr_route_plan` / `mamdr_route_pack_rows` contract
(include/mamdr_b200.h).  TEST INFRASTRUCTURE (see oracle/__init__.py): the product (mamdr_b200/sharded.py) runs the CUDA kernels.

The reference has no sharding (whole tables in one TF variable, model_zoo/DeepCTR/deepctr.py:105-126); this is the layout
north_star asks for: table row r lives on rank r % world at local index r // world, every rank sends each owner a
FIXED-capacity block of `cap` entries (-1 = padding) so that all all-to-all splits are static.
"""
import numpy as np


def route_plan(ids, world, cap):
    """-> (slot [n], send [world * cap]): slot[i] = owner * cap + (number of earlier ids with the same owner); send[slot[i]] =
    ids[i] // world; every other entry of send is -1."""
    ids = np.asarray(ids, dtype=np.int64)
    n = len(ids)
    assert n <= cap
    slot = np.zeros(n, dtype=np.int32)
    send = np.full(world * cap, -1, dtype=np.int32)
    seen = [0] * world
    for i, r in enumerate(ids):
        o = int(r % world)
        slot[i] = o * cap + seen[o]
        send[slot[i]] = r // world
        seen[o] += 1
    return slot, send


def route_plan_shared(ids_a, ids_b, world, cap):
    """Both id columns in ONE exchange buffer (block = 2 * cap): owner block r = [cap entries of a | cap entries of b].
    -> (slot_a, slot_b, send [world * 2 * cap]); the slots index the shared buffer."""
    sa, xa = route_plan(ids_a, world, cap)
    sb, xb = route_plan(ids_b, world, cap)
    send = np.full(world * 2 * cap, -1, dtype=np.int32)
    send.reshape(world, 2, cap)[:, 0, :] = xa.reshape(world, cap)
    send.reshape(world, 2, cap)[:, 1, :] = xb.reshape(world, cap)
    slot_a = (sa // cap) * 2 * cap + sa % cap
    slot_b = (sb // cap) * 2 * cap + cap + sb % cap
    return slot_a.astype(np.int32), slot_b.astype(np.int32), send


def pack_rows(src, slot, world, cap, scale=1.0):
    """dst[slot[i]] = src[i] * scale; the other rows of dst [world * cap, dim] are padding (zero here)."""
    src = np.asarray(src, dtype=np.float32)
    dst = np.zeros((world * cap, src.shape[1]), dtype=np.float32)
    dst[slot] = src * np.float32(scale)
    return dst
