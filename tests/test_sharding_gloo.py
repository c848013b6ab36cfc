"""The N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU).

Covers shard ranges (pairs / hypotheses), the packed-key all-reduce that carries the (score, index) argmax across ranks
(first-wins on ties, uint64 order preserved through the int64 transport) and the owner lookup for the model broadcast."""
import os
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "graph-cut-ransac_b200"))
from pygcransac import sharded  # noqa: E402


def pack_key(score, index):
    """Python mirror of the device key (csrc/k_select.cu): orderedFloatBits(score) << 32 | (0xFFFFFFFF - index)."""
    b = struct.unpack("<I", struct.pack("<f", score))[0]
    o = (~b & 0xFFFFFFFF) if (b & 0x80000000) else (b | 0x80000000)
    return (o << 32) | (0xFFFFFFFF - index)


def as_i64(u):
    return u - (1 << 64) if u >= (1 << 63) else u


def test_shard_ranges_cover_and_balance():
    for total in (0, 1, 7, 8, 1000, 16384, 1048576):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharded.shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
            for u in {0, total // 3, total - 1} if total else ():
                r = sharded.owner_of(u, total, world)
                assert ranges[r][0] <= u < ranges[r][1]


def test_key_packing_matches_library():
    from pygcransac import _cabi
    lib = _cabi.load_library()          # host-side helpers only; no device needed
    for score, idx in ((12.5, 800), (0.0, 5099), (1e-3, 0), (3.4e38, 2 ** 32 - 2), (-1.0, 17)):
        k = pack_key(score, idx)
        assert lib.gcr_key_index(k) == idx
        assert np.float32(lib.gcr_key_score(k)) == np.float32(score)
    # larger score wins; equal scores: lower index wins
    assert pack_key(2.0, 900) > pack_key(1.0, 3) and pack_key(1.0, 3) > pack_key(1.0, 4) > pack_key(-1.0, 0)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # three "pairs": rank-local winners; pair 1 is a cross-rank tie (lower global index must win), pair 2 has no model on rank 1
        local = {0: [pack_key(10.0, 40), pack_key(7.5, 90), pack_key(3.0, 5)],
                 1: [pack_key(11.0, 1040), pack_key(7.5, 1001), 0]}[rank]
        keys = torch.tensor([as_i64(k) for k in local], dtype=torch.int64)
        sharded.reduce_keys(keys)
        got = [int(k) & ((1 << 64) - 1) for k in keys]
        if rank == 0:
            out.put(got)
        # sharded sample ranges: every sample has exactly one owner and the owners agree with the ranges
        lo, hi = sharded.shard_range(1001, rank, world)
        mine = torch.zeros(1001, dtype=torch.int64)
        mine[lo:hi] = 1
        dist.all_reduce(mine)
        assert bool((mine == 1).all())
        assert all(sharded.owner_of(s, 1001, world) == rank for s in (lo, hi - 1))
    finally:
        dist.destroy_process_group()


def test_key_allreduce_first_wins_two_ranks():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got == [pack_key(11.0, 1040), pack_key(7.5, 90), pack_key(3.0, 5)]
    assert sharded.key_index(as_i64(got[1])) == 90
