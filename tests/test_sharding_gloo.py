"""The N > 1 host logic on CPU: world_size-2 gloo process groups (no GPU).

Covers shard ranges (pairs / hypotheses) and the winner-record reduction that carries the (FP64 score, index) argmax
across ranks: strictly larger FP64 score wins -- also when the two scores round to the SAME float, the case the round-1
packed key lost --, equal scores go to the lower 64-bit index, ranks without a model are ignored, and every rank ends up
with the identical record (model included).  The device kernel k_winners_reduce applies the same rule
(tests/test_gpu_sharded.py compares the two)."""
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


def rec(score, index, count=0, model=None, valid=1):
    from pygcransac import _cabi
    r = np.zeros((), dtype=_cabi.WINNER_DTYPE)
    if valid:
        r["score"], r["index"], r["count"], r["valid"] = score, index, count, 1
        r["model"][:9] = np.arange(9) + index if model is None else model
    return r


# pair 0: rank 1 better; pair 1: tie -> lower global index (rank 0, even though rank 1's index is > 2^32 apart);
# pair 2: rank 1 has no model; pair 3: scores 1 ulp of FP64 apart that round to the same float -> FP64 decides;
# pair 4: nobody has a model
_S = 524288.0
LOCAL = {0: [rec(10.0, 40), rec(7.5, 90), rec(3.0, 5), rec(_S, 7), rec(0, 0, valid=0)],
         1: [rec(11.0, 1040), rec(7.5, 5 * 2 ** 32 + 1), rec(0, 0, valid=0), rec(np.nextafter(_S, 2 * _S), 2 ** 33), rec(0, 0, valid=0)]}
EXPECT = [LOCAL[1][0], LOCAL[0][1], LOCAL[0][2], LOCAL[1][3], rec(0, 0, valid=0)]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        local = np.stack(LOCAL[rank])
        got = sharded.reduce_winner_records(sharded.pack_records(local))
        out.put((rank, got.tobytes()))
        # sharded sample ranges: every sample has exactly one owner and the owners agree with the ranges
        lo, hi = sharded.shard_range(1001, rank, world)
        mine = torch.zeros(1001, dtype=torch.int64)
        mine[lo:hi] = 1
        dist.all_reduce(mine)
        assert bool((mine == 1).all())
        assert all(sharded.owner_of(s, 1001, world) == rank for s in (lo, hi - 1))
    finally:
        dist.destroy_process_group()


def test_winner_record_reduce_two_ranks():
    assert np.float32(LOCAL[0][3]["score"]) == np.float32(LOCAL[1][3]["score"])      # an FP32 key could not tell them apart
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want = np.stack(EXPECT).tobytes()
    assert got[0] == want and got[1] == want                                        # bit-identical on both ranks
