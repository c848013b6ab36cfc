"""Multi-GPU plumbing for the two sharding modes of SURVEY.md 8(e) -- one process per GPU, torch.distributed.

* pairs shard (batched image pairs): contiguous ranges of pairs per rank, NO collective on the data path.
* hypotheses shard (one huge problem): rank g scores samples [g K / G, (g + 1) K / G) of the counter-based stream
  (sample i is a pure function of (seed, i), so no rank needs another's samples); the per-rank winner keys
  (orderedFloat(score) << 32 | ~index, see gcr_argmax) are all-reduced with MAX -- max is the reference's strict-'<'
  first-wins rule (GCRANSAC.h:440) across ranks as well -- and the owner broadcasts the 9-double model.

The key tensor the all-reduce runs on is the buffer K4's last kernel wrote (gcr_round's d_keys): there is no staging
copy between the kernel and NCCL.  Everything here also runs on CPU tensors over gloo (tests/test_sharding_gloo.py).
"""
import torch
import torch.distributed as dist

_FLIP = -(1 << 63)  # 0x8000...0 as int64: key ^ _FLIP maps uint64 order onto int64 order


def shard_range(total, rank, world):
    """Contiguous, balanced [start, stop) of `total` units for `rank` (first `total % world` ranks get one more)."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def owner_of(unit, total, world):
    """Inverse of shard_range: which rank owns unit `unit`."""
    base, extra = divmod(total, world)
    boundary = extra * (base + 1)
    if unit < boundary:
        return unit // (base + 1)
    return extra + (unit - boundary) // max(base, 1)


def reduce_keys(keys_u64_as_i64, group=None):
    """All-reduce packed (score, index) keys with MAX.  `keys_u64_as_i64`: int64 tensor holding the uint64 bit patterns
    gcr_round / gcr_argmax produced (device or CPU).  In place; returns the tensor."""
    keys_u64_as_i64.bitwise_xor_(_FLIP)
    dist.all_reduce(keys_u64_as_i64, op=dist.ReduceOp.MAX, group=group)
    keys_u64_as_i64.bitwise_xor_(_FLIP)
    return keys_u64_as_i64


def key_index(key_i64):
    return 0xFFFFFFFF - (int(key_i64) & 0xFFFFFFFF)


def sharded_round(ctx, problem, scoring, sampler, first_index, total_samples, seed, thr, flags, keys, best, max_models,
                  rank, world, group=None, chunk_samples=65536):
    """One fixed-K pass with the K samples sharded over the ranks.  `keys` int64 [n_pairs] and `best` float64
    [n_pairs, model_len] are device tensors; on return every rank holds the global winner key and model.
    A rank walks its share in chunks of `chunk_samples` (bounds the per-call workspace and the FP64 re-check list:
    N = 2^20 points x 65536 samples is ~2e10 evaluations, ~1e6 of them inside the threshold band) and keeps the running
    maximum of the chunk keys on the device.  Returns the model-slot x point evaluations THIS rank enqueued."""
    lo, hi = shard_range(total_samples, rank, world)
    evals = 0
    run_keys = torch.full_like(keys, _FLIP)          # flipped domain: smallest possible key
    run_best = torch.zeros_like(best)
    for c0 in range(lo, hi, chunk_samples):
        c1 = min(hi, c0 + chunk_samples)
        evals += ctx.round_device(problem, scoring, sampler, first_index + c0, c1 - c0, seed, thr, flags, c0 * max_models,
                                  keys.data_ptr(), best.data_ptr())
        flipped = keys ^ _FLIP
        better = flipped > run_keys                   # strictly: an equal key cannot occur (indices are unique)
        run_keys = torch.where(better, flipped, run_keys)
        run_best = torch.where(better[:, None], best, run_best)
    keys.copy_(run_keys ^ _FLIP)
    best.copy_(run_best)
    if world > 1:
        reduce_keys(keys, group)
        # the winner's model lives on the rank that scored it
        for p in range(keys.numel()):
            k = int(keys[p].item())
            if k == 0:
                continue
            sample = key_index(k) // max_models
            dist.broadcast(best[p], src=owner_of(sample, total_samples, world), group=group)
    return evals
