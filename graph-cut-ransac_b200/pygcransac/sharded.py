"""Multi-GPU plumbing for the two sharding modes of SURVEY.md 8(e) -- one process per GPU.

* pairs shard (batched image pairs): contiguous ranges of pairs per rank, NO collective on the data path.
* hypotheses shard (one huge problem): rank g scores samples [g K / G, (g + 1) K / G) of the counter-based stream
  (sample i is a pure function of (seed, i), so no rank needs another's samples).  The data plane is behind the C ABI:
  `gcr_round_sharded` walks the rank's range in chunks, keeps the running winner as a 128-byte `gcr_winner` record on
  the device (FP64 score, 64-bit index, exact inlier count, the model), all-gathers the records with NCCL over NVLink
  and reduces them on every rank by the reference's rule -- strictly larger FP64 score, ties to the lower index
  (GCRANSAC.h:440, scoring_function.h:63-66).  No host synchronisation, no second collective for the model.

What is left in Python is plumbing: handing the NCCL unique id to every rank (`init_comm`, over torch.distributed) and
the same reduction rule on CPU tensors (`reduce_winner_records`) so that the N > 1 host logic can be tested over gloo
without a GPU (tests/test_sharding_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous, balanced [start, stop) of `total` units for `rank` (first `total % world` ranks get one more);
    the same split gcr_round_sharded makes (csrc/gcr_comm.cu)."""
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


def init_comm(ctx, rank, world, group=None):
    """Joins `ctx` (a _cabi.Context on this rank's GPU) to the NCCL communicator of the process group: rank 0 creates the
    unique id, torch.distributed carries its 128 bytes to the other ranks."""
    from . import _cabi
    if world == 1:
        return
    box = [_cabi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(box[0], rank, world)


def sharded_round(ctx, problem, scoring, sampler, first_index, total_samples, seed, thr, flags, d_winners,
                  chunk_samples=65536):
    """One fixed-K pass with the K samples sharded over the ranks of ctx's communicator (see init_comm); on return (stream
    order) every rank's `d_winners` (device pointer, [n_pairs] gcr_winner records) holds the global winner.
    Returns the model-slot x point evaluations THIS rank enqueued."""
    return ctx.round_sharded_device(problem, scoring, sampler, first_index, total_samples, seed, thr, flags, chunk_samples,
                                    d_winners)


# ---- the reduction rule on host tensors (mirror of k_winners_reduce; CPU / gloo testable) ---------------------------
def pack_records(records):
    """numpy WINNER_DTYPE records [n_pairs] -> int64 tensor [n_pairs, 16] (the raw 128 bytes)."""
    return torch.from_numpy(np.ascontiguousarray(records).view(np.int64).reshape(len(records), 16).copy())


def unpack_records(t):
    from . import _cabi
    return t.numpy().reshape(-1).view(_cabi.WINNER_DTYPE)


def reduce_winner_records(local, group=None):
    """All-gathers every rank's records (int64 [n_pairs, 16], see pack_records) and returns the per-pair winner by
    (FP64 score, lower index): identical on every rank."""
    world = dist.get_world_size(group)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local, group=group)
    recs = [unpack_records(p) for p in parts]
    out = recs[0].copy()
    for p in range(len(out)):
        best = None
        for r in range(world):
            w = recs[r][p]
            if not w["valid"]:
                continue
            if best is None or w["score"] > best["score"] or (w["score"] == best["score"] and w["index"] < best["index"]):
                best = w
        if best is None:
            out[p] = np.zeros((), dtype=out.dtype)
        else:
            out[p] = best
    return out
