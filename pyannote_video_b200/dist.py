"""Multi-GPU plumbing: one process per GPU (torchrun), frames/shots sharded by rank, and the path's
single exchange — an all-gather of per-rank embeddings and track ids before clustering
(SURVEY.md §8e).  Works with the `nccl` backend on device tensors and with `gloo` on CPU tensors
(used by the world-size-2 CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """contiguous shard [lo, hi) of n_items for this rank (sizes differ by at most one)"""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_shots(shot_frames, world):
    """assign whole shots (list of frame counts) to ranks, balancing frame counts greedily in order;
    returns a list of rank ids, one per shot (tracking state resets at shot boundaries, so a rank
    must own whole shots: pyannote/video/tracking.py:410-417)."""
    total = float(sum(shot_frames))
    out, acc, rank = [], 0.0, 0
    for n in shot_frames:
        if rank < world - 1 and acc + n / 2.0 > total * (rank + 1) / world:
            rank += 1
        out.append(rank)
        acc += n
    return out


def gather_embeddings(emb, track_ids, group=None):
    """all-gather variable-length [n_i,128] embeddings and [n_i] track ids; track ids are made
    globally unique as rank * 2^32... (rank, local id) -> dense remap is left to the caller: the
    returned ids are `rank * stride + local` with stride = max local id + 1 over all ranks."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = emb.device
    n = torch.tensor([emb.shape[0], int(track_ids.max().item()) + 1 if track_ids.numel() else 0], device=dev,
                     dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    sizes = [int(c[0].item()) for c in counts]
    stride = max(int(c[1].item()) for c in counts)
    if track_ids.numel() and int(track_ids.min().item()) < 0:
        raise ValueError("gather_embeddings: negative track id")
    if stride * world >= 2 ** 62:
        raise ValueError("gather_embeddings: track ids up to %d cannot be made unique over %d ranks in int64 "
                         "(uninitialised ids?)" % (stride - 1, world))
    nmax = max(sizes)
    pad_e = torch.zeros(nmax, emb.shape[1], dtype=emb.dtype, device=dev)
    pad_t = torch.zeros(nmax, dtype=torch.int64, device=dev)
    pad_e[:emb.shape[0]] = emb
    pad_t[:emb.shape[0]] = track_ids.to(torch.int64) + rank * stride
    all_e = [torch.empty_like(pad_e) for _ in range(world)]
    all_t = [torch.empty_like(pad_t) for _ in range(world)]
    dist.all_gather(all_e, pad_e, group=group)
    dist.all_gather(all_t, pad_t, group=group)
    E = torch.cat([e[:s] for e, s in zip(all_e, sizes)])
    T = torch.cat([t[:s] for t, s in zip(all_t, sizes)])
    return E, T
