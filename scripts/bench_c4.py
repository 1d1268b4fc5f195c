"""BASELINE.json configs[3] (C4): synthetic 4K frames, a FIXED total frame count sharded over the ranks of one node
(`dist.shard_range`), per-rank variable face counts, detect + landmarks + embed per shard, then the path's one
exchange — a variable-count all-gather of the embeddings over NCCL (`dist.gather_embeddings`) — and the replicated
clustering tail, all inside the timed region.  STRONG scaling: total work is fixed as N grows.

    python scripts/bench_c4.py --frames 2048                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           scripts/bench_c4.py --frames 2048

Prints one JSON line on rank 0 (also appended to gpurun_out/c4_bench.jsonl): frames/s over the whole job, the frame
phase (max over ranks), the gather and the clustering tail separately (the Amdahl term SURVEY.md §8(e) warns about).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

H, W = 2160, 3840


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2048, help="total 4K frames of the job (BASELINE C4: 10000)")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--pool", type=int, default=16, help="distinct synthetic frames per rank (cycled; 16 x 25 MB > L2)")
    ap.add_argument("--threshold", type=float, default=0.6)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from pyannote_video_b200 import weights as Wt
    from pyannote_video_b200.face import Face
    from pyannote_video_b200.synth import make_frames, make_boxes
    from pyannote_video_b200.dist import shard_range, gather_embeddings
    from pyannote_video_b200.clustering import cluster

    B = a.batch
    lo, hi = shard_range(a.frames, rank, world)
    n_local = hi - lo
    face = Face(landmarks=Wt.make_shape_predictor(seed=4), embedding=Wt.make_embedder(seed=3),
                detector=Wt.make_detector(seed=2), upsample=1, device=dev, max_frames=B, max_faces=B * 6)
    pool = make_frames(a.pool, H, W, seed=100 + rank, device=dev)
    # per-frame face counts 2..6 (mean 4), seeded by the GLOBAL frame index: shards hold different numbers of faces
    rng = np.random.default_rng(7)
    counts_all = rng.integers(2, 7, size=a.frames)
    counts = counts_all[lo:hi]
    n_faces = int(counts.sum())
    emb_all = torch.zeros(n_faces, 128, dtype=torch.float32, device=dev)
    trk_all = torch.zeros(n_faces, dtype=torch.int64, device=dev)   # (uninitialised ids overflowed rank * stride in the warm-up gather at N = 8)
    # boxes of every batch, prepared before the timed region (in `extract` they come from the track file)
    batches = []
    off = 0
    for s in range(0, n_local, B):
        nb = min(B, n_local - s)
        cs = counts[s:s + nb]
        bx, _ = make_boxes(int(cs.sum()), 1, H, W, seed=1000 * rank + s, min_side=120, max_side=700)
        fi = torch.from_numpy(np.repeat(np.arange(nb), cs).astype(np.int32))
        # a face track = 8 consecutive faces of the shard (track ids are local; gather_embeddings makes them global)
        tr = torch.arange(off, off + int(cs.sum()), dtype=torch.int64) // 8
        batches.append((s, nb, bx.to(dev), fi.to(dev), tr.to(dev), off))
        off += int(cs.sum())

    def run_frames():
        for (s, nb, bx, fi, tr, o) in batches:
            idx = torch.arange(s, s + nb, device=dev) % a.pool
            fr = pool[idx]
            res = face.extract_batch(fr, bx, fi, detect=True)
            n = bx.shape[0]
            emb_all[o:o + n] = res["embeddings"]
            trk_all[o:o + n] = tr

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # warm-up: two batches + one small gather/cluster
    for (s, nb, bx, fi, tr, o) in batches[:2]:
        face.extract_batch(pool[:nb], bx, fi, detect=True)
    if world > 1:
        gather_embeddings(emb_all[:64], trk_all[:64])
    cluster(torch.randn(256, 128, device=dev), np.arange(256) // 4, threshold=a.threshold, device=dev)
    sync()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    run_frames()
    ev[1].record()
    if world > 1:
        E, T = gather_embeddings(emb_all, trk_all)
    else:
        E, T = emb_all, trk_all
    ev[2].record()
    tracks, labels, stats = cluster(E, T.cpu().numpy(), threshold=a.threshold, device=dev, return_stats=True)
    ev[3].record()
    sync()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])],
                     device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    frames_ms, gather_ms, cluster_ms, total_ms = [float(v) for v in t.tolist()]
    if rank == 0:
        line = dict(bench="C4 4K frame-sharded detect+embed -> all-gather -> clustering", n_gpus=world, frames=a.frames,
                    embeddings=int(E.shape[0]), tracks=int(len(tracks)), clusters=int(stats["n_clusters"]),
                    frames_per_s=a.frames / (total_ms * 1e-3), frame_phase_fps=a.frames / (frames_ms * 1e-3),
                    ms=dict(frames=round(frames_ms, 2), gather=round(gather_ms, 3), clustering=round(cluster_ms, 2),
                            total=round(total_ms, 2)),
                    serial_tail_frac=round((gather_ms + cluster_ms) / total_ms, 4), scaling="strong",
                    batch=B, faces_per_frame="2..6 (seeded, mean 4)", backend="nccl" if world > 1 else "none")
        print(json.dumps(line), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "c4_bench.jsonl"), "a") as f:
            f.write(json.dumps(line) + "\n")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
