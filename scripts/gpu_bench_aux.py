"""Throughput of the two non-CNN configs of BASELINE.json on one B200 (numbers for profiles/README.md):
  C3  correlation_tracker.update: 256 concurrent tracks x F frames at 1080p
  C5  agglomerative clustering of 100k x 128-d embeddings
Usage: python scripts/gpu_bench_aux.py [--frames 200]   -> JSON lines in gpurun_out/aux_bench.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def bench_tracker(n_frames):
    from pyannote_video_b200.geometry import DRect
    from pyannote_video_b200.synth import make_frames
    from pyannote_video_b200.tracker import TrackerBank
    from pyannote_video_b200 import _lib
    import ctypes as C
    dev = torch.device("cuda:0")
    frames = make_frames(min(n_frames, 64), 1080, 1920, seed=11, device=dev, shift_per_frame=(1.0, 0.5))
    bank = TrackerBank(capacity=256, device=dev)
    rects = [(100.0 + 110 * (k % 16), 60.0 + 62 * (k // 16), 196.0 + 110 * (k % 16), 156.0 + 62 * (k // 16)) for k in range(256)]
    handles = [bank.start(frames[0], DRect(*r)) for r in rects]
    bank._flush()
    ids = torch.tensor(handles, dtype=torch.int32, device=dev)
    L = _lib.lib()

    def upd(i):
        f = frames[i % frames.shape[0]]
        _lib.check(L.pv_tracker_update(bank.h, _lib.ptr(f), 1080, 1920, _lib.ptr(ids), 256, _lib.stream_ptr()))

    for i in range(1, 6):
        upd(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_frames):
        upd(i + 6)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    updates = 256 * n_frames
    state_bytes = 3 * 31 * 4096 * 8 + 2 * 4096 * 4      # A read twice + written once, B read + written
    return dict(bench="C3 tracker.update", tracks=256, frames=n_frames, ms=ms, updates_per_s=updates / ms * 1e3,
                hbm_gbs=updates * state_bytes / ms / 1e6, bytes_per_update=state_bytes)


def bench_cluster(n):
    """C5: 100k x 128-d embeddings around 2000 centroids; input (a) singleton tracks, input (b) 10 embeddings per track
    (SURVEY.md §8d).  Also times the distance matrix alone (tcgen05 Gram vs the fp32 CUDA-core kernel)."""
    from pyannote_video_b200.clustering import cluster, pairwise_distances
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    n_cent = 2000
    cent = torch.randn(n_cent, 128, generator=g)
    cent = cent / cent.norm(dim=1, keepdim=True) * 0.9
    X = (cent[:, None, :] + 0.02 * torch.randn(n_cent, n // n_cent, 128, generator=g)).reshape(-1, 128)
    perm = torch.randperm(X.shape[0], generator=g)
    X = X[perm].contiguous().to(dev)
    N = int(X.shape[0])
    out = dict(bench="C5 clustering", n=N)
    for impl in ("tcgen05", "fp32"):
        D = pairwise_distances(X, "euclidean", impl=impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        del D
        e0.record()
        D = pairwise_distances(X, "euclidean", impl=impl)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out["pdist_%s_ms" % impl] = round(ms, 2)
        if impl == "tcgen05":
            out["gram_tflops_6term"] = round(6 * 2.0 * N * N * 128 / ms / 1e9, 1)
            out["gram_write_gbs"] = round(4.0 * N * N / ms / 1e6, 1)
        del D
    torch.cuda.empty_cache()
    # tracks of 10: ten embeddings of the SAME centroid per track (row r of the shuffled matrix was row perm[r] before)
    orig = perm.numpy()
    for name, ids in (("singletons", np.arange(N)), ("tracks_of_10", orig // 10)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tracks, labels, stats = cluster(X, ids, threshold=0.6, device=dev, return_stats=True)
        torch.cuda.synchronize()
        out[name] = dict(seconds=round(time.perf_counter() - t0, 4), rounds=stats["rounds"], clusters=stats["n_clusters"],
                         tracks=int(len(tracks)))
        torch.cuda.empty_cache()
    return out


def bench_embed(batch, impl=None):
    """embed-only: chips -> 29 convs -> 128-d at a large face batch, CUDA events around every conv launch"""
    from pyannote_video_b200 import weights as W
    from pyannote_video_b200.nets import EmbedNet
    dev = torch.device("cuda:0")
    net = EmbedNet(W.make_embedder(seed=3), batch, dev, impl=impl)
    net.chips.copy_(torch.randint(0, 256, net.chips.shape, dtype=torch.uint8, device=dev))
    for _ in range(3):
        net.forward_chips(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 10
    for _ in range(reps):
        net.forward_chips(batch)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # per-layer
    evs = []
    conv_fl = net.conv_ops()
    convs = [op for op, _ in conv_fl]
    orig = [op.run for op in convs]
    for i, op in enumerate(convs):
        def timed(q=None, _r=op.run, _i=i):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); _r(q); b.record()
            evs.append((_i, a, b))
        op.run = timed
    net.forward_chips(batch); torch.cuda.synchronize(); evs.clear()
    net.forward_chips(batch); torch.cuda.synchronize()
    for op, r in zip(convs, orig):
        op.run = r
    layers = []
    for (i, a, b), (op, flf) in zip(evs, conv_fl):
        fl = float(batch) * flf
        t = a.elapsed_time(b)
        layers.append(dict(i=i, kernel=type(op).__name__, us=round(t * 1e3, 1), tflops=round(fl / t / 1e9, 1)))
    conv_ms = sum(a.elapsed_time(b) for _, a, b in evs)
    fl = batch * net.flops_per_face
    return dict(bench="embed-only", impl=net.impl, batch=batch, ms=ms, faces_per_s=batch / ms * 1e3, tflops=fl / ms / 1e9,
                frac_of_1442=fl / ms / 1e9 / 1442.3, conv_ms=conv_ms, conv_tflops=fl / conv_ms / 1e9, layers=layers)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=200)
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--only", default="", help="tracker | cluster | embed")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "aux_bench.jsonl"), "a") as f:
        for fn, arg in ((bench_tracker, a.frames), (bench_cluster, a.n), (bench_embed, a.batch)):
            if a.only and a.only not in fn.__name__:
                continue
            try:
                r = fn(arg)
            except Exception as e:  # noqa: BLE001
                r = dict(bench=fn.__name__, error=repr(e)[:500])
            f.write(json.dumps(r) + "\n")
            print(json.dumps(r), flush=True)
