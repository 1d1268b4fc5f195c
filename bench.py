#!/usr/bin/env python
"""bench.py — frames/sec of the face hot path (CNN detect -> 68-pt landmarks -> chip -> ResNet embed)
on synthetic 1080p, BASELINE.json configs[1], on N GPUs of one node.

    python bench.py --gpus 1 --steps 60 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference's dlib path

A step = one batch of `--frames-per-step` 1080p frames through the whole path, every frame
detected with upsample 1 (reference semantics, pyannote/video/face/face.py:66) plus F=4 seeded face
boxes per frame through landmarks+embed (mirrors `extract`, scripts/pyannote-face.py:290-297).
Prints ONE JSON line (contract in the task statement).  Only the `cpu_baseline` leg and
`--impl reference` execute anything under oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
FACES_PER_FRAME = 4
METRIC = "frames/sec detect+track+embed on 1080p"
WORKLOAD = ("synthetic 1080p@25fps, batched CNN detect (upsample 1, every frame) + %d faces/frame 68-pt landmarks + "
            "ResNet-v1 embed" % 4)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs"), tf_sustained=d.get("bf16_tflops_sustained"), tf_burst=d.get("bf16_tflops"),
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_sustained=1400.0, tf_burst=1590.0, source="fallback")


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, nme in enumerate(names):
            if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(nme)
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=(max(mx) if mx else None), reasons=reasons,
                    samples=len(self.rows))


# ------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle) — used by cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_reference_frames(n_frames, models, frames, boxes, fidx):
    """runs the oracle path on `n_frames` frames; returns elapsed seconds"""
    import numpy as np
    import torch
    from oracle import nets as onets, pyramid as opyr, landmarks as olm
    det, sp, emb = models
    t0 = time.perf_counter()
    for i in range(n_frames):
        rgb = frames[i].numpy()
        plane, geo = opyr.build_plane(rgb, 1)
        x = torch.from_numpy(opyr.normalize_plane(plane))[None]
        scores = onets.detector_forward(det, x)[0].numpy()
        opyr.decode(scores, geo, det["window"], det["adjust_threshold"], det["iou_thresh"], det["covered_thresh"],
                    max_candidates=4096)
        sel = (fidx == i).numpy()
        parts = olm.ert_predict(sp, rgb, boxes[sel].numpy())
        chips = olm.extract_chips(rgb, parts)
        onets.embed_forward(emb, onets.normalize_rgb(chips))
    return time.perf_counter() - t0


def make_models():
    from pyannote_video_b200 import weights as Wt
    return Wt.make_detector(seed=2), Wt.make_shape_predictor(seed=4), Wt.make_embedder(seed=3)


def run_reference(args, rank, world):
    """--impl reference: the oracle (CPU restatement of the reference's dlib path), all host threads."""
    if rank != 0:
        return
    import torch
    from pyannote_video_b200.synth import make_frames, make_boxes
    models = make_models()
    n_per_step = 1
    total = args.warmup + args.steps
    total = min(total, 8)  # bounded sample: each step is one frame; cap the whole run to a few minutes
    steps = max(1, total - args.warmup) if total > args.warmup else 1
    warm = max(0, total - steps)
    frames = make_frames(2, H, W, seed=0)
    boxes, fidx = make_boxes(2, FACES_PER_FRAME, H, W, seed=1)
    for _ in range(warm):
        cpu_reference_frames(1, models, frames, boxes, fidx)
    el = 0.0
    for s in range(steps):
        el += cpu_reference_frames(1, models, frames[s % 2:], boxes[(s % 2) * FACES_PER_FRAME:], fidx[(s % 2) * FACES_PER_FRAME:] - (s % 2))
    fps = steps * n_per_step / el
    cores = torch.get_num_threads()
    line = dict(metric=METRIC, value=fps, unit="frames/s", n_gpus=args.gpus, steps=steps, warmup=warm,
                ms_per_step=1000.0 * el / steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, frames_per_step=n_per_step, faces_per_frame=FACES_PER_FRAME,
                            sample="each step = one 1080p frame of the same workload through the CPU restatement"),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=cores, kind="port",
                                  sample="%d x 1080p frame(s), oracle restatement (torch-CPU fp32 convs + numpy)" % steps),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-convs", type=int, default=2, help="instrumented steps for the roofline leg")
    ap.add_argument("--pipeline", action="store_true",
                    help="build the pyramid of batch s+1 on a second stream under the convs of batch s (measured: +1.7 %, the "
                         "pyramid's warps take issue slots from the MMA-issuing thread; off by default)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from pyannote_video_b200 import _lib
    from pyannote_video_b200.face import Face
    from pyannote_video_b200.synth import make_frames, make_boxes

    det_m, sp_m, emb_m = make_models()
    B = args.frames_per_step
    face = Face(landmarks=sp_m, embedding=emb_m, detector=det_m, upsample=1, device=dev, max_frames=B,
                max_faces=B * FACES_PER_FRAME)
    # distinct input batches, together larger than L2 (126 MB): 4 x B x 6.2 MB
    n_sets = 4
    host_sets, dev_sets, box_sets = [], [], []
    for s in range(n_sets):
        fr = make_frames(B, H, W, seed=1000 * rank + s, device=dev)
        dev_sets.append(fr)
        host_sets.append(fr.cpu().pin_memory())
        bx, fi = make_boxes(B, FACES_PER_FRAME, H, W, seed=1 + s + 10 * rank)
        box_sets.append((bx.to(dev), fi.to(dev), bx.pin_memory(), fi.pin_memory()))

    # landmarks + chips + embed of the seeded boxes do not depend on the detector (in `extract` the boxes
    # come from the track file), so they run on a side stream and fill the SM slots the bandwidth-bound
    # pyramid kernels leave free.
    side = torch.cuda.Stream(device=dev)

    overlap = [True]

    def embed_branch(fr, bx, fi):
        main = torch.cuda.current_stream()
        if not overlap[0]:      # instrumented (roofline) steps: everything on one stream, no co-running kernels
            parts = face.shape_predictor_.predict(fr, bx, fi)
            net = face.face_recognition_
            face._chipper.extract(fr, parts, fi, net.chips)
            return parts, net.forward_chips(bx.shape[0])
        side.wait_stream(main)
        with torch.cuda.stream(side):
            parts = face.shape_predictor_.predict(fr, bx, fi)
            net = face.face_recognition_
            face._chipper.extract(fr, parts, fi, net.chips)
            emb = net.forward_chips(bx.shape[0])
        return parts, emb

    # Cross-step software pipeline: the pyramid of batch s+1 (CUDA-core, issue-bound) is built on a low-priority stream
    # into the second plane while the tensor-core convs of batch s run; the timed region still contains exactly K
    # pyramid builds and K forward passes (the first build is exposed, the last step prefetches nothing).
    det0 = face._detector_for(H, W)
    pipelined = args.pipeline
    if pipelined:
        det0.enable_double_buffer()
    lo_pri, hi_pri = torch.cuda.Stream.priority_range()      # (least, greatest) = (0, -5): lower number = higher priority
    pyr_stream = torch.cuda.Stream(device=dev, priority=lo_pri)
    if pipelined:
        # the conv / decode chain runs on a high-priority stream so that its persistent CTAs are placed first and the
        # pyramid's short CTAs fill what is left of each SM
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=hi_pri))
    built = [None, None]            # event: plane k holds the pyramid of its batch
    conv1_done = [None, None]       # event: the first conv has finished reading plane k
    prebuilt = set()

    def submit_build(s, fr, ready=None):
        k = s % 2
        with torch.cuda.stream(pyr_stream):
            if ready is not None:
                pyr_stream.wait_event(ready)
            if conv1_done[k] is not None:
                pyr_stream.wait_event(conv1_done[k])
            det0.use_plane(k)
            det0.build_plane(fr, fr.shape[0])
            ev = torch.cuda.Event()
            ev.record(pyr_stream)
            built[k] = ev
        prebuilt.add(s)

    def detect_pipelined(s, fr, nxt=None):
        """forward + decode of batch s on the current stream; `nxt` = (frames, ready event) of batch s+1 or None"""
        main = torch.cuda.current_stream()
        k = s % 2
        if s not in prebuilt:
            pyr_stream.wait_stream(main)
            submit_build(s, fr)
        prebuilt.discard(s)
        if nxt is not None:
            submit_build(s + 1, nxt[0], nxt[1])
        main.wait_event(built[k])
        det0.use_plane(k)
        det0.forward_scores(fr.shape[0])
        ev = torch.cuda.Event()
        ev.record(main)
        conv1_done[k] = ev
        return det0.decode(fr.shape[0])

    def step_resident(s, last=True):
        fr = dev_sets[s % n_sets]
        bx, fi, _, _ = box_sets[s % n_sets]
        parts, emb = embed_branch(fr, bx, fi)
        if pipelined:
            detect_pipelined(s, fr, None if last else (dev_sets[(s + 1) % n_sets], None))
        else:
            det0.detect(fr)
        torch.cuda.current_stream().wait_stream(side)
        return emb

    # end-to-end step: pinned host frames -> device (side stream, one step ahead) -> the same path ->
    # detections / landmarks / embeddings copied back to pinned host memory and read one step later.
    copy_stream = torch.cuda.Stream(device=dev)
    det_e2e = face._detector_for(H, W)
    out_host = [dict(boxes=torch.empty(B, det_e2e.MAX_DET, 4, dtype=torch.int32).pin_memory(),
                     counts=torch.empty(B, dtype=torch.int32).pin_memory(),
                     parts=torch.empty(B * FACES_PER_FRAME, 68, 2, dtype=torch.int32).pin_memory(),
                     emb=torch.empty(B * FACES_PER_FRAME, 128, dtype=torch.float32).pin_memory(),
                     ev=torch.cuda.Event()) for _ in range(2)]

    # three device staging sets allocated once (no allocator traffic, no implicit syncs inside the pipeline):
    # the copy stream refills set k only after the step that read it has been fully enqueued and finished.
    n_stage = 3
    stage = [dict(fr=torch.empty(B, H, W, 3, dtype=torch.uint8, device=dev),
                  bx=torch.empty(B * FACES_PER_FRAME, 4, dtype=torch.int32, device=dev),
                  fi=torch.empty(B * FACES_PER_FRAME, dtype=torch.int32, device=dev),
                  ready=torch.cuda.Event(), done=None) for _ in range(n_stage)]

    def upload(s):
        st = stage[s % n_stage]
        fr_h = host_sets[s % n_sets]
        _, _, bx_h, fi_h = box_sets[s % n_sets]
        with torch.cuda.stream(copy_stream):
            if st["done"] is not None:
                copy_stream.wait_event(st["done"])
            st["fr"].copy_(fr_h, non_blocking=True)
            st["bx"].copy_(bx_h, non_blocking=True)
            st["fi"].copy_(fi_h, non_blocking=True)
            st["ready"].record(copy_stream)
        return st

    def step_e2e(s, n_total):
        st = stage[s % n_stage]
        main = torch.cuda.current_stream()
        main.wait_event(st["ready"])
        ahead = 2 if pipelined else 1             # inputs travel ahead of the pyramid that runs ahead of the convs
        if s + ahead < n_total:
            upload(s + ahead)
        fr, bx, fi = st["fr"], st["bx"], st["fi"]
        parts, emb = embed_branch(fr, bx, fi)
        if pipelined:
            nxt = None
            if s + 1 < n_total:
                sn = stage[(s + 1) % n_stage]
                nxt = (sn["fr"], sn["ready"])
            boxes, scores, counts = detect_pipelined(s, fr, nxt)
        else:
            boxes, scores, counts = det_e2e.detect(fr)
        main.wait_stream(side)
        parts.record_stream(main)
        o = out_host[s % 2]
        o["boxes"].copy_(boxes, non_blocking=True)
        o["counts"].copy_(counts, non_blocking=True)
        o["parts"].copy_(parts, non_blocking=True)
        o["emb"].copy_(emb, non_blocking=True)
        o["ev"].record()
        st["done"] = o["ev"]

    def run_e2e(n_total):
        for k in range(min(2 if pipelined else 1, n_total)):
            upload(k)
        for s in range(n_total):
            step_e2e(s, n_total)
            if s > 0:
                read_result(s - 1)                # results are consumed on the host one step behind
        read_result(n_total - 1)

    def read_result(s):
        o = out_host[s % 2]
        o["ev"].synchronize()
        return int(o["counts"].sum()) + float(o["emb"][0, 0])   # touch the data on the host

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---------------- device-resident timing (value) ----------------
    for s in range(args.warmup):
        step_resident(s, last=(s == args.warmup - 1))
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    embs = None
    t_host = time.perf_counter()
    host_ms = None
    n_host = min(4, args.steps)
    for s in range(args.steps):
        embs = step_resident(s, last=(s == args.steps - 1))
        if s + 1 == n_host:
            # time the host needs to enqueue one step, taken over the first steps only: later the launch queue is full and
            # the host simply waits for the GPU
            host_ms = 1000.0 * (time.perf_counter() - t_host) / n_host
    if world > 1:
        # the path's one exchange: all-gather of the per-rank embeddings before clustering
        gathered = [torch.empty_like(embs) for _ in range(world)]
        dist.all_gather(gathered, embs.contiguous())
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - l0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    face._detector_for(H, W).check()
    face.face_recognition_.check()

    # ---------------- end-to-end timing (host buffers, H2D + D2H inside) ----------------
    run_e2e(min(args.warmup, 2))
    sync_all()
    e2e_steps = max(4, args.steps // 2)
    torch.cuda.synchronize(dev)
    e0.record()
    run_e2e(e2e_steps)
    e1.record()
    sync_all()
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = float(t.item())
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=3)

    # ---------------- roofline leg: CUDA events around every tensor-core conv launch ----------------
    # Dominant kernel family = detconv_kernel (detector convs 2..7, csrc/detconv.cu): achieved = algorithmic
    # FLOPs of those six layers / their summed launch time.  conv1_fused and the embedder's srgemm launches
    # are reported beside it ("layers", "all_convs").  `traffic` = dram bytes read+written by the six detconv
    # launches from the committed `ncu --set full` capture (profiles/ncu_traffic.json, bytes per frame x B).
    roof = None
    if rank == 0 and args.profile_convs > 0:
        det = face._detector_for(H, W)
        net = face.face_recognition_
        det_names = ["conv%d" % (i + 1) for i in range(len(det.convs))]
        conv_ops = [(n, op) for n, (op, _) in zip(det_names, det.convs)] + [("embed", a[0]) for k, a in net.ops if k == "conv"]
        evs = []
        orig = {}
        for name, op in conv_ops:
            orig[id(op)] = op.run

            def timed_run(q_rows=None, _name=name, _run=op.run):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _run(q_rows)
                b.record()
                evs.append((_name, a, b))
            op.run = timed_run
        # coarse stage timing of the same instrumented steps (CUDA events on the launch stream)
        stage_ev = {}

        def timed_stage(name, fn):
            def wrapped(*a, **k):
                e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e_a.record()
                r = fn(*a, **k)
                e_b.record()
                stage_ev.setdefault(name, []).append((e_a, e_b))
                return r
            return wrapped

        patched = [(det, "build_plane", "pyramid"), (det, "forward_scores", "convs+shift_sum"), (det, "decode", "decode+nms"),
                   (face.shape_predictor_, "predict", "landmarks"), (face._chipper, "extract", "chips"),
                   (net, "forward_chips", "embed")]
        saved = [(o, n, getattr(o, n)) for o, n, _ in patched]
        for o, n, label in patched:
            setattr(o, n, timed_stage(label, getattr(o, n)))
        overlap[0] = False
        for s in range(args.profile_convs):
            step_resident(s)
        torch.cuda.synchronize(dev)
        overlap[0] = True
        for o, n, f in saved:
            setattr(o, n, f)
        for name, op in conv_ops:
            op.run = orig[id(op)]
        P = args.profile_convs
        stage_ms = {k: sum(a.elapsed_time(b) for a, b in v) / max(1, P) for k, v in stage_ev.items()}
        layer_ms = {}
        for name, a, b in evs:
            layer_ms[name] = layer_ms.get(name, 0.0) + a.elapsed_time(b) / P
        layer_flops = {n: B * f for n, f in zip(det_names, det.algorithmic_flops_per_layer)}   # per pyramid pixel, padding excluded
        layer_flops["embed"] = B * FACES_PER_FRAME * net.flops_per_face
        peaks = load_peaks()
        peak = peaks["tf_sustained"]
        layers = {n: dict(ms=round(layer_ms[n], 4), tflops=round(layer_flops[n] / (layer_ms[n] * 1e-3) / 1e12, 1),
                          frac=round(layer_flops[n] / (layer_ms[n] * 1e-3) / 1e12 / peak, 4)) for n in layer_ms}
        dom = [n for n in det_names[1:]] if det.conv_impl in ("detconv", "rsconv") else det_names
        dom_ms = sum(layer_ms[n] for n in dom)
        dom_fl = sum(layer_flops[n] for n in dom)
        all_ms = sum(layer_ms.values())
        all_fl = sum(layer_flops.values())
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if det.conv_impl in ("detconv", "rsconv") and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("plane") == [det.geo.plane_h, det.geo.plane_w] and tj.get("impl", "detconv") == det.conv_impl:
                traffic = tj["detconv_dram_bytes_per_frame"] * B
        achieved = dom_fl / (dom_ms * 1e-3) / 1e12
        roof = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                    peak_source=peaks["source"],
                    kernel=("%s_kernel (detector convs 2..7, %d launches/step)" % (det.conv_impl, len(dom)))
                    if det.conv_impl in ("detconv", "rsconv") else "srgemm_kernel (detector convs)",
                    kernel_ms_per_step=round(dom_ms, 4), algorithmic_flops_per_step=dom_fl,
                    all_convs=dict(ms_per_step=round(all_ms, 4), tflops=round(all_fl / (all_ms * 1e-3) / 1e12, 1),
                                   frac=round(all_fl / (all_ms * 1e-3) / 1e12 / peak, 4), launches_per_step=len(evs) // P),
                    layers=layers, stage_ms_per_step={k: round(v, 3) for k, v in stage_ms.items()})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    frames_total = args.steps * B * world
    value = frames_total / (ms_max * 1e-3)
    e2e_value = e2e_steps * B * world / (ms_e2e * 1e-3)
    h2d = B * H * W * 3 + B * FACES_PER_FRAME * (16 + 4)
    det = face._detector_for(H, W)
    d2h = B * det.MAX_DET * 16 + B * 4 + B * FACES_PER_FRAME * (68 * 2 * 4 + 128 * 4)

    cpu = None
    if not args.no_cpu_baseline:
        n_cpu = 2
        fr_cpu = host_sets[0][:n_cpu].clone()
        bx, fi = make_boxes(n_cpu, FACES_PER_FRAME, H, W, seed=1)
        el = cpu_reference_frames(n_cpu, (det_m, sp_m, emb_m), fr_cpu, bx, fi)
        cpu = dict(value=n_cpu / el, unit="frames/s", cores=torch.get_num_threads(), kind="port",
                   sample="%d x 1080p frames through the oracle restatement (torch-CPU fp32 convs + numpy), %.1f s" % (n_cpu, el))

    line = dict(metric=METRIC, value=value, unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic",
                config=dict(workload=WORKLOAD, frames_per_step=B, faces_per_frame=FACES_PER_FRAME, parallelism="frame-shard x%d" % world, pipeline=("pyramid of batch s+1 under the convs of batch s" if pipelined else "none"),
                            l2="%d distinct input batches (%d MB) + ~1.4 GB/frame of plane and activation traffic per step: "
                               "inputs larger than L2" % (n_sets, n_sets * B * H * W * 3 // 1000000)),
                clocks=sampler.summary(),
                e2e=dict(value=e2e_value, unit="frames/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps),
                gpu_launches=int(launches), host_enqueue_ms_per_step=round(host_ms, 3),
                roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
