"""smoke test used by __graft_entry__.smoke(): one small detect -> landmarks -> chip -> embed pass on
cuda:0, checked against the CPU oracle (the only place outside tests/ and bench.py where the oracle
is imported — as the checker, never as the product path)."""
import numpy as np
import torch


def smoke_test(verbose=True):
    if not torch.cuda.is_available():
        raise RuntimeError("smoke_test needs a CUDA device (there is no CPU fallback)")
    from . import weights as W
    from .nets import DetectorNet, EmbedNet
    from .ops import ShapePredictor, ChipExtractor
    from .synth import make_frames, make_boxes
    from oracle import nets as onets, pyramid as opyr, landmarks as olm

    dev = torch.device("cuda:0")
    H, Wd, F = 96, 128, 2
    frames = make_frames(F, H, Wd, seed=5)
    fd = frames.to(dev)
    det_model = W.make_detector(seed=2, score_bias=0.0)
    det = DetectorNet(det_model, H, Wd, 1, max_batch=F, device=dev)
    det.build_plane(fd, F)
    scores = det.forward_scores(F).cpu()
    det.check()
    plane = det.plane[:F].cpu().numpy()
    for i in range(F):
        ref_plane, geo = opyr.build_plane(frames[i].numpy(), 1)
        assert np.array_equal(plane[i], ref_plane), "pyramid plane mismatch"
        ref = onets.detector_forward(det_model, torch.from_numpy(opyr.normalize_plane(ref_plane))[None], bf16=True)[0]
        assert float((scores[i] - ref).abs().max()) < 0.03 * max(1.0, float(ref.abs().max())), "detector scores mismatch"
    sp_model = W.make_shape_predictor(seed=4, stages=4, trees=40)
    boxes, fidx = make_boxes(F, 2, H, Wd, seed=1, min_side=30, max_side=70)
    sp = ShapePredictor(sp_model, dev)
    parts = sp.predict(fd, boxes.to(dev), fidx.to(dev))
    emb_model = W.make_embedder(seed=3)
    net = EmbedNet(emb_model, max_batch=4, device=dev)
    M = boxes.shape[0]
    ChipExtractor(dev).extract(fd, parts, fidx.to(dev), net.chips)
    net.chips[:M, :, :, 3] = 255
    emb = net.forward_chips(M).cpu()
    net.check()
    parts_c = parts.cpu().numpy()
    chips_c = net.chips[:M].cpu().numpy()
    for f in range(F):
        sel = (fidx == f).numpy()
        ref_parts = olm.ert_predict(sp_model, frames[f].numpy(), boxes[sel].numpy())
        assert np.array_equal(parts_c[sel], ref_parts), "landmarks mismatch"
        ref_chips = olm.extract_chips(frames[f].numpy(), ref_parts)
        assert np.array_equal(chips_c[sel][..., :3], ref_chips), "chip mismatch"
    ref_emb = onets.embed_forward(emb_model, onets.normalize_rgb(chips_c[..., :3]), bf16=True)
    rel = float((emb - ref_emb).norm() / ref_emb.norm())
    assert rel < 2e-2, "embedding mismatch: rel L2 %.3e" % rel
    if verbose:
        print("smoke ok: plane/landmarks/chips bit-exact, scores within 3%%, embedding rel L2 %.2e" % rel)
    return rel
