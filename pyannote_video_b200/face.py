"""Face processing — drop-in for `pyannote.video.face.face.Face` (pyannote/video/face/face.py:38-132)
with every dlib call replaced by the sm_100a kernels of this package.

Same constructor and methods as the reference (`Face(landmarks=, embedding=)`, `.iterfaces`,
`.get_landmarks`, `.get_embedding`, `.get_debug`, `.__call__`), same object duck types
(`geometry.Rect`, `geometry.FullObjectDetection`, iterable 128-d embedding), plus a batched tensor
API (`detect_batch`, `landmarks_batch`, `embed_batch`, `process_batch`) that the per-frame methods
are built on.  Differences that follow BASELINE.json's north_star, all opt-out-able:
the detector is dlib's CNN (MMOD) detector rather than the HOG one (`face/face.py:54`) and has to be named
(`detector=`); model files are this package's `.npz` containers (`weights.save_model`) or dlib `.dat` files
(`dlib_dat.load`, format restated from memory — see that module).
There is no CPU fallback: without a CUDA device the constructor raises.
"""
import os
import warnings

import numpy as np
import torch

from . import weights as W
from .geometry import Rect, FullObjectDetection
from .nets import DetectorNet, EmbedNet
from .ops import ShapePredictor, ChipExtractor

DLIB_SMALLEST_FACE = 36


def _load(model, kind):
    if model is None:
        return None
    if isinstance(model, dict):
        if model.get("kind") != kind:
            raise RuntimeError("expected a '%s' model, got '%s'" % (kind, model.get("kind")))
        return model
    path = str(model)
    if path.endswith(".dat"):
        from . import dlib_dat
        return dlib_dat.load(path, kind)
    return W.load_model(path, kind)


class Face(object):
    """Face processing

    Parameters
    ----------
    landmarks : str or dict, optional
        Path to (or in-memory) 68 facial landmarks predictor model (`.npz` container or dlib `.dat`).
    embedding : str or dict, optional
        Path to (or in-memory) face embedding model (`.npz` container or dlib `.dat`).
    detector : str or dict
        CNN (MMOD) detector model (`.npz`, dlib `mmod_human_face_detector.dat`, or an in-memory dict), or a HOG
        detector model (dict / `.npz` of kind "hog_detector": filters [D,31,10,10] + thresholds — the detector
        family the reference itself uses, `face/face.py:54`).  dlib compiles its HOG detector's trained weights
        into the library; nothing equivalent ships here, so a detector model must be named.  `"synthetic"` /
        `"synthetic-hog"` select the seeded random-weight CNN / HOG detectors used by the tests and the benchmark
        (they do not find faces).
    upsample : int
        Number of 2x upsamplings before detection (0 or 1); the reference calls `face_detector_(rgb, 1)`.
    """

    def __init__(self, landmarks=None, embedding=None, detector=None, upsample=1, device=None,
                 max_frames=8, max_faces=256):
        super(Face, self).__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("pyannote_video_b200.Face needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if int(upsample) not in (0, 1):
            raise ValueError("upsample must be 0 or 1 (got %r)" % (upsample, ))
        self.upsample = int(upsample)
        self.max_frames = int(max_frames)
        self.max_faces = int(max_faces)
        self.size = 200  # the reference's get_debug reads an attribute it never sets (face/face.py:86)

        # face detection
        # (resolved lazily: `extract` never detects, so Face(landmarks=, embedding=) stays constructible like the
        # reference's; the first detection without a model raises instead of silently using random weights)
        if detector is None:
            detector = os.environ.get("PYANNOTE_FACE_DETECTOR") or None
        if isinstance(detector, str) and detector == "synthetic":
            detector = W.make_detector()
        if isinstance(detector, str) and detector == "synthetic-hog":
            detector = W.make_hog_detector()
        if isinstance(detector, str) and detector.endswith(".npz"):
            detector = W.load_model(detector)               # either detector family
        if isinstance(detector, dict) and detector.get("kind") == "hog_detector":
            self._detector_model = detector
        else:
            self._detector_model = _load(detector, "mmod_detector")
        self._detectors = {}

        with torch.cuda.device(self.device):
            # landmark detection
            lm = _load(landmarks, "ert_shape_predictor")
            if lm is not None:
                self.shape_predictor_ = ShapePredictor(lm, self.device)

            # face embedding
            em = _load(embedding, "resnet_v1_embedder")
            if em is not None:
                self.face_recognition_ = EmbedNet(em, self.max_faces, self.device)
                self._chipper = ChipExtractor(self.device)

    # ------------------------------------------------------------------ batched tensor API
    def _detector_for(self, H, Wd):
        key = (H, Wd)
        if self._detector_model is None:
            raise RuntimeError(
                "Face: no detector model.  dlib compiles its HOG detector's weights into the library; nothing equivalent "
                "ships here.  Pass detector=<mmod_human_face_detector.dat | model.npz | dict> (or set "
                "PYANNOTE_FACE_DETECTOR), or detector='synthetic' for the seeded random-weight detector of the tests.")
        if key not in self._detectors:
            with torch.cuda.device(self.device):
                if self._detector_model.get("kind") == "hog_detector":
                    from .hog import HogDetectorNet
                    self._detectors[key] = HogDetectorNet(self._detector_model, H, Wd, self.upsample, self.max_frames, self.device)
                else:
                    self._detectors[key] = DetectorNet(self._detector_model, H, Wd, self.upsample, self.max_frames, self.device)
        return self._detectors[key]

    def _to_device_frames(self, frames):
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dim() == 3:
            frames = frames[None]
        if frames.dtype != torch.uint8 or frames.shape[-1] != 3:
            raise RuntimeError("images must be uint8 RGB arrays of shape [H,W,3]")
        return frames.to(self.device, non_blocking=True).contiguous()

    def detect_padded(self, frames):
        """frames uint8 [B,H,W,3] with B <= max_frames -> (boxes int32 [B,MAX_DET,4], scores f32 [B,MAX_DET],
        counts int32 [B]) on device, no host synchronisation.  counts[i] < 0 reports a candidate overflow."""
        frames = self._to_device_frames(frames)
        with torch.cuda.device(self.device):
            return self._detector_for(frames.shape[1], frames.shape[2]).detect(frames)

    def detect_batch(self, frames):
        """frames uint8 [B,H,W,3] -> (boxes int32 [M,4] (l,t,r,b), frame_idx int32 [M], score f32 [M]) on device."""
        frames = self._to_device_frames(frames)
        B, H, Wd, _ = frames.shape
        det = self._detector_for(H, Wd)
        all_b, all_f, all_s = [], [], []
        with torch.cuda.device(self.device):
            for s in range(0, B, det.B):
                chunk = frames[s:s + det.B]
                n = chunk.shape[0]
                boxes, scores, counts = det.detect(chunk)
                counts_h = counts.cpu()
                if int(counts_h.min()) < 0:
                    # more candidate cells than the decode kernel sorts: keep the MAX_CAND best of the frame instead of
                    # aborting the video (raise the threshold of this chunk to its MAX_CAND-th largest score)
                    warnings.warn("detector: more than %d candidate cells in a frame; keeping the best %d"
                                  % (det.MAX_CAND, det.MAX_CAND))
                    kth = det.scores[:n].reshape(n, -1).topk(det.MAX_CAND, dim=1).values[:, -1]
                    boxes, scores, counts = det.decode(n, threshold=float(kth.max()))
                    counts_h = counts.cpu()
                if int(counts_h.max()) >= det.MAX_DET:
                    warnings.warn("detector: a frame has %d or more detections; only the best %d are returned"
                                  % (det.MAX_DET, det.MAX_DET))
                cnt = counts.clamp(min=0, max=det.MAX_DET)
                mask = torch.arange(det.MAX_DET, device=self.device)[None, :] < cnt[:, None]
                fi = torch.arange(s, s + n, device=self.device, dtype=torch.int32)[:, None].expand(n, det.MAX_DET)
                all_b.append(boxes[mask])
                all_s.append(scores[mask])
                all_f.append(fi[mask])
        return torch.cat(all_b), torch.cat(all_f), torch.cat(all_s)

    def landmarks_batch(self, frames, boxes, frame_idx, out=None):
        """-> int32 [M,68,2] landmark positions (x,y) on device"""
        frames = self._to_device_frames(frames)
        boxes = torch.as_tensor(boxes, dtype=torch.int32, device=self.device).reshape(-1, 4).contiguous()
        frame_idx = torch.as_tensor(frame_idx, dtype=torch.int32, device=self.device).contiguous()
        with torch.cuda.device(self.device):
            return self.shape_predictor_.predict(frames, boxes, frame_idx, out=out)

    def embed_batch(self, frames, landmarks, frame_idx, out=None):
        """-> float32 [M,128] on device"""
        frames = self._to_device_frames(frames)
        landmarks = torch.as_tensor(landmarks, dtype=torch.int32, device=self.device).reshape(-1, 68, 2).contiguous()
        frame_idx = torch.as_tensor(frame_idx, dtype=torch.int32, device=self.device).contiguous()
        net = self.face_recognition_
        M = landmarks.shape[0]
        if out is None:
            out = torch.empty(M, W.EMB_DIM, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            for s in range(0, M, net.B):
                m = min(net.B, M - s)
                self._chipper.extract(frames, landmarks[s:s + m], frame_idx[s:s + m], net.chips)
                out[s:s + m] = net.forward_chips(m)
        return out

    def process_batch(self, frames, boxes=None, frame_idx=None):
        """detect (unless boxes are given) -> landmarks -> embed for a batch of frames, all on device."""
        frames = self._to_device_frames(frames)
        scores = None
        if boxes is None:
            boxes, frame_idx, scores = self.detect_batch(frames)
        parts = self.landmarks_batch(frames, boxes, frame_idx)
        emb = self.embed_batch(frames, parts, frame_idx)
        return boxes, frame_idx, scores, parts, emb

    # ---- pipelined host -> device staging (fixed buffers: no allocator traffic or implicit syncs in the loop) ----
    N_UPLOAD_SLOTS = 3

    def upload(self, frames, boxes=None, frame_idx=None, slot=0):
        """Asynchronous host -> device copy of one batch on a dedicated copy stream into staging slot `slot`
        (0..2): frames uint8 [B,H,W,3] (pinned host tensor or numpy), optional boxes int32 [M,4] / frame_idx int32 [M].
        Returns (frames_dev, boxes_dev, frame_idx_dev, ready_event); the consumer's stream must wait for
        `ready_event`, and gives the slot back with `release_upload(slot, done_event)` once its work is enqueued."""
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(np.ascontiguousarray(frames))
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
            raise RuntimeError("upload: frames must be uint8 [B,H,W,3]")
        if not hasattr(self, "_up"):
            self._up = dict(stream=torch.cuda.Stream(device=self.device), slots={})
        up = self._up
        key = (int(slot) % self.N_UPLOAD_SLOTS, tuple(frames.shape), None if boxes is None else int(boxes.shape[0]))
        st = up["slots"].get(key)
        if st is None:
            st = dict(fr=torch.empty(tuple(frames.shape), dtype=torch.uint8, device=self.device), done=None,
                      ready=torch.cuda.Event())
            if boxes is not None:
                st["bx"] = torch.empty(int(boxes.shape[0]), 4, dtype=torch.int32, device=self.device)
                st["fi"] = torch.empty(int(boxes.shape[0]), dtype=torch.int32, device=self.device)
            up["slots"][key] = st
        up["last"] = {key[0]: st, **{k: v for k, v in up.get("last", {}).items() if k != key[0]}}
        with torch.cuda.stream(up["stream"]):
            if st["done"] is not None:
                up["stream"].wait_event(st["done"])
            st["fr"].copy_(frames, non_blocking=True)
            if boxes is not None:
                st["bx"].copy_(torch.as_tensor(boxes, dtype=torch.int32).reshape(-1, 4), non_blocking=True)
                st["fi"].copy_(torch.as_tensor(frame_idx, dtype=torch.int32).reshape(-1), non_blocking=True)
            st["ready"].record(up["stream"])
        return st["fr"], st.get("bx"), st.get("fi"), st["ready"]

    def release_upload(self, slot, done_event):
        """`done_event`: recorded by the consumer after the last kernel that reads staging slot `slot`"""
        st = getattr(self, "_up", {}).get("last", {}).get(int(slot) % self.N_UPLOAD_SLOTS)
        if st is not None:
            st["done"] = done_event

    def extract_batch(self, frames, boxes, frame_idx, detect=True, out=None):
        """The `track` + `extract` work of one batch of frames in one call, without host synchronisation:
        detections of every frame (padded, as `detect_padded`) and landmarks + embeddings of the GIVEN face boxes
        (in `extract` they come from the track file, scripts/pyannote-face.py:290-297).  frames: uint8 [B,H,W,3]
        host (numpy / pinned tensor) or device tensor; boxes int32 [M,4]; frame_idx int32 [M].
        Returns a dict of device tensors (written into `out` when given): det_boxes, det_scores, det_counts,
        landmarks, embeddings."""
        frames = self._to_device_frames(frames)
        boxes = torch.as_tensor(boxes, dtype=torch.int32).to(self.device, non_blocking=True).reshape(-1, 4).contiguous()
        frame_idx = torch.as_tensor(frame_idx, dtype=torch.int32).to(self.device, non_blocking=True).contiguous()
        res = {} if out is None else out
        parts = self.landmarks_batch(frames, boxes, frame_idx, out=res.get("landmarks"))
        emb = self.embed_batch(frames, parts, frame_idx, out=res.get("embeddings"))
        res["landmarks"], res["embeddings"] = parts, emb
        if detect:
            det = self._detector_for(frames.shape[1], frames.shape[2])
            db, ds, dc = [], [], []
            with torch.cuda.device(self.device):
                for s in range(0, frames.shape[0], det.B):
                    b_, s_, c_ = det.detect(frames[s:s + det.B])
                    if frames.shape[0] > det.B:
                        b_, s_, c_ = b_.clone(), s_.clone(), c_.clone()
                    db.append(b_), ds.append(s_), dc.append(c_)
            res["det_boxes"] = db[0] if len(db) == 1 else torch.cat(db)
            res["det_scores"] = ds[0] if len(ds) == 1 else torch.cat(ds)
            res["det_counts"] = dc[0] if len(dc) == 1 else torch.cat(dc)
        return res

    # ------------------------------------------------------------------ reference API (per frame)
    def iterfaces(self, rgb):
        """Iterate over all detected faces"""
        boxes, _, _ = self.detect_batch(rgb)
        for l, t, r, b in boxes.cpu().tolist():
            yield Rect(l, t, r, b)

    def get_landmarks(self, rgb, face):
        box = [[face.left(), face.top(), face.right(), face.bottom()]]
        parts = self.landmarks_batch(rgb, box, [0])[0].cpu().tolist()
        return FullObjectDetection(face, parts)

    def get_embedding(self, rgb, landmarks):
        pts = [[p.x, p.y] for p in landmarks.parts()]
        if len(pts) != 68:
            raise RuntimeError("get_embedding expects a 68-point full_object_detection")
        emb = self.embed_batch(rgb, [pts], [0])[0]
        return [float(v) for v in emb.cpu().tolist()]

    def get_debug(self, image, face, landmarks):
        """Return face with overlaid landmarks"""
        import cv2
        if isinstance(image, torch.Tensor):
            image = image.cpu().numpy()
        copy = image.copy()
        for p in landmarks.parts():
            x, y = p.x, p.y
            cv2.rectangle(copy, (x, y), (x, y), (0, 255, 0), 2)
        copy = copy[max(face.top(), 0):face.bottom(), max(face.left(), 0):face.right()]
        copy = cv2.resize(copy, (self.size, self.size))
        return copy

    def __call__(self, rgb, return_landmarks=False, return_embedding=False, return_debug=False):
        """Iterate over all faces (same contract as pyannote/video/face/face.py:89-132); landmarks and
        embeddings of all faces of the frame are computed in one batched pass."""
        faces = list(self.iterfaces(rgb))
        if not (return_landmarks or return_embedding or return_debug):
            for face in faces:
                yield face
            return
        if not faces:
            return
        frames = self._to_device_frames(rgb)
        boxes = [[f.left(), f.top(), f.right(), f.bottom()] for f in faces]
        fidx = [0] * len(faces)
        parts_dev = self.landmarks_batch(frames, boxes, fidx)
        parts = parts_dev.cpu().tolist()
        embs = self.embed_batch(frames, parts_dev, fidx).cpu().tolist() if return_embedding else None
        for i, face in enumerate(faces):
            result = (face, )
            landmarks = FullObjectDetection(face, parts[i])
            if return_landmarks:
                result = result + (landmarks, )
            if return_embedding:
                result = result + ([float(v) for v in embs[i]], )
            if return_debug:
                result = result + (self.get_debug(rgb, face, landmarks), )
            yield result
