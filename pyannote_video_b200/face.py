"""Face processing — drop-in for `pyannote.video.face.face.Face` (pyannote/video/face/face.py:38-132)
with every dlib call replaced by the sm_100a kernels of this package.

Same constructor and methods as the reference (`Face(landmarks=, embedding=)`, `.iterfaces`,
`.get_landmarks`, `.get_embedding`, `.get_debug`, `.__call__`), same object duck types
(`geometry.Rect`, `geometry.FullObjectDetection`, iterable 128-d embedding), plus a batched tensor
API (`detect_batch`, `landmarks_batch`, `embed_batch`, `process_batch`) that the per-frame methods
are built on.  Differences that follow BASELINE.json's north_star, all opt-out-able:
the detector is dlib's CNN (MMOD) detector rather than the HOG one (`face/face.py:54`), and model
files are this package's `.npz` containers (`weights.save_model`) rather than dlib `.dat` files.
There is no CPU fallback: without a CUDA device the constructor raises.
"""
import numpy as np
import torch

from . import weights as W
from .geometry import Rect, FullObjectDetection
from .nets import DetectorNet, EmbedNet
from .ops import ShapePredictor, ChipExtractor

DLIB_SMALLEST_FACE = 36


def _load(model, kind, default_factory=None):
    if model is None:
        return default_factory() if default_factory else None
    if isinstance(model, dict):
        if model.get("kind") != kind:
            raise RuntimeError("expected a '%s' model, got '%s'" % (kind, model.get("kind")))
        return model
    return W.load_model(str(model), kind)


class Face(object):
    """Face processing

    Parameters
    ----------
    landmarks : str or dict, optional
        Path to (or in-memory) 68 facial landmarks predictor model.
    embedding : str or dict, optional
        Path to (or in-memory) face embedding model.
    detector : str or dict, optional
        CNN (MMOD) detector model; defaults to the seeded synthetic detector.
    upsample : int
        Number of 2x upsamplings before detection; the reference calls `face_detector_(rgb, 1)`.
    """

    def __init__(self, landmarks=None, embedding=None, detector=None, upsample=1, device=None,
                 max_frames=8, max_faces=256):
        super(Face, self).__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("pyannote_video_b200.Face needs a CUDA device (B200); there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.upsample = int(upsample)
        self.max_frames = int(max_frames)
        self.max_faces = int(max_faces)
        self.size = 200  # the reference's get_debug reads an attribute it never sets (face/face.py:86)

        # face detection
        self._detector_model = _load(detector, "mmod_detector", W.make_detector)
        self._detectors = {}

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
        if key not in self._detectors:
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

    def detect_batch(self, frames):
        """frames uint8 [B,H,W,3] -> (boxes int32 [M,4] (l,t,r,b), frame_idx int32 [M], score f32 [M]) on device."""
        frames = self._to_device_frames(frames)
        B, H, Wd, _ = frames.shape
        det = self._detector_for(H, Wd)
        all_b, all_f, all_s = [], [], []
        for s in range(0, B, det.B):
            chunk = frames[s:s + det.B]
            boxes, scores, counts = det.detect(chunk)
            counts_h = counts.cpu()
            if int(counts_h.min()) < 0:
                raise RuntimeError("detector: more than %d candidate cells in a frame" % det.MAX_CAND)
            n = chunk.shape[0]
            mask = torch.arange(det.MAX_DET, device=self.device)[None, :] < counts[:, None].clamp(max=det.MAX_DET)
            fi = torch.arange(s, s + n, device=self.device, dtype=torch.int32)[:, None].expand(n, det.MAX_DET)
            all_b.append(boxes[mask])
            all_s.append(scores[mask])
            all_f.append(fi[mask])
        return torch.cat(all_b), torch.cat(all_f), torch.cat(all_s)

    def landmarks_batch(self, frames, boxes, frame_idx):
        """-> int32 [M,68,2] landmark positions (x,y) on device"""
        frames = self._to_device_frames(frames)
        boxes = torch.as_tensor(boxes, dtype=torch.int32, device=self.device).reshape(-1, 4).contiguous()
        frame_idx = torch.as_tensor(frame_idx, dtype=torch.int32, device=self.device).contiguous()
        return self.shape_predictor_.predict(frames, boxes, frame_idx)

    def embed_batch(self, frames, landmarks, frame_idx):
        """-> float32 [M,128] on device"""
        frames = self._to_device_frames(frames)
        landmarks = torch.as_tensor(landmarks, dtype=torch.int32, device=self.device).reshape(-1, 68, 2).contiguous()
        frame_idx = torch.as_tensor(frame_idx, dtype=torch.int32, device=self.device).contiguous()
        net = self.face_recognition_
        M = landmarks.shape[0]
        out = torch.empty(M, W.EMB_DIM, dtype=torch.float32, device=self.device)
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
