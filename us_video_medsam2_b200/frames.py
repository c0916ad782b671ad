"""Frame ingest for `init_state(video_path)`: a folder of "<frame_index>.jpg" files -> normalised fp32 frames
(reference semantics: sam2/utils/misc.py:92-277 -- PIL decode, RGB, resize to SxS, /255, ImageNet mean/std).

Split of the work:
  host    JPEG decode + resize (PIL, as in the reference -- the resampling filter is part of the result) into ONE pinned
          uint8 buffer [T,S,S,3];
  device  the frames travel as uint8 (a quarter of the reference's fp32 host-to-device bytes) and are converted by
          `usvm_normalize_rgb_u8` (interleaved uint8 -> normalised fp32 planes, the reference's arithmetic).

`async_loading_frames=True` mirrors `AsyncVideoFrameLoader` (misc.py:104-169): a daemon thread decodes ahead while the
session starts; `images[t]` blocks only until frame t is decoded.  Device work stays on the caller's thread and stream:
when a frame is asked for, every frame decoded so far is uploaded and normalised in one copy + one launch.
`offload_video_to_cpu=True` keeps only the uint8 frames (in pinned host memory) and normalises a frame on the device
each time it is read.  MP4 files (misc.py:280-309, `decord` in the reference) are decoded frame by frame with OpenCV's
FFmpeg backend (bilinear resize to SxS on the host) and then take the same uint8 -> device -> normalise route; the
reference resamples inside the decoder, so pixel values can differ by the resampling filter."""
import os
import threading

import numpy as np
import torch

from .synth import IMG_MEAN, IMG_STD

_EXTS = (".jpg", ".jpeg", ".JPG", ".JPEG")


def _decode_into(path, image_size, dst):
    """PIL decode + RGB + resize (misc.py:92-101) into dst uint8 [S,S,3] (a numpy view of the pinned buffer)."""
    from PIL import Image

    img = Image.open(path)
    arr = np.asarray(img.convert("RGB").resize((image_size, image_size)))
    if arr.dtype != np.uint8:
        raise RuntimeError(f"Unknown image dtype: {arr.dtype} on {path}")
    dst[...] = arr
    w, h = img.size
    return h, w


class _Mp4Source:
    """Sequential decoder of a video file (load_video_frames_from_video_file, misc.py:280-309): frame t -> uint8 [S,S,3]
    RGB.  OpenCV (FFmpeg backend) stands in for decord; frames must be asked for in order."""

    def __init__(self, path, image_size):
        try:
            import cv2
        except ImportError as e:  # pragma: no cover
            raise NotImplementedError("MP4 input needs OpenCV (cv2) with FFmpeg; extract JPEG frames into a folder instead "
                                      "(ffmpeg -i <video>.mp4 -q:v 2 -start_number 0 <dir>/'%05d.jpg')") from e
        self.cv2, self.path, self.S = cv2, path, int(image_size)
        cap = cv2.VideoCapture(path)
        if not cap.isOpened():
            raise RuntimeError(f"cannot open video file {path}")
        n = 0
        while cap.grab():  # the container's frame count is a hint, not a promise: count what decodes
            n += 1
        cap.release()
        if n == 0:
            raise RuntimeError(f"no frames found in {path}")
        self.n, self.next, self.cap = n, 0, cv2.VideoCapture(path)

    def __len__(self):
        return self.n

    def decode_into(self, t, dst):
        if t != self.next:
            raise RuntimeError("video frames are decoded in order")
        ok, frame = self.cap.read()
        if not ok:
            raise RuntimeError(f"decoding frame {t} of {self.path} failed")
        self.next += 1
        h, w = frame.shape[:2]
        rgb = self.cv2.cvtColor(frame, self.cv2.COLOR_BGR2RGB)
        dst[...] = self.cv2.resize(rgb, (self.S, self.S), interpolation=self.cv2.INTER_LINEAR)
        if self.next == self.n:
            self.cap.release()
        return h, w


class VideoFrames:
    """List-like `inference_state["images"]`: `len()`, `images[t]` -> normalised fp32 [3,S,S] on the compute device."""

    def __init__(self, paths, image_size, compute_device, keep_on_device=True, background=False, img_mean=IMG_MEAN,
                 img_std=IMG_STD):
        from . import ops

        self._ops = ops
        # `paths`: JPEG file names, or a source object with __len__ / decode_into(t, dst) (video files)
        self.source = paths if hasattr(paths, "decode_into") else None
        self.paths = [] if self.source is not None else list(paths)
        self.T = len(self.source) if self.source is not None else len(self.paths)
        self.S, self.device = int(image_size), torch.device(compute_device)
        if self.device.type != "cuda":
            raise RuntimeError("frame ingest normalises on the GPU (usvm_normalize_rgb_u8): compute_device must be a "
                               "CUDA device; there is no CPU path")
        self.mean, self.std = tuple(float(x) for x in img_mean), tuple(float(x) for x in img_std)
        T = self.T
        self.host = torch.empty((T, self.S, self.S, 3), dtype=torch.uint8).pin_memory()
        self._np = self.host.numpy()
        self.keep = bool(keep_on_device)
        self.frames = (torch.empty((T, 3, self.S, self.S), dtype=torch.float32, device=self.device) if self.keep else None)
        self.decoded = 0          # frames [0, decoded) are in the pinned buffer
        self.uploaded = 0         # frames [0, uploaded) are normalised on the device (keep_on_device only)
        self.exception = None
        self.video_height = self.video_width = None
        self._cv = threading.Condition()
        self._decode(0)           # fills video_height / video_width; also the frame the user most likely prompts
        if background and T > 1:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        else:
            self.thread = None
            for t in range(1, T):
                self._decode(t)

    def _decode(self, t):
        if self.source is not None:
            h, w = self.source.decode_into(t, self._np[t])
        else:
            h, w = _decode_into(self.paths[t], self.S, self._np[t])
        with self._cv:
            self.video_height, self.video_width = h, w
            self.decoded = t + 1
            self._cv.notify_all()

    def _run(self):
        try:
            for t in range(1, self.T):
                self._decode(t)
        except Exception as e:  # surfaced by the next __getitem__ (misc.py:140-150)
            with self._cv:
                self.exception = e
                self._cv.notify_all()

    def __len__(self):
        return self.T

    def _wait_decoded(self, t):
        with self._cv:
            while self.decoded <= t and self.exception is None:
                self._cv.wait()
            if self.exception is not None:
                raise RuntimeError("Failure in frame loading thread") from self.exception
            return self.decoded

    def __getitem__(self, t):
        if isinstance(t, slice):
            return torch.stack([self[i] for i in range(*t.indices(len(self)))])
        t = int(t)
        if t < 0:
            t += len(self)
        if not (0 <= t < len(self)):
            raise IndexError(t)
        if self.keep and t < self.uploaded:
            return self.frames[t]
        ready = self._wait_decoded(t)
        if not self.keep:  # offload_video_to_cpu: only the uint8 frame lives on; normalise it for this read
            rgb = self.host[t:t + 1].to(self.device, non_blocking=True)
            return self._ops.normalize_rgb_u8(rgb, self.mean, self.std)[0]
        lo = self.uploaded  # everything decoded so far, in one copy + one launch
        rgb = self.host[lo:ready].to(self.device, non_blocking=True)
        self._ops.normalize_rgb_u8(rgb, self.mean, self.std, out=self.frames[lo:ready])
        self.uploaded = ready
        return self.frames[t]

    def resident(self):
        """The whole clip as ONE normalised device tensor [T,3,S,S] (waits for the decoder); None when offloaded."""
        if not self.keep:
            return None
        self[len(self) - 1]
        return self.frames


def list_jpeg_frames(video_path):
    """Sorted frame paths of a JPEG folder, with the reference's input checks (misc.py:186-243)."""
    is_str = isinstance(video_path, str)
    if not (is_str and os.path.isdir(video_path)):
        raise NotImplementedError("Only MP4 video and JPEG folder are supported at this moment")
    names = sorted(p for p in os.listdir(video_path) if os.path.splitext(p)[-1] in _EXTS)
    if not names:
        raise RuntimeError(f"no images found in {video_path}")
    return [os.path.join(video_path, n) for n in names]


def decode_jpeg_folder(video_path, image_size):
    """Host half of the ingest on its own: -> (uint8 [T,S,S,3], video_height, video_width)."""
    paths = list_jpeg_frames(video_path)
    out = np.empty((len(paths), image_size, image_size, 3), dtype=np.uint8)
    for t, p in enumerate(paths):
        h, w = _decode_into(p, image_size, out[t])
    return torch.from_numpy(out), h, w


def load_video_frames(video_path, image_size, offload_video_to_cpu, img_mean=IMG_MEAN, img_std=IMG_STD,
                      async_loading_frames=False, compute_device=torch.device("cuda")):
    """Same signature as the reference (sam2/utils/misc.py:172-211).  -> (images, video_height, video_width); `images` is a
    device tensor [T,3,S,S] (synchronous, resident) or a `VideoFrames` loader (asynchronous and / or offloaded), both
    indexable by frame."""
    if isinstance(video_path, str) and os.path.splitext(video_path)[-1] in (".mp4", ".MP4"):
        paths = _Mp4Source(video_path, image_size)  # (the reference's `bytes` input is decord-specific: not mirrored)
    else:
        paths = list_jpeg_frames(video_path)
    frames = VideoFrames(paths, image_size, compute_device, keep_on_device=not offload_video_to_cpu,
                         background=bool(async_loading_frames), img_mean=img_mean, img_std=img_std)
    if not async_loading_frames and not offload_video_to_cpu:
        return frames.resident(), frames.video_height, frames.video_width
    return frames, frames.video_height, frames.video_width
