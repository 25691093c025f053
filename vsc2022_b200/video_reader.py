"""`vsc.baseline.video_reader` mirror (video_reader.py:13-33, ffmpeg_video_reader.py:19-54): frames of a video at a fixed
rate through an `ffmpeg` binary, as decoded uint8 RGB arrays for the GPU transform (the reference yields PIL images for
torchvision's CPU transforms).  Same ffmpeg command line, same frame numbering and the same timestamps -- including the
reference's rule that `original_fps` is 1 whenever the reader cannot tell the rate (FFMpegVideoReader.fps is always None), so
frame i of ANY `--fps` gets the interval [i, i + 1].  NVDEC decoding is out of scope."""
import os
import subprocess
import tempfile
from abc import ABC, abstractmethod
from typing import Iterable, Optional, Tuple

import numpy as np


class VideoReader(ABC):
    def __init__(self, video_path: str, required_fps: float) -> None:
        self.video_path = video_path
        self.required_fps = required_fps
        self.original_fps = max(1, self.fps) if self.fps else 1
        self.video_frames = None

    @property
    @abstractmethod
    def fps(self) -> Optional[float]:
        pass

    @abstractmethod
    def frames(self) -> Iterable[Tuple[float, float, np.ndarray]]:
        """yields (start_time, end_time, uint8 [H, W, 3] RGB frame)"""


class FFMpegVideoReader(VideoReader):
    def __init__(self, video_path: str, required_fps: float, ffmpeg_path: str):
        self.ffmpeg_path = ffmpeg_path
        super().__init__(video_path, required_fps)

    @property
    def fps(self) -> Optional[float]:
        return None

    def frames(self) -> Iterable[Tuple[float, float, np.ndarray]]:
        from PIL import Image
        with tempfile.TemporaryDirectory() as scratch, open(os.devnull, "w") as null:
            subprocess.check_call(
                [self.ffmpeg_path, "-nostdin", "-y", "-i", self.video_path, "-start_number", "0", "-q", "0",
                 "-vf", "fps=%f" % self.required_fps, os.path.join(scratch, "%07d.png")], stderr=null)
            i = 0
            while True:
                frame_fn = os.path.join(scratch, f"{i:07d}.png")
                if not os.path.exists(frame_fn):
                    break
                with Image.open(frame_fn) as img:        # torchvision default_loader: PIL, converted to RGB
                    frame = np.asarray(img.convert("RGB"))
                i += 1
                yield ((i - 1) / self.original_fps, i / self.original_fps, frame)
