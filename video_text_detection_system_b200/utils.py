"""Frame source used by VideoTextPipeline.process_video.

The reference's VideoProcessor (app/ml/utils/preprocessing.py:11-98) is CPU video decode around
cv2.VideoCapture; it is OUT OF SCOPE of the accelerated path and is kept here only so the
pipeline has the same collaborator objects: same method names, same 10-fps sampling rule, same return
shapes.  ImageProcessor is instantiated by the reference pipeline (pipeliine.py:28) and never called.
"""
from __future__ import annotations

import asyncio
import logging
from pathlib import Path
from typing import Any, AsyncGenerator, Dict, Generator, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)


class VideoProcessor:
    def __init__(self):
        self.supported_formats = [".mp4", ".avi", ".mov", ".mkv", ".wmv"]

    def get_video_info(self, video_path: str) -> Dict[str, Any]:
        try:
            import cv2
            cap = cv2.VideoCapture(video_path)
            if not cap.isOpened():
                raise ValueError(f"Cannot open video: {video_path}")
            fps = cap.get(cv2.CAP_PROP_FPS)
            n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
            info = {"fps": fps, "frame_count": n, "width": int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)),
                    "height": int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), "duration": n / fps if fps > 0 else 0,
                    "format": Path(video_path).suffix.lower()}
            cap.release()
            return info
        except Exception as e:
            logger.error(f"Failed to get video info: {e}")
            return {}

    def extract_frames_at_fps(self, video_path: str, target_fps: int = 10
                              ) -> Generator[Tuple[np.ndarray, int, float], None, None]:
        try:
            import cv2
            cap = cv2.VideoCapture(video_path)
            if not cap.isOpened():
                raise ValueError(f"Cannot open video: {video_path}")
            src_fps = cap.get(cv2.CAP_PROP_FPS)
            interval = max(1, int(src_fps / target_fps))
            number = extracted = 0
            while True:
                ok, frame = cap.read()
                if not ok:
                    break
                if number % interval == 0:
                    yield frame, extracted, number / src_fps
                    extracted += 1
                number += 1
            cap.release()
        except Exception as e:
            logger.error(f"Frame extraction failed: {e}")
            return

    async def extract_frames_generator(self, video_path: str, target_fps: int = 10
                                       ) -> AsyncGenerator[Tuple[np.ndarray, int, float], None]:
        for item in self.extract_frames_at_fps(video_path, target_fps):
            yield item
            await asyncio.sleep(0)

    def extract_single_frame(self, video_path: str, frame_number: int) -> Optional[np.ndarray]:
        try:
            import cv2
            cap = cv2.VideoCapture(video_path)
            cap.set(cv2.CAP_PROP_POS_FRAMES, frame_number)
            ok, frame = cap.read()
            cap.release()
            return frame if ok else None
        except Exception as e:
            logger.error(f"Single frame extraction failed: {e}")
            return None


class ImageProcessor:
    """Placeholder collaborator (the reference never calls it on the inference path)."""
