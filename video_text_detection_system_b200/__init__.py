"""B200-native (sm_100a) detect+recognize hot path behind the app/ml call surface of
malak29/video-text-detection-system (app/ml/__init__.py:1-5 exports the same names)."""
from .models import CRNN, DBNet
from .detector import TextDetector
from .recognizer import TextRecognizer
from .pipeline import VideoTextPipeline
from .utils import ImageProcessor, VideoProcessor

__all__ = ["TextDetector", "DBNet", "TextRecognizer", "CRNN", "VideoTextPipeline", "VideoProcessor",
           "ImageProcessor"]
