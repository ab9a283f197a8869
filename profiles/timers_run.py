"""Dev tool: with the developer build (python video_text_detection_system_b200/build.py --dev; run with VTD_STORAGE=dev) print, for every tcgen05
conv launch of one bench-shaped batch, where each role warp spent its cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import port
from video_text_detection_system_b200 import _lib
B = 16
eng = _lib.Engine(device=0, backbone=18, dtype="16bit", det_h=736, det_w=1312, crop_w=128, max_batch=B, max_boxes=64,
                  max_src_h=1080, max_src_w=1920)
eng.load_detector(port.build_dbnet("resnet18", seed=0).state_dict())
eng.load_recognizer(port.build_crnn(seed=0).state_dict())
frames = np.random.default_rng(0).integers(0, 256, (B, 1080, 1920, 3), dtype=np.uint8)
bias = torch.from_numpy(port.planted_logit_bias(B, 736, 1312, seed=7, boxes=50)).cuda()
for i in range(3):
    if i == 2:
        os.environ["VTD_TIMERS"] = "1"
    eng.run_batch(list(frames), thr=0.5, recognize=True, logit_bias_dev=bias.data_ptr())
