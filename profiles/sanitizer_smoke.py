"""compute-sanitizer target: one small detect+recognize batch through every kernel family of the speed tier --
direct-window stem, halo / CTA-pair convolutions, the one-pass DB head, box extraction, crop gather, CRNN (incl. the
persistent clustered BiLSTM) and the greedy decode -- at sizes that keep a memcheck / racecheck run short.
Usage: compute-sanitizer --tool {memcheck,racecheck,synccheck,initcheck} python profiles/sanitizer_smoke.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_text_detection_system_b200 import _lib, synthetic  # noqa: E402

h, w, n = int(os.environ.get("DH", "256")), int(os.environ.get("DW", "512")), 2
det_sd, rec_sd = synthetic.random_state_dicts(seed=0)
frames = synthetic.synthetic_frames(n, 270, 540, seed=1)
bias = torch.from_numpy(synthetic.planted_logit_bias(n, h, w, seed=2, boxes=12)).cuda()
for dtype in os.environ.get("TIERS", "fp16").split(","):
    eng = _lib.Engine(device=0, backbone=18, dtype=dtype, det_h=h, det_w=w, max_batch=n, max_boxes=32, max_src_h=270, max_src_w=540)
    eng.load_detector(det_sd)
    eng.load_recognizer(rec_sd)
    r, c = eng.run_batch(list(frames), thr=0.5, recognize=True, logit_bias_dev=bias.data_ptr())
    print("sanitizer smoke %s: boxes %s, launches %d" % (dtype, c.tolist(), eng.launch_count()))
    eng.close()
