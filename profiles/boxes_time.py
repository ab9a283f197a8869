"""Dev aid: box extraction alone at the bench shape (16 planes, 50 planted boxes each); with a -DVTD_TIMERS build the
geometry kernel prints its per-phase cycle counts for the first candidates."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_text_detection_system_b200 import _lib, synthetic  # noqa: E402

B, H, W, DH, DW = 16, 1080, 1920, 736, 1312
det_sd, _ = synthetic.random_state_dicts(seed=0)
eng = _lib.Engine(backbone=18, dtype="fp16", det_h=DH, det_w=DW, max_batch=B, max_boxes=64, max_src_h=H, max_src_w=W)
eng.load_detector(det_sd)
frames = synthetic.synthetic_frames(B, H, W, seed=1)
bias = torch.from_numpy(synthetic.planted_logit_bias(B, DH, DW, seed=7, boxes=50)).cuda()
eng.preprocess(list(frames))
eng.detect_maps(B, 0.5, bias.data_ptr())
for _ in range(2):
    eng.extract_boxes(B, H, W)
eng.sync()
eng.set_profiling(True)
N = int(os.environ.get("N", "5"))
for _ in range(N):
    eng.extract_boxes(B, H, W)
eng.sync()
eng.set_profiling(False)
st = {s["name"]: s["ms"] / N for s in eng.op_profile(2) if s["launches"]}
rec, cnt = eng.read_records(B)
print("boxes %.3f ms per %d planes, counts %s" % (st.get("boxes", -1), B, cnt.tolist()))
