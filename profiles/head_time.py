"""Device time (CUDA events, one batch in flight) of the DB head and of the whole detector program at the bench shape.
Dev aid: run with a -DVTD_DEV build to try VTD_HF_BST / VTD_HF_AHEAD etc."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_text_detection_system_b200 import _lib, synthetic  # noqa: E402

B, H, W, DH, DW = 16, 1080, 1920, 736, 1312
det_sd, rec_sd = synthetic.random_state_dicts(seed=0)
eng = _lib.Engine(backbone=18, dtype=os.environ.get("DT", "fp16"), det_h=DH, det_w=DW, max_batch=B, max_boxes=64, max_src_h=H,
                  max_src_w=W, fuse_head=os.environ.get("FUSE", "1") == "1")
eng.load_detector(det_sd)
frames = synthetic.synthetic_frames(B, H, W, seed=1)
bias = torch.from_numpy(synthetic.planted_logit_bias(B, DH, DW, seed=7, boxes=50)).cuda()
eng.preprocess(list(frames))
for _ in range(3):
    eng.detect_maps(B, 0.5, bias.data_ptr())
eng.sync()
eng.set_profiling(True)
N = 10
for _ in range(N):
    eng.detect_maps(B, 0.5, bias.data_ptr())
eng.sync()
eng.set_profiling(False)
ops = eng.op_profile(0)
tot = sum(o["ms"] for o in ops) / N
head = [o for o in ops if o["Cout"] == 128 and o["KH"] == 3 and o["Cin"] == 256]
st = {s["name"]: s["ms"] / N for s in eng.op_profile(2) if s["launches"]}
print("env", {k: v for k, v in os.environ.items() if k.startswith("VTD_")}, "detector %.3f ms, head conv op %.3f ms, stages %s"
      % (tot, head[0]["ms"] / N if head else -1, {k: round(v, 3) for k, v in st.items()}))
