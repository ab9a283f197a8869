"""One-off (round 1, last seconds of GPU budget): the 640x640 deviation of the speed tier built over IEEE half
(VTD_STORAGE=f16 -> libvtd_b200_f16.so) against the oracle; same net and frames as
tests/test_gpu_parity.py::test_config1_640x640_maps_and_boxes_vs_oracle.  Prints as it goes."""
import os, sys
os.environ.setdefault("VTD_STORAGE", "f16")
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from video_text_detection_system_b200 import _lib as E

net = port.build_dbnet("resnet18", seed=0)
frames = port.synthetic_frames(2, 640, 640, seed=0)
eng = E.Engine(backbone=18, det_h=640, det_w=640, max_batch=2, dtype="bf16", max_src_h=640, max_src_w=640, max_boxes=64)
eng.load_detector(net.state_dict())
eng.preprocess(list(frames))
eng.detect_maps(2, 0.5)
p, t, m = eng.read_maps(2)
print("gpu done", os.environ["VTD_STORAGE"], flush=True)
x = torch.cat([port.preprocess(f, 640, 640) for f in frames])
with torch.no_grad():
    ref = port.dbnet_forward(net, x)
for name, got, want in (("prob", p, ref["probability"].numpy()[:, 0]), ("thr", t, ref["threshold"].numpy()[:, 0])):
    err = np.abs(got - want)
    print("%s storage=%s 640x640: max %.5f, pixels over 1e-2: %d of %d, over 1e-3: %d" %
          (name, os.environ["VTD_STORAGE"], err.max(), int((err > 1e-2).sum()), err.size, int((err > 1e-3).sum())), flush=True)
print("mask == prob > thr:", bool(np.array_equal(m, (p > 0.5).astype(np.uint8))), flush=True)
