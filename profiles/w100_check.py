"""One-off check (round 1, last GPU call): CRNN at crop width 100 (BASELINE configs[2] wording, T=24) in both tiers
against the oracle's PyTorch fp32 forward.  Prints max |dlogit| per tier; the pytest version is
tests/test_gpu_parity.py::test_crnn_w100_shapes."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port
from video_text_detection_system_b200 import _lib as E

net = port.build_crnn(seed=3)
x = np.random.default_rng(1).random((70, 3, 32, 100)).astype(np.float32)
with torch.no_grad():
    ref = net(torch.from_numpy(x)).numpy()
for dtype in ("fp32", "bf16"):
    eng = E.Engine(det_h=32, det_w=32, crop_w=100, max_batch=2, max_boxes=64, max_src_h=32, max_src_w=32, dtype=dtype)
    eng.load_recognizer(net.state_dict())
    out = eng.crnn_forward(x)
    print(dtype, out.shape, "max|dlogit| = %.3e" % np.abs(out - ref).max(), "max|ref| = %.3f" % np.abs(ref).max(),
          "argmax agree = %.4f" % (out.argmax(-1) == ref.argmax(-1)).mean(), flush=True)
