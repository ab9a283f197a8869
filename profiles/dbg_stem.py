import os, sys
import numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
from video_text_detection_system_b200 import _lib as E
from oracle import port
net = port.build_dbnet("resnet18", seed=7)
for (h, w, n) in [(736, 1312, 1), (736, 1312, 3), (736, 1024, 1), (352, 1312, 1)]:
    x = np.random.default_rng(h * w).standard_normal((n, 3, h, w)).astype(np.float32)
    outs = []
    for fuse in (True, False):
        eng = E.Engine(backbone=18, dtype="fp16", det_h=h, det_w=w, max_batch=n, fuse_stem=fuse)
        eng.load_detector(net.state_dict())
        eng.dbnet_forward(x)
        outs.append(eng.debug_tensor("c2", n))
        eng.close()
    d = np.abs(outs[0] - outs[1]).max(axis=1)     # [n, H4, W4]
    bad = d > 0
    print(h, w, "c2 shape", outs[0].shape, "bad px", int(bad.sum()), "of", bad.size, "max diff", float(d.max()), "ref max", float(np.abs(outs[1]).max()))
    for i in range(n):
        ys, xs = np.nonzero(bad[i])
        if len(ys):
            print(" img", i, "rows", np.unique(ys)[:40], "cols", np.unique(xs)[:60], "ncols", len(np.unique(xs)))
