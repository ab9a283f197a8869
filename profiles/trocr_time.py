"""Dev tool: time the TrOCR branch (base configuration, random-init) on N random crops: wall clock of vtd_trocr_generate_crops.
Under `ncu --metrics gpu__time_duration.sum --launch-skip ... --launch-count ...` the launch list shows one decode step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from video_text_detection_system_b200 import _lib, synthetic

N = int(os.environ.get("TROCR_N", "64"))
chunk = int(os.environ.get("TROCR_CHUNK", "64"))
model = synthetic.random_trocr_model("base", seed=0)
eng = _lib.Engine(device=0, dtype="fp16", det_h=32, det_w=32, max_batch=1, max_boxes=64, max_src_h=32, max_src_w=32)
eng.load_trocr(model.state_dict(), crops_per_chunk=chunk)
rng = np.random.default_rng(0)
crops = [rng.integers(0, 256, (int(rng.integers(20, 60)), int(rng.integers(60, 200)), 3), dtype=np.uint8) for _ in range(N)]
eng.trocr_generate_crops(crops[:8], 50)
for rep in range(2):
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    ids, lens = eng.trocr_generate_crops(crops, 50)
    dt = time.perf_counter() - t0
    print("trocr: %d crops in %.4f s = %.1f crops/s, %.1f tokens/crop, %d launches" % (N, dt, N / dt, lens.mean(), eng.launch_count() - l0))
