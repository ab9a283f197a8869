"""GPU, last in collection order.  (1) The two storage libraries of the speed tier -- libvtd_b200.so (IEEE half, shipped)
and libvtd_b200_bf16.so (bfloat16, build option) -- loaded side by side in ONE process do not disturb each other.
(2) An exhaustive small-alphabet check of the greedy decode kernel against the oracle.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fp16_and_bf16_libraries_side_by_side_at_640x640():
    from oracle import port
    from video_text_detection_system_b200 import _lib as E
    net = port.build_dbnet("resnet18", seed=0)
    frames = port.synthetic_frames(2, 640, 640, seed=0)
    x = torch.cat([port.preprocess(f, 640, 640) for f in frames])
    with torch.no_grad():
        ref = port.dbnet_forward(net, x)
    want_p, want_t = ref["probability"].numpy()[:, 0], ref["threshold"].numpy()[:, 0]
    worst = {}
    for dtype in ("bf16", "fp16", "bf16", "fp16"):
        eng = E.Engine(backbone=18, det_h=640, det_w=640, max_batch=2, dtype=dtype, max_src_h=640, max_src_w=640)
        eng.load_detector(net.state_dict())
        eng.preprocess(list(frames))
        eng.detect_maps(2, 0.5)
        p, t, m = eng.read_maps(2)
        assert np.array_equal(m, (p > 0.5).astype(np.uint8))
        err = max(float(np.abs(p - want_p).max()), float(np.abs(t - want_t).max()))
        worst.setdefault(dtype, []).append(err)
        del eng
    print("640x640 max deviation: bf16 %s, fp16 %s" % (worst["bf16"], worst["fp16"]))
    assert worst["fp16"][0] == worst["fp16"][1] <= 1e-2     # the north-star bar of the 16-bit tier, outright
    assert worst["bf16"][0] == worst["bf16"][1] <= 3e-2


def test_ctc_exhaustive_small_alphabet_vs_oracle():
    """Every argmax sequence of length 1..5 over {blank, '0', '1', <unk>} through vtd_ctc_decode against the oracle's
    restatement (itself equal to the reference on the same sequences: tests/test_oracle_vs_reference.py).  Exhaustive over
    the collapse logic: repeats, blanks that do not reset the previous character, <unk> dropped but remembered, the
    confidence indexed by emitted count.  (A Python transcription of the kernel's collapse() passed this on the CPU.)"""
    import itertools
    from oracle import port
    from video_text_detection_system_b200 import _lib as E
    eng = E.Engine(det_h=32, det_w=32, max_batch=1, max_src_h=32, max_src_w=32)
    rng = np.random.default_rng(0)
    V = 97
    for T in range(1, 6):
        seqs = list(itertools.product([0, 1, 2, V - 1], repeat=T))
        p = rng.random((len(seqs), T, V)).astype(np.float32) * 0.5
        for b, seq in enumerate(seqs):
            p[b, np.arange(T), list(seq)] = 0.5 + rng.random(T).astype(np.float32) * 0.5
        p /= p.sum(2, keepdims=True)
        ids, lens, conf = eng.ctc_decode(p, is_prob=True)
        for b, seq in enumerate(seqs):
            text, c, want = port.decode_prediction(torch.from_numpy(p[b]))
            assert ids[b, :lens[b]].tolist() == want, seq
            assert E.ids_to_text(want) == text
            assert conf[b] == pytest.approx(c, abs=1e-6), seq
