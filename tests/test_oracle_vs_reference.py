"""CPU, build container only: oracle/port.py against the reference's own source files loaded by path.
Skipped where /root/reference is not mounted (the GPU box)."""
import numpy as np
import pytest
import torch

from oracle import port, reference_loader as RL

pytestmark = pytest.mark.skipif(not RL.available(), reason="reference tree not mounted")


def test_reference_modules_agree_with_port():
    o = port.build_dbnet("resnet18", seed=3)
    ref = RL.reference_dbnet("resnet18", o.state_dict())
    x = torch.randn(1, 3, 64, 96, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        a = port.dbnet_forward(o, x)
        b = port.dbnet_forward(ref, x)       # reference FPN/DBHead classes, repaired wiring
    assert torch.equal(a["probability"], b["probability"]) and torch.equal(a["threshold"], b["threshold"])


def test_reference_post_process_and_transform():
    o = port.build_dbnet("resnet18", seed=0)
    D = RL.reference_detector(RL.reference_dbnet("resnet18", o.state_dict()))
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, (300, 400, 3), dtype=np.uint8)
    import cv2
    t = D.transform(cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    assert torch.equal(t, port.preprocess(frame, 640, 640)[0])
    pm = cv2.GaussianBlur(rng.random((640, 640)).astype(np.float32), (0, 0), 9)
    pm = np.clip((pm - 0.5) * 12 + 0.5, 0, 1)
    assert D._post_process(pm, 1280, 720, 0.5) == port.post_process(pm, 1280, 720, 0.5)


def test_reference_recognizer_agrees_with_port():
    net = port.build_crnn(seed=1)
    R = RL.reference_recognizer(net.state_dict())
    rng = np.random.default_rng(1)
    crops = [rng.integers(0, 256, (30, 90, 3), dtype=np.uint8), rng.integers(0, 256, (50, 20, 3), dtype=np.uint8)]
    a = R.recognize_batch(crops)
    b = port.recognize_batch(net, crops)
    for x, y in zip(a, b):
        assert x["text"] == y["text"] and x["confidence"] == pytest.approx(y["confidence"], abs=1e-7)
    assert R.vocab == port.build_vocab()
    # a 2-D crop makes the reference return the empty result
    assert R.recognize(np.zeros((20, 40), np.uint8)) == {"text": "", "confidence": 0.0}


def test_decode_exhaustive_over_a_small_alphabet():
    """Every argmax sequence of length <= 5 over {blank, '0', '1', <unk>} through the reference's _decode_prediction
    and the oracle's restatement: same text, same confidence.  Exhaustive over the collapse logic and its quirks
    (blank does not reset the previous character, <unk> is dropped but becomes the previous character, the confidence
    is indexed by emitted count)."""
    import itertools
    R = RL.reference_recognizer()
    V = len(R.vocab)
    symbols = [0, 1, 2, V - 1]
    rng = np.random.default_rng(0)
    n = 0
    for T in range(1, 6):
        for seq in itertools.product(symbols, repeat=T):
            p = rng.random((T, V)).astype(np.float32) * 0.5
            p[np.arange(T), list(seq)] = 0.5 + rng.random(T).astype(np.float32) * 0.5      # argmax = seq
            p /= p.sum(1, keepdims=True)
            t = torch.from_numpy(p)
            text_ref, conf_ref = R._decode_prediction(t)
            text, conf, ids = port.decode_prediction(t)
            assert text == text_ref and conf == pytest.approx(conf_ref, abs=1e-7), seq
            assert "".join(port.CHARS[i - 1] for i in ids) == text
            n += 1
    assert n == sum(4 ** T for T in range(1, 6))
