"""GPU parity of the transformer recogniser (SURVEY.md 8f N1 = row a12: TransformerRecognizer, text_recognizer.py:39-69)
against the HuggingFace classes the reference calls, instantiated from the checkpoint's configuration with seeded random
weights (oracle/trocr_port.py; the checkpoint itself cannot be downloaded here): encoder states, teacher-forced logits,
and greedy generate(max_length=50) token ids."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from video_text_detection_system_b200 import _lib
    return _lib


@pytest.fixture(scope="module")
def tp():
    from oracle import trocr_port
    return trocr_port


def _check(E, tp, kind, n, chunk, tol_enc, tol_logit):
    model = tp.build(kind, seed=0)
    S = tp.image_size(model)
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((n, 3, S, S)).astype(np.float32))
    eng = E.Engine(dtype="fp16", det_h=32, det_w=32, max_batch=1, max_src_h=32, max_src_w=32)
    eng.load_trocr(model.state_dict(), crops_per_chunk=chunk)
    want_ids = tp.generate(model, x, 50)
    L = 8
    teacher = want_ids[:, :L].copy()
    ref_enc, ref_logits = tp.forward_logits(model, x, teacher)
    enc, logits, ids, lens = eng.trocr_forward(x.numpy(), decoder_ids=teacher, max_length=50, want_encoder=True)
    e_enc = np.abs(enc - ref_enc).max() / np.abs(ref_enc).max()
    e_log = np.abs(logits - ref_logits).max()
    agree = float((ids == want_ids).mean())
    first_bad = [int(np.argmax(ids[b] != want_ids[b])) if (ids[b] != want_ids[b]).any() else -1 for b in range(n)]
    print("trocr %s: encoder rel %.2e, teacher-forced |dlogit| %.2e (max |logit| %.2f), greedy id agreement %.3f, first "
          "divergence %s, lengths %s" % (kind, e_enc, e_log, np.abs(ref_logits).max(), agree, first_bad, lens.tolist()))
    assert e_enc <= tol_enc
    assert e_log <= tol_logit * max(1.0, np.abs(ref_logits).max())
    # greedy decoding: identical wherever the reference's own top-2 margin exceeds the logit error
    assert (logits.argmax(-1) == ref_logits.argmax(-1)).mean() >= 0.95
    assert agree >= 0.9
    eng.close()


def test_trocr_tiny_config_vs_huggingface(E, tp):
    _check(E, tp, "tiny", n=5, chunk=2, tol_enc=5e-3, tol_logit=5e-3)       # 5 crops in chunks of 2: the chunk loop too


def test_trocr_base_config_vs_huggingface(E, tp):
    """The configuration of microsoft/trocr-base-printed: ViT-B/16 @384 (577 tokens) + 12-layer decoder, 341 M parameters."""
    _check(E, tp, "base", n=2, chunk=2, tol_enc=1e-2, tol_logit=1e-2)


def test_trocr_crops_through_the_processor(E, tp):
    """recognize()'s whole path on BGR crops of arbitrary size: the device-side Pillow resize + normalisation against the
    checkpoint's image processor, then generate."""
    model = tp.build("tiny", seed=1)
    S = tp.image_size(model)
    rng = np.random.default_rng(5)
    crops = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in [(40, 200), (17, 33), (64, 64), (90, 45), (130, 300)]]
    eng = E.Engine(dtype="fp16", det_h=32, det_w=32, max_batch=1, max_src_h=32, max_src_w=32)
    eng.load_trocr(model.state_dict(), crops_per_chunk=4)
    ids, lens = eng.trocr_generate_crops(crops, max_length=50)
    want = tp.generate(model, tp.processor_pixel_values(crops, S), 50)
    print("trocr crops: agreement %.3f" % float((ids == want).mean()))
    assert (ids == want).mean() >= 0.9
    assert ids[:, 0].tolist() == [2] * len(crops)


def test_pipeline_with_transformer_ocr(tp):
    """VideoTextPipeline(use_transformer_ocr=True), the reference's default: fused detection of the batch, then ONE
    batched pass of the transformer recogniser over all crops; the result schema is the reference's
    (pipeliine.py:127-133), 'recognition_confidence' its constant 0.95 (text_recognizer.py:64)."""
    from oracle import port
    from video_text_detection_system_b200 import VideoTextPipeline, synthetic
    model = tp.build("tiny", seed=2)
    P = VideoTextPipeline(use_transformer_ocr=True, backbone="resnet18", pretrained=False, det_size=(256, 1280),
                          trocr_state_dict=model.state_dict())
    det = port.build_dbnet("resnet18", seed=0)
    P.detector.model.load_state_dict(det.state_dict())
    frames = list(synthetic.synthetic_frames(2, 288, 1440, seed=11))
    bias = torch.from_numpy(synthetic.planted_logit_bias(2, 256, 1280, seed=6, boxes=10)).cuda()
    P.logit_bias_dev = bias.data_ptr()
    got = P.detect_and_recognize(frames)
    S = tp.image_size(model)
    assert sum(len(g) for g in got) >= 10
    for f, regions in zip(frames, got):
        crops = [f[r["bbox"][1]:r["bbox"][3], r["bbox"][0]:r["bbox"][2]] for r in regions]
        want = tp.generate(model, tp.processor_pixel_values(crops, S), 50)
        for r, w in zip(regions, want):
            assert set(r) == {"bbox", "text", "detection_confidence", "recognition_confidence", "polygon"}
            assert r["recognition_confidence"] == 0.95
            ids = [int(v) for v in w if v not in (0, 1, 2)]
            assert r["text"] == " ".join(map(str, ids))
    single = P.process_single_frame(frames[0])
    assert [d["text"] for d in single["detections"]] == [r["text"] for r in got[0]]
    # recognize(): the reference's never-raise convention
    assert P.recognizer.recognize(np.zeros((0, 0, 3), np.uint8)) == {"text": "", "confidence": 0.0}
