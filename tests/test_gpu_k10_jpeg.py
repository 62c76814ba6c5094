"""K10 through the C ABI: qc_jpeg_decode == cv2.imdecode(buf, IMREAD_COLOR) byte for byte (input side;
inference/detector.py:208-234, app.py:190-194), and the drop-in fed with encoded uploads."""
import cv2
import numpy as np
import pytest
import torch

import jpeg_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qc():
    from qcgraft import ffi
    return ffi.Context.current(0)


def test_mixed_batch_matches_cv2_bit_for_bit(qc):
    streams = [b for _, b in jpeg_cases.small_streams(2, 30, 1, 120)]
    outs = qc.jpeg_decode(streams)
    assert isinstance(outs, list) and len(outs) == len(streams) == 120
    for (desc, _), buf, o in zip(jpeg_cases.small_streams(2, 30, 1, 120), streams, outs):
        ref = cv2.imdecode(buf, cv2.IMREAD_COLOR)
        assert tuple(o.shape) == ref.shape, desc
        assert np.array_equal(o.cpu().numpy(), ref), desc


@pytest.mark.parametrize("h,w,sampling,q,restart", [(640, 640, "420", 90, 0), (640, 640, "444", 75, 7),
                                                    (1080, 1920, "420", 85, 0), (333, 500, "422", 95, 3),
                                                    (17, 1023, "420", 50, 0)])
def test_frame_batches_match_cv2_and_fold_the_rgb_flip(qc, h, w, sampling, q, restart):
    rng = np.random.default_rng(h * 7 + w)
    n = 8 if h * w < 10 ** 6 else 3
    streams = [jpeg_cases.encode(jpeg_cases.picture(rng, h, w, i % 2), q, sampling, restart) for i in range(n)]
    ref = np.stack([cv2.imdecode(b, cv2.IMREAD_COLOR) for b in streams])
    out = qc.jpeg_decode(streams)
    assert tuple(out.shape) == (n, h, w, 3) and out.dtype == torch.uint8
    assert np.array_equal(out.cpu().numpy(), ref)
    # rgb=True == the cv2.cvtColor(BGR2RGB) of inference/detector.py:233-234 folded in
    assert np.array_equal(qc.jpeg_decode(streams, rgb=True).cpu().numpy(), ref[..., ::-1])
    # the host thread count does not change a byte
    assert np.array_equal(qc.jpeg_decode(streams, threads=1).cpu().numpy(), ref)


def test_device_result_equals_oracle(qc):
    from oracle import jpeg as J
    for desc, buf in jpeg_cases.small_streams(3, 6, 8, 60):
        assert np.array_equal(qc.jpeg_decode([buf])[0].cpu().numpy(), J.decode(buf.tobytes())), desc


def test_refused_streams_raise(qc):
    from qcgraft import ffi
    rng = np.random.default_rng(9)
    im = jpeg_cases.picture(rng, 64, 64, 1)
    with pytest.raises(ffi.QcGraftError, match="progressive"):
        qc.jpeg_decode([jpeg_cases.encode(im, progressive=True)])
    with pytest.raises(ffi.QcGraftError):
        qc.jpeg_decode([np.zeros(100, np.uint8)])
    assert qc.jpeg_decode([]) == []


def test_detector_takes_encoded_uploads_like_decoded_frames():
    """QualityControlDetector.predict(np.frombuffer(file_bytes, uint8)) (app.py:190-194) == predict(cv2.imdecode(...));
    baseline JPEG goes through K10, PNG through the reference's own cv2.imdecode."""
    from qcgraft import synth
    from qcgraft.detector import QualityControlDetector
    cfg = {"model": {"confidence_threshold": 0.7, "nms_threshold": 0.5},
           "quality_control": {"defect_classes": ["crack", "scratch", "dent", "discoloration", "contamination"]},
           "qcgraft": {"yolo_scale": "n", "device": 0, "synthetic_weights": True}}
    det = QualityControlDetector("none.pt", "none.pt", cfg)
    frames = synth.make_frames(11, 3)
    heads = [l.cuda() for l in synth.make_head(77, 3, 640, 640, 40)]
    streams = [jpeg_cases.encode(np.ascontiguousarray(f[..., ::-1]), 92, "420") for f in frames]
    decoded = [cv2.imdecode(b, cv2.IMREAD_COLOR) for b in streams]
    a = det.predict_batch(streams, _head_override=heads)
    assert det.decode_stats == {"device_jpeg": 3, "host_imdecode": 0}
    b = det.predict_batch(decoded, _head_override=heads)
    for x, y in zip(a, b):
        assert "error" not in x and "error" not in y, (x.get("error"), y.get("error"))
        assert len(x["detections"]) == len(y["detections"]) > 0
        for dx, dy in zip(x["detections"], y["detections"]):
            assert dx["bbox"] == dy["bbox"] and dx["class"] == dy["class"]
            assert dx["confidence"] == dy["confidence"]
            assert dx.get("has_segmentation") == dy.get("has_segmentation")
            assert dx.get("defect_area_pixels") == dy.get("defect_area_pixels")
        assert x["image_metadata"]["original_shape"] == (640, 640, 3)
    ok, png = cv2.imencode(".png", decoded[0])
    c = det.predict(png, _head_override=[l[:1] for l in heads])
    assert "error" not in c and det.decode_stats["host_imdecode"] == 1
    assert [d["bbox"] for d in c["detections"]] == [d["bbox"] for d in b[0]["detections"]]
