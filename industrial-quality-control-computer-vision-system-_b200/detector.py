"""Drop-in host classes with the reference's own names, signatures, result dictionaries and error
convention, running on the CUDA kernels through ``ffi``:

    YOLODetector            models/yolo_model.py:17-260
    ResNetClassifier        models/resent_model.py:21-287
    EnsemblePredictor       models/ensemble.py:20-448
    ImageSegmentator        inference/segmentation.py:21-146
    QualityControlDetector  inference/detector.py:21-395

so ``app.py`` / ``api/routes.py`` / ``PostProcessor`` / ``SPCAnalyzer`` run on top unchanged.  Every
public method catches and returns ``{"error": str(e)}`` like the reference (never raises); there is
no CPU fallback -- without a CUDA device the constructors record the error and every call returns it.

Channel-order bookkeeping follows the reference literally: each class performs the same
``cv2.cvtColor(BGR2RGB)`` flips the reference performs at the same place (as a K1/K3/K4 ``swap_rb``
flag instead of a copy), so whatever order the caller hands in, each stage sees what the reference's
stage would have seen.
"""
from __future__ import annotations

import logging
import os
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Tuple

import cv2
import numpy as np
import torch

from . import ffi, nets
from .pipeline import CLASS_NAMES, DARK_CLASSES, RESNET_MEAN, RESNET_STD, InspectionPipeline

logger = logging.getLogger(__name__)

_shared_pipelines: Dict[tuple, InspectionPipeline] = {}


def _qcfg(config: Optional[Dict]) -> Dict:
    return (config or {}).get("qcgraft", {}) or {}


def _pipeline(config: Optional[Dict], conf=0.7, iou=0.5, nc=5) -> InspectionPipeline:
    """One InspectionPipeline (qc_ctx, streams, staging buffers; nets are installed by the model wrappers) per
    (device, class count, image size, weight files) shared by the wrapper classes of one detector."""
    cfg = _qcfg(config)
    dev = int(cfg.get("device", 0))
    key = (dev, nc, int(cfg.get("imgsz", 640)), cfg.get("_weights"), bool(cfg.get("synthetic_weights", False)),
           cfg.get("yolo_scale", "n"))
    if key not in _shared_pipelines:
        _shared_pipelines[key] = InspectionPipeline(cfg.get("yolo_scale", "n"), nc, device=dev, conf=conf, iou=iou,
                                                    imgsz=int(cfg.get("imgsz", 640)),
                                                    use_graphs=bool(cfg.get("cuda_graphs", True)), build_nets=False)
    return _shared_pipelines[key]


def _install_yolo(pipe: InspectionPipeline, model_path: str, config: Optional[Dict], nc: int):
    """models/yolo_model.py:45-69: the checkpoint file is loaded or load_model() fails.  Seeded random-init weights
    only on the explicit ``config['qcgraft']['synthetic_weights'] = True`` (benchmarks / tests; no network here)."""
    cfg = _qcfg(config)
    if cfg.get("synthetic_weights", False):
        src = f"synthetic:yolov8{cfg.get('yolo_scale', 'n')}"
        if pipe.yolo_source != src:
            pipe.install_yolo(nets.build_yolo(cfg.get("yolo_scale", "n"), nc, 0), src)
        return
    src = os.path.abspath(model_path)
    if pipe.yolo_source != src:
        if not os.path.exists(model_path):
            raise FileNotFoundError(f"YOLO weights not found: {model_path}")
        module, _scale, ck_nc = nets.load_yolo_checkpoint(model_path)
        if ck_nc != nc:
            raise ValueError(f"{model_path}: checkpoint has {ck_nc} classes, configuration has {nc}")
        pipe.install_yolo(module, src)


def _install_resnet(pipe: InspectionPipeline, model_path: str, config: Optional[Dict], nc: int):
    """models/resent_model.py:85-107.  The reference falls back to ImageNet weights (a download) when the file is
    missing; offline that does not exist, so a missing file fails unless synthetic weights are asked for."""
    cfg = _qcfg(config)
    if cfg.get("synthetic_weights", False):
        if pipe.resnet_source != "synthetic:resnet50":
            pipe.install_resnet(nets.build_resnet50(nc, 0), "synthetic:resnet50")
        return
    src = os.path.abspath(model_path)
    if pipe.resnet_source != src:
        if not os.path.exists(model_path):
            raise FileNotFoundError(f"ResNet weights not found: {model_path} (no network for the ImageNet fallback)")
        pipe.install_resnet(nets.load_resnet_checkpoint(model_path, nc), src)


def _decode(image: np.ndarray) -> np.ndarray:
    if isinstance(image, np.ndarray) and image.ndim == 1:
        image = cv2.imdecode(image, cv2.IMREAD_COLOR)          # models/yolo_model.py:74-76
    return image


def _jpeg_info(image, device: int = 0):
    """qc_jpeg_info of an encoded upload (the 1-D uint8 buffer of app.py:190-194) when K10 decodes it on the device
    (baseline JPEG), else None: other encodings -- PNG, progressive JPEG, ... -- keep the reference's own host call,
    cv2.imdecode, like before."""
    if not (isinstance(image, np.ndarray) and image.ndim == 1 and image.dtype == np.uint8 and image.size >= 4
            and image[0] == 0xFF and image[1] == 0xD8):
        return None
    try:
        return ffi.Context.current(device).jpeg_info(image)
    except ffi.QcGraftError:
        return None


def _as_rgb3(image: np.ndarray) -> np.ndarray:
    if image.ndim == 2:
        image = np.repeat(image[:, :, None], 3, 2)
    return np.ascontiguousarray(image[:, :, :3], dtype=np.uint8)


# ----------------------------------------------------------------------------------------------
class YOLODetector:
    """models/yolo_model.py:17 -- K1 + YOLOv8 + K2 instead of ``ultralytics.YOLO.__call__``."""

    def __init__(self, model_path: str = "yolov8n.pt", confidence_threshold: float = 0.7,
                 nms_threshold: float = 0.5, device: str = "auto", config: Optional[Dict] = None):
        self.model_path = model_path
        self.confidence_threshold = confidence_threshold
        self.nms_threshold = nms_threshold
        self.device = "cuda" if device == "auto" else device
        self.class_names = list(CLASS_NAMES)
        self.model = None
        self._config = config
        self.load_model()

    def load_model(self) -> bool:
        try:
            pipe = _pipeline(self._config, self.confidence_threshold, self.nms_threshold, len(self.class_names))
            _install_yolo(pipe, self.model_path, self._config, len(self.class_names))
            self.model = pipe
            return True
        except Exception as e:                                    # models/yolo_model.py:67-69
            logger.error(f"Error loading YOLO model: {str(e)}")
            self._load_error = str(e)
            self.model = None
            return False

    def preprocess_image(self, image: np.ndarray) -> np.ndarray:
        """models/yolo_model.py:71-82 (kept for API parity; predict() folds the flip into K1)."""
        image = _decode(image)
        if isinstance(image, np.ndarray) and image.ndim == 3 and image.shape[2] == 3:
            image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
        return image

    def predict(self, image: np.ndarray, _head_override=None, _frames: Optional[torch.Tensor] = None) -> Dict:
        if self.model is None:
            logger.error("Model not loaded")
            return {"error": "Model not loaded"}
        try:
            start_time = time.time()
            img = _as_rgb3(_decode(image))
            pipe = self.model
            pipe.conf, pipe.iou = self.confidence_threshold, self.nms_threshold
            frames = _frames if _frames is not None else torch.from_numpy(img[None]).to(pipe.dev)
            # reference: one flip here (:80) and one inside ultralytics ([..., ::-1]) -> net sees `image` order
            det, anchor, count, _ = pipe.detect(frames, head_override=_head_override, swap_rb=False)
            k = int(count[0])
            rows = det[0, :k].cpu().numpy()
            inference_time = time.time() - start_time
            detections = self._parse_results(rows, img.shape)
            return {"detections": detections, "inference_time_ms": inference_time * 1000,
                    "image_shape": img.shape[:2], "total_detections": len(detections)}
        except Exception as e:
            logger.error(f"Error during YOLO prediction: {str(e)}")
            return {"error": str(e)}

    def _parse_results(self, rows: np.ndarray, image_shape: Tuple) -> List[Dict]:
        """models/yolo_model.py:128-160 on K2's [k,6] rows (x1,y1,x2,y2,conf,cls)."""
        detections = []
        for i, row in enumerate(np.asarray(rows, np.float32)):
            x1, y1, x2, y2 = map(int, row[:4])
            conf, class_id = row[4], int(row[5])
            detections.append({
                "id": i,
                "class": self.class_names[class_id] if class_id < len(self.class_names) else f"class_{class_id}",
                "confidence": float(conf),
                "bbox": {"x1": x1, "y1": y1, "x2": x2, "y2": y2, "width": x2 - x1, "height": y2 - y1,
                         "center_x": (x1 + x2) / 2, "center_y": (y1 + y2) / 2},
                "area": (x2 - x1) * (y2 - y1),
                "severity": self._classify_severity(conf, (x2 - x1) * (y2 - y1)),
            })
        return detections

    def _classify_severity(self, confidence: float, area: float) -> str:
        normalized_area = area / (1024 * 1024)                  # models/yolo_model.py:162-172
        if confidence > 0.9 or normalized_area > 0.1:
            return "critical"
        elif confidence > 0.8 or normalized_area > 0.05:
            return "major"
        return "minor"

    def batch_predict(self, images: List[np.ndarray]) -> List[Dict]:
        results = []
        for i, image in enumerate(images):
            result = self.predict(image)
            result["batch_index"] = i
            results.append(result)
        return results

    def get_model_info(self) -> Dict:
        return {"model_path": self.model_path, "device": self.device,
                "confidence_threshold": self.confidence_threshold, "nms_threshold": self.nms_threshold,
                "class_names": self.class_names, "model_loaded": self.model is not None}

    def update_thresholds(self, confidence: Optional[float] = None, nms: Optional[float] = None):
        if confidence is not None:
            self.confidence_threshold = confidence
        if nms is not None:
            self.nms_threshold = nms


# ----------------------------------------------------------------------------------------------
class ResNetClassifier:
    """models/resent_model.py:21 -- K3b + ResNet-50.  The whole-frame input is the reference's own transform chain
    (PIL antialiased BILINEAR Resize((224,224)) -> ToTensor -> Normalize, :57-66) computed bit-exactly on device."""

    def __init__(self, model_path: str = "models/resnet50_qc.pt", num_classes: int = 5, device: str = "auto",
                 config: Optional[Dict] = None):
        self.model_path = model_path
        self.num_classes = num_classes
        self.device = torch.device("cuda" if device == "auto" else device)
        self.class_names = list(CLASS_NAMES)
        self.severity_mapping = {0: "minor", 1: "major", 2: "critical"}
        self.model = None
        self._config = config
        self.load_model()

    def load_model(self) -> bool:
        try:
            self._pipe = _pipeline(self._config, nc=self.num_classes)
            _install_resnet(self._pipe, self.model_path, self._config, self.num_classes)
            self.model = self._pipe.resnet
            return True
        except Exception as e:
            logger.error(f"Error loading ResNet model: {str(e)}")
            self.model = None
            return False

    def _classify(self, frames: torch.Tensor, rois: np.ndarray, swap_rb: bool) -> np.ndarray:
        pipe = self._pipe
        r = torch.from_numpy(np.ascontiguousarray(rois, np.float32)).to(pipe.dev)
        x = pipe.qc.roi_align(frames, r, swap_rb=swap_rb, mean=RESNET_MEAN, std=RESNET_STD, out_dtype=pipe.net_dtype)
        with torch.no_grad():
            logits = pipe.resnet(x.contiguous(memory_format=torch.channels_last))
        return torch.softmax(logits.float(), 1).cpu().numpy()

    def _classify_whole(self, frames: torch.Tensor, swap_rb: bool) -> np.ndarray:
        pipe = self._pipe
        x = pipe.qc.pil_resize(frames, swap_rb=swap_rb, mean=RESNET_MEAN, std=RESNET_STD, out_dtype=pipe.net_dtype)
        with torch.no_grad():
            logits = pipe.resnet(x.contiguous(memory_format=torch.channels_last))
        return torch.softmax(logits.float(), 1).cpu().numpy()

    def _result(self, probs: np.ndarray) -> Dict:
        cid = int(np.argmax(probs))
        conf = float(probs[cid])
        return {"predicted_class": self.class_names[cid], "confidence": conf,
                "class_probabilities": {self.class_names[i]: float(p) for i, p in enumerate(probs)},
                "severity": self._predict_severity(cid, conf)}

    def predict(self, image: np.ndarray, _frames: Optional[torch.Tensor] = None) -> Dict:
        if self.model is None:
            logger.error("Model not loaded")
            return {"error": "Model not loaded"}
        try:
            start_time = time.time()
            img = _as_rgb3(_decode(image))
            frames = _frames if _frames is not None else torch.from_numpy(img[None]).to(self._pipe.dev)
            # reference flips once (models/resent_model.py:123) -> swap_rb=True relative to `image`
            probs = self._classify_whole(frames, swap_rb=True)[0]
            result = self._result(probs)
            result["inference_time_ms"] = (time.time() - start_time) * 1000
            return result
        except Exception as e:
            logger.error(f"Error during ResNet prediction: {str(e)}")
            return {"error": str(e)}

    def predict_batch(self, images: List[np.ndarray]) -> List[Dict]:
        if self.model is None:
            return [{"error": "Model not loaded"} for _ in images]
        try:
            start_time = time.time()
            results = []
            for i, im in enumerate(images):          # shapes may differ: one K3 launch per shape group
                r = self.predict(im)
                r["batch_index"] = i
                results.append(r)
            total = (time.time() - start_time) * 1000
            for r in results:
                r["batch_inference_time_ms"] = total
                r["avg_time_per_image_ms"] = total / max(len(images), 1)
            return results
        except Exception as e:
            logger.error(f"Error during batch prediction: {str(e)}")
            return [{"error": str(e)} for _ in images]

    def _predict_severity(self, class_id: int, confidence: float) -> str:
        critical_classes, major_classes = [0, 2], [1, 3]      # models/resent_model.py:237-255
        if class_id in critical_classes and confidence > 0.8:
            return "critical"
        elif class_id in critical_classes and confidence > 0.6:
            return "major"
        elif class_id in major_classes and confidence > 0.8:
            return "major"
        return "minor"

    def get_model_info(self) -> Dict:
        return {"model_path": self.model_path, "device": str(self.device), "num_classes": self.num_classes,
                "class_names": self.class_names, "model_loaded": self.model is not None, "input_size": (224, 224)}


# ----------------------------------------------------------------------------------------------
class EnsemblePredictor:
    """models/ensemble.py:20.  ``_classify_region`` really crops (K3) and classifies every detection --
    the step the reference mocks (:181-195); the reference's confidence arithmetic is kept so
    downstream numbers (ensemble_confidence, grades, SPC) are the reference's."""

    def __init__(self, yolo_weights: str = "models/yolov8_qc.pt", resnet_weights: str = "models/resnet50_qc.pt",
                 config: Dict = None):
        self.config = config or {}
        self.yolo_detector = None
        self.resnet_classifier = None
        self.ensemble_weights = {"yolo": 0.6, "resnet": 0.4}
        self.confidence_threshold = 0.5
        self.load_models(yolo_weights, resnet_weights)

    def load_models(self, yolo_weights: str, resnet_weights: str) -> bool:
        try:
            yolo_config = self.config.get("model", {})
            # both wrappers share one pipeline: tell them about both weight files
            self.config = dict(self.config)
            self.config["qcgraft"] = dict(_qcfg(self.config), _weights=(str(yolo_weights), str(resnet_weights)))
            self.yolo_detector = YOLODetector(model_path=yolo_weights,
                                              confidence_threshold=yolo_config.get("confidence_threshold", 0.7),
                                              nms_threshold=yolo_config.get("nms_threshold", 0.5), config=self.config)
            self.resnet_classifier = ResNetClassifier(
                model_path=resnet_weights,
                num_classes=len(self.config.get("quality_control", {}).get("defect_classes", [])) or 5,
                config=self.config)
            if self.yolo_detector.model is None or self.resnet_classifier.model is None:
                self.yolo_detector = self.resnet_classifier = None
                return False
            return True
        except Exception as e:
            logger.error(f"Error loading ensemble models: {str(e)}")
            return False

    def predict(self, image: np.ndarray, _head_override=None, _frames: Optional[torch.Tensor] = None) -> Dict:
        if self.yolo_detector is None or self.resnet_classifier is None:
            return {"error": "Models not loaded"}
        try:
            start_time = time.time()
            if _frames is None and isinstance(image, np.ndarray) and image.ndim != 1:
                # one upload shared by YOLO, the whole-frame classifier and the ROI crops
                _frames = torch.from_numpy(_as_rgb3(image)[None]).to(self.resnet_classifier._pipe.dev)
            yolo_results = self.yolo_detector.predict(image, _head_override=_head_override, _frames=_frames)
            resnet_results = self.resnet_classifier.predict(image, _frames=_frames)
            ensemble_results = self._combine_predictions(yolo_results, resnet_results, image.shape, image, _frames=_frames)
            ensemble_results["total_inference_time_ms"] = (time.time() - start_time) * 1000
            return ensemble_results
        except Exception as e:
            logger.error(f"Error in ensemble prediction: {str(e)}")
            return {"error": str(e)}

    def _combine_predictions(self, yolo_results: Dict, resnet_results: Dict, image_shape: Tuple,
                             image: Optional[np.ndarray] = None, _frames: Optional[torch.Tensor] = None,
                             _roi_probs: Optional[Dict[int, np.ndarray]] = None) -> Dict:
        combined = {"detections": [], "global_classification": resnet_results, "detection_summary": {},
                    "quality_assessment": {}, "ensemble_confidence": 0.0}
        if "error" in yolo_results or "error" in resnet_results:
            combined["error"] = "Error in one or more models"
            return combined
        try:
            dets = yolo_results.get("detections", [])
            if _roi_probs is not None:
                roi_probs = _roi_probs
            else:
                roi_probs = self._classify_regions(image, dets, _frames) if image is not None else {}
            enhanced = []
            for d in dets:
                self._extract_region(image_shape, d["bbox"])
                cls = self._classify_region_mock(d)
                if d["id"] in roi_probs:                       # the real K3 -> ResNet-50 hand-off
                    p = roi_probs[d["id"]]
                    cls["roi_classification"] = {"predicted_class": CLASS_NAMES[int(np.argmax(p))],
                                                 "class_probabilities": {CLASS_NAMES[i]: float(v) for i, v in enumerate(p)}}
                enhanced.append(self._merge_detection_classification(d, cls))
            combined["detections"] = enhanced
            combined["detection_summary"] = self._create_detection_summary(enhanced)
            combined["quality_assessment"] = self._assess_overall_quality(enhanced, resnet_results)
            combined["ensemble_confidence"] = self._calculate_ensemble_confidence(yolo_results, resnet_results, enhanced)
        except Exception as e:
            logger.error(f"Error combining predictions: {str(e)}")
            combined["error"] = str(e)
        return combined

    def _classify_regions(self, image: np.ndarray, dets: List[Dict], _frames=None) -> Dict[int, np.ndarray]:
        valid = [d for d in dets if d["bbox"]["x2"] > d["bbox"]["x1"] and d["bbox"]["y2"] > d["bbox"]["y1"]]
        if not valid:
            return {}
        frames = _frames
        if frames is None:
            frames = torch.from_numpy(_as_rgb3(_decode(image))[None]).to(self.resnet_classifier._pipe.dev)
        rois = np.array([[0, d["bbox"]["x1"], d["bbox"]["y1"], d["bbox"]["x2"], d["bbox"]["y2"]] for d in valid], np.float32)
        probs = self.resnet_classifier._classify(frames, rois, swap_rb=True)
        return {d["id"]: probs[i] for i, d in enumerate(valid)}

    def _extract_region(self, image_shape: Tuple, bbox: Dict) -> Dict:
        height, width = image_shape[:2] if len(image_shape) > 2 else (bbox["y2"], bbox["x2"])
        return {"relative_position": {"x": bbox["center_x"] / width, "y": bbox["center_y"] / height},
                "relative_size": {"width": bbox["width"] / width, "height": bbox["height"] / height},
                "absolute_area": bbox["width"] * bbox["height"]}

    def _classify_region_mock(self, detection: Dict) -> Dict:
        return {"predicted_class": detection["class"], "confidence": min(detection["confidence"] * 1.1, 1.0),
                "region_severity": detection["severity"], "classification_source": "ensemble_refined"}

    def _merge_detection_classification(self, detection: Dict, classification: Dict) -> Dict:
        merged = detection.copy()
        yolo_conf, resnet_conf = detection["confidence"], classification["confidence"]
        merged.update({
            "ensemble_confidence": yolo_conf * self.ensemble_weights["yolo"] + resnet_conf * self.ensemble_weights["resnet"],
            "yolo_confidence": yolo_conf, "resnet_confidence": resnet_conf,
            "classification_details": classification,
            "final_severity": self._determine_final_severity(detection, classification)})
        return merged

    def _determine_final_severity(self, detection: Dict, classification: Dict) -> str:
        scores = {"minor": 1, "major": 2, "critical": 3}
        final = max(scores.get(detection["severity"], 1), scores.get(classification["region_severity"], 1))
        for severity, score in scores.items():
            if score == final:
                return severity
        return "minor"

    def _create_detection_summary(self, detections: List[Dict]) -> Dict:
        if not detections:
            return {"total_defects": 0, "defect_counts": {}, "severity_distribution": {},
                    "average_confidence": 0.0, "max_severity": "none"}
        counts, sev, confs = {}, {"minor": 0, "major": 0, "critical": 0}, []
        for d in detections:
            counts[d["class"]] = counts.get(d["class"], 0) + 1
            sev[d["final_severity"]] += 1
            confs.append(d["ensemble_confidence"])
        max_sev = "critical" if sev["critical"] else "major" if sev["major"] else "minor" if sev["minor"] else "none"
        return {"total_defects": len(detections), "defect_counts": counts, "severity_distribution": sev,
                "average_confidence": np.mean(confs) if confs else 0.0, "max_severity": max_sev}

    def _assess_overall_quality(self, detections: List[Dict], global_classification: Dict) -> Dict:
        if not detections:                                       # models/ensemble.py:282-337
            return {"quality_grade": "A", "pass_fail": "PASS", "defect_density": 0.0, "risk_level": "low",
                    "recommended_action": "accept"}
        crit = sum(1 for d in detections if d["final_severity"] == "critical")
        major = sum(1 for d in detections if d["final_severity"] == "major")
        minor = sum(1 for d in detections if d["final_severity"] == "minor")
        if crit > 0:
            g = ("F", "FAIL", "high", "reject")
        elif major > 2:
            g = ("D", "FAIL", "high", "reject")
        elif major > 0:
            g = ("C", "CONDITIONAL", "medium", "review")
        elif minor > 3:
            g = ("B", "CONDITIONAL", "low", "review")
        else:
            g = ("A", "PASS", "low", "accept")
        return {"quality_grade": g[0], "pass_fail": g[1], "defect_density": len(detections), "risk_level": g[2],
                "recommended_action": g[3], "defect_breakdown": {"critical": crit, "major": major, "minor": minor}}

    def _calculate_ensemble_confidence(self, yolo_results, resnet_results, detections) -> float:
        try:
            yolo_conf = np.mean([d["yolo_confidence"] for d in detections]) if detections else 0.0
            resnet_conf = resnet_results.get("confidence", 0.0)
            return float(yolo_conf * self.ensemble_weights["yolo"] + resnet_conf * self.ensemble_weights["resnet"])
        except Exception as e:
            logger.error(f"Error calculating ensemble confidence: {str(e)}")
            return 0.0

    def batch_predict(self, images: List[np.ndarray]) -> List[Dict]:
        results = []
        for i, image in enumerate(images):
            result = self.predict(image)
            result["batch_index"] = i
            results.append(result)
        return results

    def update_ensemble_weights(self, yolo_weight: float, resnet_weight: float):
        total = yolo_weight + resnet_weight
        self.ensemble_weights = {"yolo": yolo_weight / total, "resnet": resnet_weight / total}

    def get_model_info(self) -> Dict:
        info = {"ensemble_weights": self.ensemble_weights, "confidence_threshold": self.confidence_threshold,
                "models_loaded": {"yolo": self.yolo_detector is not None, "resnet": self.resnet_classifier is not None}}
        if self.yolo_detector:
            info["yolo_info"] = self.yolo_detector.get_model_info()
        if self.resnet_classifier:
            info["resnet_info"] = self.resnet_classifier.get_model_info()
        return info


# ----------------------------------------------------------------------------------------------
class _LazyFullMask:
    """Materialises the reference's full-frame per-detection mask (inference/segmentation.py:93-95) only
    when a consumer touches it: n_det x H x W bytes of D2H would dwarf the kernels."""

    def __init__(self, shape, y1, x1, local):
        self.shape, self._y1, self._x1, self._local, self._arr = shape, y1, x1, local, None

    def __array__(self, dtype=None, copy=None):
        if self._arr is None:
            m = np.zeros(self.shape, np.uint8)
            h, w = self._local.shape
            m[self._y1:self._y1 + h, self._x1:self._x1 + w] = self._local
            self._arr = m
        return self._arr if dtype is None else self._arr.astype(dtype)

    def __gt__(self, other):
        return np.asarray(self) > other

    def extent(self):
        """(y_min, y_max, x_min, x_max) of mask > 0 in frame coordinates, or None -- what
        inference/postprocess.py:784-790 reads off the full-frame mask, without materialising it."""
        if "_ext" in self.__dict__:
            return self._ext
        self._ext = self._extent()
        return self._ext

    def _extent(self):
        loc = self._local
        if loc.dtype == np.uint8 and loc.ndim == 2 and loc.flags.c_contiguous:
            x, y, w, h = cv2.boundingRect(loc)                     # bounding box of the non-zero pixels
            if w == 0 or h == 0:
                return None
            return y + self._y1, y + h - 1 + self._y1, x + self._x1, x + w - 1 + self._x1
        m = loc > 0
        rows = np.flatnonzero(m.any(1))
        if rows.size == 0:
            return None
        cols = np.flatnonzero(m.any(0))
        return rows[0] + self._y1, rows[-1] + self._y1, cols[0] + self._x1, cols[-1] + self._x1

    def __getitem__(self, k):
        return np.asarray(self)[k]


class _LazyContours:
    """The contour list of one region (inference/segmentation.py:102-110).  Its LENGTH comes from the device (K8 counts
    the external contours with contourArea >= 10 without tracing them); the point arrays are traced by cv2 on the
    ROI-sized mask only when somebody reads them."""

    def __init__(self, local, x1, y1, count):
        self._local, self._x1, self._y1, self._n, self._items = local, x1, y1, int(count), None

    def _get(self):
        if self._items is None:
            cs, _ = cv2.findContours(np.ascontiguousarray(self._local), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
            out = []
            for c in cs:
                if cv2.contourArea(c) >= 10:                                    # :498-500
                    g = c.copy(); g[:, 0, 0] += self._x1; g[:, 0, 1] += self._y1
                    out.append(g)
            self._items = out
        return self._items

    def __len__(self):
        return self._n

    def __iter__(self):
        return iter(self._get())

    def __getitem__(self, i):
        return self._get()[i]


class _LazyConcat:
    """``segmentation_results['contours']``: all regions' contours in order (:131), traced on first use."""

    def __init__(self, parts):
        self._parts = parts

    def __len__(self):
        return sum(len(p) for p in self._parts)

    def __iter__(self):
        for p in self._parts:
            yield from p

    def __getitem__(self, i):
        return list(self)[i]


class ImageSegmentator:
    """inference/segmentation.py:21 -- the reference's per-class methods on device (K7, csrc/k7_refseg.cuh):

        crack / scratch          'adaptive'   cv2.adaptiveThreshold(GAUSSIAN_C) -> 3x3 close/open -> CC filter(20)   :248-298
        discoloration / others   'threshold'  sigma=1 blur -> Otsu x class factor -> CC filter(10) -> fill holes ->
                                              disk(2) close/open                                            :197-246,449-462
        dent / contamination     the reference's watershed (:300-346) and region growing (:348-447) are out of scope
                                 (SURVEY.md §2 row 3); these two classes take the reference's own fallback (:175-183):
                                 the threshold method, reported as 'threshold_fallback'.

    All detections of a frame go through at most two launch sequences (one per method)."""

    METHOD_BY_CLASS = {"crack": "adaptive", "scratch": "adaptive", "dent": "watershed",
                       "discoloration": "threshold", "contamination": "region_growing"}           # :187-193
    THRESHOLD_FACTORS = {"crack": 0.9, "scratch": 0.9, "dent": 1.1, "discoloration": 0.8, "contamination": 0.85}  # :218-224
    DARK = ("crack", "scratch")                                                                   # :229, :278

    def __init__(self, config: Dict = None):
        self.config = config or {}
        self.segmentation_methods = {"threshold": None, "watershed": None, "adaptive": None, "region_growing": None}
        try:
            self._pipe = _pipeline(self.config)
        except Exception as e:
            logger.error(f"Error initialising segmentator: {str(e)}")
            self._pipe = None

    def _select_segmentation_method(self, defect_class: str) -> str:
        return self.METHOD_BY_CLASS.get(defect_class, "threshold")

    def _select(self, detections: List[Dict], H: int, W: int):
        """(detection index, x1, y1, x2, y2, method name) of every detection with a non-empty ROI."""
        sel = []
        for i, d in enumerate(detections):
            b = d["bbox"]
            x1, y1 = max(b["x1"], 0), max(b["y1"], 0)
            x2, y2 = min(b["x2"], W), min(b["y2"], H)             # numpy slicing clips the same way (:80)
            if x2 <= x1 or y2 <= y1:
                continue                                           # roi.size == 0, :82-83
            m = self._select_segmentation_method(d["class"])
            sel.append((i, x1, y1, x2, y2, m if m in ("adaptive", "threshold") else "threshold_fallback"))
        return sel

    def segment_defects(self, image: np.ndarray, detections: List[Dict], _frames: Optional[torch.Tensor] = None) -> Dict:
        try:
            if not detections:
                return {"segmented_regions": [], "masks": [], "contours": [], "area_analysis": {},
                        "total_defect_area": 0, "defect_density": 0.0}
            if self._pipe is None:
                return {"error": "CUDA segmentation backend not available"}
            qc, dev = self._pipe.qc, self._pipe.dev
            img = np.ascontiguousarray(image if image.ndim == 2 else image[:, :, :3], np.uint8)
            H, W = img.shape[:2]
            sel = self._select(detections, H, W)
            frames = _frames if _frames is not None else torch.from_numpy(img[None]).to(dev)
            local_masks, areas, sums, cstats = {}, {}, {}, {}
            for method, code, iters in (("adaptive", ffi.QC_SEG_REF_ADAPTIVE, 1), ("threshold", ffi.QC_SEG_REF_THRESHOLD, 2)):
                grp = [k for k, s_ in enumerate(sel) if (s_[5] == "adaptive") == (method == "adaptive")]
                if not grp:
                    continue
                cls = [detections[sel[k][0]]["class"] for k in grp]
                regs = dict(img=[0] * len(grp), x=[sel[k][1] for k in grp], y=[sel[k][2] for k in grp],
                            w=[sel[k][3] - sel[k][1] for k in grp], h=[sel[k][4] - sel[k][2] for k in grp],
                            polarity=[0 if c in self.DARK else 1 for c in cls])
                if method == "adaptive":
                    regs["min_area"] = 20                                                  # :289
                else:
                    regs["min_area"] = 10                                                  # :452
                    regs["thr_scale"] = [self.THRESHOLD_FACTORS.get(c, 1.0) for c in cls]
                    regs["max_comp"] = [2] * len(grp)
                res = qc.segment(frames, regs, morph_iters=iters, want_labels=False, swap_rb=False, method=code,
                                 want_sums=True, want_contours=True)
                mask_all = res["mask"].cpu().numpy()
                area = res["area"].cpu().numpy()
                sm = res["sums"].cpu().numpy() if res["sums"] is not None else None
                ct = res["contours"].cpu().numpy()
                for j, k in enumerate(grp):
                    _, x1, y1, x2, y2, _m = sel[k]
                    o = int(res["regions"][j].out_offset)
                    local_masks[k] = mask_all[o:o + (x2 - x1) * (y2 - y1)].reshape(y2 - y1, x2 - x1)
                    areas[k] = int(area[j])
                    sums[k] = sm[j] if sm is not None else None
                    cstats[k] = ct[j]
            return self._assemble(img, detections, sel, local_masks, areas, sums, cstats=cstats)
        except Exception as e:
            logger.error(f"Error in defect segmentation: {str(e)}")
            return {"error": str(e)}

    def _assemble(self, img: np.ndarray, detections: List[Dict], sel, local_masks, areas, sums, bgr: bool = False,
                  cstats=None) -> Dict:
        """The reference's result dictionary (:113-138) from the device outputs of every selected ROI.  ``img`` is the
        frame the masks were computed on (``bgr``: in the caller's channel order, i.e. before detector.py:234).
        ``cstats[k]`` = K8's (n_axis, n_diag, n_contours): perimeter, compactness and the contour count come from the
        device and the contour point arrays are traced lazily; without it cv2 traces every mask here."""
        out = {"segmented_regions": [], "masks": [], "contours": [], "area_analysis": {},
               "total_defect_area": 0, "defect_density": 0.0}
        H, W = img.shape[:2]
        total_image_area = H * W
        total_defect_area = 0
        parts = []
        for k, (i, x1, y1, x2, y2, method) in enumerate(sel):
            d = detections[i]
            local = local_masks[k]
            defect_area = areas[k]
            total_defect_area += defect_area
            if cstats is not None:
                n_axis, n_diag, n_cont = (int(v) for v in cstats[k][:3])
                gcontours = _LazyContours(local, x1, y1, n_cont)
                perimeter = n_axis + float(np.sqrt(2.0)) * n_diag                # sum of cv2.arcLength, :508-513
                if n_cont == 0 or defect_area == 0 or perimeter == 0:            # :515-525
                    compactness = 0.0
                else:
                    compactness = min((4 * np.pi * defect_area) / (perimeter ** 2), 1.0)
            else:
                contours = self._find_contours(local)
                gcontours = []
                for c in contours:
                    g = c.copy(); g[:, 0, 0] += x1; g[:, 0, 1] += y1
                    gcontours.append(g)
                perimeter = self._calculate_perimeter(contours)
                compactness = self._calculate_compactness(defect_area, contours)
            parts.append(gcontours)
            gmask = _LazyFullMask((H, W), y1, x1, local)
            if method == "adaptive" and sums.get(k) is not None:
                conf = self._adaptive_confidence_from_counts(sums[k], defect_area)      # K9: Canny overlap on device
            elif method == "adaptive":
                roi = img[y1:y2, x1:x2]
                if roi.ndim == 3:            # gray of the ROI == ROI of the gray frame (per-pixel transform)
                    roi = cv2.cvtColor(np.ascontiguousarray(roi), cv2.COLOR_BGR2GRAY if bgr else cv2.COLOR_RGB2GRAY)
                conf = self._adaptive_confidence(roi, local, defect_area)
            else:
                conf = self._threshold_confidence(sums[k], defect_area)
            out["segmented_regions"].append({
                "detection_id": i, "defect_class": d["class"], "confidence": d.get("confidence", 0.0),
                "bbox": d["bbox"], "mask": gmask, "local_mask": local, "contours": gcontours,
                "area_pixels": defect_area, "area_percentage": (defect_area / total_image_area) * 100,
                "perimeter": perimeter, "compactness": compactness,
                "segmentation_method": method, "confidence_score": conf})
            out["masks"].append(gmask)
        out["contours"] = _LazyConcat(parts) if cstats is not None else [c for p_ in parts for c in p_]
        out["total_defect_area"] = total_defect_area
        out["defect_density"] = (total_defect_area / total_image_area) * 100
        out["area_analysis"] = self._analyze_defect_areas(out["segmented_regions"])
        return out

    @staticmethod
    def _threshold_confidence(sums, area: int) -> float:
        """inference/segmentation.py:527-555 from K7's integer sums (sum of gray over the mask, sum of gray, sum of
        gray^2, pixels): |mean_fg - mean_bg| / (std(roi) + 1e-6) / 3, clipped to 1."""
        s_fg, s_all, s_sq, n = (int(v) for v in sums)
        if area == 0 or area == n:
            return 0.0
        fg_mean, bg_mean = s_fg / area, (s_all - s_fg) / (n - area)
        std = float(np.sqrt(n * s_sq - s_all * s_all) / n)
        return float(min(abs(fg_mean - bg_mean) / (std + 1e-6) / 3.0, 1.0))

    @staticmethod
    def _adaptive_confidence_from_counts(counts, area: int) -> float:
        """inference/segmentation.py:557-578 from K9's integers (Canny edge pixels on the mask, Canny edge pixels, -,
        pixels of the ROI)."""
        n_both, _n_edge, _, size = (int(v) for v in counts)
        ratio = area / size
        if ratio < 0.01 or ratio > 0.8:
            return 0.3
        return float(min(0.5 + 0.5 * (n_both / area), 1.0))

    @staticmethod
    def _adaptive_confidence(roi: np.ndarray, local: np.ndarray, area: int) -> float:
        """inference/segmentation.py:557-578 (Canny edge overlap; host cv2 on the ROI)."""
        try:
            ratio = area / local.size
            if ratio < 0.01 or ratio > 0.8:
                return 0.3
            edges = cv2.Canny(np.ascontiguousarray(roi), 50, 150)
            overlap = np.sum((edges > 0) & (local > 0)) / area
            return float(min(0.5 + 0.5 * overlap, 1.0))
        except Exception:
            return 0.5

    def _find_contours(self, mask: np.ndarray) -> List[np.ndarray]:
        if mask is None:
            return []
        contours, _ = cv2.findContours(np.ascontiguousarray(mask), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        return [c for c in contours if cv2.contourArea(c) >= 10]      # :498-500

    def _calculate_perimeter(self, contours) -> float:
        return float(sum(cv2.arcLength(c, True) for c in contours))

    def _calculate_compactness(self, area: float, contours) -> float:
        if not contours or area == 0:
            return 0.0
        p = self._calculate_perimeter(contours)
        return 0.0 if p == 0 else min((4 * np.pi * area) / (p ** 2), 1.0)

    def _analyze_defect_areas(self, regions: List[Dict]) -> Dict:
        if not regions:
            return {}
        areas = [r["area_pixels"] for r in regions]
        pct = [r["area_percentage"] for r in regions]
        return {"total_regions": len(regions), "total_area_pixels": sum(areas), "total_area_percentage": sum(pct),
                "average_area_pixels": np.mean(areas), "median_area_pixels": np.median(areas),
                "max_area_pixels": max(areas), "min_area_pixels": min(areas), "area_std": np.std(areas),
                "size_distribution": {"small_defects": sum(1 for a in areas if a < 100),
                                      "medium_defects": sum(1 for a in areas if 100 <= a < 1000),
                                      "large_defects": sum(1 for a in areas if a >= 1000)}}


# ----------------------------------------------------------------------------------------------
class _PassThroughPostProcessor:
    """Stand-in with PostProcessor's interface (inference/postprocess.py:32-109).  The reference's rule
    engine is plain numpy/sklearn on <= 300 dicts and runs unchanged on these results: pass it as
    ``config['qcgraft']['postprocessor']`` (SURVEY.md §8f row f1 is the device version)."""

    def __init__(self, config):
        self.config = config

    def process_results(self, ensemble_results, segmentation_results, image_shape):
        seg = {r["detection_id"]: r for r in (segmentation_results or {}).get("segmented_regions", [])}
        dets = []
        for d in ensemble_results.get("detections", []):
            d = dict(d)
            if d["id"] in seg:
                r = seg[d["id"]]
                d["segmentation"] = {k: r[k] for k in ("area_pixels", "area_percentage", "perimeter", "compactness",
                                                       "confidence_score")}
            dets.append(d)
        return {"detections": dets, "quality_assessment": ensemble_results.get("quality_assessment", {}),
                "risk_analysis": {}, "recommendations": [],
                "metadata": {"total_detections": len(dets), "image_shape": tuple(image_shape)}}

    def update_config(self, new_config):
        self.config.update(new_config)


class RulesPostProcessor(_PassThroughPostProcessor):
    """PostProcessor.process_results (inference/postprocess.py:32-109) with the detection chain
    refine -> filter -> merge -> validate computed by ``rules.DetectionRules`` (the array form of row f1, pinned
    against the reference's own class), the quality assessment (:466-599), the risk analysis (:601-682) and the
    recommendations (:684-762; the reference walks a Python *set* of class names there, i.e. in no reproducible
    order -- here the classes come in order of first appearance).  This is the default post-processor of
    ``QualityControlDetector``; the reference's own object can be plugged through ``config['qcgraft']['postprocessor']``."""

    TYPE_RECOMMENDATIONS = {                                                                       # :737-762
        "crack": ["Check material stress levels and handling procedures",
                  "Verify temperature control during manufacturing"],
        "scratch": ["Review handling and packaging procedures",
                    "Check for abrasive contact points in production line"],
        "dent": ["Inspect handling equipment for damage", "Review impact protection during transport"],
        "discoloration": ["Check chemical process parameters",
                          "Verify environmental conditions (humidity, temperature)"],
        "contamination": ["Review cleaning procedures and protocols",
                          "Check for foreign material sources in production area"]}

    def __init__(self, config):
        super().__init__(config)
        self.rules = self._make_rules()

    def _make_rules(self):
        """K6 on the device (``qc_rules_apply``) is the default wherever a GPU is present;
        ``config['qcgraft']['rules_on_device'] = False`` selects the host tables of qcgraft/rules.py, which produce the
        same dictionaries (they are what the CPU tests pin against the reference's PostProcessor)."""
        from .rules import DetectionRules, DeviceDetectionRules
        q = _qcfg(self.config)
        if q.get("rules_on_device", torch.cuda.is_available()):
            return DeviceDetectionRules(self.config, int(q.get("device", 0)))
        return DetectionRules(self.config)

    def _generate_recommendations(self, detections, quality, risk) -> List[str]:
        rec = []
        if quality["pass_fail_status"] == "FAIL":
            rec.append("REJECT: Product does not meet quality standards")
            if quality["severity_breakdown"]["critical"] > 0:
                rec.append("Critical defects detected - investigate root cause immediately")
            if quality["severity_breakdown"]["major"] > 2:
                rec.append("Multiple major defects - review manufacturing process")
        elif quality["pass_fail_status"] == "CONDITIONAL":
            rec.append("CONDITIONAL PASS: Monitor closely and consider rework")
            rec.append("Increase inspection frequency for similar products")
        if risk["overall_risk_level"] == "high":
            rec.append("High risk detected - implement immediate corrective actions")
        if risk["defect_clustering"]["has_clusters"]:
            rec.append("Defect clustering detected - check for systematic manufacturing issues")
        seen = []
        for d in detections:
            if d["class"] not in seen:
                seen.append(d["class"])
                rec.extend(self.TYPE_RECOMMENDATIONS.get(d["class"], []))
        if quality["average_confidence"] < 0.7:
            rec.append("Low detection confidence - consider additional inspection methods")
        if quality["defect_density"] > 5.0:
            rec.append("High defect density - review entire manufacturing process")
        return rec

    def process_results_many(self, items) -> List[Dict]:
        """process_results for every (ensemble_results, segmentation_results, image_shape) of a batch; with the device
        rules the detection chains of all images run in ONE kernel launch."""
        chains: List[Optional[List[Dict]]] = [None] * len(items)
        if hasattr(self.rules, "process_many"):
            try:
                idx = [i for i, it in enumerate(items) if it[0].get("detections")]
                done = self.rules.process_many([(items[i][0]["detections"], items[i][1] or {}, items[i][2]) for i in idx])
                for i, d in zip(idx, done):
                    chains[i] = d
            except Exception as e:
                logger.error(f"Error in batched post-processing: {str(e)}")
                return [{"error": str(e)} for _ in items]
        return [self.process_results(*it, _chain=chains[i]) for i, it in enumerate(items)]

    def process_results(self, ensemble_results, segmentation_results, image_shape, _chain=None):
        try:
            from datetime import datetime
            meta = {"processing_timestamp": datetime.now().isoformat(), "image_shape": image_shape,
                    "post_processing_version": "1.0.0"}
            raw = ensemble_results.get("detections", [])
            if not raw:                                                  # early return of :62-64
                return {"detections": [], "quality_assessment": self.rules._assess_quality([], {}),
                        "risk_analysis": {}, "recommendations": [], "metadata": dict(meta, total_detections=0)}
            dets = _chain if _chain is not None else self.rules.process(raw, segmentation_results or {}, image_shape)
            quality = self.rules._assess_quality(dets, segmentation_results or {})
            risk = self.rules._analyze_risks(dets, quality)
            rec = self._generate_recommendations(dets, quality, risk)
        except Exception as e:                                    # the reference's convention: never raise
            logger.error(f"Error in post-processing: {str(e)}")
            return {"error": str(e)}
        return {"detections": dets, "quality_assessment": quality, "risk_analysis": risk, "recommendations": rec,
                "metadata": dict(meta, total_detections=len(dets))}

    def update_config(self, new_config):
        super().update_config(new_config)
        self.rules = self._make_rules()


class QualityControlDetector:
    """inference/detector.py:21 -- same constructor, methods, attributes and result keys."""

    def __init__(self, yolo_weights: str, resnet_weights: str, config: Dict):
        self.config = config
        self.ensemble_predictor = None
        self.segmentator = None
        self.postprocessor = None
        self.performance_stats = {"total_predictions": 0, "total_time": 0.0, "average_time": 0.0}
        # encoded uploads: how many were decoded by K10 on the device / by cv2.imdecode on the host (non-JPEG,
        # progressive, ...: the reference's own call)
        self.decode_stats = {"device_jpeg": 0, "host_imdecode": 0}
        self.initialize_components(yolo_weights, resnet_weights)

    def initialize_components(self, yolo_weights: str, resnet_weights: str):
        self.ensemble_predictor = EnsemblePredictor(yolo_weights=yolo_weights, resnet_weights=resnet_weights,
                                                    config=self.config)
        self.segmentator = ImageSegmentator(self.config)
        # the reference always runs PostProcessor (inference/detector.py:59-64): the array-form restatement is the
        # default; 'passthrough' or any object with process_results/update_config can be plugged instead
        pp = _qcfg(self.config).get("postprocessor", "rules")
        if pp == "rules" or pp is None:
            pp = RulesPostProcessor(self.config)
        elif pp == "passthrough":
            pp = _PassThroughPostProcessor(self.config)
        self.postprocessor = pp

    def predict(self, image: np.ndarray, include_segmentation: bool = True, _head_override=None, _inject=None) -> Dict:
        start_time = time.time()
        try:
            if not self._validate_image(image):
                return {"error": "Invalid image input"}
            if self._fused():
                return self._predict_group([image], include_segmentation, _head_override, _inject)[0]
            processed_image = self._preprocess_image(image)
            ensemble_results = self.ensemble_predictor.predict(processed_image)
            if "error" in ensemble_results:
                return ensemble_results
            segmentation_results = {}
            if include_segmentation and ensemble_results.get("detections"):
                segmentation_results = self.segmentator.segment_defects(processed_image, ensemble_results["detections"])
            final_results = self.postprocessor.process_results(ensemble_results, segmentation_results,
                                                               processed_image.shape)
            inference_time = time.time() - start_time
            self._update_performance_stats(inference_time)
            final_results.update({"total_inference_time_ms": inference_time * 1000,
                                  "processing_pipeline": "ensemble + segmentation + postprocess",
                                  "image_metadata": self._get_image_metadata(image)})
            return final_results
        except Exception as e:
            logger.error(f"Error in QC prediction: {str(e)}")
            return {"error": str(e)}

    # ---- the batched device path behind predict / predict_batch ------------------------------------------------
    def _fused(self) -> bool:
        """True when the stock drop-in stages are in place (a user-replaced stage keeps the staged path above)."""
        ens = self.ensemble_predictor
        return (type(ens) is EnsemblePredictor and type(self.segmentator) is ImageSegmentator
                and ens.yolo_detector is not None and ens.resnet_classifier is not None
                and ens.yolo_detector.model is not None and ens.resnet_classifier.model is not None
                and ens.yolo_detector.model is ens.resnet_classifier._pipe)

    def _predict_group(self, images: List[np.ndarray], include_segmentation: bool, head_override=None,
                       inject=None) -> List[Dict]:
        """Same-shape validated images -> reference result dicts.  One upload of the frames, one pass of
        InspectionPipeline (K1 -> YOLO -> K2 -> K3b/K3 -> ResNet-50 -> K7), one download of the packed results; the
        per-image dictionaries are then built by the same host code the single-stage classes use."""
        t0 = time.time()
        ens, seg = self.ensemble_predictor, self.segmentator
        ydet, rcls = ens.yolo_detector, ens.resnet_classifier
        pipe = ydet.model
        pipe.conf, pipe.iou = ydet.confidence_threshold, ydet.nms_threshold
        pre = self.config.get("preprocessing", {}) if isinstance(self.config, dict) else {}
        raw_ok = not ("resize" in pre or pre.get("denoise", False) or pre.get("enhance_contrast", False))
        dev_index = pipe.dev.index or 0
        infos = [_jpeg_info(im, dev_index) for im in images] if raw_ok else [None]
        if all(i is not None for i in infos):
            # encoded uploads: K10 decodes the whole group on the device (bit-exact with cv2.imdecode); no decoded
            # frame ever exists on the host, so the assembly below only gets the frame's shape
            batch_dev = pipe.qc.jpeg_decode(images, rgb=False, infos=infos)
            shapes = [(i.height, i.width, 3) for i in infos]
            frames_np = [np.broadcast_to(np.uint8(0), sh) for sh in shapes]
            self.decode_stats["device_jpeg"] += len(images)
        else:
            self.decode_stats["host_imdecode"] += sum(1 for im in images if isinstance(im, np.ndarray) and im.ndim == 1)
            decoded = [_decode(im) for im in images]
            if raw_ok:
                # no optional host preprocessing: the three channel flips are folded into the kernels' swap flags
                frames_np = [_as_rgb3(im) for im in decoded]
                shapes = [self._processed_shape(im) for im in decoded]
            else:
                processed = [self._preprocess_image(im) for im in images]
                frames_np = [_as_rgb3(im) for im in processed]
                shapes = [im.shape for im in processed]
            # one pinned staging tensor per batch shape (kept: page-locking costs milliseconds): the frames are copied
            # into it once and the pipeline uploads it asynchronously on its copy stream
            key = (len(frames_np),) + tuple(frames_np[0].shape)
            cache = self.__dict__.setdefault("_stage_cache", {})
            stage = cache.get(key)
            if stage is None:
                if len(cache) >= 4:
                    cache.clear()
                stage = cache[key] = torch.empty(key, dtype=torch.uint8, pin_memory=True)
            view = stage.numpy()
            for i, f in enumerate(frames_np):
                np.copyto(view[i], f)
            batch_dev = stage
        pend = pipe.run_async(batch_dev, head_override=head_override, inject=inject,
                              include_segmentation=include_segmentation,
                              keep_seg_tensors=include_segmentation, processed=not raw_ok)
        n = len(images)
        # ---- phase 1, while the classifier graphs are still running: the detections are on the host already (the
        # pipeline planned its ROI tables from them), the segmentation finishes early on its own stream
        res = pend.res
        roi_det_l, roi_box_l = res.roi_det.tolist(), res.roi_box.tolist()
        ks_of = [np.nonzero(res.roi_img == i)[0].tolist() for i in range(n)]
        dets_of, errs = [None] * n, [None] * n
        for i in range(n):
            try:
                dets_of[i] = ydet._parse_results(res.det[i, :int(res.count[i])], shapes[i])
            except Exception as e:
                errs[i] = e
        res = pend.seg_result()
        # ROI rows of every image; the packed masks of a method group came down in one (pinned, asynchronous) copy
        masks_host = {}
        if res.seg is not None:
            for g in res.seg:
                m = g["mask_host"] if "mask_host" in g else g["seg"]["mask"].cpu().numpy()
                regs = g["seg"]["regions"]
                offs, ws, hs = regs["out_offset"].tolist(), regs["w"].tolist(), regs["h"].tolist()
                for j, k in enumerate(np.asarray(g["idx"]).tolist()):
                    o = offs[j]
                    masks_host[k] = m[o:o + ws[j] * hs[j]].reshape(hs[j], ws[j])
        # plain Python numbers for the per-ROI scalars (numpy scalar indexing costs ~100 ns a piece, x 10 per ROI)
        seg_method_l = res.seg_method.tolist() if getattr(res, "seg_method", None) is not None else None
        seg_area_l = res.seg_area.tolist() if getattr(res, "seg_area", None) is not None else None
        seg_sums_l = res.seg_sums.tolist() if getattr(res, "seg_sums", None) is not None else None
        seg_cont_l = res.seg_contours.tolist() if getattr(res, "seg_contours", None) is not None else None
        segs_of = [{} for _ in range(n)]
        for i in range(n):
            dets = dets_of[i]
            if errs[i] is not None or not (include_segmentation and dets):
                continue
            try:
                sel, lm, ar, sm, cs_ = [], {}, {}, {}, {}
                for k in ks_of[i]:
                    x1, y1, x2, y2 = roi_box_l[k]
                    code = seg_method_l[k]
                    name = seg._select_segmentation_method(dets[roi_det_l[k]]["class"])
                    name = name if name in ("adaptive", "threshold") else "threshold_fallback"
                    j = len(sel)
                    sel.append((roi_det_l[k], x1, y1, x2, y2, name))
                    lm[j], ar[j], sm[j] = masks_host[k], seg_area_l[k], seg_sums_l[k]
                    cs_[j] = seg_cont_l[k]
                    assert (code == ffi.QC_SEG_REF_ADAPTIVE) == (name == "adaptive")
                segs_of[i] = seg._assemble(frames_np[i], dets, sel, lm, ar, sm, bgr=raw_ok, cstats=cs_)
                for r in segs_of[i]["segmented_regions"]:          # the rule chain's mask extents (memoised), off the tail
                    r["mask"].extent()
            except Exception as e:
                errs[i] = e
        # ---- phase 2: everything else of the batch (classifier probabilities, SPC vector)
        res = pend.result()
        per_image = (time.time() - t0) / n
        out, pending = [], []
        for i in range(n):
            try:
                if errs[i] is not None:
                    raise errs[i]
                shape = shapes[i]
                dets = dets_of[i]
                yolo_results = {"detections": dets, "inference_time_ms": per_image * 1000,
                                "image_shape": tuple(shape[:2]), "total_detections": len(dets)}
                resnet_results = rcls._result(res.global_probs[i])
                resnet_results["inference_time_ms"] = per_image * 1000
                roi_probs = {roi_det_l[k]: res.roi_probs[k] for k in ks_of[i]} if res.roi_probs is not None else {}
                ensemble_results = ens._combine_predictions(yolo_results, resnet_results, shape, None, _roi_probs=roi_probs)
                ensemble_results["total_inference_time_ms"] = per_image * 1000
                if "error" in ensemble_results:
                    out.append(ensemble_results)
                    continue
                out.append(None)
                pending.append((i, (ensemble_results, segs_of[i], shape)))
            except Exception as e:
                logger.error(f"Error in QC prediction: {str(e)}")
                out.append({"error": str(e)})
        # the post-NMS rule chains of the whole group: one K6 launch with the default post-processor
        try:
            if hasattr(self.postprocessor, "process_results_many"):
                finals = self.postprocessor.process_results_many([it for _, it in pending])
            else:
                finals = [self.postprocessor.process_results(*it) for _, it in pending]
        except Exception as e:
            logger.error(f"Error in QC prediction: {str(e)}")
            finals = [{"error": str(e)} for _ in pending]
        if not self.__dict__.get("_gc_frozen") and _qcfg(self.config).get("gc_freeze", True):
            # The nets, CUDA graphs and plans built by the first pass are long-lived: move them out of the collector's
            # reach once, or every few calls a full collection walks ~10^5 objects (80 ms on a 64-frame batch call).
            import gc
            gc.collect()
            gc.freeze()
            self._gc_frozen = True
        per_image = (time.time() - t0) / n
        for (i, _), final in zip(pending, finals):
            self._update_performance_stats(per_image)
            final.update({"total_inference_time_ms": per_image * 1000,
                          "processing_pipeline": "ensemble + segmentation + postprocess",
                          "image_metadata": self._get_image_metadata(images[i])})
            out[i] = final
        return out

    @staticmethod
    def _processed_shape(decoded: np.ndarray) -> tuple:
        return decoded.shape

    def predict_batch(self, images: List[np.ndarray], max_workers: int = 4, _head_override=None, _inject=None) -> List[Dict]:
        """Same signature / result as inference/detector.py:128-180 (results in input order, ``batch_index`` and
        ``batch_statistics`` on every result).  Instead of ``max_workers`` threads calling predict(), valid images
        are grouped by shape and every group goes through the device pipeline as ONE batch."""
        start_time = time.time()
        try:
            results: List[Optional[Dict]] = [None] * len(images)
            groups: Dict[tuple, List[int]] = {}
            fused = self._fused()
            for i, img in enumerate(images):
                if not self._validate_image(img):
                    results[i] = {"error": "Invalid image input"}
                elif not fused:
                    results[i] = self.predict(img)
                else:
                    try:
                        info = _jpeg_info(img, int(_qcfg(self.config).get("device", 0)))
                        shape = (info.height, info.width, 3) if info is not None else tuple(_decode(img).shape)
                        groups.setdefault(shape, []).append(i)
                    except Exception as e:
                        results[i] = {"error": str(e)}
            chunk = int(_qcfg(self.config).get("max_batch", 64))
            for shape, idx in groups.items():
                for c0 in range(0, len(idx), chunk):
                    part = idx[c0:c0 + chunk]
                    try:
                        ho = ij = None
                        if _head_override is not None:
                            sel = torch.as_tensor(part, device=_head_override[0].device)
                            ho = [l.index_select(0, sel) for l in _head_override]
                        if _inject is not None:
                            sel = torch.as_tensor(part, device=_inject[0].device)
                            ij = [l.index_select(0, sel) for l in _inject]
                        rs = self._predict_group([images[i] for i in part], True, ho, ij)
                    except Exception as e:
                        rs = [{"error": str(e)} for _ in part]
                    for i, r in zip(part, rs):
                        results[i] = r
            for i, r in enumerate(results):
                r["batch_index"] = i
            total_time = time.time() - start_time
            stats = {"batch_size": len(images), "total_batch_time_ms": total_time * 1000,
                     "average_time_per_image_ms": (total_time * 1000) / max(len(images), 1),
                     "throughput_images_per_second": len(images) / total_time if total_time > 0 else 0.0}
            for r in results:
                if r:
                    r["batch_statistics"] = stats
            return results
        except Exception as e:
            logger.error(f"Error in batch prediction: {str(e)}")
            return [{"error": str(e)} for _ in images]

    def predict_stream(self, image_generator, callback=None):
        try:
            for i, image in enumerate(image_generator):
                result = self.predict(image)
                result["stream_index"] = i
                result["timestamp"] = time.time()
                if callback:
                    callback(result)
                else:
                    yield result
        except Exception as e:
            logger.error(f"Error in stream prediction: {str(e)}")
            if callback:
                callback({"error": str(e)})
            else:
                yield {"error": str(e)}

    def _validate_image(self, image: np.ndarray) -> bool:
        if image is None or not isinstance(image, np.ndarray):
            return False
        if image.ndim == 1:
            if self._fused() and _jpeg_info(image, int(_qcfg(self.config).get("device", 0))) is not None:
                return True                      # a stream K10 decodes: the header parse is the validation
            try:
                return cv2.imdecode(image, cv2.IMREAD_COLOR) is not None
            except Exception:
                return False
        return image.ndim in (2, 3)

    def _preprocess_image(self, image: np.ndarray) -> np.ndarray:
        """inference/detector.py:226-254 (decode, BGR2RGB, optional stretch resize, bilateral, CLAHE)."""
        if image.ndim == 1:
            image = cv2.imdecode(image, cv2.IMREAD_COLOR)
        if image.ndim == 3 and image.shape[2] == 3:
            image = cv2.cvtColor(image, cv2.COLOR_BGR2RGB)
        pre = self.config.get("preprocessing", {})
        if "resize" in pre:
            ts = pre["resize"]
            image = cv2.resize(image, (ts[1], ts[0]))
        if pre.get("denoise", False):
            image = cv2.bilateralFilter(image, 9, 75, 75)
        if pre.get("enhance_contrast", False):
            lab = cv2.cvtColor(image, cv2.COLOR_RGB2LAB)
            lab[:, :, 0] = cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(lab[:, :, 0])
            image = cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)
        return image

    def _get_image_metadata(self, image: np.ndarray) -> Dict:
        if image.ndim == 1:
            info = _jpeg_info(image, int(_qcfg(self.config).get("device", 0))) if self._fused() else None
            if info is not None:
                shape = (info.height, info.width, 3)
            else:
                decoded = cv2.imdecode(image, cv2.IMREAD_COLOR)
                shape = decoded.shape if decoded is not None else (0, 0, 0)
        else:
            shape = image.shape
        return {"original_shape": shape, "channels": shape[2] if len(shape) > 2 else 1,
                "dtype": str(image.dtype), "size_bytes": image.nbytes if hasattr(image, "nbytes") else 0}

    def _update_performance_stats(self, inference_time: float):
        s = self.performance_stats
        s["total_predictions"] += 1
        s["total_time"] += inference_time
        s["average_time"] = s["total_time"] / s["total_predictions"]

    def get_performance_stats(self) -> Dict:
        stats = self.performance_stats.copy()
        if stats["total_predictions"] > 0:
            stats.update({"average_time_ms": stats["average_time"] * 1000,
                          "throughput_images_per_second": 1.0 / stats["average_time"] if stats["average_time"] > 0 else 0,
                          "total_time_minutes": stats["total_time"] / 60})
        return stats

    def reset_performance_stats(self):
        self.performance_stats = {"total_predictions": 0, "total_time": 0.0, "average_time": 0.0}

    def get_system_info(self) -> Dict:
        info = {"detector_status": "operational",
                "components_loaded": {"ensemble_predictor": self.ensemble_predictor is not None,
                                      "segmentator": self.segmentator is not None,
                                      "postprocessor": self.postprocessor is not None},
                "performance_stats": self.get_performance_stats(), "configuration": self.config}
        if self.ensemble_predictor:
            info["ensemble_info"] = self.ensemble_predictor.get_model_info()
        return info

    def update_config(self, new_config: Dict):
        self.config.update(new_config)
        if self.postprocessor:
            self.postprocessor.update_config(new_config)

    def benchmark(self, test_images: List[np.ndarray], iterations: int = 1) -> Dict:
        all_times, all_results = [], []
        for _ in range(iterations):
            for image in test_images:
                t = time.time()
                result = self.predict(image)
                all_times.append(time.time() - t)
                all_results.append(result)
        times_ms = [t * 1000 for t in all_times]
        n = len(test_images) * iterations
        return {"total_images": n, "iterations": iterations,
                "timing_statistics": {"mean_ms": np.mean(times_ms), "median_ms": np.median(times_ms),
                                      "min_ms": np.min(times_ms), "max_ms": np.max(times_ms), "std_ms": np.std(times_ms),
                                      "p95_ms": np.percentile(times_ms, 95), "p99_ms": np.percentile(times_ms, 99)},
                "throughput": {"images_per_second": n / sum(all_times), "images_per_minute": n / sum(all_times) * 60},
                "accuracy_metrics": self._calculate_accuracy_metrics(all_results)}

    def _calculate_accuracy_metrics(self, results: List[Dict]) -> Dict:
        total = len(results)
        ok = [r for r in results if "error" not in r]
        dets = sum(len(r.get("detections", [])) for r in ok)
        avg = np.mean([r.get("ensemble_confidence", 0.0) for r in ok]) if ok else 0.0
        return {"success_rate": len(ok) / total if total else 0.0,
                "average_detections_per_image": dets / len(ok) if ok else 0.0, "average_confidence": float(avg)}
