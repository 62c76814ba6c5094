"""K6 parity: same-class overlap clustering on device (through the C ABI) vs the host form in qcgraft/rules.py, which
tests/test_rules_cpu.py pins bit for bit against the reference's PostProcessor (sklearn DBSCAN on 1 - IoU).  Labels
are integers: exact."""
import numpy as np
import pytest
import torch

from gpu_common import qc  # noqa: F401

pytestmark = pytest.mark.gpu


def _host_labels(box, cls, keep, count):
    from qcgraft import rules
    n, cap = cls.shape
    out = np.full((n, cap), -2, np.int32)
    for i in range(n):
        k = int(count[i])
        for c in np.unique(cls[i, :k]):
            idx = np.nonzero((cls[i, :k] == c) & (keep[i, :k] != 0))[0]
            if len(idx):
                out[i, idx] = rules.cluster_labels(rules.iou_matrix(box[i, idx]))
    return out


def _random_case(rng, n, cap, kmax, cluster):
    box = np.zeros((n, cap, 4), np.int32)
    cls = rng.integers(0, 5, (n, cap)).astype(np.int32)
    keep = (rng.random((n, cap)) < 0.85).astype(np.uint8)
    count = rng.integers(0, kmax + 1, n).astype(np.int32)
    count[0] = kmax
    for i in range(n):
        ctr = rng.uniform(50, 1200, (max(1, kmax // 8), 2))
        for j in range(cap):
            c = ctr[rng.integers(0, len(ctr))] + rng.normal(0, 8, 2) if rng.random() < cluster else rng.uniform(0, 1280, 2)
            w, h = rng.integers(4, 160, 2)
            x1, y1 = int(c[0]), int(c[1])
            box[i, j] = (x1, y1, x1 + w, y1 + h)
    return box, cls, keep, count


@pytest.mark.parametrize("cap,kmax,cluster", [(300, 300, 0.8), (300, 40, 0.5), (64, 64, 0.0), (1024, 1000, 0.9)])
def test_cluster_labels_match_host_rules(qc, cap, kmax, cluster):
    rng = np.random.default_rng(cap + kmax)
    box, cls, keep, count = _random_case(rng, 5, cap, kmax, cluster)
    lab = qc.rules_cluster(torch.from_numpy(box).cuda(), torch.from_numpy(cls).cuda(), torch.from_numpy(count).cuda(),
                           keep=torch.from_numpy(keep).cuda()).cpu().numpy()
    ref = _host_labels(box, cls, keep, count)
    assert np.array_equal(lab, ref), np.argwhere(lab != ref)[:10]


def test_threshold_edges_and_quirk(qc):
    """IoU exactly on the link threshold, chains (A-B, B-C linked, A-C not), the all-isolated class and the golden
    cases of the reference's own PostProcessor."""
    #            A            B (IoU 1/3 with A)   C (IoU 1/3 with B)   far          3/10+3/10: IoU = 9/191 < .3
    b = np.array([[[0, 0, 10, 10], [5, 0, 15, 10], [10, 0, 20, 10], [500, 500, 510, 510], [100, 100, 110, 110],
                   [107, 107, 117, 117], [0, 0, 10, 10], [0, 0, 13, 10], [900, 0, 913, 10]]], np.int32)
    c = np.array([[0, 0, 0, 0, 1, 1, 2, 2, 2]], np.int32)
    lab = qc.rules_cluster(torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda(),
                           torch.tensor([9], dtype=torch.int32).cuda()).cpu().numpy()
    assert lab[0].tolist() == [0, 0, 0, -1, -1, -1, 0, 0, -1]          # 10/13 >= 0.3 links the class-2 pair
    # no-keep-mask path == all kept; count < cap -> -2 beyond
    lab2 = qc.rules_cluster(torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda(),
                            torch.tensor([4], dtype=torch.int32).cuda()).cpu().numpy()
    assert lab2[0].tolist() == [0, 0, 0, -1, -2, -2, -2, -2, -2]


def test_golden_cases_through_device_labels(qc):
    """The merged output of the reference's PostProcessor is reproduced when the cluster labels come from K6."""
    import gzip, json, os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from rules_cases import CASES, CLASSES, build_case, jsonable
    from qcgraft.rules import DetectionRules
    with gzip.open(os.path.join(here, "golden", "rules_golden.json.gz"), "rt") as f:
        gold = json.load(f)
    for name in ("clustered", "dense_150", "isolated_quirk", "edge_boxes"):
        dets, seg, shape, cfg = build_case(**CASES[name])
        r = DetectionRules(cfg)
        filtered = r._filter_detections(r._refine_detections(dets, seg, shape))
        k = len(filtered)
        box = np.array([[d["bbox"]["x1"], d["bbox"]["y1"], d["bbox"]["x2"], d["bbox"]["y2"]] for d in filtered], np.int32)[None]
        cls = np.array([CLASSES.index(d["class"]) for d in filtered], np.int32)[None]
        lab = qc.rules_cluster(torch.from_numpy(np.ascontiguousarray(box)).cuda(), torch.from_numpy(cls).cuda(),
                               torch.tensor([k], dtype=torch.int32).cuda()).cpu().numpy()[0]
        merged = r._merge_overlapping_detections(filtered, labels=lab)
        assert json.loads(json.dumps(jsonable(merged))) == gold[name]["merged"], name
