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


def _load_gold():
    import gzip, json, os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    with gzip.open(os.path.join(here, "golden", "rules_golden.json.gz"), "rt") as f:
        return json.load(f)


def test_whole_chain_on_device_matches_reference_golden(qc):
    """qc_rules_apply (refine -> filter -> merge -> validate in one launch) through DeviceDetectionRules == the
    reference's PostProcessor output, every key and every float64 bit (inference/postprocess.py:64-81)."""
    import json
    from rules_cases import CASES, build_case, jsonable
    from qcgraft.rules import DeviceDetectionRules
    gold = _load_gold()
    for name in sorted(CASES):
        dets, seg, shape, cfg = build_case(**CASES[name])
        out = DeviceDetectionRules(cfg).process(dets, seg, shape)
        assert json.loads(json.dumps(jsonable(out))) == gold[name]["validated"], name
    # all cases as ONE batch (different image shapes, thresholds of the first config): same rows per image
    names = [n for n in sorted(CASES) if "thresholds" not in CASES[n]]
    built = [build_case(**CASES[n]) for n in names]
    outs = DeviceDetectionRules(built[0][3]).process_many([(d, s, sh) for d, s, sh, _ in built])
    for n, o in zip(names, outs):
        assert json.loads(json.dumps(jsonable(o))) == gold[n]["validated"], n


def test_whole_chain_on_device_equals_host_tables_on_random_batches(qc):
    """Seeded stress against the host form (itself pinned to the reference): up to 300 detections per image, big
    clusters (np.mean's pairwise order beyond 8 and beyond 128 members), slivers, edge boxes, no-mask rows."""
    import json
    from rules_cases import build_case, jsonable
    from qcgraft.rules import DetectionRules, DeviceDetectionRules
    specs = [dict(seed=100 + i, k=int(k), shape=shape, cluster=c, seg_frac=sf, one_class=oc, slivers=sl, edges=ed)
             for i, (k, shape, c, sf, oc, sl, ed) in enumerate([
                 (300, (2400, 2400, 3), 0.0, 0.0, True, False, False),    # mostly isolated boxes of one class: one cluster of > 128
                 (300, (640, 640, 3), 0.9, 0.5, False, False, False),
                 (200, (1080, 1920, 3), 0.95, 0.7, True, False, True),
                 (120, (333, 500, 3), 0.5, 0.4, False, True, True),
                 (17, (64, 64, 3), 0.3, 1.0, False, False, True),
                 (1, (640, 640, 3), 0.0, 0.0, False, False, False),
                 (0, (640, 640, 3), 0.0, 0.0, False, False, False),
                 (64, (640, 640, 3), 0.7, 0.9, False, True, False)])]
    ref = []
    for s in specs:
        d, seg, shape, cfg = build_case(**s)
        ref.append(json.loads(json.dumps(jsonable(DetectionRules(cfg).process(d, seg, shape)))))
    built = [build_case(**s) for s in specs]
    got = DeviceDetectionRules(built[0][3]).process_many([(d, s, sh) for d, s, sh, _ in built])
    merged_sizes = []
    for s, r, g in zip(specs, ref, got):
        g = json.loads(json.dumps(jsonable(g)))
        assert g == r, s
        merged_sizes += [d.get("merged_from", 1) for d in r]
    assert max(merged_sizes) > 128 and any(8 < m <= 128 for m in merged_sizes)
