"""Result hand-off.  CPU side: the oracle's restatement of the reference loop body against a fixture
produced by executing the reference's BoxList.resize / convert + json.dump; and the library's native JSON writer (a
host function of the C ABI -- no GPU involved) byte for byte against CPython's json.dump."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc


def load_coco(golden_dir, name="mixed"):
    z = np.load(os.path.join(golden_dir, f"coco_{name}.npz"))
    e = len(z["counts"])
    boxes = [torch.from_numpy(z[f"boxes{i}"]) for i in range(e)]
    scores = [torch.from_numpy(z[f"scores{i}"]) for i in range(e)]
    det = [tuple(int(v) for v in r) for r in z["image_sizes_wh"]]
    orig = [tuple(int(v) for v in r) for r in z["orig_sizes_wh"]]
    return z, boxes, scores, det, orig, [int(c) for c in z["category_ids"]], bytes(z["json"]).decode()


def test_oracle_matches_executed_reference(golden_dir):
    z, boxes, scores, det, orig, cats, text = load_coco(golden_dir)
    res = orc.coco_detection_results(boxes, scores, det, orig, list(range(len(boxes))), cats)
    assert orc.coco_results_json(res) == text


def test_native_writer_is_byte_identical_to_json_dump(tmp_path):
    from oneshotdet_b200.evaluation import write_coco_json

    rng = np.random.RandomState(0)
    special = [0.0, 1.0, 2.5, 1e-4, 9.999e-5, 1e-5, 1e15, 1e16, 123456789.0, 0.1, 100.0, 16777216.0, 3.4e38, 1e-45, 5e-324,
               0.30000001192092896, 1e22, 1e23]
    vals = np.concatenate([rng.uniform(0, 1500, 400), 10.0 ** rng.uniform(-9, 18, 200), special]).astype(np.float32)
    vals = np.concatenate([vals, -vals[:60]])
    n = len(vals) // 5
    rec = vals[:n * 5].reshape(n, 5)
    ep = np.sort(rng.randint(0, 4, n)).astype(np.int32)
    img, cat = [10, 11, 12, 2 ** 40], [7, 3, 15, 1]
    want = orc.coco_results_json([{"image_id": img[e], "category_id": cat[e], "bbox": [float(x) for x in r[:4]],
                                   "score": float(r[4])} for r, e in zip(rec, ep)])
    path = tmp_path / "coco_custom_result.json"
    write_coco_json(rec, ep, img, cat, path)
    assert path.read_text() == want
    write_coco_json(np.zeros((0, 5), np.float32), np.zeros((0,), np.int32), [], [], path)
    assert path.read_text() == orc.coco_results_json([]) == "[]"


def test_writer_reports_bad_arguments(tmp_path):
    from oneshotdet_b200.evaluation import write_coco_json

    with pytest.raises(RuntimeError, match="names episode"):
        write_coco_json(np.zeros((1, 5), np.float32), np.array([3], np.int32), [0], [1], tmp_path / "x.json")
    with pytest.raises(RuntimeError, match="cannot open"):
        write_coco_json(np.zeros((1, 5), np.float32), np.array([0], np.int32), [0], [1], tmp_path / "no" / "dir" / "x.json")


def test_native_writer_on_random_float32_bit_patterns(tmp_path):
    """Every finite float32 must print exactly as CPython's repr of the widened double: 100 000 random bit patterns
    (normals of every exponent, denormals, both signs); 1 M were checked once by hand."""
    from oneshotdet_b200.evaluation import write_coco_json

    rng = np.random.RandomState(321)
    vals = rng.randint(0, 2 ** 32, size=100_000, dtype=np.uint64).astype(np.uint32).view(np.float32)
    vals = vals[np.isfinite(vals)]
    n = len(vals) // 5
    rec = vals[:n * 5].reshape(n, 5).copy()
    path = tmp_path / "fuzz.json"
    write_coco_json(rec, np.zeros(n, np.int32), [1], [2], path)
    want = orc.coco_results_json([{"image_id": 1, "category_id": 2, "bbox": [float(x) for x in r[:4]], "score": float(r[4])}
                                  for r in rec])
    assert path.read_text() == want
