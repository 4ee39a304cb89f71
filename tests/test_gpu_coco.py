"""GPU parity of the result hand-off: osd_coco_records (resize + xywh + compaction, fp32, bit-exact) and the file the
native writer produces from it, against the executed-reference fixture and the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc
from test_oracle_coco import load_coco

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_reference_fixture_byte_identical(golden_dir, tmp_path):
    import oneshotdet_b200 as osd
    from oneshotdet_b200.evaluation import prepare_for_coco_detection

    z, boxes, scores, det, orig, cats, text = load_coco(golden_dir)
    preds = []
    for b, s, (w, h) in zip(boxes, scores, det):
        bl = osd.BoxList(b.to(DEV), (w, h), mode="xyxy")
        bl.add_field("scores", s.to(DEV))
        preds.append(bl)
    infos = [{"width": w, "height": h} for w, h in orig]
    path = tmp_path / "coco_custom_result.json"
    n = prepare_for_coco_detection(preds, infos, cats, str(path))
    assert n == int(z["counts"].sum())
    assert path.read_text() == text


def test_records_from_padded_stage_output_vs_oracle(tmp_path):
    """[E,K,4] + counts as the post-processing stages emit them (rows beyond count are garbage and must be ignored)."""
    from oneshotdet_b200.evaluation import coco_records, write_coco_json

    rng = np.random.RandomState(3)
    e, k = 16, 2000
    det = [(1333, 800)] * e
    orig = [(int(rng.randint(300, 1400)), int(rng.randint(200, 1000))) for _ in range(e)]
    orig[3] = (2666, 1600)                              # equal ratios -> the single-factor branch
    counts = rng.randint(0, k + 1, e).astype(np.int32)
    counts[5] = 0
    counts[7] = k
    boxes = rng.uniform(0, 800, (e, k, 4)).astype(np.float32)
    boxes[..., 2:] += boxes[..., :2]
    scores = rng.uniform(0, 1, (e, k)).astype(np.float32)
    rec, ep = coco_records(torch.from_numpy(boxes).to(DEV), torch.from_numpy(scores).to(DEV),
                           torch.from_numpy(counts).to(DEV), det, orig)
    want = orc.coco_detection_results([torch.from_numpy(boxes[i, :counts[i]]) for i in range(e)],
                                      [torch.from_numpy(scores[i, :counts[i]]) for i in range(e)], det, orig,
                                      list(range(e)), list(range(100, 100 + e)))
    assert rec.size(0) == len(want) == int(counts.sum())
    got = rec.cpu().numpy()
    np.testing.assert_array_equal(got[:, :4], np.asarray([r["bbox"] for r in want], dtype=np.float32))
    np.testing.assert_array_equal(got[:, 4], np.asarray([r["score"] for r in want], dtype=np.float32))
    np.testing.assert_array_equal(ep.cpu().numpy(), np.asarray([r["image_id"] for r in want], dtype=np.int32))
    path = tmp_path / "r.json"
    write_coco_json(rec, ep, list(range(e)), list(range(100, 100 + e)), path)
    assert path.read_text() == orc.coco_results_json(want)
