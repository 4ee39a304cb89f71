"""Generates the committed golden fixtures by EXECUTING THE UNMODIFIED REFERENCE in the build
container (needs /root/reference; cannot run on the GPU box -- that is why the vectors are
committed).  Run:  python tests/golden/make_golden.py

What is executed (paths relative to /root/reference):
  * maskrcnn_benchmark/csrc/cpu/nms_cpu.cpp            via oracle/_ref/osd_ref_C.so, injected as
                                                        ``maskrcnn_benchmark._C`` before anything imports it
  * tests/test_nms.py                                   the reference's own known-answer tests; their inputs and
                                                        expected indices are recorded into nms_kat.json
  * tests/test_box_coder.py                             the reference's known-answer vector for BoxCoder.decode
                                                        (modeling/box_coder.py:52-95)          -> box_coder_kat.json
  * maskrcnn_benchmark/modeling/rpn/fcos/inference.py   FCOSPostProcessor.forward            -> fcos_post_*.npz
  * maskrcnn_benchmark/modeling/rpn/fcos/fcos.py        FCOSModule.compute_locations_per_level (unbound) -> same
  * modeling/detector/generalized_rcnn.py:100-104,306-311   batch_pooling + product expression -> match_*.npz
  * modeling/roi_heads/box_head/box_head.py:43-54,147-149   concat + compress_dim_conv (the stock nn.Sequential
                                                        the reference builds there)          -> match_*.npz

  * modeling/roi_heads/box_head/inference.py:46-167     PostProcessor.forward (with modeling/box_coder.py)  -> box_post_*.npz

  * modeling/poolers.py:93-125 with layers/roi_align.py  Pooler.forward + LevelMapper                         -> pooler_*.npz

  * modeling/roi_heads/box_head/box_head.py:81-257  ROIBoxHead.forward at eval (pooler, compress_dim_conv,
    feature_aggreg, fc6/fc7, FPNPredictor, PostProcessor), config from config/defaults.py                   -> box_head_*.npz
  * layers/scale.py + torch.exp as in modeling/rpn/fcos/fcos.py:95-97                                       -> fcos_head_tail.npz
  * structures/bounding_box.py:55-127 (BoxList.resize / convert) inside the loop body of
    data/datasets/evaluation/coco/coco_eval.py:137-165                                                      -> coco_*.npz

Inputs are seeded numpy; they are stored next to the outputs so tests never need the reference.
"""
from __future__ import annotations

import json
import os
import sys
import types
import unittest

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("OSD_REFERENCE_ROOT", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import build_ref  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def import_reference():
    ref_c = build_ref.load_ref()
    assert ref_c is not None, "reference ops did not build"
    sys.path.insert(0, REF)
    import maskrcnn_benchmark  # noqa: PLC0415

    maskrcnn_benchmark._C = ref_c
    sys.modules["maskrcnn_benchmark._C"] = ref_c
    return ref_c


def record_nms_kat(ref_c):
    """Run the reference's tests/test_nms.py with a recording shim around the reference nms."""
    import maskrcnn_benchmark.layers as layers  # noqa: PLC0415

    calls = []

    def recording_nms(boxes, scores, thr):
        keep = ref_c.nms(boxes, scores, float(thr))
        calls.append({"boxes": boxes.numpy().astype(np.float32).tolist(),
                      "scores": scores.numpy().astype(np.float32).tolist(),
                      "thresh": float(thr),
                      "keep_sorted": np.sort(keep.numpy()).tolist()})
        return keep

    layers.nms = recording_nms
    sys.path.insert(0, os.path.join(REF, "tests"))
    import importlib  # noqa: PLC0415

    mod = importlib.import_module("test_nms")
    mod.box_nms = recording_nms
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    assert result.wasSuccessful(), "reference test_nms.py failed against its own nms_cpu"
    # the assertions passed, so keep_sorted == the test file's expected indices
    with open(os.path.join(HERE, "nms_kat.json"), "w") as f:
        json.dump({"source": "tests/test_nms.py (5-box x 5 thresholds, 53-box @0.5)", "cases": calls}, f)
    print("nms_kat.json:", len(calls), "cases")


def fcos_case(name, batch, height, width, image_sizes, params, seed, quantize=None):
    from maskrcnn_benchmark.modeling.rpn.fcos.fcos import FCOSModule  # noqa: PLC0415
    from maskrcnn_benchmark.modeling.rpn.fcos.inference import FCOSPostProcessor  # noqa: PLC0415

    cfg = types.SimpleNamespace(MODEL=types.SimpleNamespace(RPN_ONLY=False),
                                FEW_SHOT=types.SimpleNamespace(ADD_ARTIFICIAL_PROPOSALS=False))
    post = FCOSPostProcessor(cfg, params.pre_nms_thresh, params.pre_nms_top_n, params.nms_thresh,
                             params.fpn_post_nms_top_n, params.min_size, num_classes=2, dense_points=1,
                             score_calculator="BINARY").eval()
    cls, reg, ctr = orc.synth_head_outputs(batch, height, width, seed, distinct=quantize is None)
    if quantize is not None:  # tie set: many equal scores
        cls = [torch.round(c * quantize) / quantize for c in cls]
        ctr = [torch.zeros_like(c) for c in ctr]
    # scale the synthetic boxes down to the small images used here
    fake = types.SimpleNamespace(dense_points=1)
    fake.get_dense_locations = lambda loc, stride, device: loc
    locations = []
    for c, s in zip(cls, orc.FPN_STRIDES):
        h, w = c.shape[-2:]
        locations.append(FCOSModule.compute_locations_per_level(fake, h, w, s, torch.device("cpu")))
    with torch.no_grad():
        out = post(locations, [c.clone() for c in cls], [r.clone() for r in reg], [c.clone() for c in ctr],
                   image_sizes)
    data = {"height": height, "width": width, "batch": batch, "seed": seed,
            "image_sizes": np.asarray(image_sizes, dtype=np.int64),
            "params": np.asarray([params.pre_nms_thresh, params.pre_nms_top_n, params.nms_thresh,
                                  params.fpn_post_nms_top_n, params.min_size], dtype=np.float64)}
    for l, (c, r, t, loc) in enumerate(zip(cls, reg, ctr, locations)):
        data[f"cls{l}"] = c.numpy()
        data[f"reg{l}"] = r.numpy()
        data[f"ctr{l}"] = t.numpy()
        data[f"loc{l}"] = loc.numpy()
    for i, bl in enumerate(out):
        assert bl.mode == "xyxy"
        data[f"out_boxes{i}"] = bl.bbox.numpy().astype(np.float32)
        data[f"out_scores{i}"] = bl.get_field("scores").numpy().astype(np.float32)
        data[f"out_size{i}"] = np.asarray(bl.size, dtype=np.int64)  # (w, h)
        data[f"out_fields{i}"] = np.asarray(sorted(bl.fields()))
    np.savez_compressed(os.path.join(HERE, f"fcos_post_{name}.npz"), **data)
    print(f"fcos_post_{name}.npz:", [len(b) for b in out], "detections")


def match_case(name, batch, shots, channels, height, width, seed):
    feats, supp = orc.synth_features(batch, shots, channels, height, width, seed)
    # --- generalized_rcnn.py:100-104 / :306-311, restated verbatim as expressions on tensors
    pooled = []
    for s in supp:
        d, c, h, w = s.shape
        x = s.view(batch, int(d / batch), c, h, w)
        pooled.append(torch.mean(x, dim=1, keepdim=False))
    product = []
    for f, p in zip(feats, pooled):
        _, _, d1, d2 = f.shape
        product.append(f * p.expand(-1, -1, d1, d2))
    # --- box_head.py:147 concat, :43-54 compress_dim_conv (stock torch.nn, weights N(0, .01))
    from torch import nn  # noqa: PLC0415

    torch.manual_seed(seed)
    oc = channels
    compress = nn.Sequential(
        nn.Conv2d(oc * 2, oc * 2, 1), nn.GroupNorm(32, oc * 2), nn.LeakyReLU(0.2),
        nn.Conv2d(oc * 2, oc, 1), nn.GroupNorm(32, oc), nn.LeakyReLU(0.2))
    for layer in compress:
        if isinstance(layer, nn.Conv2d):
            torch.nn.init.normal_(layer.weight, std=0.01)
        if isinstance(layer, nn.GroupNorm):  # exercise the affine terms
            torch.nn.init.normal_(layer.weight, mean=1.0, std=0.1)
            torch.nn.init.normal_(layer.bias, std=0.1)
    compress.eval()
    concat, conv1, fused = [], [], []
    with torch.no_grad():
        for f, p in zip(feats, pooled):
            x = torch.cat((f, p.expand_as(f)), dim=1)
            concat.append(x)
            conv1.append(compress[0](x))
            fused.append(compress(x))
    data = {"batch": batch, "shots": shots, "channels": channels, "height": height, "width": width}
    for l in range(len(feats)):
        data[f"feat{l}"] = feats[l].numpy()
        data[f"supp{l}"] = supp[l].numpy()
        data[f"product{l}"] = product[l].numpy()
        data[f"conv1_{l}"] = conv1[l].numpy()
        data[f"fused{l}"] = fused[l].numpy()
        # concat is feat || broadcast(pooled): store the pooled vector only
        data[f"pooled{l}"] = pooled[l].numpy()
        assert torch.equal(concat[l][:, :channels], feats[l])
    for k, v in compress.state_dict().items():
        data["w_" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, f"match_{name}.npz"), **data)
    print(f"match_{name}.npz written")


def box_post_case(name, batch, rois, image_sizes, params, seed, cls_loss="ce_loss", num_logits=2, agnostic=False):
    """The reference's second-stage PostProcessor.forward (modeling/roi_heads/box_head/inference.py:46-104), executed."""
    from maskrcnn_benchmark.modeling.box_coder import BoxCoder  # noqa: PLC0415
    from maskrcnn_benchmark.modeling.roi_heads.box_head.inference import PostProcessor  # noqa: PLC0415
    from maskrcnn_benchmark.structures.bounding_box import BoxList  # noqa: PLC0415

    cfg = types.SimpleNamespace(FEW_SHOT=types.SimpleNamespace(SECOND_STAGE_CLS_LOSS=cls_loss))
    post = PostProcessor(cfg, params.score_thresh, params.nms_thresh, params.detections_per_img,
                         BoxCoder(weights=params.weights), agnostic).eval()
    logits, reg, props = orc.synth_box_head_outputs(batch, rois, image_sizes, seed, num_logits=num_logits)
    boxes = [BoxList(props[i].clone(), (int(image_sizes[i][1]), int(image_sizes[i][0])), mode="xyxy") for i in range(batch)]
    target_ids = [7 + i for i in range(batch)]
    with torch.no_grad():
        out = post((logits.clone(), reg.clone()), boxes, target_ids=target_ids)
    data = {"batch": batch, "rois": rois, "seed": seed, "image_sizes": np.asarray(image_sizes, dtype=np.int64),
            "params": np.asarray([params.score_thresh, params.nms_thresh, params.detections_per_img], dtype=np.float64),
            "weights": np.asarray(params.weights, dtype=np.float64), "cls_loss": np.asarray(cls_loss),
            "agnostic": np.asarray(int(agnostic)), "target_ids": np.asarray(target_ids, dtype=np.int64),
            "logits": logits.numpy(), "reg": reg.numpy(), "props": props.numpy()}
    for i, bl in enumerate(out):
        assert bl.mode == "xyxy"
        data[f"out_boxes{i}"] = bl.bbox.numpy().astype(np.float32)
        data[f"out_scores{i}"] = bl.get_field("scores").numpy().astype(np.float32)
        data[f"out_labels{i}"] = bl.get_field("labels").numpy().astype(np.int64)
        data[f"out_size{i}"] = np.asarray(bl.size, dtype=np.int64)
        data[f"out_fields{i}"] = np.asarray(sorted(bl.fields()))
    np.savez_compressed(os.path.join(HERE, f"box_post_{name}.npz"), **data)
    print(f"box_post_{name}.npz:", [len(b) for b in out], "detections")


def box_post_cases():
    BP = orc.BoxPostParams
    box_post_case("default", 2, 400, [(480, 640), (500, 375)], BP(0.0, 0.5, 2000), seed=41)
    box_post_case("cut_thresh", 3, 300, [(300, 400)] * 3, BP(0.05, 0.7, 40), seed=42)
    box_post_case("focal_agnostic", 2, 200, [(256, 256), (200, 300)], BP(0.3, 0.5, 2000, (10., 10., 5., 5.), "sigmoid"),
                  seed=43, cls_loss="focal_loss", num_logits=1, agnostic=True)


def pooler_case(name, batch, rois, channels, height, width, image_sizes, seed, resolution=7, sampling=2, lo=12.0):
    """The reference's Pooler.forward (modeling/poolers.py:93-125: LevelMapper + per-level ROIAlign modules), executed."""
    from maskrcnn_benchmark.modeling.poolers import Pooler  # noqa: PLC0415
    from maskrcnn_benchmark.structures.bounding_box import BoxList  # noqa: PLC0415

    scales = tuple(1.0 / s for s in orc.FPN_STRIDES)
    pooler = Pooler(output_size=(resolution, resolution), scales=scales, sampling_ratio=sampling)
    feats, _ = orc.synth_features(batch, 1, channels, height, width, seed)
    boxes_t = orc.synth_rois(batch, rois, image_sizes, seed + 1, lo)
    boxes = [BoxList(boxes_t[i].clone(), (int(image_sizes[i][1]), int(image_sizes[i][0])), mode="xyxy") for i in range(batch)]
    with torch.no_grad():
        out = pooler(tuple(feats), boxes)
        levels = pooler.map_levels(boxes)
    assert tuple(out.shape) == (batch, rois, channels, resolution, resolution)
    data = {"batch": batch, "rois": rois, "channels": channels, "height": height, "width": width, "seed": seed,
            "resolution": resolution, "sampling": sampling, "lo": lo, "image_sizes": np.asarray(image_sizes, dtype=np.int64),
            "scales": np.asarray(scales, dtype=np.float64), "boxes": boxes_t.numpy(),
            "pooled": out.numpy().astype(np.float32), "levels": levels.numpy().astype(np.int64)}
    np.savez_compressed(os.path.join(HERE, f"pooler_{name}.npz"), **data)
    print(f"pooler_{name}.npz: levels", np.bincount(levels.numpy().astype(np.int64), minlength=5).tolist())


def pooler_cases():
    pooler_case("r7_s2", 2, 48, 8, 512, 640, [(500, 640), (512, 600)], seed=71)
    pooler_case("all_levels", 1, 64, 2, 2048, 2304, [(2048, 2300)], seed=75, lo=150.0)
    pooler_case("r3_adaptive", 1, 32, 4, 384, 384, [(384, 384)], seed=73, resolution=3, sampling=0)


def coco_case(name, image_sizes_wh, orig_sizes_wh, counts, seed):
    """The per-image body of prepare_for_coco_detection (data/datasets/evaluation/coco/coco_eval.py:137-156) with the
    reference's own BoxList.resize / convert executed, and the json.dump of :163-165.  (The function itself needs
    pycocotools and a dataset on disk; its dataset bookkeeping is not part of the hand-off format.)"""
    import json  # noqa: PLC0415

    from maskrcnn_benchmark.structures.bounding_box import BoxList  # noqa: PLC0415

    rng = np.random.RandomState(seed)
    coco_results = []
    data = {"image_sizes_wh": np.asarray(image_sizes_wh, dtype=np.int64), "orig_sizes_wh": np.asarray(orig_sizes_wh, dtype=np.int64),
            "counts": np.asarray(counts, dtype=np.int64), "seed": seed}
    cats = [int(c) for c in rng.randint(1, 21, len(counts))]
    data["category_ids"] = np.asarray(cats, dtype=np.int64)
    for image_id, ((w, h), (ow, oh), n) in enumerate(zip(image_sizes_wh, orig_sizes_wh, counts)):
        x1 = rng.uniform(0, w - 2, n); y1 = rng.uniform(0, h - 2, n)
        x2 = np.minimum(x1 + rng.uniform(1, w / 2, n), w - 1); y2 = np.minimum(y1 + rng.uniform(1, h / 2, n), h - 1)
        boxes = np.stack([x1, y1, x2, y2], 1).astype(np.float32)
        scores = rng.uniform(0, 1, n).astype(np.float32) ** 3
        data[f"boxes{image_id}"], data[f"scores{image_id}"] = boxes, scores
        prediction = BoxList(torch.from_numpy(boxes.copy()), (w, h), mode="xyxy")
        prediction.add_field("scores", torch.from_numpy(scores.copy()))
        if len(prediction) == 0:
            continue
        prediction = prediction.resize((ow, oh))
        prediction = prediction.convert("xywh")
        bl = prediction.bbox.tolist()
        sl = prediction.get_field("scores").tolist()
        coco_results.extend([{"image_id": image_id, "category_id": cats[image_id], "bbox": box, "score": sl[k]}
                             for k, box in enumerate(bl)])
    text = json.dumps(coco_results, sort_keys=True, indent=4, separators=(',', ':'))
    data["json"] = np.frombuffer(text.encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"coco_{name}.npz"), **data)
    print(f"coco_{name}.npz:", len(coco_results), "records,", len(text), "bytes of JSON")


def box_head_case(name="c64", batch=2, rois=24, channels=64, height=256, width=320, seed=101):
    """The reference's whole second stage, executed: ROIBoxHead.forward at eval
    (modeling/roi_heads/box_head/box_head.py:81-257) = FPN2ROIFeatureExtractor (Pooler) -> cat with the expanded support
    -> compress_dim_conv -> feature_aggreg -> fc6 -> fc7 -> FPNPredictor -> PostProcessor.  Built from the reference's own
    config defaults (a stub replaces the missing yacs package) with the shipped yaml's ROI_BOX_HEAD settings, 64 channels
    and MLP_HEAD_DIM 64 to keep the fixture small.  Every stage's output is recorded through forward hooks."""
    sys.path.insert(0, os.path.join(HERE, "_stubs"))
    from maskrcnn_benchmark.config import cfg as ref_cfg  # noqa: PLC0415
    from maskrcnn_benchmark.modeling.roi_heads.box_head.box_head import ROIBoxHead  # noqa: PLC0415
    from maskrcnn_benchmark.structures.bounding_box import BoxList  # noqa: PLC0415

    cfg = ref_cfg.clone()
    cfg.merge_from_list(["MODEL.DEVICE", "cpu", "MODEL.ROI_HEADS.USE_FPN", True,
                         "MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION", 7,
                         "MODEL.ROI_BOX_HEAD.POOLER_SCALES", (0.125, 0.0625, 0.03125, 0.015625, 0.0078125),
                         "MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO", 2,
                         "MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR", "FPN2ROIFeatureExtractor",
                         "MODEL.ROI_BOX_HEAD.PREDICTOR", "FPNPredictor", "MODEL.ROI_BOX_HEAD.NUM_CLASSES", 2,
                         "MODEL.ROI_BOX_HEAD.MLP_HEAD_DIM", 64,
                         "FEW_SHOT.SECOND_STAGE_METHOD", "concat", "FEW_SHOT.POOLING", "ROI"])
    assert cfg.FEW_SHOT.SECOND_STAGE_CLS_LOSS == "ce_loss" and not cfg.FEW_SHOT.NEG_SUPPORT.TURN_ON
    torch.manual_seed(seed)
    head = ROIBoxHead(cfg, channels).eval()
    with torch.no_grad():   # exercise the affine terms and the biases
        for m in head.modules():
            if isinstance(m, torch.nn.GroupNorm):
                torch.nn.init.normal_(m.weight, mean=1.0, std=0.1)
                torch.nn.init.normal_(m.bias, std=0.1)
        torch.nn.init.normal_(head.predictor.cls_score.bias, std=0.5)
    image_sizes = [(height - 6, width), (height, width - 20)][:batch]
    feats, _ = orc.synth_features(batch, 1, channels, height, width, seed)
    boxes_t = orc.synth_rois(batch, rois, image_sizes, seed + 1, lo=24.0)
    proposals = [BoxList(boxes_t[i].clone(), (int(image_sizes[i][1]), int(image_sizes[i][0])), mode="xyxy") for i in range(batch)]
    g = torch.Generator().manual_seed(seed + 2)
    supp = torch.randn((batch, 1, channels, 7, 7), generator=g)
    target_ids = [3 + i for i in range(batch)]
    taps = {}

    def tap(key):
        def hook(_m, _inp, out):
            taps[key] = [o.detach().clone() for o in out] if isinstance(out, tuple) else out.detach().clone()
        return hook

    head.feature_extractor.register_forward_hook(tap("pooled"))
    head.compress_dim_conv.register_forward_hook(tap("compressed"))
    head.feature_aggreg.register_forward_hook(tap("aggregated"))
    head.fc7.register_forward_hook(tap("fc7"))
    head.predictor.register_forward_hook(tap("predictor"))
    with torch.no_grad():
        _x, result, _ = head(tuple(feats), proposals, features_supp_roipooled=supp, target_ids=target_ids)
    data = {"batch": batch, "rois": rois, "channels": channels, "height": height, "width": width, "seed": seed,
            "image_sizes": np.asarray(image_sizes, dtype=np.int64), "boxes": boxes_t.numpy(), "supp": supp.numpy(),
            "target_ids": np.asarray(target_ids, dtype=np.int64),
            "pooled": taps["pooled"].numpy(), "compressed": taps["compressed"].numpy(),
            "aggregated": taps["aggregated"].numpy(), "fc7_pre_relu": taps["fc7"].numpy(),
            "class_logits": taps["predictor"][0].numpy(), "box_regression": taps["predictor"][1].numpy(),
            "params": np.asarray([cfg.MODEL.ROI_HEADS.SCORE_THRESH, cfg.MODEL.ROI_HEADS.NMS,
                                  cfg.MODEL.ROI_HEADS.DETECTIONS_PER_IMG], dtype=np.float64),
            "weights": np.asarray(cfg.MODEL.ROI_HEADS.BBOX_REG_WEIGHTS, dtype=np.float64)}
    for k, v in head.state_dict().items():
        data["w_" + k] = v.numpy()
    for i, bl in enumerate(result):
        data[f"out_boxes{i}"] = bl.bbox.numpy().astype(np.float32)
        data[f"out_scores{i}"] = bl.get_field("scores").numpy().astype(np.float32)
        data[f"out_labels{i}"] = bl.get_field("labels").numpy().astype(np.int64)
    np.savez_compressed(os.path.join(HERE, f"box_head_{name}.npz"), **data)
    print(f"box_head_{name}.npz:", [len(b) for b in result], "detections;",
          {k: tuple(v.shape) for k, v in data.items() if k in ("pooled", "compressed", "aggregated", "class_logits")})


def record_box_coder_kat():
    """Run the reference's tests/test_box_coder.py (its own known-answer vector for BoxCoder.decode) with a recording
    shim around the reference decode."""
    import importlib  # noqa: PLC0415

    from maskrcnn_benchmark.modeling import box_coder as ref_bc  # noqa: PLC0415

    calls = []
    orig = ref_bc.BoxCoder.decode

    def recording_decode(self, rel_codes, boxes):
        out = orig(self, rel_codes, boxes)
        calls.append({"weights": [float(w) for w in self.weights], "deltas": rel_codes.numpy().astype(np.float32).tolist(),
                      "boxes": boxes.numpy().astype(np.float32).tolist(), "decoded": out.numpy().astype(np.float32).tolist()})
        return out

    ref_bc.BoxCoder.decode = recording_decode
    sys.path.insert(0, os.path.join(REF, "tests"))
    mod = importlib.import_module("test_box_coder")
    result = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(mod))
    ref_bc.BoxCoder.decode = orig
    assert result.wasSuccessful() and calls, "reference test_box_coder.py failed against its own BoxCoder"
    # the test's own expected values (it asserts |decoded - gt| <= 1e-4): keep them next to what the reference computed
    import re  # noqa: PLC0415

    src = open(os.path.join(REF, "tests", "test_box_coder.py")).read()
    gt_block = src[src.index("gt_bbox = ("):src.index("results = box_coder.decode")]
    gt = [float(v) for v in re.findall(r"-?\d+\.\d+", gt_block)]
    calls[0]["expected_by_the_test"] = np.asarray(gt, np.float32).reshape(-1, 4).tolist()
    with open(os.path.join(HERE, "box_coder_kat.json"), "w") as f:
        json.dump({"source": "tests/test_box_coder.py (TestBoxCoder.test_box_decoder, atol 1e-4)", "cases": calls}, f)
    print("box_coder_kat.json:", len(calls), "case(s)")


def head_tail_case():
    """The tail of FCOSHead.forward for the regression branch (modeling/rpn/fcos/fcos.py:95-97): the reference's own
    Scale module (layers/scale.py) and torch.exp, executed on seeded raw bbox_pred maps."""
    from maskrcnn_benchmark.layers import Scale  # noqa: PLC0415

    rng = np.random.RandomState(17)
    data = {}
    for l, (h, w, init) in enumerate([(25, 34, 1.0), (13, 17, 0.83), (7, 9, 1.37)]):
        raw = torch.from_numpy(rng.normal(1.5 + l, 0.8, (2, 4, h, w)).astype(np.float32))
        scale = Scale(init_value=init)
        with torch.no_grad():
            out = torch.exp(scale(raw))
        data[f"raw{l}"], data[f"scale{l}"], data[f"out{l}"] = raw.numpy(), np.float32(init), out.numpy()
    np.savez_compressed(os.path.join(HERE, "fcos_head_tail.npz"), **data)
    print("fcos_head_tail.npz written")


def coco_cases():
    # equal ratios (one factor), unequal ratios, an empty image, an upscale
    coco_case("mixed", [(1333, 800), (1066, 800), (800, 800), (640, 480)], [(500, 300), (500, 375), (400, 400), (1280, 961)],
              [40, 25, 0, 30], seed=91)


def main():
    torch.set_num_threads(1)
    ref_c = import_reference()
    if "--only-box-post" in sys.argv:
        box_post_cases()
        return
    if "--only-box-coder" in sys.argv:
        record_box_coder_kat()
        return
    if "--only-box-head" in sys.argv:
        box_head_case()
        return
    if "--only-head-tail" in sys.argv:
        head_tail_case()
        return
    if "--only-coco" in sys.argv:
        coco_cases()
        return
    if "--only-pooler" in sys.argv:
        pooler_cases()
        return
    record_nms_kat(ref_c)
    P = orc.PostParams
    # small padded inputs (multiples of 128 so every level is non-empty and regular)
    fcos_case("two_stage_small", 2, 256, 384, [(250, 380), (256, 333)], P(0.0, 300, 0.8, 100, 0.0), seed=11)
    fcos_case("stress_small", 3, 256, 256, [(256, 256)] * 3, P(0.01, 120, 0.6, 4000, 0.0), seed=12)
    fcos_case("minsize_small", 2, 128, 256, [(120, 250), (128, 256)], P(0.02, 1000, 0.5, 50, 24.0), seed=13)
    fcos_case("nonms_small", 1, 128, 128, [(128, 128)], P(0.05, 50, 0.0, 30, 0.0), seed=14)
    match_case("s1_c64", 2, 1, 64, 96, 160, seed=21)
    match_case("s3_c64", 2, 3, 64, 64, 96, seed=22)
    box_post_cases()
    pooler_cases()
    coco_cases()
    head_tail_case()
    box_head_case()
    record_box_coder_kat()


if __name__ == "__main__":
    main()
