"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/osd_b200.h declares;
host-side argument validation works without touching a device."""
import ctypes
import os
import re

import pytest

from oneshotdet_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "osd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osd_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in osd_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names  # the ctypes table covers the header, nothing more


def test_version_and_error_string(lib):
    assert lib.osd_version() >= 100
    assert isinstance(lib.osd_last_error(), bytes)


def test_plans_need_no_gpu(lib):
    plan = _lib.NmsPlan()
    assert lib.osd_batched_nms_plan(16, 11600, ctypes.byref(plan)) == 0
    assert plan.padded_len == 11648 and plan.mask_words == 182
    assert plan.workspace_bytes > 16 * 11648 * 182 * 8
    cfg = _lib.FcosConfig()
    cfg.num_levels, cfg.batch = 5, 16
    for l, ((h, w), s) in enumerate(zip([(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)], (8, 16, 32, 64, 128))):
        cfg.height[l], cfg.width[l], cfg.stride[l] = h, w, s
    cfg.pre_nms_thresh, cfg.pre_nms_top_n, cfg.nms_thresh, cfg.post_nms_top_n = 0.0, 6000, 0.8, 2000
    fp = _lib.FcosPlan()
    assert lib.osd_fcos_postprocess_plan(ctypes.byref(cfg), ctypes.byref(fp)) == 0
    assert fp.cand_capacity == 6000 + 4200 + 1050 + 273 + 77 == 11600
    assert fp.out_capacity == 2000
    assert [fp.level_slot[i] for i in range(5)] == [0, 6000, 10200, 11250, 11523]


def test_invalid_arguments_are_reported(lib):
    cfg = _lib.FcosConfig()
    cfg.num_levels = 0
    fp = _lib.FcosPlan()
    assert lib.osd_fcos_postprocess_plan(ctypes.byref(cfg), ctypes.byref(fp)) == -1
    assert b"num_levels" in lib.osd_last_error()
    d = _lib.MatchDesc()
    d.num_levels, d.batch, d.shots, d.channels, d.mode = 1, 1, 1, 8, 7
    assert lib.osd_match_forward(ctypes.byref(d), None) == -1
    assert b"mode" in lib.osd_last_error()


def test_product_path_refuses_cpu_tensors():
    import torch

    import oneshotdet_b200 as osd

    with pytest.raises(RuntimeError, match="no CPU path"):
        osd.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)
    with pytest.raises(RuntimeError, match="B200"):
        osd.match_forward([torch.zeros(1, 8, 2, 2)], [torch.zeros(1, 8, 1, 1)], 1)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "oneshotdet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_torch_extension_loads_and_refuses_cpu_tensors():
    """The thin torch C++ extension over the C ABI (csrc/torch_ext.cpp) builds on the CPU box, imports without a GPU
    and has no CPU path."""
    import torch

    from oneshotdet_b200 import build as osd_build

    osd_build.build_torch_extension()
    from oneshotdet_b200 import _C_torch

    assert _C_torch.version() == _lib.load().osd_version()
    with pytest.raises(RuntimeError, match="no CPU path"):
        _C_torch.nms(torch.zeros(4, 4), torch.zeros(4), 0.5)


def test_second_stage_and_handoff_entry_points_validate_without_a_gpu(lib, tmp_path):
    """Plans and argument checks of the 'next'-row entry points run on the CPU box (no kernel is launched)."""
    cfg = _lib.BoxPostConfig()
    plan = _lib.BoxPostPlan()
    cfg.batch, cfg.rois_per_image, cfg.num_logits, cfg.reg_columns, cfg.reg_offset, cfg.score_mode = 16, 2000, 2, 8, 4, 0
    for k, w in enumerate((10.0, 10.0, 5.0, 5.0)):
        cfg.weights[k] = w
    cfg.detections_per_img = 100
    assert lib.osd_box_postprocess_plan(ctypes.byref(cfg), ctypes.byref(plan)) == 0
    assert plan.cand_capacity == 2000 and plan.out_capacity == 100 and plan.workspace_bytes > 16 * 2000 * 24
    cfg.reg_columns = 6                                    # class 1 would read columns [4, 8)
    assert lib.osd_box_postprocess_plan(ctypes.byref(cfg), ctypes.byref(plan)) == -1
    assert b"regression columns" in lib.osd_last_error()
    cfg.reg_columns, cfg.num_logits = 8, 1                 # softmax needs two logits
    assert lib.osd_box_postprocess_plan(ctypes.byref(cfg), ctypes.byref(plan)) == -1
    assert b"num_logits" in lib.osd_last_error()

    d = _lib.RoiPoolDesc()
    d.num_levels, d.batch, d.rois_per_image, d.channels, d.pooled_size, d.sampling_ratio = 2, 2, 10, 8, 7, 2
    d.k_min, d.k_max = 3, 4
    for l, (h, w) in enumerate(((32, 40), (16, 20))):
        d.height[l], d.width[l], d.spatial_scale[l] = h, w, 1.0 / (8 << l)
    need = ctypes.c_size_t(0)
    assert lib.osd_roi_pool_workspace_bytes(ctypes.byref(d), ctypes.byref(need)) == 0
    assert need.value >= 2 * 8 * (32 * 40 + 16 * 20) * 4
    d.k_max = 7                                            # 5 mapper levels for 2 feature levels
    assert lib.osd_roi_pool(ctypes.byref(d), None) == -1
    assert b"LevelMapper" in lib.osd_last_error()

    # the JSON writer is a host function: it works here
    import numpy as np

    rec = np.asarray([[1.5, 2.0, 3.25, 4.0, 0.5]], np.float32)
    ep = np.asarray([0], np.int32)
    img, cat = np.asarray([42], np.int64), np.asarray([7], np.int64)
    path = str(tmp_path / "r.json").encode()
    assert lib.osd_coco_write_json(rec.ctypes.data, ep.ctypes.data, 1, img.ctypes.data, cat.ctypes.data, 1, path) == 0
    import json

    assert json.load(open(path)) == [{"bbox": [1.5, 2.0, 3.25, 4.0], "category_id": 7, "image_id": 42, "score": 0.5}]


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/osd_b200.h compiles as C99 and a C program links libosd_b200.so and uses the GPU-free entry points."""
    import json
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    src = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(ROOT, "oneshotdet_b200", "lib")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                        "-L", libdir, "-losd_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "r.json")
    r = subprocess.run([exe, out], capture_output=True, text=True)
    assert r.returncode == 0 and "c-abi ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
    assert json.load(open(out)) == [{"bbox": [1.5, 2.0, 3.25, 4.0], "category_id": 3, "image_id": 7, "score": 0.10000000149011612}]
