"""BoxList: the return type of the post-processor.

Interface-compatible with the subset of maskrcnn_benchmark/structures/bounding_box.py:9-255 the hot path
touches (constructor, fields, ``convert``, ``clip_to_image``, indexing, ``to``, ``area``), written from
scratch.  Boxes are fp32, ``size`` is (image_width, image_height), ``mode`` is 'xyxy' or 'xywh'; widths
follow the reference's legacy +1 convention (TO_REMOVE = 1)."""
from __future__ import annotations

import torch

_LEGACY_ONE = 1  # inclusive pixel coordinates: width = x2 - x1 + 1


class BoxList:
    def __init__(self, bbox, image_size, mode: str = "xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2:
            raise ValueError(f"bbox should have 2 dimensions, got {bbox.ndimension()}")
        if bbox.size(-1) != 4:
            raise ValueError(f"last dimension of bbox should have a size of 4, got {bbox.size(-1)}")
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (w, h)
        self.mode = mode
        self.extra_fields = {}

    # ---- fields ------------------------------------------------------------------------------
    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, other):
        for k, v in other.extra_fields.items():
            self.extra_fields[k] = v

    # ---- geometry ----------------------------------------------------------------------------
    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        a, b, c, d = self.bbox.split(1, dim=-1)
        if mode == "xywh":  # from xyxy
            new = torch.cat((a, b, c - a + _LEGACY_ONE, d - b + _LEGACY_ONE), dim=-1)
        else:               # xywh -> xyxy
            new = torch.cat((a, b, a + (c - _LEGACY_ONE).clamp(min=0), b + (d - _LEGACY_ONE).clamp(min=0)), dim=-1)
        out = BoxList(new, self.size, mode=mode)
        out._copy_extra_fields(self)
        return out

    def clip_to_image(self, remove_empty: bool = True):
        w, h = self.size
        self.bbox[:, 0].clamp_(min=0, max=w - _LEGACY_ONE)
        self.bbox[:, 1].clamp_(min=0, max=h - _LEGACY_ONE)
        self.bbox[:, 2].clamp_(min=0, max=w - _LEGACY_ONE)
        self.bbox[:, 3].clamp_(min=0, max=h - _LEGACY_ONE)
        if remove_empty:
            keep = (self.bbox[:, 3] > self.bbox[:, 1]) & (self.bbox[:, 2] > self.bbox[:, 0])
            return self[keep]
        return self

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + _LEGACY_ONE) * (b[:, 3] - b[:, 1] + _LEGACY_ONE)
        return b[:, 2] * b[:, 3]

    # ---- tensor-like -------------------------------------------------------------------------
    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def copy_with_fields(self, fields, skip_missing: bool = False):
        out = BoxList(self.bbox, self.size, self.mode)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for f in fields:
            if self.has_field(f):
                out.add_field(f, self.get_field(f))
            elif not skip_missing:
                raise KeyError(f"Field '{f}' not found in {self}")
        return out

    def __repr__(self):
        return (f"{self.__class__.__name__}(num_boxes={len(self)}, image_width={self.size[0]}, "
                f"image_height={self.size[1]}, mode={self.mode})")
