"""Times the second stage's dense head (csrc/box_head.cu) at the reference's widths: C=256, MLP 1024, R=2000 ROIs per
episode, B=16 episodes, and prints the library's per-kernel event timeline for one chunk.
Usage: python tools/box_head_time.py [--batch 16] [--rois 2000] [--chunk 0] [--steps 3]"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oneshotdet_b200 as osd  # noqa: E402
from oneshotdet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--rois", type=int, default=2000)
    ap.add_argument("--channels", type=int, default=256)
    ap.add_argument("--mlp", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    b, r, c = a.batch, a.rois, a.channels
    g = torch.Generator(device="cuda").manual_seed(1)
    pooled = torch.empty(b, r, c, 7, 7, device=dev).normal_(generator=g)
    supp = torch.empty(b, 1, c, 7, 7, device=dev).normal_(generator=g)
    torch.manual_seed(3)
    head = osd.BoxHeadDense(c, a.mlp, roi_chunk=a.chunk).to(dev).eval()
    for _ in range(2):
        head(pooled, supp)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        head(pooled, supp)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[-1]) / a.steps
    n = b * r
    ch = c // 2
    flops = 2.0 * n * (49 * (2 * c) * (2 * c) + 49 * (2 * c) * c + 49 * 9 * c * ch + 49 * ch * a.mlp + a.mlp * a.mlp + a.mlp * 10)
    executed = flops * (128.0 / 98.0 if True else 1.0)   # ROI layers run 128-row tiles for 98 useful rows
    print(f"box head: {ms:.3f} ms/step for {n} ROIs  {b / ms * 1e3:.0f} episodes/s  useful {flops / ms / 1e9:.1f} TFLOP/s "
          f"(reference-form FLOPs; ROI-layer tiles are 98/128 full)")
    lib = _lib.load()
    lib.osd_timeline_enable(1)
    buf = ctypes.create_string_buffer(1 << 20)
    head(pooled, supp)
    torch.cuda.synchronize()
    lib.osd_timeline_read(buf, len(buf))
    rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
    tot, prev = {}, 0.0
    for name, t in rows:
        tot[name] = tot.get(name, 0.0) + (float(t) - prev) * 1000
        prev = float(t)
    for k, v in tot.items():
        print(f"{v:10.1f} us  {k}")
    lib.osd_timeline_enable(0)


if __name__ == "__main__":
    main()
