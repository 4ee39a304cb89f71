"""Per-kernel timeline of one overlapped step (configs[1] workload) from the library's event marks
(osd_timeline_enable / osd_timeline_read).  Usage: python tools/timeline.py [--serial] [--steps N]"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oneshotdet_b200 as osd  # noqa: E402
from oneshotdet_b200 import _lib  # noqa: E402
from oneshotdet_b200.pipeline import EpisodePipeline  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--serial", action="store_true")
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    b = 16
    pipe = EpisodePipeline(b, 800, 1344, [(800, 1333)] * b)
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in pipe.features + pipe.supp:
        t.normal_(generator=g)
    for l, s in enumerate((8, 16, 32, 64, 128)):
        pipe.cls[l].normal_(-4.0, 2.0, generator=g)
        pipe.ctr[l].normal_(generator=g)
        pipe.reg[l].normal_(0.0, 0.5, generator=g).exp_().mul_(4.0 * s)
    step = pipe.run if a.serial else pipe.run_overlapped
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.osd_timeline_enable(1)
    buf = ctypes.create_string_buffer(1 << 16)
    for i in range(a.steps):
        step()
        n = lib.osd_timeline_read(buf, len(buf))
        print(f"--- step {i} ({n} marks; ms since the first mark; an event fires when the kernel before it has finished)")
        rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
        for name, ms in sorted(rows, key=lambda r: float(r[1])):
            print(f"{float(ms) * 1000:9.1f} us  {name}")
    lib.osd_timeline_enable(0)


if __name__ == "__main__":
    main()
