"""Times the 1x1 fusion-conv matching mode (BASELINE configs[1] shapes: 16 episodes, 800x1344, C=256) and prints the
library's per-kernel event timeline.  Usage: python tools/fusion_time.py [--batch 16] [--steps 10] [--check]"""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oneshotdet_b200 as osd  # noqa: E402
from oneshotdet_b200 import _lib  # noqa: E402
from oneshotdet_b200.fusion import PreparedFusion  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--channels", type=int, default=256)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--h", type=int, default=800)
    ap.add_argument("--w", type=int, default=1344)
    ap.add_argument("--check", action="store_true", help="compare one episode against the torch fp32 module on the GPU")
    a = ap.parse_args()
    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    b, c = a.batch, a.channels
    g = torch.Generator(device="cuda").manual_seed(1)
    feats, supp = [], []
    for s in (8, 16, 32, 64, 128):
        h, w = -(-a.h // s), -(-a.w // s)
        feats.append(torch.empty(b, c, h, w, device=dev).normal_(generator=g))
        supp.append(torch.empty(b, c, 1, 1, device=dev).normal_(generator=g))
    torch.manual_seed(7)
    mm = osd.MatchingModule("fusion", channels=c).to(dev)
    pf = PreparedFusion(feats, supp, b, mm.compress_dim_conv, "full")
    for _ in range(3):
        pf()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        pf()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[-1]) / a.steps
    locs = sum(f.shape[2] * f.shape[3] for f in feats)
    flops = 2.0 * locs * b * (c * 2 * c + 2 * c * c)
    print(f"fusion full: {ms:.4f} ms/step  {b / ms * 1e3:.0f} episodes/s  executed {flops / ms / 1e9:.1f} TFLOP/s")
    lib = _lib.load()
    lib.osd_timeline_enable(1)
    buf = ctypes.create_string_buffer(1 << 16)
    for i in range(2):
        pf()
        n = lib.osd_timeline_read(buf, len(buf))
        rows = [ln.split() for ln in buf.value.decode().strip().splitlines()]
        prev = 0.0
        print(f"--- timeline {i}")
        for name, t in rows:
            print(f"{(float(t) - prev) * 1000:9.1f} us  {name}")
            prev = float(t)
    lib.osd_timeline_enable(0)
    if a.check:
        out = pf()
        m = mm.compress_dim_conv
        worst = 0.0
        for f, s, o in zip(feats, supp, out):
            x = torch.cat((f[:1], s[:1].expand(-1, -1, f.shape[2], f.shape[3])), 1)
            ref = m(x)
            worst = max(worst, float((o[:1] - ref).abs().max()))
        print(f"max |err| vs torch fp32 module (episode 0): {worst:.4f}")


if __name__ == "__main__":
    main()
