"""Where does the bf16 tier's deviation at 640x640 come from?  (gap 7.)

CPU-only emulation of the bf16 tier's rounding points on the oracle's DBNet-ResNet18 (seed 0, randomised BN, the frames
of tests/test_gpu_parity.py::test_config1_640x640_maps_and_boxes_vs_oracle): eval-mode BN folded into the conv, folded
weights rounded to bf16, fp32 accumulation, every stored activation rounded to bf16 (residuals are read back as bf16),
ConvT1's output and ConvT2 kept in fp32 as the fused head tail does.  Each rounding group can be switched off, which
gives (a) the error with only that group rounded and (b) the error with everything but that group rounded.
Test infrastructure / analysis only: imports oracle/, never the product.

  python profiles/bf16_error_budget.py > profiles/r01_bf16_error_budget.md
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port  # noqa: E402

GROUPS = ["input", "stem", "layer1", "layer2", "layer3", "layer4", "lateral", "fpn_out", "head_feat"]
WGROUPS = ["w_stem", "w_layer1", "w_layer2", "w_layer3", "w_layer4", "w_lateral", "w_fpn_out", "w_head3x3", "w_convT1"]


ROUND_TO = [torch.bfloat16]          # the storage type being emulated (fp16 for the what-if rows)
ABSMAX = [0.0]


def bf(x):
    ABSMAX[0] = max(ABSMAX[0], float(x.abs().max()))
    return x.to(ROUND_TO[0]).float()


def fold(conv, bn):
    w, b = conv.weight, conv.bias
    if bn is None:
        return w, b
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shape = (-1, 1, 1, 1) if isinstance(conv, nn.Conv2d) else (1, -1, 1, 1)
    b0 = b if b is not None else torch.zeros_like(bn.running_mean)
    return w * s.view(shape), bn.bias + (b0 - bn.running_mean) * s


class Emu:
    def __init__(self, net, on):
        self.net, self.on = net, set(on)

    def q(self, x, g):
        return bf(x) if g in self.on else x

    def conv(self, x, conv, bn, g):
        w, b = fold(conv, bn)
        if "w_" + g in self.on:
            w = bf(w)
        return F.conv2d(x, w, b, conv.stride, conv.padding)

    def block(self, x, blk, g):
        o = self.q(F.relu(self.conv(x, blk.conv1, blk.bn1, g)), g)
        o = self.conv(o, blk.conv2, blk.bn2, g)
        idn = x if blk.downsample is None else self.q(self.conv(x, blk.downsample[0], blk.downsample[1], g), g)
        return self.q(F.relu(o + idn), g)

    def forward(self, x):
        n = self.net
        b = n.backbone
        x = self.q(x, "input")
        x = self.q(F.relu(self.conv(x, b[0], b[1], "stem")), "stem")
        x = F.max_pool2d(x, 3, 2, 1)
        cs = []
        for li in range(4):
            for blk in b[4 + li]:
                x = self.block(x, blk, "layer%d" % (li + 1))
            cs.append(x)
        feats = cs[::-1]
        last = self.q(self.conv(feats[0], n.fpn.inner_blocks[0], None, "lateral"), "lateral")
        for i in range(1, 4):
            lat = self.conv(feats[i], n.fpn.inner_blocks[i], None, "lateral")
            last = self.q(lat + F.interpolate(last, scale_factor=2, mode="nearest"), "lateral")
        p2 = self.q(self.conv(last, n.fpn.layer_blocks[3], None, "fpn_out"), "fpn_out")
        outs = []
        for head in (n.head.probability_head, n.head.threshold_head):
            f = self.q(F.relu(self.conv(p2, head[0], head[1], "head3x3")), "head_feat")
            w, bb = fold(head[3], head[4])
            if "w_convT1" in self.on:
                w = bf(w)
            h = F.relu(F.conv_transpose2d(f, w, bb, stride=2))
            outs.append(torch.sigmoid(F.conv_transpose2d(h, head[6].weight, head[6].bias, stride=2)))
        return outs


def stats(got, want):
    e = (got - want).abs()
    return float(e.max()), float((e > 1e-2).float().mean())


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    net = port.build_dbnet("resnet18", seed=0)
    frames = port.synthetic_frames(2, 640, 640, seed=0)
    x = torch.cat([port.preprocess(f, 640, 640) for f in frames])
    with torch.no_grad():
        ref = Emu(net, []).forward(x)
        chk = port.dbnet_forward(net, x)
        print("# bf16 error budget, DBNet-ResNet18 seed 0 @640x640 (2 frames), CPU emulation of the tier's rounding points\n")
        print("folded fp32 emulation vs oracle forward: max |dprob| = %.2e (sanity)\n" %
              float((ref[0] - chk["probability"]).abs().max()))
        print("| rounded groups | prob max | prob >1e-2 | thr max | thr >1e-2 |")
        print("|---|---|---|---|---|")

        def row(name, on):
            p, t = Emu(net, on).forward(x)
            a, b = stats(p, ref[0]), stats(t, ref[1])
            print("| %s | %.4f | %.3f %% | %.4f | %.3f %% |" % (name, a[0], 100 * a[1], b[0], 100 * b[1]), flush=True)

        allg = GROUPS + WGROUPS
        row("all (the tier as built)", allg)
        row("activations only", GROUPS)
        row("weights only", WGROUPS)
        for g in GROUPS + WGROUPS:
            row("only " + g, [g])
        for g in GROUPS + WGROUPS:
            row("all but " + g, [k for k in allg if k != g])
        tail = ["fpn_out", "head_feat", "w_fpn_out", "w_head3x3", "w_convT1"]
        row("all but the last three layers (fpn_out, head 3x3, ConvT1: acts + weights)", [k for k in allg if k not in tail])
        row("all but lateral+fpn_out+head_feat activations", [k for k in allg if k not in ("lateral", "fpn_out", "head_feat")])
        # what-if: the same rounding points with IEEE half (10 mantissa bits; tcgen05 kind::f16 takes it at the same rate)
        ROUND_TO[0] = torch.float16
        ABSMAX[0] = 0.0
        row("WHAT-IF fp16 storage, all groups", allg)
        print("\nlargest |value| rounded in the fp16 run (activations and folded weights): %.1f (fp16 max 65504)" % ABSMAX[0])


if __name__ == "__main__":
    main()
