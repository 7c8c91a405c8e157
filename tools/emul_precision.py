"""Design tool (not product): CPU emulation of the descriptor kernels' ROUNDING POINTS, to attribute the device's L2
error against the fp64 oracle and to choose where operand splits / storage formats pay.

  storage format of an activation tensor   'h'  fp16            'q'  unorm16 fixed point, step 6/65535 (ReLU6 range)
  depthwise -> MMA A operand               'h'  fp16            's'  fp16 hi + fp16 lo (two tiles)
  pointwise weights (B operand)            'h'  fp16            's'  fp16 hi + lo

python tools/emul_precision.py [model] [H W]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cerebro_b200.keras_weights import fold_model  # noqa: E402
from oracle import netvlad as NV  # noqa: E402
from tests import golden_io, synth  # noqa: E402


def r16(x):
    return x.half().float()


def split16(x):
    hi = r16(x)
    return hi, r16(x - hi)


def store(x, fmt):
    if fmt == "h":
        return r16(x)
    if fmt == "q":
        return torch.round(x * (65535.0 / 6.0)) * np.float32(6.0 / 65535.0)
    if fmt == "p":  # unorm15
        return torch.round(x * (32767.0 / 6.0)) * np.float32(6.0 / 32767.0)
    return x


def emulate(net, imgs, cfg):
    """cfg: dict(store=[fmt per layer output: stem, block1..], a=[per block], b=[per block])"""
    x = torch.as_tensor(imgs).float() - 128.0  # exact integers
    x = x.permute(0, 3, 1, 2)
    w1 = torch.as_tensor(net["conv1_w"]) * np.float32(2.0 / 255.0)
    hi, lo = split16(w1)
    k = (hi + lo).permute(3, 2, 0, 1)
    y = F.conv2d(F.pad(x, (0, 1, 0, 1)), k, stride=2) + torch.as_tensor(net["conv1_b"]).view(1, -1, 1, 1)
    y = store(torch.clamp(y, 0, 6), cfg["store"][0])
    for i, b in enumerate(net["blocks"]):
        dw = torch.as_tensor(b["dw_w"]).permute(2, 0, 1).unsqueeze(1)
        C = dw.shape[0]
        if b["stride"] == 2:
            y = F.conv2d(F.pad(y, (0, 1, 0, 1)), dw, stride=2, groups=C)
        else:
            y = F.conv2d(y, dw, stride=1, padding=1, groups=C)
        y = torch.clamp(y + torch.as_tensor(b["dw_b"]).view(1, -1, 1, 1), 0, 6)
        if b["pw_w"] is None:
            y = store(y, cfg["store"][i + 1])
            continue
        pw = torch.as_tensor(b["pw_w"])  # [C, Cout]
        a_mode, b_mode = cfg["a"][i], cfg["b"][i]
        if a_mode == "h":
            a_hi, a_lo = r16(y), None
        elif a_mode == "s":
            a_hi, a_lo = split16(y)
        else:
            a_hi, a_lo = y, None
        if b_mode == "h":
            b_hi, b_lo = r16(pw), None
        elif b_mode == "s":
            b_hi, b_lo = split16(pw)
        else:
            b_hi, b_lo = pw, None
        mm = lambda a, w: torch.einsum("nchw,cd->ndhw", a.double(), w.double()).float()
        z = mm(a_hi, b_hi)
        if a_lo is not None:
            z = z + mm(a_lo, b_hi)
        if b_lo is not None:
            z = z + mm(a_hi, b_lo)
        y = torch.clamp(z + torch.as_tensor(b["pw_b"]).view(1, -1, 1, 1), 0, 6)
        y = store(y, cfg["store"][i + 1])
    N, D, H, W = y.shape
    xf = y.permute(0, 2, 3, 1).reshape(N, H * W, D)
    s = xf @ torch.as_tensor(net["vlad_w"]) + torch.as_tensor(net["vlad_b"])
    a = torch.softmax(s, dim=-1)
    v = torch.einsum("npk,npd->ndk", a, xf) + torch.as_tensor(net["vlad_c"]).unsqueeze(0) * a.sum(1).unsqueeze(1)
    v = v.permute(0, 2, 1)
    v = v * torch.rsqrt(torch.clamp((v * v).sum(-1, keepdim=True), min=1e-12))
    v = v.reshape(N, -1)
    v = v * torch.rsqrt(torch.clamp((v * v).sum(-1, keepdim=True), min=1e-12))
    return v.numpy()


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "mobilenet_conv7"
    h, w = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (240, 320)
    raw = golden_io.raw_weights(model)
    net = fold_model(raw)
    c = net["conv1_w"].shape[2]
    imgs = synth.band_limited_images(2, h, w, c, seed=h + c)
    ref = NV.describe(imgs, raw, dtype="float64")
    nb = len(net["blocks"])

    def cfg(st, a, b):
        return {"store": list(st), "a": list(a), "b": list(b)}

    def upto(n, x, y):  # first n entries x, rest y
        return [x if i < n else y for i in range(nb + 1)]

    cases = {
        "all fp16 (round 1 device)": cfg("h" * (nb + 1), "h" * nb, "h" * nb),
        "exact everywhere (fp32)": cfg("x" * (nb + 1), "x" * nb, "x" * nb),
        "A,B split blocks 1-4": cfg("h" * (nb + 1), upto(4, "s", "h"), upto(4, "s", "h")),
        "A,B split blocks 1-4 + q16 store through block 4": cfg(upto(5, "q", "h"), upto(4, "s", "h"), upto(4, "s", "h")),
        "A,B split blocks 1-4 + q16 store through block 3": cfg(upto(4, "q", "h"), upto(4, "s", "h"), upto(4, "s", "h")),
        "A,B split all + q16 all but last": cfg(upto(nb, "q", "h"), "s" * nb, "s" * nb),
        "A,B split all + q16 all": cfg("q" * (nb + 1), "s" * nb, "s" * nb),
        "B split all, A split 1-4, q16 through 4": cfg(upto(5, "q", "h"), upto(4, "s", "h"), "s" * nb),
        "A split only blocks 1-4": cfg("h" * (nb + 1), upto(4, "s", "h"), "h" * nb),
        "q16 all, no splits": cfg("q" * (nb + 1), "h" * nb, "h" * nb),
        "A,B split 1-5 + q16 through 5": cfg(upto(6, "q", "h"), upto(5, "s", "h"), upto(5, "s", "h")),
        "A,B split 1-4 + q15 through block 4": cfg(upto(5, "p", "h"), upto(4, "s", "h"), upto(4, "s", "h")),
        "A,B split 1-4 + q15 all but last": cfg(upto(nb, "p", "h"), upto(4, "s", "h"), upto(4, "s", "h")),
        "A,B split all + q15 all but last": cfg(upto(nb, "p", "h"), "s" * nb, "s" * nb),
        "A,B split 1-5 + q15 all but last": cfg(upto(nb, "p", "h"), upto(5, "s", "h"), upto(5, "s", "h")),
        "A,B split 1-6 + q15 all but last": cfg(upto(nb, "p", "h"), upto(6, "s", "h"), upto(6, "s", "h")),
        "no split + q15 all but last": cfg(upto(nb, "p", "h"), "h"*nb, "h"*nb),
    }
    for name, cf in cases.items():
        d = emulate(net, imgs, cf)
        err = np.linalg.norm(d - ref, axis=1)
        print("%-55s L2 err %s" % (name, np.array2string(err, precision=5)))


if __name__ == "__main__":
    with torch.no_grad():
        main()
