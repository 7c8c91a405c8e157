"""ORACLE (test infrastructure, not product): CPU restatement of the reference's
whole-image descriptor forward pass.

PARITY UNPINNED: the reference holds no golden vectors for this path and Keras
2.2.4 / TensorFlow 1.11 cannot be installed here, so this restatement is checked
only against its own invariants (unit norm, per-cluster norm 1/sqrt(K), K-major
layout, fp32-vs-fp64 agreement).  See DESIGN.md.

Follows, line by line:
  * pre-processing            scripts/whole_image_desc_compute_server.py:629
                              ``(img.astype('float32') - 128.) * 2.0 / 255.``
  * backbone layer list       scripts/keras.models/model.json (MobileNet-v1 prefix):
                              ZeroPadding2D((0,1),(0,1)) + Conv2D 3x3 s2 'valid' bias-free
                              + BN(eps=1e-3) + ReLU(max 6); then DepthwiseConv2D 3x3
                              (s1 'same' | pad bottom/right + s2 'valid') + BN + ReLU6,
                              Conv2D 1x1 + BN + ReLU6 ...
  * NetVLADLayer.call         scripts/predict_utils.py:36-64
  * reply                     server.py:648  ``result.desc = u[0,:]``

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.  Input is the RAW Keras weight dict ("layer/weight" -> array), BN is
applied un-folded, so the product's BN folding is checked too.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3  # model.json: every BatchNormalization has epsilon 0.001
L2_EPS = 1e-12  # tf.nn.l2_normalize default epsilon: x * rsqrt(max(sum(x^2), eps))


def _bn(x, w, prefix, dt):
    g = torch.as_tensor(w[prefix + "/gamma"], dtype=dt)
    b = torch.as_tensor(w[prefix + "/beta"], dtype=dt)
    m = torch.as_tensor(w[prefix + "/moving_mean"], dtype=dt)
    v = torch.as_tensor(w[prefix + "/moving_variance"], dtype=dt)
    sh = (1, -1, 1, 1)
    return (x - m.view(sh)) / torch.sqrt(v.view(sh) + BN_EPS) * g.view(sh) + b.view(sh)


def _relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def preprocess(images_u8: np.ndarray, dt=torch.float32) -> torch.Tensor:
    """server.py:629 -- uint8 [N,H,W,C] -> float in [-1,1], NCHW."""
    x = torch.as_tensor(np.ascontiguousarray(images_u8)).to(dt)
    x = (x - 128.0) * 2.0 / 255.0
    return x.permute(0, 3, 1, 2).contiguous()


def backbone(x: torch.Tensor, w: dict, return_all: bool = False):
    """MobileNet-v1 prefix as listed in model.json. x: NCHW. Returns NCHW feature map."""
    dt = x.dtype
    acts = []
    # conv1_pad ((0,1),(0,1)) then 3x3 s2 valid
    k = torch.as_tensor(w["conv1/kernel"], dtype=dt).permute(3, 2, 0, 1)  # (kh,kw,Cin,Cout)->(Cout,Cin,kh,kw)
    x = F.pad(x, (0, 1, 0, 1))
    x = F.conv2d(x, k, stride=2)
    x = _relu6(_bn(x, w, "conv1_bn", dt))
    acts.append(x)
    i = 1
    while ("conv_dw_%d/depthwise_kernel" % i) in w:
        dk = torch.as_tensor(w["conv_dw_%d/depthwise_kernel" % i], dtype=dt)  # (3,3,C,1)
        C = dk.shape[2]
        dk = dk.permute(2, 3, 0, 1)  # (C,1,3,3)
        if i % 2 == 0:  # conv_pad_i ((0,1),(0,1)) + stride 2 valid
            x = F.pad(x, (0, 1, 0, 1))
            x = F.conv2d(x, dk, stride=2, groups=C)
        else:  # 'same', stride 1
            x = F.conv2d(x, dk, stride=1, padding=1, groups=C)
        x = _relu6(_bn(x, w, "conv_dw_%d_bn" % i, dt))
        acts.append(x)
        if ("conv_pw_%d/kernel" % i) in w:
            pk = torch.as_tensor(w["conv_pw_%d/kernel" % i], dtype=dt).permute(3, 2, 0, 1)
            x = F.conv2d(x, pk)
            x = _relu6(_bn(x, w, "conv_pw_%d_bn" % i, dt))
            acts.append(x)
        i += 1
    return (x, acts) if return_all else x


def backbone_v2(x: torch.Tensor, w: dict, return_all: bool = False):
    """MobileNetV2 prefix of the June2019 models (layer list = the ``model_config`` embedded in
    scripts/keras.models/June2019/*mobilenetv2-block_9_add*/modelarch_and_weights.*.h5, cut at ``block_<n>_add``):
    Conv1_pad ((0,1),(0,1)) + Conv1 3x3 s2 'valid' + bn_Conv1 + ReLU6; expanded_conv (depthwise 'same' + BN + ReLU6,
    project 1x1 + BN, linear); block_i: expand 1x1 + BN + ReLU6, [block_i_pad ((0,1),(0,1)) + depthwise s2 'valid' |
    depthwise s1 'same'] + BN + ReLU6, project 1x1 + BN (linear), Add with the block input when stride 1 and the channel
    counts agree (keras_applications MobileNetV2 _inverted_res_block; strides (1,2,1,2,1,1,2,1,1,1,...) for blocks 0..).
    x: NCHW.  Returns the NCHW feature map."""
    dt = x.dtype
    acts = []

    def conv(x, name, stride=1):
        k = torch.as_tensor(w[name + "/kernel"], dtype=dt).permute(3, 2, 0, 1)
        return F.conv2d(x, k, stride=stride)

    def dwise(x, name, stride):
        dk = torch.as_tensor(w[name + "/depthwise_kernel"], dtype=dt)
        C = dk.shape[2]
        dk = dk.permute(2, 3, 0, 1)
        if stride == 2:
            return F.conv2d(F.pad(x, (0, 1, 0, 1)), dk, stride=2, groups=C)
        return F.conv2d(x, dk, stride=1, padding=1, groups=C)

    x = F.pad(x, (0, 1, 0, 1))
    x = _relu6(_bn(conv(x, "Conv1", 2), w, "bn_Conv1", dt))
    acts.append(x)
    x = _relu6(_bn(dwise(x, "expanded_conv_depthwise", 1), w, "expanded_conv_depthwise_BN", dt))
    acts.append(x)
    x = _bn(conv(x, "expanded_conv_project"), w, "expanded_conv_project_BN", dt)
    acts.append(x)
    i = 1
    while ("block_%d_expand/kernel" % i) in w:
        pre = "block_%d_" % i
        stride = MOBILENETV2_STRIDES[i]
        inp = x
        x = _relu6(_bn(conv(x, pre + "expand"), w, pre + "expand_BN", dt))
        acts.append(x)
        x = _relu6(_bn(dwise(x, pre + "depthwise", stride), w, pre + "depthwise_BN", dt))
        acts.append(x)
        x = _bn(conv(x, pre + "project"), w, pre + "project_BN", dt)
        if stride == 1 and inp.shape[1] == x.shape[1]:
            x = inp + x
        acts.append(x)
        i += 1
    return (x, acts) if return_all else x


# depthwise stride of inverted-residual block i (0 = expanded_conv) in keras_applications' MobileNetV2
MOBILENETV2_STRIDES = [1, 2, 1, 2, 1, 1, 2, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1]


def is_mobilenetv2(w: dict) -> bool:
    return "Conv1/kernel" in w and "expanded_conv_depthwise/depthwise_kernel" in w


def netvlad(x: torch.Tensor, w: dict, num_ghost_clusters: int = 0) -> torch.Tensor:
    """NetVLADLayer.call, predict_utils.py:36-64.  x: NCHW feature map -> [N, K*D], index k*D+d.
    ``num_ghost_clusters`` > 0 restates GhostVLADLayer.call (predict_utils.py:110-141): the weights carry K + ghosts
    clusters, all take part in the softmax, and ``v = v[:, 0:num_clusters, :]`` (:133) drops the ghosts before the norms."""
    dt = x.dtype
    name = [k.split("/")[0] for k in w if k.endswith("/cluster_centers")][0]
    kern = torch.as_tensor(w[name + "/kernel"], dtype=dt)[0, 0]  # (D,K)
    bias = torch.as_tensor(w[name + "/bias"], dtype=dt).reshape(-1)  # (K,)
    cent = torch.as_tensor(w[name + "/cluster_centers"], dtype=dt)[0, 0, 0]  # (D,K)
    N, D, H, W = x.shape
    xf = x.permute(0, 2, 3, 1).reshape(N, H * W, D)  # NHWC pixels
    s = xf @ kern + bias  # :38   K.conv2d(x, kernel) + bias
    a = torch.softmax(s, dim=-1)  # :39
    # :47-52  v[d,k] = sum_hw a[hw,k] * (x[hw,d] + C[d,k])    (PLUS, as written at :47)
    v = torch.einsum("npk,npd->ndk", a, xf) + cent.unsqueeze(0) * a.sum(dim=1).unsqueeze(1)
    v = v.permute(0, 2, 1)  # :54  -> N x K x D
    if num_ghost_clusters:
        v = v[:, : v.shape[1] - num_ghost_clusters, :]  # :133
    v = v * torch.rsqrt(torch.clamp((v * v).sum(-1, keepdim=True), min=L2_EPS))  # :59
    v = v.reshape(N, -1)  # :60 batch_flatten (K-major)
    v = v * torch.rsqrt(torch.clamp((v * v).sum(-1, keepdim=True), min=L2_EPS))  # :61
    return v


def describe(images_u8: np.ndarray, w: dict, dtype: str = "float32", threads: int | None = None, num_ghost_clusters: int = 0) -> np.ndarray:
    """uint8 [N,H,W,C] (or [N,H,W]) -> descriptors [N, K*D] as float32/float64 numpy."""
    if images_u8.ndim == 3:
        images_u8 = images_u8[..., None]  # server.py:603-605
    dt = torch.float32 if dtype == "float32" else torch.float64
    if threads is not None:
        torch.set_num_threads(threads)
    with torch.no_grad():
        x = preprocess(images_u8, dt)
        f = backbone_v2(x, w) if is_mobilenetv2(w) else backbone(x, w)
        d = netvlad(f, w, num_ghost_clusters)
    return d.numpy()
