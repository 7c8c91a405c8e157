"""Reader for the reference's shipped Keras 2.2.4 weight files, without h5py.

The reference descriptor server is started with a ``kerasmodel_file`` path
(reference: scripts/whole_image_desc_compute_server.py:485-553,
``keras.models.load_model(kerasmodel_file, custom_objects=...)``).  To stay a
drop-in for that argument this module parses the same files directly.

The shipped files are HDF5 superblock-v0 / symbol-table groups / version-1
object headers with contiguous little-endian float32 datasets (SURVEY.md
section 8c).  Only that subset of HDF5 is understood; anything else raises
``KerasWeightsError`` loudly rather than guessing.

Also defines the repo's own flat container (``.cbw``: JSON header + raw
float32 payload) so that converted weights can travel to machines where the
reference tree is absent.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class KerasWeightsError(RuntimeError):
    pass


# ----------------------------------------------------------------------------
# minimal HDF5 walk
# ----------------------------------------------------------------------------
class _H5:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise KerasWeightsError("not an HDF5 file")
        if buf[8] != 0:
            raise KerasWeightsError("only superblock v0 supported, got v%d" % buf[8])
        if buf[13] != 8 or buf[14] != 8:
            raise KerasWeightsError("only 8-byte offsets/lengths supported")
        # root symbol table entry lives at byte 56 (v0 superblock, 8-byte offsets)
        self.root_header = struct.unpack_from("<Q", buf, 56 + 8)[0]

    # -- object header (version 1) -> list of (type, bytes)
    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self.b
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise KerasWeightsError("object header v%d unsupported" % ver)
        out: List[Tuple[int, bytes]] = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8 : p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((caddr, clen))
                out.append((mtype, body))
        return out

    def _heap_data(self, heap_addr: int) -> int:
        if self.b[heap_addr : heap_addr + 4] != b"HEAP":
            raise KerasWeightsError("bad local heap signature")
        return struct.unpack_from("<Q", self.b, heap_addr + 8 + 16)[0]

    def _name(self, heap_data: int, off: int) -> str:
        e = self.b.index(b"\x00", heap_data + off)
        return self.b[heap_data + off : e].decode("utf8")

    def _walk_btree(self, addr: int, heap_data: int, out: Dict[str, int]):
        b = self.b
        if b[addr : addr + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", b, addr + 6)[0]
            p = addr + 8
            for _ in range(nsym):
                name_off, obj = struct.unpack_from("<QQ", b, p)
                out[self._name(heap_data, name_off)] = obj
                p += 40
            return
        if b[addr : addr + 4] != b"TREE":
            raise KerasWeightsError("bad b-tree node signature")
        ntype, _level, nent = struct.unpack_from("<BBH", b, addr + 4)
        if ntype != 0:
            raise KerasWeightsError("unexpected b-tree type %d" % ntype)
        p = addr + 8 + 16  # skip siblings
        p += 8  # key 0
        for _ in range(nent):
            child = struct.unpack_from("<Q", b, p)[0]
            p += 16  # child + next key
            self._walk_btree(child, heap_data, out)

    def children(self, header_addr: int) -> Dict[str, int]:
        """name -> object header address, for a group; None for a dataset."""
        for mtype, body in self.messages(header_addr):
            if mtype == 0x11:  # symbol table message
                btree, heap = struct.unpack_from("<QQ", body, 0)
                out: Dict[str, int] = {}
                self._walk_btree(btree, self._heap_data(heap), out)
                return out
        return None

    def dataset(self, header_addr: int):
        shape = None
        dtype = None
        data_addr = None
        data_size = None
        for mtype, body in self.messages(header_addr):
            if mtype == 0x01:  # dataspace
                ver, rank, flags = struct.unpack_from("<BBB", body, 0)
                if ver == 1:
                    off = 8
                elif ver == 2:
                    off = 4
                else:
                    raise KerasWeightsError("dataspace v%d" % ver)
                shape = struct.unpack_from("<%dQ" % rank, body, off)
            elif mtype == 0x03:  # datatype
                cls = body[0] & 0x0F
                size = struct.unpack_from("<I", body, 4)[0]
                bits0 = body[1]
                if cls == 1 and size == 4 and (bits0 & 1) == 0:
                    dtype = "<f4"
                elif cls == 1 and size == 8 and (bits0 & 1) == 0:
                    dtype = "<f8"
                else:
                    dtype = ("other", cls, size)
            elif mtype == 0x08:  # layout
                ver = body[0]
                if ver != 3:
                    raise KerasWeightsError("layout v%d unsupported" % ver)
                lclass = body[1]
                if lclass == 1:
                    data_addr, data_size = struct.unpack_from("<QQ", body, 2)
                elif lclass == 0:  # compact
                    sz = struct.unpack_from("<H", body, 2)[0]
                    data_addr, data_size = ("compact", body[4 : 4 + sz]), sz
                else:
                    raise KerasWeightsError("chunked dataset unsupported")
        return shape, dtype, data_addr, data_size

    def read_array(self, header_addr: int) -> np.ndarray:
        shape, dtype, addr, size = self.dataset(header_addr)
        if shape is None or not isinstance(dtype, str):
            raise KerasWeightsError("dataset is not a plain float array")
        n = int(np.prod(shape)) if len(shape) else 1
        if isinstance(addr, tuple):
            raw = addr[1]
        else:
            if addr == UNDEF:
                raise KerasWeightsError("dataset has no storage")
            raw = self.b[addr : addr + size]
        arr = np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)
        return np.array(arr, dtype=np.float32)

    def attribute_string(self, header_addr: int, name: str):
        """Return a scalar string attribute (fixed-length only) or None."""
        for mtype, body in self.messages(header_addr):
            if mtype != 0x0C:
                continue
            ver = body[0]
            if ver != 1:
                continue
            nsz, dtsz, dssz = struct.unpack_from("<HHH", body, 2)
            pad = lambda x: (x + 7) & ~7
            p = 8
            aname = body[p : p + nsz].split(b"\x00")[0].decode("utf8")
            p += pad(nsz)
            dt = body[p : p + dtsz]
            p += pad(dtsz)
            p += pad(dssz)
            if aname != name:
                continue
            cls = dt[0] & 0x0F
            if cls == 3:  # fixed string
                size = struct.unpack_from("<I", dt, 4)[0]
                return body[p : p + size].split(b"\x00")[0].decode("utf8")
            return None
        return None


def load_keras_file(path: str) -> Dict[str, np.ndarray]:
    """Return {"<layer>/<weight>": float32 array} for a shipped Keras file.

    Dataset paths in the file are ``/model_weights/<layer>/<layer>/<weight>:0``
    (full-model files) or ``/<layer>/<layer>/<weight>:0`` (weights-only files).
    """
    with open(path, "rb") as f:
        h5 = _H5(f.read())
    root = h5.children(h5.root_header)
    top = h5.children(root["model_weights"]) if "model_weights" in root else root
    out: Dict[str, np.ndarray] = {}

    def rec(addr: int, prefix: List[str]):
        kids = h5.children(addr)
        if kids is None:
            leaf = prefix[-1]
            if leaf.endswith(":0"):
                leaf = leaf[:-2]
            # prefix = [layer, layer, weight:0] -> "layer/weight"
            out[prefix[0] + "/" + leaf] = h5.read_array(addr)
            return
        for k, a in kids.items():
            rec(a, prefix + [k])

    for lname, laddr in top.items():
        rec(laddr, [lname])
    if not out:
        raise KerasWeightsError("no datasets found in %s" % path)
    return out


# ----------------------------------------------------------------------------
# architecture recovery from weight names/shapes (MobileNet-v1 prefix + NetVLAD)
# ----------------------------------------------------------------------------
BN_EPS = 1e-3  # reference: scripts/keras.models/model.json, every BatchNormalization epsilon


def fold_mobilenet_netvlad(w: Dict[str, np.ndarray]) -> dict:
    """Fold BN into the bias-free convs and describe the network.

    Follows the layer list in scripts/keras.models/model.json: conv1 (3x3 s2,
    pad bottom/right) then blocks i=1.. of depthwise 3x3 (stride 2 for even i,
    with bottom/right zero pad + 'valid'; stride 1 'same' for odd i) and
    pointwise 1x1, each followed by BN(eps 1e-3) and ReLU6; then NetVLADLayer.
    Returns float32 arrays:
      conv1_w (3,3,Cin,32), conv1_b (32,)
      blocks: list of dict(dw_w (3,3,C), dw_b (C,), pw_w (C,Cout), pw_b (Cout,), stride);
              pw_w/pw_b are None when the model is cut after a depthwise layer
      vlad_w (D,K), vlad_b (K,), vlad_c (D,K)
    """

    def bn(prefix):
        g = w[prefix + "/gamma"].astype(np.float64)
        b = w[prefix + "/beta"].astype(np.float64)
        m = w[prefix + "/moving_mean"].astype(np.float64)
        v = w[prefix + "/moving_variance"].astype(np.float64)
        s = g / np.sqrt(v + BN_EPS)
        return s, b - m * s

    if "conv1/kernel" not in w:
        raise KerasWeightsError("not a MobileNet-v1 style model (no conv1/kernel)")
    s, o = bn("conv1_bn")
    net = {
        "conv1_w": (w["conv1/kernel"].astype(np.float64) * s).astype(np.float32),
        "conv1_b": o.astype(np.float32),
        "blocks": [],
    }
    i = 1
    while ("conv_dw_%d/depthwise_kernel" % i) in w:
        dk = w["conv_dw_%d/depthwise_kernel" % i].astype(np.float64)[:, :, :, 0]
        s, o = bn("conv_dw_%d_bn" % i)
        blk = {"dw_w": (dk * s).astype(np.float32), "dw_b": o.astype(np.float32)}
        blk["stride"] = 2 if i % 2 == 0 else 1
        if ("conv_pw_%d/kernel" % i) in w:
            pk = w["conv_pw_%d/kernel" % i].astype(np.float64)[0, 0]
            s, o = bn("conv_pw_%d_bn" % i)
            blk["pw_w"] = (pk * s).astype(np.float32)
            blk["pw_b"] = o.astype(np.float32)
        else:
            # e.g. Apr2019/gray_conv6_K16: the backbone is cut after conv_dw_6_relu
            blk["pw_w"] = None
            blk["pw_b"] = None
        net["blocks"].append(blk)
        i += 1
    vl = [k.split("/")[0] for k in w if k.endswith("/cluster_centers")]
    if len(vl) != 1:
        raise KerasWeightsError("expected exactly one NetVLAD layer, found %r" % vl)
    v = vl[0]
    net["vlad_w"] = np.ascontiguousarray(w[v + "/kernel"][0, 0]).astype(np.float32)
    net["vlad_b"] = np.ascontiguousarray(w[v + "/bias"].reshape(-1)).astype(np.float32)
    net["vlad_c"] = np.ascontiguousarray(w[v + "/cluster_centers"][0, 0, 0]).astype(np.float32)
    last = net["blocks"][-1]
    d_feat = last["pw_w"].shape[1] if last["pw_w"] is not None else last["dw_w"].shape[2]
    if net["vlad_w"].shape[0] != d_feat:
        raise KerasWeightsError("NetVLAD input dim %d != backbone output %d" % (net["vlad_w"].shape[0], d_feat))
    return net


# depthwise stride of inverted-residual block i (0 = expanded_conv) in keras_applications' MobileNetV2, the builder the
# June2019 models were made with (keras_helpers.py / the model_config embedded in each .h5)
MOBILENETV2_STRIDES = [1, 2, 1, 2, 1, 1, 2, 1, 1, 1, 1, 1, 1, 2, 1, 1, 1]


def is_mobilenetv2(w: Dict[str, np.ndarray]) -> bool:
    return "Conv1/kernel" in w and "expanded_conv_depthwise/depthwise_kernel" in w


def fold_mobilenetv2_netvlad(w: Dict[str, np.ndarray]) -> dict:
    """Fold BN and describe a MobileNetV2-prefix + NetVLAD model (the June2019 ``mobilenetv2-block_9_add`` files the
    mynteye launch files select, launch/mynteye_vinsfusion.launch:100).  Layer list = the embedded ``model_config``:
    Conv1 (3x3 s2, pad bottom/right) + bn_Conv1 + ReLU6, then inverted-residual blocks: [expand 1x1 + BN + ReLU6],
    depthwise 3x3 (s1 'same' | pad bottom/right + s2 'valid') + BN + ReLU6, project 1x1 + BN (linear), Add with the block
    input when stride 1 and the channel counts agree.  Returns
      arch 'mobilenetv2', conv1_w (3,3,Cin,32), conv1_b (32,),
      ir_blocks: list of dict(expand_w (Cin,Cexp)|None, expand_b, dw_w (3,3,Cexp), dw_b, project_w (Cexp,Cout), project_b,
                              stride, residual), vlad_w (D,K), vlad_b (K,), vlad_c (D,K)."""

    def bn(prefix):
        g = w[prefix + "/gamma"].astype(np.float64)
        b = w[prefix + "/beta"].astype(np.float64)
        m = w[prefix + "/moving_mean"].astype(np.float64)
        v = w[prefix + "/moving_variance"].astype(np.float64)
        sc = g / np.sqrt(v + BN_EPS)
        return sc, b - m * sc

    def pw(name):
        k = w[name + "/kernel"].astype(np.float64)[0, 0]
        sc, o = bn(name + "_BN")
        return (k * sc).astype(np.float32), o.astype(np.float32)

    def dw(name):
        k = w[name + "/depthwise_kernel"].astype(np.float64)[:, :, :, 0]
        sc, o = bn(name + "_BN")
        return (k * sc).astype(np.float32), o.astype(np.float32)

    sc, o = bn("bn_Conv1")
    net = {
        "arch": "mobilenetv2",
        "conv1_w": (w["Conv1/kernel"].astype(np.float64) * sc).astype(np.float32),
        "conv1_b": o.astype(np.float32),
        "ir_blocks": [],
    }
    names = ["expanded_conv"]
    i = 1
    while ("block_%d_depthwise/depthwise_kernel" % i) in w:
        names.append("block_%d" % i)
        i += 1
    c_in = 32
    for i, nm in enumerate(names):
        blk = {"stride": MOBILENETV2_STRIDES[i]}
        if (nm + "_expand/kernel") in w:
            blk["expand_w"], blk["expand_b"] = pw(nm + "_expand")
        else:
            blk["expand_w"], blk["expand_b"] = None, None
        blk["dw_w"], blk["dw_b"] = dw(nm + "_depthwise")
        blk["project_w"], blk["project_b"] = pw(nm + "_project")
        c_out = blk["project_w"].shape[1]
        blk["residual"] = int(blk["stride"] == 1 and c_in == c_out and blk["expand_w"] is not None)
        net["ir_blocks"].append(blk)
        c_in = c_out
    vl = [k.split("/")[0] for k in w if k.endswith("/cluster_centers")]
    if len(vl) != 1:
        raise KerasWeightsError("expected exactly one NetVLAD layer, found %r" % vl)
    v = vl[0]
    net["vlad_w"] = np.ascontiguousarray(w[v + "/kernel"][0, 0]).astype(np.float32)
    net["vlad_b"] = np.ascontiguousarray(w[v + "/bias"].reshape(-1)).astype(np.float32)
    net["vlad_c"] = np.ascontiguousarray(w[v + "/cluster_centers"][0, 0, 0]).astype(np.float32)
    if net["vlad_w"].shape[0] != c_in:
        raise KerasWeightsError("NetVLAD input dim %d != backbone output %d" % (net["vlad_w"].shape[0], c_in))
    return net


def fold_model(w: Dict[str, np.ndarray]) -> dict:
    """BN-fold whichever shipped architecture the raw weight dict belongs to."""
    return fold_mobilenetv2_netvlad(w) if is_mobilenetv2(w) else fold_mobilenet_netvlad(w)


def random_mobilenet_netvlad(in_ch: int = 3, n_blocks: int = 7, K: int = 16, seed: int = 0) -> dict:
    """Random-init weights of the shipped architecture (for synthetic benches)."""
    rng = np.random.default_rng(seed)
    widths = [64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512, 512, 1024]
    c = 32
    net = {
        "conv1_w": (rng.standard_normal((3, 3, in_ch, 32)) * np.sqrt(2.0 / (9 * in_ch))).astype(np.float32),
        "conv1_b": (rng.standard_normal(32) * 0.1).astype(np.float32),
        "blocks": [],
    }
    for i in range(1, n_blocks + 1):
        co = widths[i - 1]
        net["blocks"].append(
            {
                "dw_w": (rng.standard_normal((3, 3, c)) * np.sqrt(2.0 / 9)).astype(np.float32),
                "dw_b": (rng.standard_normal(c) * 0.1).astype(np.float32),
                "pw_w": (rng.standard_normal((c, co)) * np.sqrt(2.0 / c)).astype(np.float32),
                "pw_b": (rng.standard_normal(co) * 0.1).astype(np.float32),
                "stride": 2 if i % 2 == 0 else 1,
            }
        )
        c = co
    net["vlad_w"] = (rng.standard_normal((c, K)) * 0.05).astype(np.float32)
    net["vlad_b"] = (rng.standard_normal(K) * 0.05).astype(np.float32)
    net["vlad_c"] = (rng.standard_normal((c, K)) * 0.05).astype(np.float32)
    return net


# ----------------------------------------------------------------------------
# flat container
# ----------------------------------------------------------------------------
_MAGIC = b"CBW1"


_V1_KEYS = ("dw_w", "dw_b", "pw_w", "pw_b")
_V2_KEYS = ("expand_w", "expand_b", "dw_w", "dw_b", "project_w", "project_b")


def _flatten(net: dict) -> List[Tuple[str, np.ndarray]]:
    items = [("conv1_w", net["conv1_w"]), ("conv1_b", net["conv1_b"])]
    v2 = net.get("arch") == "mobilenetv2"
    for i, b in enumerate(net["ir_blocks"] if v2 else net["blocks"]):
        for k in _V2_KEYS if v2 else _V1_KEYS:
            if b[k] is not None:
                items.append(("b%d_%s" % (i, k), b[k]))
    for k in ("vlad_w", "vlad_b", "vlad_c"):
        items.append((k, net[k]))
    return items


def save_cbw(path: str, net: dict, meta: dict | None = None) -> None:
    items = _flatten(net)
    v2 = net.get("arch") == "mobilenetv2"
    hdr = {
        "arch": "mobilenetv2" if v2 else "mobilenet",
        "strides": [b["stride"] for b in (net["ir_blocks"] if v2 else net["blocks"])],
        "residual": [b["residual"] for b in net["ir_blocks"]] if v2 else [],
        "arrays": [[n, list(a.shape)] for n, a in items],
        "vlad_ghost": int(net.get("vlad_ghost", 0)),  # GhostVLADLayer: trailing clusters dropped before the norms
        "meta": meta or {},
    }
    hj = json.dumps(hdr).encode("utf8")
    with open(path, "wb") as f:
        f.write(_MAGIC + struct.pack("<I", len(hj)) + hj)
        for _, a in items:
            f.write(np.ascontiguousarray(a, dtype="<f4").tobytes())


def load_cbw(path: str) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != _MAGIC:
        raise KerasWeightsError("not a CBW1 file: %s" % path)
    n = struct.unpack_from("<I", buf, 4)[0]
    hdr = json.loads(buf[8 : 8 + n].decode("utf8"))
    p = 8 + n
    arrs = {}
    for name, shape in hdr["arrays"]:
        cnt = int(np.prod(shape))
        arrs[name] = np.frombuffer(buf, dtype="<f4", count=cnt, offset=p).reshape(shape).copy()
        p += 4 * cnt
    if hdr.get("arch") == "mobilenetv2":
        net = {"arch": "mobilenetv2", "conv1_w": arrs["conv1_w"], "conv1_b": arrs["conv1_b"], "ir_blocks": []}
        for i, s in enumerate(hdr["strides"]):
            net["ir_blocks"].append({k: arrs.get("b%d_%s" % (i, k)) for k in _V2_KEYS})
            net["ir_blocks"][-1]["stride"] = s
            net["ir_blocks"][-1]["residual"] = hdr["residual"][i]
        for k in ("vlad_w", "vlad_b", "vlad_c"):
            net[k] = arrs[k]
        net["meta"] = hdr.get("meta", {})
        net["vlad_ghost"] = int(hdr.get("vlad_ghost", 0))
        return net
    net = {"conv1_w": arrs["conv1_w"], "conv1_b": arrs["conv1_b"], "blocks": []}
    for i, s in enumerate(hdr["strides"]):
        net["blocks"].append({k: arrs.get("b%d_%s" % (i, k)) for k in _V1_KEYS})
        net["blocks"][-1]["stride"] = s
    for k in ("vlad_w", "vlad_b", "vlad_c"):
        net[k] = arrs[k]
    net["meta"] = hdr.get("meta", {})
    net["vlad_ghost"] = int(hdr.get("vlad_ghost", 0))
    return net


def load_model(path: str) -> dict:
    """Accept either a reference Keras HDF5 file or a .cbw container."""
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:4] == _MAGIC:
        return load_cbw(path)
    return fold_model(load_keras_file(path))
