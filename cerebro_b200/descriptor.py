"""Host-side mirror of the reference's whole-image descriptor server.

``HDF5ModelImageDescriptor`` keeps the constructor arguments and the ``handle_req`` contract of
the class of the same name in scripts/whole_image_desc_compute_server.py:485-650 (the one the
reference selects at :734): it is built from a ``kerasmodel_file`` path plus the expected image
shape, asserts the shape of every request, and answers with ``desc`` (float64 list semantics of
``float64[] desc``, srv/WholeImageDescriptorCompute.srv:4) and ``model_type``.  The forward pass
itself runs in ``libcerebro_b200.so`` (no Keras/TensorFlow, no CPU fallback).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib, keras_weights
from ._lib import IrBlock, NetvladV2Weights, NetvladWeights, check, ptr


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class NetvladDescriptor:
    """Thin owner of a ``cb_descriptor`` handle.  ``net`` is the folded-weight dict produced by
    ``keras_weights.load_model`` / ``fold_mobilenet_netvlad`` / ``random_mobilenet_netvlad``."""

    def __init__(self, net: dict, rows: int, cols: int, chnls: int, max_batch: int = 1, device: int = 0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.rows, self.cols, self.chnls, self.max_batch = rows, cols, chnls, max_batch
        keep = []  # keep numpy arrays alive until create returns

        def f32(a):
            a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
            return a

        if net.get("arch") == "mobilenetv2":  # June2019 models: inverted-residual blocks
            nb = len(net["ir_blocks"])
            blocks = (IrBlock * nb)()
            c_in = 32
            for i, b in enumerate(net["ir_blocks"]):
                blocks[i].c_in = c_in
                blocks[i].c_exp = int(b["dw_w"].shape[2])
                blocks[i].c_out = int(b["project_w"].shape[1])
                blocks[i].stride = int(b["stride"])
                blocks[i].residual = int(b["residual"])
                if b["expand_w"] is not None:
                    blocks[i].expand_w = _fptr(f32(b["expand_w"]))
                    blocks[i].expand_b = _fptr(f32(b["expand_b"]))
                blocks[i].dw_w = _fptr(f32(b["dw_w"]))
                blocks[i].dw_b = _fptr(f32(b["dw_b"]))
                blocks[i].project_w = _fptr(f32(b["project_w"]))
                blocks[i].project_b = _fptr(f32(b["project_b"]))
                c_in = blocks[i].c_out
            w2 = NetvladV2Weights()
            w2.in_channels = int(net["conv1_w"].shape[2])
            w2.conv1_w = _fptr(f32(net["conv1_w"]))
            w2.conv1_b = _fptr(f32(net["conv1_b"]))
            w2.n_blocks = nb
            w2.blocks = blocks
            w2.vlad_k = int(net["vlad_w"].shape[1])
            w2.vlad_d = int(net["vlad_w"].shape[0])
            w2.vlad_w = _fptr(f32(net["vlad_w"]))
            w2.vlad_b = _fptr(f32(net["vlad_b"]))
            w2.vlad_c = _fptr(f32(net["vlad_c"]))
            w2.vlad_ghost = int(net.get("vlad_ghost", 0))
            check(self._lib.cb_descriptor_create_v2(C.byref(self._h), C.byref(w2), rows, cols, chnls, max_batch, device))
            self.dim = int(self._lib.cb_descriptor_dim(self._h))
            return
        nb = len(net["blocks"])

        w = NetvladWeights()
        w.in_channels = int(net["conv1_w"].shape[2])
        w.n_blocks = nb
        w.conv1_w = _fptr(f32(net["conv1_w"]))
        w.conv1_b = _fptr(f32(net["conv1_b"]))
        PF = C.POINTER(C.c_float)
        dw_w, dw_b, pw_w, pw_b = (PF * nb)(), (PF * nb)(), (PF * nb)(), (PF * nb)()
        strides = (C.c_int * nb)()
        couts = (C.c_int * nb)()
        for i, b in enumerate(net["blocks"]):
            dw_w[i] = _fptr(f32(b["dw_w"]))
            dw_b[i] = _fptr(f32(b["dw_b"]))
            if b["pw_w"] is not None:
                pw_w[i] = _fptr(f32(b["pw_w"]))
                pw_b[i] = _fptr(f32(b["pw_b"]))
                couts[i] = int(b["pw_w"].shape[1])
            else:
                pw_w[i] = PF()
                pw_b[i] = PF()
                couts[i] = int(b["dw_w"].shape[2])
            strides[i] = int(b["stride"])
        w.dw_w, w.dw_b, w.pw_w, w.pw_b = dw_w, dw_b, pw_w, pw_b
        w.dw_stride, w.channels_out = strides, couts
        w.vlad_k = int(net["vlad_w"].shape[1])
        w.vlad_d = int(net["vlad_w"].shape[0])
        w.vlad_w = _fptr(f32(net["vlad_w"]))
        w.vlad_b = _fptr(f32(net["vlad_b"]))
        w.vlad_c = _fptr(f32(net["vlad_c"]))
        w.vlad_ghost = int(net.get("vlad_ghost", 0))  # GhostVLADLayer: trailing clusters dropped before the norms
        check(self._lib.cb_descriptor_create(C.byref(self._h), C.byref(w), rows, cols, chnls, max_batch, device))
        self.dim = int(self._lib.cb_descriptor_dim(self._h))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.cb_descriptor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compute(self, images_u8: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """uint8 [n, rows, cols, chnls] (or [n, rows, cols] for 1 channel) -> float32 [n, dim].  ``out`` may be a
        caller-owned (e.g. pinned) float32 [n, dim] array."""
        if images_u8.ndim == 3:
            images_u8 = images_u8[..., None]
        assert images_u8.dtype == np.uint8
        n = images_u8.shape[0]
        assert images_u8.shape[1:] == (self.rows, self.cols, self.chnls), images_u8.shape
        images_u8 = np.ascontiguousarray(images_u8)
        if out is None:
            out = np.empty((n, self.dim), dtype=np.float32)
        assert out.dtype == np.float32 and out.shape == (n, self.dim) and out.flags["C_CONTIGUOUS"]
        check(self._lib.cb_descriptor_compute(self._h, n, ptr(images_u8), 0, ptr(out)))
        return out

    def compute_f64(self, images_u8: np.ndarray, row_stride_bytes: int = 0) -> np.ndarray:
        """``cb_descriptor_compute_f64``: the service reply's ``float64[] desc`` written directly (srv:4)."""
        if images_u8.ndim == 3:
            images_u8 = images_u8[..., None]
        assert images_u8.dtype == np.uint8
        n = images_u8.shape[0]
        images_u8 = np.ascontiguousarray(images_u8)
        out = np.empty((n, self.dim), dtype=np.float64)
        check(self._lib.cb_descriptor_compute_f64(self._h, n, ptr(images_u8), row_stride_bytes, ptr(out)))
        return out

    def compute_device(self, images_u8, out=None):
        """CUDA uint8 tensor [n, rows, cols, chnls] -> CUDA float32 [n, dim], async on torch's stream."""
        import torch

        n = images_u8.shape[0]
        assert images_u8.is_cuda and images_u8.dtype == torch.uint8 and images_u8.is_contiguous()
        if out is None:
            out = torch.empty((n, self.dim), dtype=torch.float32, device=images_u8.device)
        check(self._lib.cb_descriptor_compute_device(self._h, n, ptr(images_u8), ptr(out), _lib.current_stream_ptr()))
        return out

    def get_activation(self, layer: int) -> np.ndarray:
        buf = np.empty(self.rows * self.cols * 64, dtype=np.float32)
        n = self._lib.cb_descriptor_get_activation(self._h, layer, ptr(buf), buf.size)
        if n < 0:
            check(int(n))
        return buf[:n].copy()


class WholeImageDescriptorComputeResponse:
    """srv/WholeImageDescriptorCompute.srv:4-5"""

    def __init__(self):
        self.desc = []
        self.model_type = ""


class HDF5ModelImageDescriptor:
    """Same constructor and ``handle_req`` semantics as the reference class (server.py:485-650).

    ``req`` needs ``.ima`` -- here a numpy uint8 array [rows, cols] or [rows, cols, chnls] standing in
    for ``CvBridge().imgmsg_to_cv2(req.ima)`` (server.py:601) -- and optionally ``.a`` (ignored,
    the reference client always sends 986, src/Cerebro.cpp:258)."""

    def __init__(self, kerasmodel_file, im_rows=600, im_cols=960, im_chnls=3, device: int = 0, max_batch: int = 1):
        self.im_rows, self.im_cols, self.im_chnls = int(im_rows), int(im_cols), int(im_chnls)
        assert os.path.isfile(kerasmodel_file), (
            "The model weights file doesnot exists or there is a permission issue." + "kerasmodel_file=" + kerasmodel_file
        )  # server.py:552
        log_dir = "/".join(kerasmodel_file.split("/")[0:-1])
        self.model_type = log_dir.split("/")[-1]  # server.py:532
        net = keras_weights.load_model(kerasmodel_file)
        self.model = NetvladDescriptor(net, self.im_rows, self.im_cols, self.im_chnls, max_batch=max_batch, device=device)
        self.request_count = 0
        # server.py:580-581: a zeros image is pushed through once at start-up
        self.model.compute(np.zeros((1, self.im_rows, self.im_cols, self.im_chnls), dtype=np.uint8) + 128)

    def handle_req(self, req):
        cv_image = np.asarray(req.ima)
        if cv_image.ndim == 2:  # server.py:603-605
            cv_image = np.expand_dims(cv_image, -1)
        elif cv_image.ndim != 3:
            assert False
        assert (
            cv_image.shape[0] == self.im_rows and cv_image.shape[1] == self.im_cols and cv_image.shape[2] == self.im_chnls
        ), "\n[whole_image_descriptor_compute_server] Input shape of the image \
                does not match with the allocated GPU memory. Expecting an input image of \
                size %dx%dx%d, but received : %s" % (self.im_rows, self.im_cols, self.im_chnls, str(cv_image.shape))  # :614-619
        u = self.model.compute_f64(cv_image.astype(np.uint8)[None])  # float64[] desc (srv:4) straight from the C ABI
        self.request_count += 1
        result = WholeImageDescriptorComputeResponse()
        result.desc = u[0, :]
        result.model_type = self.model_type
        return result
