"""Error behaviour of the C ABI on a real device: bad arguments come back as negative status codes with a
message (no exceptions/aborts cross the boundary), and the handle stays usable afterwards."""
import ctypes as C

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def test_index_argument_validation(native_lib, cuda_device):
    from cerebro_b200._lib import CerebroB200Error
    from cerebro_b200.index import IndexFlatIP

    h = C.c_void_p()
    assert native_lib.cb_index_create(C.byref(h), 100, 10, 0, 0, 1) == -1  # dim not a multiple of 128
    assert b"multiple of 128" in native_lib.cb_last_error()
    assert native_lib.cb_index_create(C.byref(h), 128, 10, 99, 0, 1) == -1  # no such device
    assert native_lib.cb_index_create(C.byref(h), 128, 10, 0, 3, 2) == -1  # rank >= world
    ix = IndexFlatIP(128, capacity=10)
    db = synth.unit_rows(12, 128, seed=1)
    ix.add(db[:10])
    with pytest.raises(CerebroB200Error) as e:
        ix.add(db[10:])  # capacity exceeded
    assert e.value.code == -4 and ix.ntotal == 10
    with pytest.raises(CerebroB200Error):
        ix.search(db[:1], 33)  # k > 32
    with pytest.raises(CerebroB200Error):
        ix.search(db[:1], 0)
    D, I = ix.search(db[:2], 3)  # still works
    assert list(I[:, 0]) == [0, 1]
    ix.reset()
    assert ix.ntotal == 0
    D, I = ix.search(db[:1], 3)
    assert np.all(I == -1) and np.all(np.isinf(D))
    ix.close()


def test_descriptor_and_pnp_argument_validation(native_lib, cuda_device):
    from cerebro_b200._lib import CerebroB200Error
    from cerebro_b200.descriptor import NetvladDescriptor
    from cerebro_b200.keras_weights import random_mobilenet_netvlad
    from cerebro_b200.pnp import PnpBatch, default_params

    net = random_mobilenet_netvlad(1, 6, 16, seed=1)
    with pytest.raises(CerebroB200Error):
        NetvladDescriptor(net, 96, 128, 3, max_batch=1)  # channel mismatch with the model
    nd = NetvladDescriptor(net, 96, 128, 1, max_batch=2)
    imgs = synth.band_limited_images(3, 96, 128, 1, seed=2)
    with pytest.raises(CerebroB200Error):
        nd.compute(imgs)  # batch 3 > max_batch 2
    d = nd.compute(imgs[:2])
    assert d.shape == (2, nd.dim) and np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-5)
    nd.close()
    pb = PnpBatch(max_candidates=2, max_points_total=100, max_hypotheses=10)
    X = np.random.default_rng(0).standard_normal((30, 3))
    uv = np.random.default_rng(1).standard_normal((30, 2))
    with pytest.raises(CerebroB200Error):
        pb.solve([X, X, X], [uv, uv, uv])  # 3 candidates > max 2
    with pytest.raises(CerebroB200Error):
        pb.solve([X], [uv], default_params(max_iterations=11))  # hypotheses > max 10
    r = pb.solve([X], [uv], default_params(max_iterations=10))  # garbage data: no crash, low confidence
    assert -1.0 <= float(r["confidence"][0]) <= 1.0
    pb.close()
