"""state.json descriptor persistence (DataManager::saveStateToDisk / loadStateFromDisk wire format)."""
import json
import os

import numpy as np
import pytest

from cerebro_b200 import state_io
from tests import synth


def test_vector_csv_is_eigen_fullprecision_and_roundtrips():
    v = np.array([0.1, -1.0 / 3.0, 1e-12, 123456.789012345678, 0.0, 1.0])
    s = state_io.vector_to_csv(v)
    assert s.split("\n")[0] == "0.1" and s.split("\n")[1] == "-0.333333333333333"  # 15 significant digits, one value per line
    assert "," not in s and not s.endswith("\n")
    back = state_io.csv_to_vector({"rows": v.size, "cols": 1, "data": s})
    assert np.allclose(back, v, rtol=1e-14, atol=0)
    # fp32 descriptors survive the text round trip exactly (15 digits > the 9 an fp32 needs)
    d32 = synth.unit_rows(3, 256, seed=4)
    for row in d32:
        back = state_io.csv_to_vector({"rows": 256, "cols": 1, "data": state_io.vector_to_csv(row)})
        assert np.array_equal(back.astype(np.float32), row)


def test_reader_refusals_match_reference():
    with pytest.raises(ValueError):
        state_io.csv_to_vector({"rows": 0, "cols": 1, "data": ""})
    with pytest.raises(ValueError):
        state_io.csv_to_vector({"rows": 2, "cols": 2, "data": "1, 2\n3, 4"})
    with pytest.raises(ValueError):
        state_io.csv_to_vector({"rows": 3, "cols": 1, "data": "1\n2"})
    with pytest.raises(ValueError):
        state_io.csv_to_vector({"rows": 2, "cols": 1, "data": "1, 5\n2"})


def test_state_json_schema_and_order(tmp_path):
    d = synth.unit_rows(5, 128, seed=1).astype(np.float64)
    d[3] = np.nan  # a node without a descriptor (non-keyframe)
    stamps = [5_000_000_000, 1_000_000_000, 3_000_000_000, 2_000_000_000, 4_000_000_000]
    path = state_io.save_state(str(tmp_path), stamps, d, is_keyframe=[1, 1, 1, 0, 1], n_tracked=[50, 60, 70, 10, 80])
    obj = json.load(open(path))
    nodes = obj["DataNodes"]
    assert [n["seq"] for n in nodes] == [0, 1, 2, 3, 4]
    assert [n["stampNSec"] for n in nodes] == sorted(stamps)
    n0 = nodes[0]
    for key in ("stampNSec", "stamp_relative", "seq", "isKeyFrame", "getNumberOfSuccessfullyTrackedFeatures",
                "isWholeImageDescriptorAvailable", "isPoseAvailable", "wholeImageDescriptor"):
        assert key in n0, key
    assert n0["wholeImageDescriptor"]["rows"] == 128 and n0["wholeImageDescriptor"]["cols"] == 1
    assert "wholeImageDescriptor" not in nodes[1] and nodes[1]["isWholeImageDescriptorAvailable"] is False  # stamp 2e9 = the NaN row
    st, mat, raw = state_io.load_state(str(tmp_path))
    assert st == [1_000_000_000, 3_000_000_000, 4_000_000_000, 5_000_000_000]
    assert np.allclose(mat, d[[1, 2, 4, 0]], rtol=1e-14, atol=0)
    with pytest.raises(FileNotFoundError):
        state_io.load_state(os.path.join(str(tmp_path), "nope"))


@pytest.mark.gpu
def test_cerebro_resume_from_state_json(native_lib, cuda_device, tmp_path):
    """A stream stopped after 330 keyframes, saved, resumed in a fresh Cerebro and continued gives the same loop
    candidates as the uninterrupted stream (the reference's loadStateFromDisk launch files)."""
    from cerebro_b200.loop_detector import Cerebro

    class FakeDesc:  # descriptors are supplied directly; only .dim is used by the mirror
        dim = 512

    n = 420
    base = synth.unit_rows(n, 512, seed=9)
    for i in range(6):
        base[360 + 3 * i: 363 + 3 * i] = synth.planted_queries(base, [40 + i, 40 + i, 40 + i], seed=50 + i, score=0.95)
    stamps = [10_000_000_000 + 100_000_000 * i for i in range(n)]

    def run(c, a, b):
        out = []
        for i in range(a, b, 3):
            c.index.add(base[i: i + 3])
            c._whole.extend(stamps[i: i + 3])
            e = c.run_step()
            if e:
                out.append(e)
        return out

    full = Cerebro(FakeDesc(), capacity=n)
    expect = run(full, 0, n)
    assert len(expect) >= 3
    first = Cerebro(FakeDesc(), capacity=n)
    got = run(first, 0, 330)
    first.save_state(str(tmp_path))
    second = Cerebro(FakeDesc(), capacity=n)
    assert second.load_state(str(tmp_path)) == 330
    assert second.wholeImageComputedList_size() == 330 and second.wholeImageComputedList_at(7) == stamps[7]
    assert np.array_equal(second.index.get_rows(0, 330), first.index.get_rows(0, 330))
    got += run(second, 330, n)
    assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in expect]
    assert np.allclose([s for *_, s in got], [s for *_, s in expect], atol=1e-12)
