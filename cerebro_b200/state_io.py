"""Persistence wire format of the keyframe descriptor database: the ``DataNodes`` array of ``state.json``.

Mirrors the slice of ``DataManager::saveStateToDisk`` / ``loadStateFromDisk`` (src/DataManager.cpp:1107-1215,
1218-1350) that the loop-detection hot path depends on: per data node ``stampNSec``, ``seq``, ``isKeyFrame``,
``getNumberOfSuccessfullyTrackedFeatures``, ``isWholeImageDescriptorAvailable`` and
``wholeImageDescriptor = {rows: D, cols: 1, data: "v0\\nv1\\n..."}`` -- Eigen's
``IOFormat(FullPrecision, DontAlignCols, ", ", "\\n")`` of a VectorXd (:1121, :1157-1168; FullPrecision =
15 significant digits), parsed back by ``RawFileIO::read_eigen_vector_fromjson`` (src/utils/RawFileIO.cpp:418-457:
split on '\\n', one value per row, ``std::stod``).  On resume the reference re-lists every node that has a
descriptor into ``wholeImageComputedList`` (src/Cerebro.cpp:128-160); ``Cerebro.load_state`` does the same and
bulk-loads the rows into the device index in one call.

Poses and the ImageDataManager block are carried through untouched when present (they are not on this path).
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, Optional, Tuple

import numpy as np


def vector_to_csv(v: np.ndarray) -> str:
    """Eigen ``VectorXd.format(CSVFormat)`` with FullPrecision: one value per line, 15 significant digits."""
    return "\n".join("%.15g" % float(x) for x in np.asarray(v, dtype=np.float64).reshape(-1))


def csv_to_vector(desc_ifo: dict) -> np.ndarray:
    """RawFileIO::read_eigen_vector_fromjson (RawFileIO.cpp:418-457), same refusals (raises ValueError)."""
    ncols, nrows, data = int(desc_ifo["cols"]), int(desc_ifo["rows"]), desc_ifo["data"]
    if nrows <= 0:
        raise ValueError("[read_eigen_vector_fromjson] nrows should be positive")
    if ncols != 1:
        raise ValueError("[read_eigen_vector_fromjson] json cols != 1")
    rows = data.split("\n")
    if len(rows) != nrows:
        raise ValueError("[read_eigen_vector_fromjson] requested %d but actually are %d" % (nrows, len(rows)))
    out = np.empty(nrows, dtype=np.float64)
    for r, line in enumerate(rows):
        cols = line.split(",")
        if len(cols) != 1:
            raise ValueError("[read_eigen_vector_fromjson] %d columns in row %d" % (len(cols), r))
        out[r] = float(cols[0])
    return out


def make_datanode(seq: int, stamp_nsec: int, desc: Optional[np.ndarray], is_keyframe: bool = True, n_tracked: int = -1,
                  pose0_nsec: int = 0) -> dict:
    """One entry of ``DataNodes`` as saveStateToDisk writes it (DataManager.cpp:1129-1178), without the pose block."""
    obj = {
        "stampNSec": int(stamp_nsec),
        "stamp_relative": (int(stamp_nsec) - int(pose0_nsec)) * 1e-9,
        "seq": int(seq),
        "isKeyFrame": bool(is_keyframe),
        "getNumberOfSuccessfullyTrackedFeatures": int(n_tracked),
        "isWholeImageDescriptorAvailable": desc is not None,
        "isPoseAvailable": False,
    }
    if desc is not None:
        d = np.asarray(desc, dtype=np.float64).reshape(-1)
        obj["wholeImageDescriptor"] = {"rows": int(d.shape[0]), "cols": 1, "data": vector_to_csv(d)}
    return obj


def save_state(save_folder_name: str, stamps_nsec: Iterable[int], descriptors: np.ndarray, is_keyframe=None, n_tracked=None,
               extra: Optional[dict] = None) -> str:
    """Write ``<save_folder_name>/state.json`` holding one data node per stamp (ascending, like the std::map the
    reference iterates, DataManager.cpp:1127).  ``descriptors`` is [n, D]; a row of NaNs means "no descriptor"."""
    stamps = [int(s) for s in stamps_nsec]
    descriptors = np.asarray(descriptors)
    assert descriptors.ndim == 2 and descriptors.shape[0] == len(stamps)
    order = np.argsort(np.asarray(stamps, dtype=np.int64), kind="stable")
    nodes = []
    for seq, i in enumerate(order):
        d = descriptors[i]
        has = not bool(np.isnan(d[0]))
        nodes.append(make_datanode(seq, stamps[i], d if has else None,
                                   True if is_keyframe is None else bool(is_keyframe[i]),
                                   -1 if n_tracked is None else int(n_tracked[i]), pose0_nsec=stamps[order[0]]))
    state = dict(extra or {})
    state["DataNodes"] = nodes
    os.makedirs(save_folder_name, exist_ok=True)
    path = os.path.join(save_folder_name, "state.json")
    with open(path, "w") as f:
        json.dump(state, f)
    return path


def load_state(save_folder_name: str) -> Tuple[List[int], np.ndarray, List[dict]]:
    """Read ``state.json`` (a folder, or the file itself).  Returns (stamps_nsec of the nodes that have a descriptor,
    in map order; their descriptors as float64 [n, D]; all raw node dicts).  A missing file raises FileNotFoundError --
    the reference prints "Cannot load from previous state" and exits (DataManager.cpp:1232-1238)."""
    path = save_folder_name if save_folder_name.endswith(".json") else os.path.join(save_folder_name, "state.json")
    with open(path) as f:
        obj = json.load(f)
    nodes = obj["DataNodes"]
    # data_map is a std::map keyed by stamp: iteration order on resume is ascending stamp (Cerebro.cpp:143)
    nodes_sorted = sorted(nodes, key=lambda n: int(n["stampNSec"]))
    stamps, descs = [], []
    for n in nodes_sorted:
        if n.get("isWholeImageDescriptorAvailable"):
            stamps.append(int(n["stampNSec"]))
            descs.append(csv_to_vector(n["wholeImageDescriptor"]))
    if descs:
        dim = descs[0].shape[0]
        if any(d.shape[0] != dim for d in descs):
            raise ValueError("descriptors of different lengths in %s" % path)
        mat = np.stack(descs)
    else:
        mat = np.zeros((0, 0), dtype=np.float64)
    return stamps, mat, nodes_sorted
