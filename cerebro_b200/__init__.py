"""cerebro-b200: the loop-detection hot path of mpkuse/cerebro (NetVLAD descriptor -> descriptor search -> DLS-PnP RANSAC, plus
the correspondence front end) on B200, behind the C ABI of ``include/cerebro_b200.h``.

The reference-named host mirrors are importable from the package root; the native library is loaded on first use of a handle
(there is no CPU fallback: without the built ``_native/libcerebro_b200.so`` or without an sm_100 device the constructors raise).

    from cerebro_b200 import HDF5ModelImageDescriptor, IndexFlatIP, StaticTheiaPoseCompute, Cerebro
"""
_EXPORTS = {
    # reference name -> module (scripts/whole_image_desc_compute_server.py, faiss::IndexFlatIP, src/DlsPnpWithRansac.h, src/Cerebro.h)
    "HDF5ModelImageDescriptor": "descriptor",
    "NetvladDescriptor": "descriptor",
    "IndexFlatIP": "index",
    "ShardedIndex": "index",
    "StaticTheiaPoseCompute": "pnp",
    "PnpBatch": "pnp",
    "Cerebro": "loop_detector",
    "ProcessedLoopCandidate": "loop_detector",
    "LoopEdge": "loop_detector",
    "HypothesisManager": "loop_detector",
    "StaticPointFeatureMatching": "frontend",
    "FrontEnd": "frontend",
    "Features": "features",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):  # lazy: importing the package neither imports torch-free modules eagerly nor touches the native library
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError("module 'cerebro_b200' has no attribute %r" % name)
    import importlib

    return getattr(importlib.import_module("." + mod, __name__), name)
