"""The built library really contains hand-written Blackwell code on the hot path: `cuobjdump -sass` of the in-tree .so must show
tcgen05 MMAs (UTCHMMA / UTCIMMA), TMA tensor loads / stores (UTMALDG / UTMASTG) and TMEM loads (LDTM) inside the kernels that
DESIGN.md says use them, and warp reductions (REDUX) in the register-resident PnP elimination and the StereoBM decision.
Runs without a GPU (the driver's build check compiles the library here)."""
import re
import shutil
import subprocess

import pytest

from cerebro_b200 import build as cb_build

EXPECT = {
    "scores_tc2_kernel<64>": ["UTCHMMA", "UTMALDG", "LDTM"],
    "scores_tc2_kernel<128>": ["UTCHMMA", "UTMALDG", "LDTM"],
    "dwpw_halo_kernel<32, 64, 1, 7, 8>": ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"],
    "dwpw_halo_kernel<512, 512, 1, 9, 8>": ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"],
    "conv1_tc_kernel<3, true>": ["UTCHMMA", "LDTM"],
    "vlad_assign_tc_kernel": ["UTCHMMA", "UTMALDG", "LDTM"],
    "hamming_tc_kernel": ["UTCIMMA", "UTMALDG", "LDTM"],
    "dls_eliminate2_kernel": ["REDUX", "DFMA"],
    "sbm_vsad2_kernel": ["REDUX"],
}


@pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None, reason="cuobjdump / c++filt not installed")
def test_hot_kernels_contain_blackwell_instructions(native_lib):
    sass = subprocess.run(["cuobjdump", "-sass", cb_build.LIB], capture_output=True, text=True, check=True).stdout
    sass = subprocess.run(["c++filt"], input=sass, capture_output=True, text=True, check=True).stdout
    found = {}
    name = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (.*)$", ln)
        if m:
            name = re.sub(r"\(anonymous namespace\)::", "", re.sub(r"^void ", "", m.group(1)))
            name = re.sub(r"\(.*$", "", name)
            found.setdefault(name, set())
            continue
        if name is None:
            continue
        mm = re.search(r"\b(UTCHMMA|UTCIMMA|UTMALDG|UTMASTG|LDTM|C?REDUX|DFMA)\b", ln)
        if mm:
            found[name].add(mm.group(1).replace("CREDUX", "REDUX"))
    for kern, need in EXPECT.items():
        assert kern in found, "kernel %s not in the library (have e.g. %s)" % (kern, sorted(found)[:5])
        missing = [m for m in need if m not in found[kern]]
        assert not missing, "%s lacks %s (has %s)" % (kern, missing, sorted(found[kern]))
