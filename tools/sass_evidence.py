"""Raw SASS evidence: runs `cuobjdump -sass` on the built library and writes, per kernel of the bench step (and the front-end
tensor-core kernels), the first raw lines of every tensor-core / TMA / TMEM / warp-reduction mnemonic, with their addresses.
  python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cerebro_b200", "_native", "libcerebro_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTCIMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|UTCATOMSWS|SYNCS|REDUX|CREDUX|DFMA|FFMA2)\b")
KEEP = re.compile(r"scores_tc2|dwpw_halo_kernel<(32, 64|128, 128|512, 512)|conv1_tc_kernel<3|vlad_assign_tc|hamming_tc|dls_eliminate2|dls_roots|sbm_vsad2|pw_gemm_kernel<128, 64")


def main():
    cmd = ["cuobjdump", "-sass", LIB]
    out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
    filt = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout
    print("# " + " ".join(cmd) + " | c++filt   (sm_100a; raw lines, first 2 per mnemonic and kernel, then the totals)")
    print("# UTC*MMA = tcgen05.mma (UTCHMMA kind::f16, UTCIMMA kind::i8), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,")
    print("# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier, (C)REDUX = warp reduction, FFMA2 = packed fp32 FMA")
    name, rows, counts = None, collections.OrderedDict(), collections.Counter()

    def flush():
        if name and KEEP.search(name):
            print("\n== " + name[:150])
            for m, ls in rows.items():
                for ln in ls[:2]:
                    print("   " + ln)
            print("   totals: " + "  ".join("%s x%d" % kv for kv in sorted(counts.items())))

    for ln in filt.splitlines():
        m = re.match(r"\s*Function : (.*)$", ln)
        if m:
            flush()
            name, rows, counts = m.group(1), collections.OrderedDict(), collections.Counter()
            continue
        mm = PAT.search(ln)
        if mm and "/*" in ln:
            txt = re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln).strip()
            rows.setdefault(mm.group(1), []).append(txt)
            counts[mm.group(1)] += 1
    flush()


if __name__ == "__main__":
    sys.exit(main())
