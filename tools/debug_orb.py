import sys, os, collections
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import golden_io, synth_orb
from cerebro_b200.features import Features
g = golden_io.load("orb_golden.npz")
name, h, w, n, kind, seed = synth_orb.CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 3]
img = synth_orb.image(kind, h, w, seed)
fe = Features(h, w, max_images=1, max_keypoints=8000)
r = fe.orb(img[None], n)[0]
gk = g[name + "_kps"]
print("device octaves", sorted(collections.Counter(r["octave"].tolist()).items()))
print("opencv octaves", sorted(collections.Counter(gk[:, 5].astype(int).tolist()).items()))
dev = set((float(a), float(b), int(c)) for (a, b), c in zip(r["pt"], r["octave"]))
ref = set((float(a), float(b), int(c)) for a, b, c in zip(gk[:, 0], gk[:, 1], gk[:, 5]))
print("common", len(dev & ref), "only device", len(dev - ref), "only opencv", len(ref - dev))
miss = sorted(ref - dev, key=lambda t: (t[2], t[1], t[0]))[:20]
print("missing (x,y,octave):", miss)
extra = sorted(dev - ref, key=lambda t: (t[2], t[1], t[0]))[:20]
print("extra:", extra)
from oracle import orb as O
lv = O.build_pyramid(img)
pyr = fe.debug_read(0); sc = fe.debug_read(1)
off = 0
for L in range(8):
    hh, ww = lv[L].shape
    dp = pyr[off:off + hh * ww].reshape(hh, ww); ds = sc[off:off + hh * ww].reshape(hh, ww)
    so = O.fast_score_map(lv[L], 0).astype(np.uint8)
    print("level", L, (hh, ww), "pyramid mismatches", int((dp != lv[L]).sum()), "score mismatches", int((ds != so).sum()), "(where pyramid ok:", int(((ds != so) & True).sum()), ")")
    if L == 0 and (ds != so).any():
        ys, xs = np.nonzero(ds != so)
        print("   first score diffs", [(int(x), int(y), int(ds[y, x]), int(so[y, x])) for x, y in zip(xs[:8], ys[:8])])
    off += hh * ww
