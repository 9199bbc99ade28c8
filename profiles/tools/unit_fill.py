# Fill distribution of the listed 128-px units of the synchronised masks (bench.py default workload): python profiles/tools/unit_fill.py
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, bench
from roft_b200 import api
r = bench.Runner(api, "cuda:0", 0, 64, 12, 0.25, 6, 1, "f32", "auto")
for _ in range(15): r.do_step()
raw, thr = r.trk.mask()
m = (raw > 1).reshape(raw.shape[0], -1, 128)
cnt = m.sum(-1)
listed = cnt[(raw.reshape(raw.shape[0], -1, 128) != 0).any(-1)]
h = np.histogram(listed, bins=[0,1,8,16,32,48,64,80,96,112,127,128,129])
print("listed units", listed.size, "mean fill", listed.mean()/128)
for a,b,c in zip(h[1][:-1], h[1][1:], h[0]): print(f"[{a:3d},{b:3d}) {c:8d} {100*c/listed.size:5.1f}%")
# per-quad occupancy: fraction of quads (4 px) with any candidate, and mean candidates per nonempty quad
q = m.reshape(raw.shape[0], -1, 32, 4)
qa = q.any(-1)
lu = (raw.reshape(raw.shape[0], -1, 128) != 0).any(-1)
print("nonempty quads per listed unit", qa[lu].sum(-1).mean(), "cand per nonempty quad", q[lu].sum()/max(qa[lu].sum(),1))
