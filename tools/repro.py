"""Developer aid: hammer the deferred-clear scenes to catch rare races (runs on the GPU box)."""
import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from harness import abi, scenes
gpu = abi.backend("vb200", 0)
vor = abi.backend("vor", 0)
def mk():
    a = scenes.random_triangles(300, 200, 100, 90); a.clear_depth = None
    b = scenes.random_triangles(300, 200, 100, 91, depth_op=abi.CMP_ALWAYS, depth_write=False)
    c = scenes.random_triangles(333, 211, 100, 92)
    c.draws += scenes.random_triangles(333, 211, 100, 93, blend=(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0), depth_op=abi.CMP_LESS, depth_write=False).draws
    d = scenes.random_triangles(64, 48, 3, 94, max_size=0.05)
    return [("a", a), ("b", b), ("c", c), ("d", d)]
ref = {n: scenes.render(vor, s) for n, s in mk()}
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 150):
    for n, s in mk():
        c, d = scenes.render(gpu, s)
        c2, d2 = ref[n]
        mc = (c != c2).any(-1)
        md = (d.view(np.uint32) != d2.view(np.uint32)) if d is not None else np.zeros(1, bool)
        if mc.any() or md.any():
            bad += 1
            ys, xs = np.nonzero(mc) if mc.any() else np.nonzero(md)
            print(f"iter {it} scene {n}: colour diff {int(mc.sum())} depth diff {int(md.sum())} bbox x {xs.min()}..{xs.max()} y {ys.min()}..{ys.max()}",
                  "gpu", c[ys[0], xs[0]], "cpu", c2[ys[0], xs[0]], flush=True)
print("mismatching renders:", bad)
