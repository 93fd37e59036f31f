import sys
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from harness import abi, scenes
gpu = abi.backend("vb200", 0)
vor = abi.backend("vor", 0)
def chk(name, sc):
    c, d = scenes.render(gpu, sc)
    c2, d2 = scenes.render(vor, sc)
    bad = (c != c2).any(-1)
    print(name, "colour diff px", int(bad.sum()), "depth diff", None if d is None else int((d.view(np.uint32) != d2.view(np.uint32)).sum()), flush=True)
    if bad.any():
        ys, xs = np.nonzero(bad)
        print("  first", ys[:5], xs[:5], c[ys[0], xs[0]], c2[ys[0], xs[0]])
a = scenes.random_triangles(300, 200, 100, 90); a.clear_depth = None
chk("a", a)
b = scenes.random_triangles(300, 200, 100, 91, depth_op=abi.CMP_ALWAYS, depth_write=False)
chk("b", b)
c = scenes.random_triangles(333, 211, 100, 92)
c.draws += scenes.random_triangles(333, 211, 100, 93, blend=(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, 0), depth_op=abi.CMP_LESS, depth_write=False).draws
chk("c", c)
d = scenes.random_triangles(64, 48, 3, 94, max_size=0.05)
chk("d", d)
