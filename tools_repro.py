import sys
import os; sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from harness import abi, scenes
gpu = abi.backend("vb200", 0)
for sc in (scenes.c4_particles(640, 360, 20000), scenes.c5_textured(960, 540, 250, 125, tex_size=256)):
    c, d = scenes.render(gpu, sc)
    print(sc.name, "ok", flush=True)
