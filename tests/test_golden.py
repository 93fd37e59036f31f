"""Committed golden vectors (tests/golden/golden.json, generated from the reference's own rasterizer
by tests/golden/make_golden.py).  CPU: the oracle restatement must reproduce them.  GPU: the CUDA path
must reproduce them — this is the parity check that needs neither /root/reference nor oracle/_ref."""
import hashlib
import json
import os

import numpy as np
import pytest

from harness import scenes
from tests.golden.make_golden import SCENES, crop

with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as f:
    GOLDEN = json.load(f)["scenes"]

SMALL = [n for n in GOLDEN if "3840" not in n and "200000" not in n and "7680" not in n]
FULL = [n for n in GOLDEN if n not in SMALL]


def _check(be, name):
    g = GOLDEN[name]
    sc = SCENES[name]()
    c, d = scenes.render(be, sc)
    assert crop(c).tobytes().hex() == g["color_crop_hex"], f"{name}: centre crop differs"
    assert hashlib.sha256(c.tobytes()).hexdigest() == g["color_sha256"], f"{name}: colour hash differs"
    if d is not None:
        assert hashlib.sha256(d.tobytes()).hexdigest() == g["depth_sha256"], f"{name}: depth hash differs"


@pytest.mark.parametrize("name", SMALL)
def test_oracle_reproduces_golden(vor, name):
    _check(vor, name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SMALL + FULL)
def test_gpu_reproduces_golden(gpu, name):
    _check(gpu, name)
