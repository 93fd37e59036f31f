"""Seeded synthetic scenes for the five BASELINE.json configs plus random stress
scenes, and a backend-agnostic render() that issues the reference's operator sequence
(ClearTarget.. DrawTriangles.., cmd_exec.cpp:35-142) to any of the three libraries.

All data is generated once in float32 and fed bit-identically to every backend.  Scenes stay
inside the reference's defined domain: w > 0 for every vertex, window coordinates within
[-8192, 16383], texture sizes multiples of 4, supported blend factors only (SURVEY.md §7 hard
part 6).
"""
from __future__ import annotations

import ctypes as C
import hashlib
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi, shaders
from .abi import Backend


# ------------------------------------------------------------------------------------------------
@dataclass
class PipelineDesc:
    vs: np.ndarray
    fs: np.ndarray
    vattrs: Sequence[Tuple[int, int, int, int, int]]  # (location, format, stride, offset, vb)
    topology: int = abi.TOPO_LIST
    front_face: int = abi.FRONT_CCW
    cull_mode: int = abi.CULL_NONE
    depth_op: int = abi.CMP_ALWAYS
    depth_write: bool = False
    blend: Optional[Tuple[int, int, int]] = None  # (src, dst, op)


@dataclass
class Draw:
    pipe: PipelineDesc
    count: int
    first: int = 0
    indexed: bool = False
    vbs: Sequence[Tuple[np.ndarray, int]] = ()  # (buffer, offset) per slot
    ib: Optional[Tuple[np.ndarray, int, int]] = None  # (buffer, offset, index_type)
    ubos: Sequence[Tuple[int, int, np.ndarray, int]] = ()  # (set, binding, buffer, offset)
    textures: Sequence[Tuple[int, int, np.ndarray, int, int, int, int, int]] = ()
    # (set, binding, bytes, width, height, format, bpp, layers)
    push: bytes = b""


@dataclass
class Scene:
    name: str
    width: int
    height: int
    draws: List[Draw]
    depth: bool = False
    clear_color: Optional[Tuple[float, float, float, float]] = (0.2, 0.2, 0.2, 1.0)
    clear_depth: Optional[float] = 1.0
    notes: Dict[str, object] = field(default_factory=dict)

    def triangles(self) -> int:
        n = 0
        for d in self.draws:
            n += (d.count // 3) if d.pipe.topology == abi.TOPO_LIST else max(d.count - 2, 0)
        return n

    def algorithmic_bytes(self) -> int:
        """every input byte read once, every output byte written once."""
        seen = set()
        total = 0
        for d in self.draws:
            bufs = [v for v, _ in d.vbs] + ([d.ib[0]] if d.ib is not None else []) + \
                [t[2] for t in d.textures]
            for b in bufs:
                if id(b) not in seen:
                    seen.add(id(b))
                    total += b.nbytes
        px = self.width * self.height * 4
        total += px * (1 + (0 if self.clear_color is not None else 1))
        if self.depth:
            writes = any(d.pipe.depth_write for d in self.draws)
            total += px * ((1 if writes else 0) + (0 if self.clear_depth is not None else 1))
        return total


# ------------------------------------------------------------------------------------------------
class _ShaderCache:
    def __init__(self) -> None:
        self.mods: Dict[Tuple[str, str], Tuple[int, int]] = {}

    def entry(self, be: Backend, words: np.ndarray) -> int:
        key = (be.kind, hashlib.sha1(words.tobytes()).hexdigest())
        if key not in self.mods:
            mod = be.CompileFunction(words)
            self.mods[key] = (mod, be.GetFuncPointer(mod, "main"))
        return self.mods[key][1]


_shader_cache = _ShaderCache()


class BoundScene:
    """A Scene lowered to ctypes structs for one backend; keeps every buffer alive."""

    def __init__(self, be: Backend, scene: Scene, color: Optional[np.ndarray] = None,
                 depth: Optional[np.ndarray] = None, color_device_ptr: Optional[int] = None,
                 depth_device_ptr: Optional[int] = None) -> None:
        """color/depth: host arrays to render into (allocated if omitted). color_device_ptr /
        depth_device_ptr: CUDA device addresses to use as attachments instead (vb200 only)."""
        self.be, self.scene = be, scene
        w, h = scene.width, scene.height
        self.color = color if color is not None else np.full((h, w, 4), 0xCD, dtype=np.uint8)
        self.depth = None
        if scene.depth:
            self.depth = depth if depth is not None else np.full((h, w), 0.75, dtype=np.float32)
        self.color_img = abi.make_image(self.color, w, h, abi.FMT_B8G8R8A8_UNORM)
        self.depth_img = abi.make_image(self.depth, w, h, abi.FMT_D32_SFLOAT)
        if color_device_ptr is not None:
            self.color_img.pixels = color_device_ptr
        if depth_device_ptr is not None and scene.depth:
            self.depth_img.pixels = depth_device_ptr
        self.keep: List[object] = []
        self.calls: List[Tuple[abi.DrawState, Draw]] = []
        pipes: Dict[int, abi.Pipeline] = {}
        for d in scene.draws:
            if id(d.pipe) not in pipes:
                pipes[id(d.pipe)] = self._pipeline(d.pipe)
            st = abi.DrawState()
            if d.ib is not None:
                st.ib.buffer = abi.make_buffer(d.ib[0])
                st.ib.offset = d.ib[1]
                st.ib.index_type = d.ib[2]
            for slot, (buf, off) in enumerate(d.vbs):
                st.vbs[slot].buffer = abi.make_buffer(buf)
                st.vbs[slot].offset = off
            st.color = self.color_img
            st.depth = self.depth_img
            st.pipeline = C.pointer(pipes[id(d.pipe)])
            nb = len(d.ubos) + len(d.textures)
            arr = (abi.Binding * max(nb, 1))()
            i = 0
            for (s, b, buf, off) in d.ubos:
                arr[i].set, arr[i].binding, arr[i].type, arr[i].is_image = s, b, abi.DESC_UNIFORM_BUFFER, 0
                arr[i].buffer = abi.make_buffer(buf)
                arr[i].offset = off
                i += 1
            for (s, b, data, tw, th, fmt, bpp, layers) in d.textures:
                arr[i].set, arr[i].binding, arr[i].is_image = s, b, 1
                arr[i].type = abi.DESC_COMBINED_IMAGE_SAMPLER
                arr[i].image = abi.make_image(data, tw, th, fmt, bpp, layers)
                i += 1
            st.bindings = C.cast(arr, C.POINTER(abi.Binding))
            st.num_bindings = nb
            if d.push:
                C.memmove(st.pushconsts, d.push, min(len(d.push), 128))
            self.keep.append(arr)
            self.calls.append((st, d))
        self.keep.append(pipes)

    def _pipeline(self, p: PipelineDesc) -> abi.Pipeline:
        pl = abi.Pipeline()
        for (loc, fmt, stride, off, vb) in p.vattrs:
            pl.vattrs[loc].format, pl.vattrs[loc].stride = fmt, stride
            pl.vattrs[loc].offset, pl.vattrs[loc].vb = off, vb
        pl.topology, pl.front_face, pl.cull_mode = p.topology, p.front_face, p.cull_mode
        pl.depth_compare_op, pl.depth_write_enable = p.depth_op, int(p.depth_write)
        if p.blend is not None:
            pl.blend_enable = 1
            pl.src_color_blend_factor, pl.dst_color_blend_factor, pl.color_blend_op = p.blend
        pl.vs = _shader_cache.entry(self.be, p.vs)
        pl.fs = _shader_cache.entry(self.be, p.fs)
        return pl

    def submit(self) -> None:
        """One command buffer: BeginRenderPass clears, then the draws (no flush)."""
        be, sc = self.be, self.scene
        if sc.clear_color is not None:
            be.ClearTarget(self.color_img, sc.clear_color)
        if sc.depth and sc.clear_depth is not None:
            be.ClearTarget(self.depth_img, sc.clear_depth)
        for st, d in self.calls:
            be.DrawTriangles(st, d.count, d.first, d.indexed)

    def run(self) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        self.submit()
        self.be.flush()
        return self.color, self.depth


def render(be: Backend, scene: Scene) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    return BoundScene(be, scene).run()


def image_hash(color: np.ndarray, depth: Optional[np.ndarray]) -> str:
    h = hashlib.sha256(np.ascontiguousarray(color).tobytes())
    if depth is not None:
        h.update(np.ascontiguousarray(depth).tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------------------------------------
# float32 matrix helpers (host side; column-major storage for the UBO = transpose of row-major np)
def _perspective(fovy_deg: float, aspect: float, near: float, far: float) -> np.ndarray:
    f = np.float32(1.0 / np.tan(np.radians(fovy_deg) / 2.0))
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = f / np.float32(aspect)
    m[1, 1] = f
    m[2, 2] = np.float32(far / (near - far))
    m[2, 3] = np.float32(near * far / (near - far))
    m[3, 2] = -1.0
    return m


def _look_at(eye, target, up) -> np.ndarray:
    eye, target, up = (np.asarray(v, dtype=np.float64) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    s = np.cross(f, up)
    s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m.astype(np.float32)


def _rot_y(deg: float) -> np.ndarray:
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    m = np.eye(4)
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    return m.astype(np.float32)


def _rot_x(deg: float) -> np.ndarray:
    c, s = np.cos(np.radians(deg)), np.sin(np.radians(deg))
    m = np.eye(4)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m.astype(np.float32)


def _ubo_mat(m: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(m.astype(np.float32).T).reshape(-1)


def _texture(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    # smooth-ish random RGBA so bilinear filtering has structure
    base = rng.integers(0, 256, size=(size // 4, size // 4, 4), dtype=np.uint8)
    tex = np.kron(base, np.ones((4, 4, 1), dtype=np.uint8))
    noise = rng.integers(0, 32, size=(size, size, 4), dtype=np.uint8)
    return np.ascontiguousarray((tex // 2 + noise + 32).astype(np.uint8))


# ------------------------------------------------------------------------------------------------
def c1_triangle(width: int = 1280, height: int = 720) -> Scene:
    """C1: single vkCmdDraw triangle, passthrough VS/FS, no depth."""
    v = np.array([[0.0, -0.5, 0.5, 1.0, 1, 0, 0, 1],
                  [0.5, 0.5, 0.5, 1.0, 0, 1, 0, 1],
                  [-0.5, 0.5, 0.5, 1.0, 0, 0, 1, 1]], dtype=np.float32)
    pipe = PipelineDesc(shaders.vs_passthrough(), shaders.fs_color(),
                        [(0, abi.FMT_R32G32B32A32_SFLOAT, 32, 0, 0), (1, abi.FMT_R32G32B32A32_SFLOAT, 32, 16, 0)])
    return Scene("c1_triangle", width, height, [Draw(pipe, 3, vbs=[(v, 0)])], depth=False)


_CUBE_FACES = [  # (normal axis, sign)
    ((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)),
    ((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, -1, 0), (1, 0, 0), (0, 0, 1)),
    ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0)),
]


def c2_cube(width: int = 1920, height: int = 1080, frame: int = 17, tex_size: int = 256) -> Scene:
    """C2: textured rotating cube, D32 depth LESS+write, bilinear RGBA8 sampling."""
    verts = []
    for n, a, b in _CUBE_FACES:
        n, a, b = (np.array(x, dtype=np.float32) for x in (n, a, b))
        quad = [(-1, -1), (1, -1), (1, 1), (-1, -1), (1, 1), (-1, 1)]
        for (s, t) in quad:
            p = n + s * a + t * b
            verts.append([p[0], p[1], p[2], 1.0, (s + 1) * 0.5 * 2.0, (t + 1) * 0.5 * 2.0])
    v = np.array(verts, dtype=np.float32)  # 36 x 6 floats = 24 B stride
    mvp = _perspective(45.0, width / height, 0.5, 20.0) @ _look_at((0, 1.8, 5.0), (0, 0, 0), (0, 1, 0)) \
        @ _rot_y(float(frame)) @ _rot_x(float(frame) * 0.5)
    ubo = _ubo_mat(mvp)
    tex = _texture(tex_size, 2)
    pipe = PipelineDesc(shaders.vs_mvp_uv(), shaders.fs_texture(),
                        [(0, abi.FMT_R32G32B32A32_SFLOAT, 24, 0, 0), (1, abi.FMT_R32G32_SFLOAT, 24, 16, 0)],
                        depth_op=abi.CMP_LESS, depth_write=True)
    d = Draw(pipe, 36, vbs=[(v, 0)], ubos=[(0, 0, ubo, 0)],
             textures=[(0, 1, tex, tex_size, tex_size, abi.FMT_R8G8B8A8_UNORM, 4, 1)])
    return Scene("c2_cube", width, height, [d], depth=True)


def _value_noise(nx: int, ny: int, cells: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    g = rng.random((cells + 2, cells + 2)).astype(np.float64)
    xs = np.linspace(0, cells, nx, endpoint=False)
    ys = np.linspace(0, cells, ny, endpoint=False)
    x0, y0 = xs.astype(int), ys.astype(int)
    fx, fy = xs - x0, ys - y0
    fx, fy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
    a = g[np.ix_(y0, x0)]
    b = g[np.ix_(y0, x0 + 1)]
    c = g[np.ix_(y0 + 1, x0)]
    d = g[np.ix_(y0 + 1, x0 + 1)]
    top = a * (1 - fx)[None, :] + b * fx[None, :]
    bot = c * (1 - fx)[None, :] + d * fx[None, :]
    return top * (1 - fy)[:, None] + bot * fy[:, None]


def _grid_mesh(qx: int, qy: int, seed: int, amp: float = 0.22) -> Tuple[np.ndarray, np.ndarray]:
    """(qx x qy) quads -> vertices {pos3, normal3, uv2} (32 B) and u32 triangle-list indices."""
    nx, ny = qx + 1, qy + 1
    xs = np.linspace(-2.45, 2.45, nx)
    zs = np.linspace(-1.75, 1.05, ny)
    hgt = amp * (_value_noise(nx, ny, 24, seed) + 0.5 * _value_noise(nx, ny, 61, seed + 100))
    X, Z = np.meshgrid(xs, zs)
    dx = np.gradient(hgt, xs, axis=1)
    dz = np.gradient(hgt, zs, axis=0)
    nrm = np.stack([-dx, np.ones_like(dx), -dz], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    # strictly positive uv: for v in (-2^-25, 0) the reference's `v - floor(v)` rounds to exactly 1.0
    # and it then reads the texel row at y == height, past the end of the texture (texture_sampling.cpp
    # :142-149) — undefined behaviour the scenes must stay clear of (SURVEY.md §7 hard part 6)
    uv = np.stack([(X + 2.5) * 2.0, (Z + 2.0) * 2.0], axis=-1)
    v = np.concatenate([np.stack([X, hgt, Z], axis=-1), nrm, uv], axis=-1).astype(np.float32)
    v = np.ascontiguousarray(v.reshape(-1, 8))
    j, i = np.meshgrid(np.arange(qy), np.arange(qx), indexing="ij")
    a = (j * nx + i).astype(np.uint32)
    b, c, d = a + 1, a + nx, a + nx + 1
    idx = np.stack([a, c, b, b, c, d], axis=-1).reshape(-1)
    return v, np.ascontiguousarray(idx.astype(np.uint32))


def _mesh_ubo(width: int, height: int) -> np.ndarray:
    mvp = _perspective(50.0, width / height, 0.1, 10.0) @ _look_at((0.0, 2.0, 1.55), (0, 0, -0.1), (0, 1, 0))
    light = np.array([0.35, 0.85, 0.4, 0.0], dtype=np.float32)
    light[:3] /= np.linalg.norm(light[:3])
    albedo = np.array([0.8, 0.7, 0.5, 1.0], dtype=np.float32)
    ambient = np.array([0.1, 0.12, 0.15, 0.0], dtype=np.float32)
    return np.concatenate([_ubo_mat(mvp), light, albedo, ambient]).astype(np.float32)


def _front_face_for(v: np.ndarray, idx: np.ndarray, ubo: np.ndarray, w: int, h: int) -> int:
    """Choose frontFace so most of the mesh is front-facing under rasterizer.cpp:401-424."""
    m = ubo[:16].reshape(4, 4).T.astype(np.float64)
    tri = idx[: (min(idx.size, 30000) // 3) * 3].reshape(-1, 3)
    p = np.concatenate([v[tri.reshape(-1), :3].astype(np.float64), np.ones((tri.size, 1))], axis=1) @ m.T
    wx = ((p[:, 0] / p[:, 3] + 1) * 0.5 * w).astype(np.int64).reshape(-1, 3)
    wy = ((-p[:, 1] / p[:, 3] + 1) * 0.5 * h).astype(np.int64).reshape(-1, 3)
    area2 = (wx[:, 1] - wx[:, 0]) * (wy[:, 2] - wy[:, 0]) - (wy[:, 1] - wy[:, 0]) * (wx[:, 2] - wx[:, 0])
    return abi.FRONT_CCW if (area2 > 0).sum() >= (area2 < 0).sum() else abi.FRONT_CW


def c3_mesh(width: int = 3840, height: int = 2160, qx: int = 1000, qy: int = 500, seed: int = 3) -> Scene:
    """C3: indexed lit mesh (qx*qy*2 triangles), per-vertex lighting, early-Z, cull back."""
    v, idx = _grid_mesh(qx, qy, seed)
    ubo = _mesh_ubo(width, height)
    pipe = PipelineDesc(shaders.vs_lit(False), shaders.fs_color(),
                        [(0, abi.FMT_R32G32B32_SFLOAT, 32, 0, 0), (1, abi.FMT_R32G32B32_SFLOAT, 32, 12, 0)],
                        front_face=_front_face_for(v, idx, ubo, width, height), cull_mode=abi.CULL_BACK,
                        depth_op=abi.CMP_LESS, depth_write=True)
    d = Draw(pipe, idx.size, indexed=True, vbs=[(v, 0)], ib=(idx, 0, abi.INDEX_U32), ubos=[(0, 0, ubo, 0)])
    return Scene("c3_mesh", width, height, [d], depth=True)


def c5_textured(width: int = 7680, height: int = 4320, qx: int = 2000, qy: int = 1000, seed: int = 5,
                tex_size: int = 2048) -> Scene:
    """C5: textured lit mesh at 8K (4M triangles at the default size)."""
    v, idx = _grid_mesh(qx, qy, seed)
    ubo = _mesh_ubo(width, height)
    tex = _texture(tex_size, 5)
    pipe = PipelineDesc(shaders.vs_lit(True), shaders.fs_lit_tex(),
                        [(0, abi.FMT_R32G32B32_SFLOAT, 32, 0, 0), (1, abi.FMT_R32G32B32_SFLOAT, 32, 12, 0),
                         (2, abi.FMT_R32G32_SFLOAT, 32, 24, 0)],
                        front_face=_front_face_for(v, idx, ubo, width, height), cull_mode=abi.CULL_BACK,
                        depth_op=abi.CMP_LESS, depth_write=True)
    d = Draw(pipe, idx.size, indexed=True, vbs=[(v, 0)], ib=(idx, 0, abi.INDEX_U32), ubos=[(0, 0, ubo, 0)],
             textures=[(0, 1, tex, tex_size, tex_size, abi.FMT_R8G8B8A8_UNORM, 4, 1)])
    return Scene("c5_textured", width, height, [d], depth=True)


def c4_particles(width: int = 1920, height: int = 1080, n: int = 200_000, seed: int = 4,
                 side: Tuple[int, int] = (6, 11)) -> Scene:
    """C4: alpha-blended screen-aligned quads, draw order = index order, no depth.
    Vertex = {pos float4, rgba8, 4 B pad} = 24 B; 4 verts + 6 u32 indices per quad."""
    rng = np.random.default_rng(seed)
    s = rng.integers(side[0], side[1] + 1, size=n)
    x0 = rng.integers(-4, width - 2, size=n)
    y0 = rng.integers(-4, height - 2, size=n)
    X = np.stack([x0, x0 + s, x0 + s, x0], axis=1).astype(np.float64)
    Y = np.stack([y0, y0, y0 + s, y0 + s], axis=1).astype(np.float64)
    ndcx = (X + 0.5) / width * 2.0 - 1.0
    ndcy = -((Y + 0.5) / height * 2.0 - 1.0)
    verts = np.zeros((n, 4), dtype=[("pos", np.float32, 4), ("rgba", np.uint8, 4), ("pad", np.uint32)])
    verts["pos"][..., 0] = ndcx
    verts["pos"][..., 1] = ndcy
    verts["pos"][..., 2] = 0.5
    verts["pos"][..., 3] = 1.0
    rgb = rng.integers(0, 256, size=(n, 1, 3), dtype=np.uint8)
    alpha = rng.integers(26, 128, size=(n, 1, 1), dtype=np.uint8)
    verts["rgba"] = np.concatenate([np.broadcast_to(rgb, (n, 4, 3)), np.broadcast_to(alpha, (n, 4, 1))], axis=2)
    vb = np.ascontiguousarray(verts.reshape(-1)).view(np.uint8)
    base = (np.arange(n, dtype=np.uint32) * 4)[:, None]
    idx = np.ascontiguousarray((base + np.array([0, 1, 2, 0, 2, 3], dtype=np.uint32)[None, :]).reshape(-1))
    pipe = PipelineDesc(shaders.vs_passthrough(), shaders.fs_color(),
                        [(0, abi.FMT_R32G32B32A32_SFLOAT, 24, 0, 0), (1, abi.FMT_R8G8B8A8_UNORM, 24, 16, 0)],
                        blend=(abi.BF_SRC_ALPHA, abi.BF_ONE_MINUS_SRC_ALPHA, abi.BLEND_ADD))
    d = Draw(pipe, idx.size, indexed=True, vbs=[(vb, 0)], ib=(idx, 0, abi.INDEX_U32))
    return Scene("c4_particles", width, height, [d], depth=False)


# ------------------------------------------------------------------------------------------------
def random_triangles(width: int, height: int, n: int, seed: int, *, depth_op: int = abi.CMP_LESS,
                     depth_write: bool = True, blend: Optional[Tuple[int, int, int]] = None,
                     cull: int = abi.CULL_NONE, front: int = abi.FRONT_CCW, max_size: float = 0.35,
                     has_depth: bool = True, topology: int = abi.TOPO_LIST, index_type: Optional[int] = None,
                     perspective: bool = True, offscreen: float = 0.15) -> Scene:
    """Overlapping random triangles with random w (perspective-correct interpolation), random
    colours and alpha: the stress scene for ordering, depth ties, culling and bbox clamping."""
    rng = np.random.default_rng(seed)
    nv = n * 3 if topology == abi.TOPO_LIST else n + 2
    if topology == abi.TOPO_LIST:
        c = rng.uniform(-1 - offscreen, 1 + offscreen, size=(n, 1, 2))
        p = c + rng.uniform(-max_size, max_size, size=(n, 3, 2))
        p = p.reshape(-1, 2)
    else:
        t = np.arange(nv)
        p = np.stack([-0.9 + 1.8 * t / max(nv - 1, 1) + rng.uniform(-0.05, 0.05, nv),
                      np.where(t % 2 == 0, -0.6, 0.6) + rng.uniform(-0.3, 0.3, nv)], axis=1)
    w = rng.uniform(0.5, 3.0, size=(nv, 1)) if perspective else np.ones((nv, 1))
    # quantise depth to a few levels so equal-depth ties between triangles really occur
    z = np.round(rng.uniform(0.05, 0.95, size=(nv, 1)) * 16) / 16
    pos = np.concatenate([p * w, z * w, w], axis=1).astype(np.float32)
    col = rng.uniform(0, 1, size=(nv, 4)).astype(np.float32)
    v = np.ascontiguousarray(np.concatenate([pos, col], axis=1))
    pipe = PipelineDesc(shaders.vs_passthrough(), shaders.fs_color(),
                        [(0, abi.FMT_R32G32B32A32_SFLOAT, 32, 0, 0), (1, abi.FMT_R32G32B32A32_SFLOAT, 32, 16, 0)],
                        topology=topology, front_face=front, cull_mode=cull, depth_op=depth_op,
                        depth_write=depth_write, blend=blend)
    if index_type is None:
        d = Draw(pipe, nv, vbs=[(v, 0)])
    else:
        dt = np.uint16 if index_type == abi.INDEX_U16 else np.uint32
        perm = rng.permutation(nv) if topology == abi.TOPO_LIST else np.arange(nv)
        # shuffle the vertex buffer and index back into draw order: exercises the index fetch
        vb = np.empty_like(v)
        vb[perm] = v
        idx = np.ascontiguousarray(perm.astype(dt))
        d = Draw(pipe, nv, indexed=True, vbs=[(vb, 0)], ib=(idx, 0, index_type))
    return Scene(f"random_{n}_{seed}", width, height, [d], depth=has_depth)


def split_draws(scene: Scene, parts: int, *, order: Optional[Sequence[int]] = None) -> Scene:
    """The same scene with its (single, triangle-list) draw cut into `parts` vkCmdDraw*s of consecutive
    triangle ranges out of the same buffers, as an application that draws a mesh chunk by chunk would
    record it (the reference replays them one after the other, cmd_exec.cpp:129-142). `order`: a
    permutation of the parts (submission order matters for blended and equal-depth fragments)."""
    import dataclasses
    (d,) = scene.draws
    assert d.pipe.topology == abi.TOPO_LIST
    tris = d.count // 3
    cuts = [tris * i // parts for i in range(parts + 1)]
    chunks = [(cuts[i], cuts[i + 1]) for i in range(parts) if cuts[i + 1] > cuts[i]]
    if order is not None:
        chunks = [chunks[i] for i in order if i < len(chunks)]
    draws = [dataclasses.replace(d, first=d.first + 3 * a, count=3 * (b - a)) for a, b in chunks]
    return dataclasses.replace(scene, name=f"{scene.name}_x{len(draws)}", draws=draws)


def c6_many_draws(width: int = 3840, height: int = 2160, draws: int = 1000, qx: int = 1000, qy: int = 500) -> Scene:
    """Diagnostic workload (not a BASELINE config): the C3 mesh recorded as `draws` indexed draws of
    consecutive index ranges, same pipeline and bindings."""
    return split_draws(c3_mesh(width, height, qx, qy), draws)


def bc_blocks(rng, w, h, corner_cases=True):
    """random 16-byte BC blocks for a w x h texture (1 byte per texel, images.cpp:31-33)"""
    blocks = rng.integers(0, 256, size=((h // 4) * (w // 4), 16), dtype=np.uint8)
    if corner_cases:
        blocks[0, 8:12] = [0x34, 0x12, 0x34, 0x12]      # color0 == color1 (BC1 three-colour mode in BC2)
        blocks[1, 8:12] = [0x00, 0x00, 0xff, 0xff]      # color0 < color1
        blocks[2, 8:12] = [0xff, 0xff, 0x00, 0x00]      # color0 > color1
        blocks[3, 0:2] = [10, 200]                      # BC3 alpha0 < alpha1: codes 6/7 -> 0/255
        blocks[4, 0:2] = [200, 10]                      # BC3 alpha0 > alpha1
        blocks[5, 0:2] = [77, 77]
    return np.ascontiguousarray(blocks.reshape(-1))
