"""Host-side mirror of the reference's scene API (src/math/mod.rs:10-19).

Same names and argument meaning as the Rust types: `Sphere`, `MovingSphere`,
`XY/XZ/YZ.rectangle(material, a0..a1, b0..b1, k)`, `Cube.new(min, max, material)`,
`List`, `BvhTree.from_list`, `.translate(offset)`, `.rotate_y(angle)`,
`ConstantMedium.new(boundary, density, texture)`, the five materials and four
textures, `CameraDescriptor`. `Scene.to_desc()` serialises the object graph into
the plain-data `rtx_scene_desc` of include/rttnw_b200.h; shared objects (the
reference's `Arc`s) are emitted once, by identity. Nothing here computes
intersections or shading — that happens only in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import abi


def _v3(x) -> tuple:
    if isinstance(x, (int, float)):
        return (float(x),) * 3
    x = tuple(float(v) for v in x)
    assert len(x) == 3
    return x


# ----------------------------------------------------------------------------
# textures (src/math/texture.rs)
# ----------------------------------------------------------------------------
class Texture:
    pass


@dataclass(eq=False)
class SolidColor(Texture):  # `impl Texture for Vec3f<Color>`, texture.rs:9-13
    color: tuple

    def __post_init__(self):
        self.color = _v3(self.color)


@dataclass(eq=False)
class CheckerTexture(Texture):  # texture.rs:15-30
    odd: Texture
    even: Texture

    def __post_init__(self):
        self.odd, self.even = _tex(self.odd), _tex(self.even)


@dataclass(eq=False)
class NoiseTexture(Texture):  # texture.rs:32-59
    scale: float = 1.0
    seed: int = 0
    table: Optional[abi.Perlin] = None  # explicit tables win over `seed`

    @staticmethod
    def scaled(scale: float, seed: int = 0) -> "NoiseTexture":
        return NoiseTexture(scale=scale, seed=seed)


@dataclass(eq=False)
class ImageTexture(Texture):  # texture.rs:61-107
    rgba: Optional[np.ndarray] = None  # (H, W, 4) uint8; None = failed load -> cyan

    @staticmethod
    def new(path: str) -> "ImageTexture":
        """Decodes with the library's own PNG reader; an unreadable file gives cyan (texture.rs:96-99)."""
        from .render import png_read_rgba8
        try:
            return ImageTexture(png_read_rgba8(path))
        except abi.RtxError:
            return ImageTexture(None)


def _tex(t) -> Texture:
    return t if isinstance(t, Texture) else SolidColor(t)


# ----------------------------------------------------------------------------
# materials (src/math/material.rs)
# ----------------------------------------------------------------------------
class Material:
    pass


@dataclass(eq=False)
class Lambertian(Material):  # material.rs:23-100
    albedo: Texture

    def __post_init__(self):
        self.albedo = _tex(self.albedo)

    arc = classmethod(lambda cls, albedo: cls(albedo))
    boxed = arc


@dataclass(eq=False)
class Metal(Material):  # material.rs:102-149
    albedo: tuple
    fuzz: float

    def __post_init__(self):
        self.albedo = _v3(self.albedo)
        self.fuzz = min(float(self.fuzz), 1.0)

    arc = classmethod(lambda cls, albedo, fuzz: cls(albedo, fuzz))
    boxed = arc


@dataclass(eq=False)
class Dielectric(Material):  # material.rs:151-204
    refraction_index: float
    arc = classmethod(lambda cls, ir: cls(ir))
    boxed = arc


@dataclass(eq=False)
class DiffuseLight(Material):  # material.rs:206-250
    emit: Texture

    def __post_init__(self):
        self.emit = _tex(self.emit)

    arc = classmethod(lambda cls, emit: cls(emit))
    boxed = arc


@dataclass(eq=False)
class Isotropic(Material):  # material.rs:252-266
    albedo: Texture

    def __post_init__(self):
        self.albedo = _tex(self.albedo)


# ----------------------------------------------------------------------------
# hittables (src/math/hittable.rs)
# ----------------------------------------------------------------------------
class Hittable:
    def translate(self, offset) -> "Translate":  # hittable.rs:51-59
        return Translate(self, _v3(offset))

    def rotate_y(self, angle: float) -> "YRotate":  # hittable.rs:60-65
        return YRotate(self, float(angle))


@dataclass(eq=False)
class Sphere(Hittable):
    center: tuple
    radius: float
    material: Material

    def __post_init__(self):
        self.center = _v3(self.center)


@dataclass(eq=False)
class MovingSphere(Hittable):
    center: tuple  # (center_start, center_end)  — `center: Range<Vec3f>`
    time: tuple  # (time_start, time_end)
    radius: float
    material: Material

    def __post_init__(self):
        self.center = (_v3(self.center[0]), _v3(self.center[1]))
        self.time = (float(self.time[0]), float(self.time[1]))


@dataclass(eq=False)
class Rectangle(Hittable):
    plane: int  # abi.NODE_RECT_XY / XZ / YZ
    material: Material
    p0: tuple
    p1: tuple
    k: float


class _Plane:
    kind = 0

    @classmethod
    def rectangle(cls, material, p0, p1, k) -> Rectangle:  # Plane::rectangle, hittable.rs:401-411
        return Rectangle(cls.kind, material, (float(p0[0]), float(p0[1])), (float(p1[0]), float(p1[1])), float(k))


class XY(_Plane):
    kind = abi.NODE_RECT_XY


class XZ(_Plane):
    kind = abi.NODE_RECT_XZ


class YZ(_Plane):
    kind = abi.NODE_RECT_YZ


Xy, Xz, Yz = XY, XZ, YZ  # the reference spells both


@dataclass(eq=False)
class Cube(Hittable):
    box_min: tuple
    box_max: tuple
    material: Material

    def __post_init__(self):
        self.box_min, self.box_max = _v3(self.box_min), _v3(self.box_max)

    new = classmethod(lambda cls, a, b, m: cls(a, b, m))


@dataclass(eq=False)
class List(Hittable):
    list: list = field(default_factory=list)

    def push(self, item: Hittable) -> None:
        self.list.append(item)

    new = classmethod(lambda cls: cls())
    with_capacity = classmethod(lambda cls, n: cls())


@dataclass(eq=False)
class BvhTree(Hittable):
    items: List

    from_list = classmethod(lambda cls, l: cls(l))  # `BvhTree::from(list)`


@dataclass(eq=False)
class Translate(Hittable):
    item: Hittable
    offset: tuple


@dataclass(eq=False)
class YRotate(Hittable):
    item: Hittable
    angle: float


@dataclass(eq=False)
class ConstantMedium(Hittable):
    boundary: Hittable
    density: float
    phase_function: Texture

    def __post_init__(self):
        self.phase_function = _tex(self.phase_function)

    new = classmethod(lambda cls, b, d, t: cls(b, d, t))


@dataclass
class CameraDescriptor:  # camera.rs:5-15
    lookfrom: tuple = (0.0, 0.0, 0.0)
    lookat: tuple = (0.0, 0.0, -1.0)
    view_up: tuple = (0.0, 1.0, 0.0)
    vertical_fov: float = 40.0
    aspect_ratio: float = 1.0
    aperture: float = 0.0
    focus_distance: float = 10.0
    open_time: float = 0.0
    close_time: float = 1.0


def perlin_table(seed: int) -> abi.Perlin:
    """Perlin::new() (noise.rs:12-47) with SplitMix64(seed) standing in for thread_rng()."""
    mask = (1 << 64) - 1
    state = seed & mask

    def nxt():
        nonlocal state
        state = (state + 0x9E3779B97F4A7C15) & mask
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & mask
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & mask
        return z ^ (z >> 31)

    def gen():
        return (nxt() >> 11) * (1.0 / 9007199254740992.0)

    t = abi.Perlin()
    for i in range(256):
        for c in range(3):
            t.ranvec[i][c] = -1.0 + 2.0 * gen()
    for perm in (t.perm_x, t.perm_y, t.perm_z):
        p = list(range(256))
        for i in range(255, 0, -1):
            j = int(gen() * (i + 1))
            p[i], p[j] = p[j], p[i]
        for i in range(256):
            perm[i] = p[i]
    return t


class SceneDescHandle:
    """Owns the ctypes arrays an rtx_scene_desc points into."""

    def __init__(self, desc: abi.SceneDesc, keep: list):
        self.desc = desc
        self._keep = keep

    def __getattr__(self, name):
        return getattr(self.desc, name)


@dataclass
class Scene:  # `struct Scene`, main.rs:47-55
    world: Hittable
    camera: CameraDescriptor = field(default_factory=CameraDescriptor)
    background: tuple = (0.0, 0.0, 0.0)

    def to_desc(self) -> SceneDescHandle:
        nodes: list = []
        children: list = []
        mats: list = []
        texs: list = []
        perlins: list = []
        images: list = []
        keep: list = []
        node_ids: dict = {}
        mat_ids: dict = {}
        tex_ids: dict = {}

        def tex_id(t: Texture) -> int:
            if id(t) in tex_ids:
                return tex_ids[id(t)]
            r = abi.Texture()
            if isinstance(t, SolidColor):
                r.kind = abi.TEX_SOLID
                r.f[0], r.f[1], r.f[2] = t.color
            elif isinstance(t, CheckerTexture):
                r.kind = abi.TEX_CHECKER
                r.a, r.b = tex_id(t.odd), tex_id(t.even)
            elif isinstance(t, NoiseTexture):
                r.kind = abi.TEX_NOISE
                r.a = len(perlins)
                perlins.append(t.table if t.table is not None else perlin_table(t.seed))
                r.f[0] = t.scale
            elif isinstance(t, ImageTexture):
                r.kind = abi.TEX_IMAGE
                r.a = len(images)
                img = abi.Image()
                if t.rgba is not None:
                    arr = np.ascontiguousarray(t.rgba, dtype=np.uint8)
                    assert arr.ndim == 3 and arr.shape[2] == 4
                    keep.append(arr)
                    img.height, img.width = arr.shape[0], arr.shape[1]
                    img.rgba = arr.ctypes.data_as(C.POINTER(C.c_uint8))
                images.append(img)
            else:
                raise TypeError(f"not a texture: {t!r}")
            texs.append(r)
            tex_ids[id(t)] = len(texs) - 1
            keep.append(t)
            return len(texs) - 1

        def mat_id(m: Material) -> int:
            if id(m) in mat_ids:
                return mat_ids[id(m)]
            r = abi.Material()
            r.texture = -1
            if isinstance(m, Lambertian):
                r.kind, r.texture = abi.MAT_LAMBERTIAN, tex_id(m.albedo)
            elif isinstance(m, Metal):
                r.kind, r.param = abi.MAT_METAL, m.fuzz
                r.albedo[0], r.albedo[1], r.albedo[2] = m.albedo
            elif isinstance(m, Dielectric):
                r.kind, r.param = abi.MAT_DIELECTRIC, m.refraction_index
            elif isinstance(m, DiffuseLight):
                r.kind, r.texture = abi.MAT_DIFFUSE_LIGHT, tex_id(m.emit)
            elif isinstance(m, Isotropic):
                r.kind, r.texture = abi.MAT_ISOTROPIC, tex_id(m.albedo)
            else:
                raise TypeError(f"not a material: {m!r}")
            mats.append(r)
            mat_ids[id(m)] = len(mats) - 1
            keep.append(m)
            return len(mats) - 1

        def node_id(h: Hittable) -> int:
            if id(h) in node_ids:
                return node_ids[id(h)]
            n = abi.Node()
            n.material, n.child, n.n_children = -1, -1, 0
            if isinstance(h, Sphere):
                n.kind, n.material = abi.NODE_SPHERE, mat_id(h.material)
                n.f[0], n.f[1], n.f[2], n.f[3] = (*h.center, h.radius)
            elif isinstance(h, MovingSphere):
                n.kind, n.material = abi.NODE_MOVING_SPHERE, mat_id(h.material)
                vals = (*h.center[0], *h.center[1], h.radius, *h.time)
                for i, v in enumerate(vals):
                    n.f[i] = v
            elif isinstance(h, Rectangle):
                n.kind, n.material = h.plane, mat_id(h.material)
                for i, v in enumerate((*h.p0, *h.p1, h.k)):
                    n.f[i] = v
            elif isinstance(h, Cube):
                n.kind, n.material = abi.NODE_CUBE, mat_id(h.material)
                for i, v in enumerate((*h.box_min, *h.box_max)):
                    n.f[i] = v
            elif isinstance(h, (List, BvhTree)):
                items = h.list if isinstance(h, List) else h.items.list
                n.kind = abi.NODE_LIST if isinstance(h, List) else abi.NODE_BVH
                ids = [node_id(c) for c in items]
                n.child, n.n_children = len(children), len(ids)
                children.extend(ids)
            elif isinstance(h, Translate):
                n.kind, n.child = abi.NODE_TRANSLATE, node_id(h.item)
                n.f[0], n.f[1], n.f[2] = h.offset
            elif isinstance(h, YRotate):
                n.kind, n.child = abi.NODE_ROTATE_Y, node_id(h.item)
                n.f[0] = h.angle
            elif isinstance(h, ConstantMedium):
                n.kind, n.child = abi.NODE_MEDIUM, node_id(h.boundary)
                n.material = tex_id(h.phase_function)
                n.f[0] = h.density
            else:
                raise TypeError(f"not a hittable: {h!r}")
            nodes.append(n)
            node_ids[id(h)] = len(nodes) - 1
            keep.append(h)
            return len(nodes) - 1

        root = node_id(self.world)
        d = abi.SceneDesc()

        def arr(ctype, items):
            a = (ctype * max(1, len(items)))(*items)
            keep.append(a)
            return a

        d.nodes, d.n_nodes, d.root = arr(abi.Node, nodes), len(nodes), root
        d.children, d.n_children = arr(C.c_int32, children), len(children)
        d.materials, d.n_materials = arr(abi.Material, mats), len(mats)
        d.textures, d.n_textures = arr(abi.Texture, texs), len(texs)
        d.perlins, d.n_perlins = arr(abi.Perlin, perlins), len(perlins)
        d.images, d.n_images = arr(abi.Image, images), len(images)
        d.background[0], d.background[1], d.background[2] = _v3(self.background)
        cam = self.camera
        for i in range(3):
            d.camera.lookfrom[i] = cam.lookfrom[i]
            d.camera.lookat[i] = cam.lookat[i]
            d.camera.view_up[i] = cam.view_up[i]
        d.camera.vertical_fov, d.camera.aspect_ratio = cam.vertical_fov, cam.aspect_ratio
        d.camera.aperture, d.camera.focus_distance = cam.aperture, cam.focus_distance
        d.camera.open_time, d.camera.close_time = cam.open_time, cam.close_time
        return SceneDescHandle(d, keep)
