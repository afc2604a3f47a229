//! The reference's scene API surface (`src/math/mod.rs:10-19`) as a host-side mirror whose objects DESCRIBE
//! themselves to the device instead of intersecting rays on the CPU.
//!
//! Same names, same constructors, same field names as the reference, so that its `src/scenes.rs` compiles against
//! this module unchanged (`use rttnw_b200_sys::math::{...}` instead of `use crate::math::{...}`) — including the
//! `XY, XZ, YZ` spellings `scenes.rs:6` imports although `math/mod.rs:13` exports `Xy, Xz, Yz` (the reason HEAD of
//! the reference does not type-check): both spellings exist here.
//!
//! What differs from the reference: `Hittable::hit`, `Material::scatter` and `Texture::value` — the hot path — are
//! not here at all. Their replacement is one method per trait, `describe`, which writes the object as a plain
//! `rtx_node` / `rtx_material` / `rtx_texture` (include/rttnw_b200.h) into a `SceneBuilder`; the CUDA library
//! flattens that description, builds its own BVH and answers `hit` / `scatter` / `value` on the GPU. Shared objects
//! (`Arc`) are described once, keyed by pointer identity, like the reference shares them by reference count.
//!
//! NOT BUILT in this repository's image (no Rust toolchain): reviewed source only. tests/test_host_cpu.py checks the
//! struct layouts of `ffi.rs` against the C header's ctypes view field by field.
use std::marker::PhantomData;
use std::ops::{Add, Mul, Neg, Range, Sub};
use std::path::Path;
use std::sync::Arc;

use rand::seq::SliceRandom;
use rand::Rng;

use crate::ffi;
use crate::SceneBuilder;

// ---------------------------------------------------------------------------------------------
// Vec3f<T> (src/math/vec3.rs:12-141,185-260): only what scene construction touches
// ---------------------------------------------------------------------------------------------
pub trait Phantom {}
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct Color;
#[derive(Clone, Copy, Debug, Default, PartialEq)]
pub struct Position;
impl Phantom for Color {}
impl Phantom for Position {}

#[derive(Clone, Copy, Debug, Eq, PartialEq)]
pub enum Coordinate {
    X = 0,
    Y = 1,
    Z = 2,
}

#[derive(Debug, PartialEq)]
pub struct Vec3f<T> {
    items: [f64; 3],
    _phantom: PhantomData<T>,
}
impl<T: Phantom> Copy for Vec3f<T> {}
impl<T: Phantom> Clone for Vec3f<T> {
    fn clone(&self) -> Self {
        *self
    }
}
impl<T: Phantom> Default for Vec3f<T> {
    fn default() -> Self {
        Self::repeat(0.0)
    }
}
impl<T: Phantom> Vec3f<T> {
    pub fn new(x: f64, y: f64, z: f64) -> Self {
        Self { items: [x, y, z], _phantom: PhantomData }
    }
    pub fn repeat(x: f64) -> Self {
        Self::new(x, x, x)
    }
    /// vec3.rs:58-64: three draws of `thread_rng().gen_range(range)`
    pub fn random(range: Range<f64>) -> Self {
        let mut rng = rand::thread_rng();
        Self::new(rng.gen_range(range.clone()), rng.gen_range(range.clone()), rng.gen_range(range))
    }
    pub fn x(&self) -> f64 {
        self.items[0]
    }
    pub fn y(&self) -> f64 {
        self.items[1]
    }
    pub fn z(&self) -> f64 {
        self.items[2]
    }
    pub fn magnitude(&self) -> f64 {
        (self.items[0] * self.items[0] + self.items[1] * self.items[1] + self.items[2] * self.items[2]).sqrt()
    }
    pub fn to_array(self) -> [f64; 3] {
        self.items
    }
}
impl<T: Phantom> From<(f64, f64, f64)> for Vec3f<T> {
    fn from(t: (f64, f64, f64)) -> Self {
        Self::new(t.0, t.1, t.2)
    }
}
impl<T: Phantom> Add for Vec3f<T> {
    type Output = Self;
    fn add(self, r: Self) -> Self {
        Self::new(self.items[0] + r.items[0], self.items[1] + r.items[1], self.items[2] + r.items[2])
    }
}
impl<T: Phantom> Sub for Vec3f<T> {
    type Output = Self;
    fn sub(self, r: Self) -> Self {
        Self::new(self.items[0] - r.items[0], self.items[1] - r.items[1], self.items[2] - r.items[2])
    }
}
impl<T: Phantom> Mul<f64> for Vec3f<T> {
    type Output = Self;
    fn mul(self, r: f64) -> Self {
        Self::new(self.items[0] * r, self.items[1] * r, self.items[2] * r)
    }
}
impl<T: Phantom> Neg for Vec3f<T> {
    type Output = Self;
    fn neg(self) -> Self {
        Self::new(-self.items[0], -self.items[1], -self.items[2])
    }
}

// ---------------------------------------------------------------------------------------------
// Textures (src/math/texture.rs, noise.rs)
// ---------------------------------------------------------------------------------------------
/// `Texture::value` (texture.rs:5-7) runs on the device; the host side only describes.
pub trait Texture: Send + Sync {
    /// Writes this texture as an `rtx_texture`; returns its index.
    fn describe(&self, b: &mut SceneBuilder) -> i32;
}

/// A texture behind an `Arc`, described once however many materials share it.
fn shared_texture<T: Texture + ?Sized>(t: &Arc<T>, b: &mut SceneBuilder) -> i32 {
    let key = Arc::as_ptr(t) as *const u8;
    if let Some(i) = b.seen(2, key) {
        return i;
    }
    let i = t.describe(b);
    b.remember(2, key, i)
}

/// `impl Texture for Vec3f<Color>` (texture.rs:9-13): a solid colour.
impl Texture for Vec3f<Color> {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        b.solid(self.to_array())
    }
}

/// texture.rs:15-30 (Q21): `sin(10x) sin(10y) sin(10z) < 0` picks `odd`.
pub struct CheckerTexture {
    pub odd: Arc<dyn Texture>,
    pub even: Arc<dyn Texture>,
}
impl Texture for CheckerTexture {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let odd = shared_texture(&self.odd, b);
        let even = shared_texture(&self.even, b);
        b.texture(ffi::RTX_TEX_CHECKER, odd, even, [0.0; 4])
    }
}

/// noise.rs:5-47: 256 gradients from `Vec3f::random(-1..1)` (not normalised, Q22) and three shuffled permutations.
/// The tables are drawn on the host exactly as the reference draws them and handed to the device as they are.
pub struct Perlin {
    random_points: Vec<Vec3f<Position>>,
    x: Vec<usize>,
    y: Vec<usize>,
    z: Vec<usize>,
}
impl Perlin {
    const POINT_COUNT: usize = 256;
    fn generate_permutation() -> Vec<usize> {
        let mut points: Vec<usize> = (0..Self::POINT_COUNT).collect();
        points.shuffle(&mut rand::thread_rng());
        points
    }
    pub fn new() -> Self {
        Self {
            random_points: (0..Self::POINT_COUNT).map(|_| Vec3f::random(-1.0..1.0)).collect(),
            x: Self::generate_permutation(),
            y: Self::generate_permutation(),
            z: Self::generate_permutation(),
        }
    }
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let ranvec: Vec<[f64; 3]> = self.random_points.iter().map(|p| p.to_array()).collect();
        b.perlin(&ranvec, &self.x, &self.y, &self.z)
    }
}
impl Default for Perlin {
    fn default() -> Self {
        Self::new()
    }
}

/// texture.rs:32-59: grey `0.5 (1 + sin(scale z + 10 turbulence(p, 7)))`.
pub struct NoiseTexture {
    noise: Perlin,
    scale: f64,
}
impl NoiseTexture {
    pub fn new() -> Self {
        Self { noise: Perlin::new(), scale: 1.0 }
    }
    pub fn scaled(scale: f64) -> Self {
        Self { noise: Perlin::new(), scale }
    }
}
impl Default for NoiseTexture {
    fn default() -> Self {
        Self::new()
    }
}
impl Texture for NoiseTexture {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let perlin = self.noise.describe(b);
        b.texture(ffi::RTX_TEX_NOISE, perlin, 0, [self.scale, 0.0, 0.0, 0.0])
    }
}

/// texture.rs:61-107 (Q23): nearest texel of an RGBA8 image; a file that does not load is cyan, not an error.
pub struct ImageTexture {
    data: Option<image::RgbaImage>,
}
impl ImageTexture {
    pub fn new<T: AsRef<Path>>(file: T) -> Self {
        let data = image::io::Reader::open(file).ok().and_then(|x| x.decode().map(|x| x.to_rgba8()).ok());
        Self { data }
    }
}
impl Texture for ImageTexture {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let image = match &self.data {
            Some(img) => b.image(img.width(), img.height(), Some(img.as_raw().clone())),
            None => b.image(0, 0, None),
        };
        b.texture(ffi::RTX_TEX_IMAGE, image, 0, [0.0; 4])
    }
}

// ---------------------------------------------------------------------------------------------
// Materials (src/math/material.rs)
// ---------------------------------------------------------------------------------------------
/// `Material::scatter` / `emitted` (material.rs:6-21) run on the device; the host side only describes.
pub trait Material: Send + Sync {
    /// Writes this material as an `rtx_material`; returns its index.
    fn describe(&self, b: &mut SceneBuilder) -> i32;
    fn arc(self) -> Arc<Self>
    where
        Self: Sized,
    {
        Arc::new(self)
    }
    fn boxed(self) -> Box<Self>
    where
        Self: Sized,
    {
        Box::new(self)
    }
}
fn shared_material<M: Material + ?Sized>(m: &Arc<M>, b: &mut SceneBuilder) -> i32 {
    let key = Arc::as_ptr(m) as *const u8;
    if let Some(i) = b.seen(1, key) {
        return i;
    }
    let i = m.describe(b);
    b.remember(1, key, i)
}

/// material.rs:23-100 (Q2: scattered direction = normal + a point IN the unit ball).
#[derive(Clone)]
pub struct Lambertian<T: Texture> {
    albedo: Arc<T>,
}
impl<T: 'static + Texture> From<T> for Lambertian<T> {
    fn from(albedo: T) -> Self {
        Self { albedo: Arc::new(albedo) }
    }
}
impl<T: 'static + Texture> Lambertian<T> {
    pub fn new<A: Into<Arc<T>>>(albedo: A) -> Self {
        Self { albedo: albedo.into() }
    }
    pub fn boxed<A: Into<Arc<T>>>(albedo: A) -> Box<Self> {
        Box::new(Self::new(albedo))
    }
    pub fn arc<A: Into<Arc<T>>>(albedo: A) -> Arc<Self> {
        Arc::new(Self::new(albedo))
    }
}
impl<T: 'static + Texture> Material for Lambertian<T> {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let texture = shared_texture(&self.albedo, b);
        b.material(ffi::RTX_MAT_LAMBERTIAN, texture, [0.0; 3], 0.0)
    }
}

/// material.rs:102-149 (Q4: fuzz clamped to 1; absorbed when the fuzzed reflection points into the surface).
#[derive(Clone, Copy)]
pub struct Metal {
    albedo: Vec3f<Color>,
    fuzz: f64,
}
impl Metal {
    pub fn new(albedo: Vec3f<Color>, fuzz: f64) -> Self {
        Self { albedo, fuzz: fuzz.min(1.0) }
    }
    pub fn boxed(albedo: Vec3f<Color>, fuzz: f64) -> Box<Self> {
        Box::new(Self::new(albedo, fuzz))
    }
    pub fn arc(albedo: Vec3f<Color>, fuzz: f64) -> Arc<Self> {
        Arc::new(Self::new(albedo, fuzz))
    }
}
impl Material for Metal {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        b.material(ffi::RTX_MAT_METAL, -1, self.albedo.to_array(), self.fuzz)
    }
}

/// material.rs:151-204 (Q5, Q6).
#[derive(Clone, Copy)]
pub struct Dielectric {
    refraction_index: f64,
}
impl Dielectric {
    pub fn new(refraction_index: f64) -> Self {
        Self { refraction_index }
    }
    pub fn arc(refraction_index: f64) -> Arc<Self> {
        Arc::new(Self { refraction_index })
    }
    pub fn boxed(refraction_index: f64) -> Box<Self> {
        Box::new(Self { refraction_index })
    }
}
impl Material for Dielectric {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        b.material(ffi::RTX_MAT_DIELECTRIC, -1, [0.0; 3], self.refraction_index)
    }
}

/// material.rs:206-250 (Q24: emits on both sides, never scatters).
#[derive(Clone)]
pub struct DiffuseLight {
    emit: Arc<dyn Texture>,
}
impl<T: 'static + Texture> From<T> for DiffuseLight {
    fn from(albedo: T) -> Self {
        Self { emit: Arc::new(albedo) }
    }
}
impl DiffuseLight {
    pub fn new<T: 'static + Texture>(albedo: &Arc<T>) -> Self {
        Self { emit: albedo.clone() }
    }
    pub fn boxed<T: 'static + Texture>(albedo: T) -> Box<Self> {
        Box::new(Self { emit: Arc::new(albedo) })
    }
    pub fn arc<T: 'static + Texture>(albedo: T) -> Arc<Self> {
        Arc::new(Self { emit: Arc::new(albedo) })
    }
}
impl Material for DiffuseLight {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let texture = shared_texture(&self.emit, b);
        b.material(ffi::RTX_MAT_DIFFUSE_LIGHT, texture, [0.0; 3], 0.0)
    }
}

/// material.rs:252-266 (Q3: a uniformly random direction from the unit ball).
pub struct Isotropic {
    pub albedo: Arc<dyn Texture>,
}
impl Material for Isotropic {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let texture = shared_texture(&self.albedo, b);
        b.material(ffi::RTX_MAT_ISOTROPIC, texture, [0.0; 3], 0.0)
    }
}

// ---------------------------------------------------------------------------------------------
// Hittables (src/math/hittable.rs)
// ---------------------------------------------------------------------------------------------
/// `Hittable::hit` / `bounding_box` (hittable.rs:47-50) run on the device (which builds its own BVH over world-space
/// bounds); the host side only describes. `translate` / `rotate_y` are the reference's (hittable.rs:51-65).
pub trait Hittable: Send + Sync {
    /// Writes this object as an `rtx_node` (children first); returns its index.
    fn describe(&self, b: &mut SceneBuilder) -> i32;
    fn translate(self, offset: Vec3f<Position>) -> Translate
    where
        Self: 'static + Sized,
    {
        Translate { item: Box::new(self), offset }
    }
    fn rotate_y(self, angle: f64) -> YRotate
    where
        Self: 'static + Sized,
    {
        YRotate::new(Box::new(self), angle)
    }
}
fn shared_hittable(h: &Arc<dyn Hittable>, b: &mut SceneBuilder) -> i32 {
    let key = Arc::as_ptr(h) as *const u8;
    if let Some(i) = b.seen(0, key) {
        return i;
    }
    let i = h.describe(b);
    b.remember(0, key, i)
}
fn f10(values: &[f64]) -> [f64; 10] {
    let mut f = [0.0; 10];
    f[..values.len()].copy_from_slice(values);
    f
}

/// hittable.rs:68-131 (Q9: both ends of [t_min, t_max] inclusive; u, v from the outward normal).
#[derive(Clone)]
pub struct Sphere {
    pub center: Vec3f<Position>,
    pub radius: f64,
    pub material: Arc<dyn Material>,
}
impl Hittable for Sphere {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let m = shared_material(&self.material, b);
        b.node(ffi::RTX_NODE_SPHERE, m, -1, f10(&[self.center.x(), self.center.y(), self.center.z(), self.radius]))
    }
}

/// hittable.rs:179-245 (Q10: centre moves linearly over `time`; u = v = 0).
pub struct MovingSphere {
    pub center: Range<Vec3f<Position>>,
    pub time: Range<f64>,
    pub radius: f64,
    pub material: Box<dyn Material>,
}
impl Hittable for MovingSphere {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let m = self.material.describe(b); // a Box is never shared
        let (c0, c1) = (self.center.start, self.center.end);
        b.node(ffi::RTX_NODE_MOVING_SPHERE, m, -1,
               f10(&[c0.x(), c0.y(), c0.z(), c1.x(), c1.y(), c1.z(), self.radius, self.time.start, self.time.end]))
    }
}

/// hittable.rs:133-177: a linear scan on the CPU; on the device its items join the BVH of whatever encloses it.
#[derive(Default)]
pub struct List {
    pub list: Vec<Box<dyn 'static + Hittable>>,
}
impl List {
    pub fn push<T: 'static + Hittable>(&mut self, item: T) {
        self.list.push(Box::new(item))
    }
    pub fn new() -> Self {
        Self { list: Vec::new() }
    }
    pub fn with_capacity(capacity: usize) -> Self {
        Self { list: Vec::with_capacity(capacity) }
    }
}
impl Hittable for List {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let items: Vec<i32> = self.list.iter().map(|i| i.describe(b)).collect();
        b.list(ffi::RTX_NODE_LIST, &items)
    }
}

/// hittable.rs:247-373. The reference's tree (random axes, `remove(0)`, Q17/Q18) is an acceleration structure, not
/// an answer: only the closest hit over the items is the contract, and the device builds its own BVH (binned SAH, or
/// LBVH on the GPU). So the mirror keeps the LIST it was built from — the one field the real `BvhTree` would gain.
pub struct BvhTree {
    items: List,
}
impl From<List> for BvhTree {
    fn from(list: List) -> Self {
        Self::from_time(list, 0., 1.)
    }
}
impl BvhTree {
    pub fn from_time(list: List, _initial_time: f64, _final_time: f64) -> Self {
        Self { items: list }
    }
}
impl Hittable for BvhTree {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let items: Vec<i32> = self.items.list.iter().map(|i| i.describe(b)).collect();
        b.list(ffi::RTX_NODE_BVH, &items)
    }
}

/// hittable.rs:375-433: the plane a `Rectangle` lies in, as a type.
#[derive(Copy, Clone, Eq, PartialEq)]
pub struct PlaneCoordinates {
    pub axis0: Coordinate,
    pub axis1: Coordinate,
    pub k: Coordinate,
}
pub trait Plane: Send + Sync {
    fn points(p0: Vec3f<Position>, p1: Vec3f<Position>) -> (Range<f64>, Range<f64>, f64, f64);
    fn axes() -> PlaneCoordinates;
    /// Which `rtx_node_kind` a rectangle in this plane is.
    fn node_kind() -> i32;
    fn rectangle<M: Material>(material: Arc<M>, p0: Range<f64>, p1: Range<f64>, k: f64) -> Rectangle<M, Self>
    where
        Self: Sized,
    {
        Rectangle::new(material, p0, p1, k)
    }
    fn rectangles<M: Material>(p0: Vec3f<Position>, p1: Vec3f<Position>, material: &Arc<M>) -> (Rectangle<M, Self>, Rectangle<M, Self>)
    where
        Self: Sized,
    {
        let (r0, r1, k0, k1) = Self::points(p0, p1);
        (Rectangle::new(material.clone(), r0.clone(), r1.clone(), k0), Rectangle::new(material.clone(), r0, r1, k1))
    }
}
pub struct Xy(());
pub struct Xz(());
pub struct Yz(());
/// `src/scenes.rs:6` imports these spellings, `src/math/mod.rs:13` exports the ones above (Q28): both work here.
pub use self::Xy as XY;
pub use self::Xz as XZ;
pub use self::Yz as YZ;
impl Plane for Xy {
    fn points(p0: Vec3f<Position>, p1: Vec3f<Position>) -> (Range<f64>, Range<f64>, f64, f64) {
        (p0.x()..p1.x(), p0.y()..p1.y(), p0.z(), p1.z())
    }
    fn axes() -> PlaneCoordinates {
        PlaneCoordinates { axis0: Coordinate::X, axis1: Coordinate::Y, k: Coordinate::Z }
    }
    fn node_kind() -> i32 {
        ffi::RTX_NODE_RECT_XY
    }
}
impl Plane for Xz {
    fn points(p0: Vec3f<Position>, p1: Vec3f<Position>) -> (Range<f64>, Range<f64>, f64, f64) {
        (p0.x()..p1.x(), p0.z()..p1.z(), p0.y(), p1.y())
    }
    fn axes() -> PlaneCoordinates {
        PlaneCoordinates { axis0: Coordinate::X, axis1: Coordinate::Z, k: Coordinate::Y }
    }
    fn node_kind() -> i32 {
        ffi::RTX_NODE_RECT_XZ
    }
}
impl Plane for Yz {
    fn points(p0: Vec3f<Position>, p1: Vec3f<Position>) -> (Range<f64>, Range<f64>, f64, f64) {
        (p0.y()..p1.y(), p0.z()..p1.z(), p0.x(), p1.x())
    }
    fn axes() -> PlaneCoordinates {
        PlaneCoordinates { axis0: Coordinate::Y, axis1: Coordinate::Z, k: Coordinate::X }
    }
    fn node_kind() -> i32 {
        ffi::RTX_NODE_RECT_YZ
    }
}

/// hittable.rs:434-547 (Q11: t inclusive, the two ranges half-open, normal = +k axis flipped against the ray).
pub struct Rectangle<M: Material, P: Plane> {
    pub material: Arc<M>,
    pub p0: Range<f64>,
    pub p1: Range<f64>,
    pub k: f64,
    _phantom: PhantomData<P>,
}
impl<M: Material, P: Plane> Rectangle<M, P> {
    pub fn new(material: Arc<M>, p0: Range<f64>, p1: Range<f64>, k: f64) -> Self {
        Self { material, p0, p1, k, _phantom: PhantomData }
    }
}
impl<M: 'static + Material, P: Plane> Hittable for Rectangle<M, P> {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let m = shared_material(&self.material, b);
        b.node(P::node_kind(), m, -1, f10(&[self.p0.start, self.p0.end, self.p1.start, self.p1.end, self.k]))
    }
}

/// hittable.rs:549-592 (Q12: six rectangles in the order xy(min.z), xy(max.z), xz(min.y), xz(max.y), yz(min.x),
/// yz(max.x); the device keeps that order for the primitive ids and tests the box as one slab computation).
pub struct Cube {
    box_min: Vec3f<Position>,
    box_max: Vec3f<Position>,
    material: Arc<dyn Material>,
}
impl Cube {
    pub fn new<T: 'static + Material>(box_min: Vec3f<Position>, box_max: Vec3f<Position>, material: Arc<T>) -> Self {
        Self { box_min, box_max, material }
    }
}
impl Hittable for Cube {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let m = shared_material(&self.material, b);
        let (lo, hi) = (self.box_min, self.box_max);
        b.node(ffi::RTX_NODE_CUBE, m, -1, f10(&[lo.x(), lo.y(), lo.z(), hi.x(), hi.y(), hi.z()]))
    }
}

/// hittable.rs:594-629 (Q13: face_normal is run a second time on the already flipped normal).
pub struct Translate {
    pub item: Box<dyn Hittable>,
    pub offset: Vec3f<Position>,
}
impl Hittable for Translate {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let child = self.item.describe(b);
        b.node(ffi::RTX_NODE_TRANSLATE, -1, child, f10(&[self.offset.x(), self.offset.y(), self.offset.z()]))
    }
}

/// hittable.rs:631-722 (Q14: the hit point and normal are rotated back with a sequential update — [0] is overwritten
/// before it feeds [2]; the device reproduces that; Q15: the bounds `new` computes are not used for culling there
/// either). The angle is kept in degrees: the device recomputes sin / cos like :641-645.
pub struct YRotate {
    item: Box<dyn Hittable>,
    angle: f64,
}
impl YRotate {
    pub fn new(item: Box<dyn Hittable>, angle: f64) -> Self {
        Self { item, angle }
    }
}
impl Hittable for YRotate {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let child = self.item.describe(b);
        b.node(ffi::RTX_NODE_ROTATE_Y, -1, child, f10(&[self.angle]))
    }
}

/// hittable.rs:724-801 (Q16: free flight `-ln(u) / density` inside a convex boundary; the hit record carries an
/// arbitrary normal and the medium's own `Isotropic`).
pub struct ConstantMedium {
    boundary: Arc<dyn Hittable>,
    phase_function: Isotropic,
    density: f64,
}
impl ConstantMedium {
    pub fn new(boundary: Arc<dyn Hittable>, density: f64, phase_function: Arc<dyn Texture>) -> Self {
        Self { boundary, phase_function: Isotropic { albedo: phase_function }, density }
    }
}
impl Hittable for ConstantMedium {
    fn describe(&self, b: &mut SceneBuilder) -> i32 {
        let boundary = shared_hittable(&self.boundary, b);
        let texture = shared_texture(&self.phase_function.albedo, b);
        b.node(ffi::RTX_NODE_MEDIUM, texture, boundary, f10(&[self.density]))
    }
}

// ---------------------------------------------------------------------------------------------
// Camera (src/math/camera.rs:5-15): the descriptor is handed over as it is; Camera::new runs on the host side of the
// library in f64 exactly as camera.rs:32-61 does.
// ---------------------------------------------------------------------------------------------
#[derive(Default)]
pub struct CameraDescriptor {
    pub lookfrom: Vec3f<Position>,
    pub lookat: Vec3f<Position>,
    pub view_up: Vec3f<Position>,
    pub vertical_fov: f64,
    pub aspect_ratio: f64,
    pub aperture: f64,
    pub focus_distance: f64,
    pub open_time: f64,
    pub close_time: f64,
}
impl CameraDescriptor {
    pub fn to_ffi(&self) -> ffi::rtx_camera {
        ffi::rtx_camera {
            lookfrom: self.lookfrom.to_array(),
            lookat: self.lookat.to_array(),
            view_up: self.view_up.to_array(),
            vertical_fov: self.vertical_fov,
            aspect_ratio: self.aspect_ratio,
            aperture: self.aperture,
            focus_distance: self.focus_distance,
            open_time: self.open_time,
            close_time: self.close_time,
        }
    }
}
