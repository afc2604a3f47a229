// device_types.h — the flattened scene as it lives in HBM (shared by the host
// flattener and the CUDA kernels). See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

namespace rtx {

// ---- BVH: binary, both child boxes stored in the parent (one 64-byte fetch = 4 x LDG.128
// decides both children). Boxes are fp32 and rounded OUTWARD from the f64 bounds, so the
// fp32 slab test can only over-accept (the exact f64 primitive test decides). ----
struct alignas(16) BvhNode {
    float c0x[2], c0y[2];  // child 0: lo.x, hi.x, lo.y, hi.y
    float c1x[2], c1y[2];  // child 1
    float c0z[2], c1z[2];  // lo.z, hi.z of child 0, then of child 1
    int32_t child0;        // >= 0: inner node index; < 0: leaf, ~child = (first_record << 4) | count
    int32_t child1;
    int32_t _pad[2];
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");
constexpr int kTraversalStack = 64;   // per-thread traversal stack entries; the flattener refuses deeper scenes
constexpr int32_t kEmptyChild = ~0;  // leaf with count 0 (its box is inverted, never hit)

enum RecType : int32_t {
    REC_SPHERE = 0,    // d = {cx, cy, cz, r}
    REC_MSPHERE = 1,   // d = {c0x, c0y, c0z, dcx, dcy, dcz, r, t0, 1 / (t1 - t0)}
    REC_RECT_XY = 2,   // d = {a0, a1, b0, b1, k}  (axis0, axis1, k_axis) = (0,1,2)
    REC_RECT_XZ = 3,   //                                                   (0,2,1)
    REC_RECT_YZ = 4,   //                                                   (1,2,0)
    REC_BOX = 7,       // Cube: d = {lo.x, lo.y, lo.z, hi.x, hi.y, hi.z}; a = index of the first of its six rectangle
                       // records (Cube::new order), b = first primitive id
    REC_MEDIUM = 6     // a = phase texture, b = prim id, c = boundary BVH root, or -1: the boundary is the
                       // untransformed sphere d[4..7] = {cx, cy, cz, r}, or <= -2: it is the box
                       // d[4..9] = {lo, hi} in the space of chain (-2 - c);
                       // d = {neg_inv_density, medium ordinal, outer chain index}.
                       // Medium records are not BVH leaves: they are listed in SceneView::media.
};

// One 96-byte record per leaf primitive / box face / medium. Geometry is f64: the reference
// computes in f64 (vec3.rs:12) and the fixed-ray contract (t, normal, u, v within 1e-5 on
// radius-1000 and radius-5000 spheres) cannot be met by an fp32 quadratic.
struct alignas(16) Record {
    int32_t type;
    int32_t a;  // geometry: material index
    int32_t b;  // geometry: primitive id
    int32_t c;  // geometry: wrapper chain index (SceneView::chains; 0 = none, the primitive is in world space)
    double d[10];
};
static_assert(sizeof(Record) == 96, "Record must be 96 bytes");

// Transform chain ops, outermost first (the order the reference's wrappers see the ray).
enum XformKind : int32_t { XF_TRANSLATE = 0, XF_ROTATE_Y = 1 };
struct alignas(16) XformOp {
    int32_t kind;
    int32_t _pad;
    double v[3];  // TRANSLATE: offset; ROTATE_Y: {sin, cos, -}
};
static_assert(sizeof(XformOp) == 32, "XformOp must be 32 bytes");

// One entry per distinct wrapper path. Any sequence of Translate / YRotate ops composes to a
// rotation about y plus an offset: ray-into-object-space is o' = Ry(cs, sn) o + t, d' = Ry d (what
// the chain of Translate::hit / YRotate::hit computes, hittable.rs:600-604,687-692, in one step).
// The individual ops (xforms[begin .. begin + len), outermost first) are still needed on the way
// out, where the reference post-processes the hit record op by op (Q13 / Q14). Chain 0 = identity.
struct alignas(16) DChain {
    double cs, sn;
    double tx, ty, tz;
    int32_t begin, len;
    double _pad[2];
};
static_assert(sizeof(DChain) == 64, "DChain must be 64 bytes");

struct alignas(16) DMaterial {
    int32_t kind;  // rtx_material_kind
    int32_t texture;
    float albedo[3];
    float param;
    int32_t _pad[2];
};
static_assert(sizeof(DMaterial) == 32, "DMaterial must be 32 bytes");

struct alignas(16) DTexture {
    int32_t kind;  // rtx_texture_kind
    int32_t a, b;
    int32_t _pad;  // 1: evaluating this texture needs the surface (u, v)
    float f[4];
};
static_assert(sizeof(DTexture) == 32, "DTexture must be 32 bytes");

// Perlin tables in fp32: 256 gradients (xyz + pad) and three byte permutations.
struct alignas(16) DPerlin {
    float ranvec[256][4];
    uint8_t perm[3][256];
};

// ConstantMedium (hittable.rs:724-801) entries. Every ray evaluates every medium whose (fp32,
// outward-rounded) world bounds it crosses BEFORE the surface traversal: the medium's scatter
// distance is a candidate like any other hit and the closest candidate wins, which is what the
// reference's `closest`-clipped test computes whatever the list order (SURVEY.md hard parts).
struct alignas(16) DMedium {
    float lo[3];
    int32_t record;  // index of the REC_MEDIUM record
    float hi[3];
    int32_t ordinal;  // which Philox MEDIUM word this medium draws (the record's d[1], as an integer)
    // 1 / density rounded toward zero: a LOWER bound of the distance -ln(u) / density the medium lets a ray fly
    // (hittable.rs:765), for the fp32 filter that skips the f64 boundary test when the ray cannot scatter in front of
    // the surface it already hit
    float inv_density_lb;
    int32_t _pad[3];
};
static_assert(sizeof(DMedium) == 48, "DMedium must be 48 bytes");

struct DImage {
    unsigned long long tex;  // cudaTextureObject_t (0: failed load -> cyan)
    int32_t width, height;
};

// Everything a kernel needs, passed by value.
struct SceneView {
    const BvhNode* nodes;
    const Record* records;
    const XformOp* xforms;
    const DChain* chains;
    const DMaterial* materials;
    const DTexture* textures;
    const DPerlin* perlins;
    const DImage* images;
    const DMedium* media;
    int32_t world_root;
    int32_t n_media;  // > 0: rays draw one Philox MEDIUM block per 4 media
    int32_t wide_root;  // >= 0: the world BVH once more, 4-wide (a node = the two consecutive entries at this index)
};

// Camera::new precomputed on the host in f64 exactly as camera.rs:32-61 does.
struct CameraView {
    double origin[3], lower_left[3], horizontal[3], vertical[3], u[3], v[3];
    double lens_radius, time0, time1;
    float background[3];
};

}  // namespace rtx
