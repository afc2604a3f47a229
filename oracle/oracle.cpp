// oracle.cpp — CPU oracle for the rttnw path-tracing hot path. TEST INFRASTRUCTURE ONLY
// (see oracle.h for who may load it and for what the parity is pinned against).
//
// Every function restates, in C++17 / f64, what the cited lines of the reference
// (luliic2/rttnw, paths relative to /root/reference) compute — including the
// behaviours listed in SURVEY.md §2.3 (Q1..Q28). It deliberately keeps the
// reference's *structure*: a trait-object tree walked by virtual calls, a
// recursive `color`, a flat `List` with a shrinking `closest`, the reference's
// BVH construction. Two things are injected because the reference takes them
// from an OS-seeded thread_rng(): the scene randomness (SplitMix64 from a seed)
// and the per-sample randomness (Philox4x32-10 counters, shared with the GPU).
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace orc {

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11; Random123). Replaces rand 0.8.3's
// thread_rng() (ChaCha12, OS seeded — unreproducible by design).
// ---------------------------------------------------------------------------
static inline void philox4x32_10(const uint32_t c_in[4], const uint32_t k_in[2], uint32_t out[4]) {
    uint32_t c0 = c_in[0], c1 = c_in[1], c2 = c_in[2], c3 = c_in[3];
    uint32_t k0 = k_in[0], k1 = k_in[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// 24-bit uniform in [0,1): exactly representable in fp32, so the CUDA kernels
// and the oracle see the same variates. (rand's gen::<f64>() has 53 bits.)
static inline double u01(uint32_t w) { return (double)(w >> 8) * (1.0 / 16777216.0); }

enum Purpose : uint32_t { P_CAMERA = 0, P_LENS = 1, P_SCATTER = 2, P_MEDIUM = 3 };

// Per-thread sampling context: stands in for `rand::thread_rng()`.
struct Sampler {
    bool fixed = true;   // fixed-ray mode: media draw `xi`; no other draw is legal
    double xi = 0.5;
    uint32_t key[2] = {0, 0};
    uint32_t pixel = 0, sample = 0;
    uint32_t bounce = 0;
    uint64_t rays = 0;
    void block(Purpose p, uint32_t j, uint32_t out[4]) const {
        uint32_t ctr[4] = {pixel, sample, (bounce << 8) | (uint32_t)p, j};
        philox4x32_10(ctr, key, out);
    }
    double medium_u(int medium_id) const {
        if (fixed) return xi;
        uint32_t w[4];
        block(P_MEDIUM, (uint32_t)medium_id >> 2, w);
        return u01(w[medium_id & 3]);
    }
    double dielectric_u() const {
        uint32_t w[4];
        block(P_SCATTER, 0, w);
        return u01(w[3]);
    }
};
static thread_local Sampler g_sampler;

// Fragility probe: records whether any comparison the reference makes while answering a fixed
// ray was within 1e-9 relative of flipping, and the distance `t` of the candidate hit the
// comparison was about. The closest hit is the minimum over accepted candidates, so a marginal
// decision about a candidate FARTHER than the final hit cannot change the answer (e.g. a box face
// coplanar with the floor behind the face that is actually hit); orc_trace_rays reports a ray as
// fragile only if a marginal candidate lies at or before the final t (or nothing was hit).
// Comparisons made inside a ConstantMedium's boundary queries decide the medium's own segment,
// not a candidate at their t, so they always count (nested > 0).
struct Probe {
    bool on = false;
    bool fragile = false;
    int nested = 0;
    double min_t = INFINITY;
};
static thread_local Probe g_probe;
static inline void probe_mark(double t) {
    g_probe.fragile = true;
    if (g_probe.nested > 0 || !(t == t)) t = -INFINITY;
    g_probe.min_t = std::fmin(g_probe.min_t, t);
}
static inline void probe_cmp(double a, double b, double t) {
    if (!g_probe.on) return;
    if (std::isinf(a) || std::isinf(b)) return;  // a bound of +-inf (ConstantMedium's queries, t_max) is never marginal
    double m = std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b)));
    if (std::fabs(a - b) <= 1e-9 * m) probe_mark(t);
}

// ---------------------------------------------------------------------------
// SplitMix64: the seeded stand-in for thread_rng() at scene-construction time.
// ---------------------------------------------------------------------------
struct SceneRng {
    uint64_t s;
    explicit SceneRng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double gen() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double range(double a, double b) { return a + (b - a) * gen(); }
    uint32_t below(uint32_t n) { return (uint32_t)(gen() * n); }
};

// ---------------------------------------------------------------------------
// vec3.rs
// ---------------------------------------------------------------------------
struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
    static Vec3 repeat(double v) { return Vec3(v, v, v); }
    double at(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    double& ref(int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double dot(const Vec3& r) const { return x * r.x + y * r.y + z * r.z; }            // vec3.rs:77-79
    Vec3 cross(const Vec3& r) const {                                                   // vec3.rs:82-88
        return Vec3(y * r.z - z * r.y, -(x * r.z - z * r.x), x * r.y - y * r.x);
    }
    double squared_length() const { return x * x + y * y + z * z; }                     // vec3.rs:93-95 (powf(2.0) == x*x)
    double magnitude() const { return std::sqrt(x * x + y * y + z * z); }               // vec3.rs:90-92
    Vec3 operator+(const Vec3& r) const { return Vec3(x + r.x, y + r.y, z + r.z); }
    Vec3 operator-(const Vec3& r) const { return Vec3(x - r.x, y - r.y, z - r.z); }
    Vec3 operator*(const Vec3& r) const { return Vec3(x * r.x, y * r.y, z * r.z); }
    Vec3 operator*(double s) const { return Vec3(x * s, y * s, z * s); }
    Vec3 operator/(double s) const { return Vec3(x / s, y / s, z / s); }
    Vec3 operator-() const { return Vec3(-x, -y, -z); }
    Vec3 unit() const {                                                                 // vec3.rs:97-100
        double k = 1.0 / magnitude();
        return *this * k;
    }
    Vec3 reflect(const Vec3& n) const { return *this - n * (2.0 * dot(n)); }            // vec3.rs:112-114
    Vec3 refract(const Vec3& n, double etai_over_etat) const {                          // vec3.rs:116-121 (Q6)
        double cos_theta = std::fmin((-*this).dot(n), 1.0);
        Vec3 perp = (*this + n * cos_theta) * etai_over_etat;
        Vec3 par = n * (-std::sqrt(std::fabs(1.0 - perp.squared_length())));
        return perp + par;
    }
};
static inline Vec3 operator*(double s, const Vec3& v) { return Vec3(v.x * s, v.y * s, v.z * s); }

// vec3.rs:149-160 (Q2/Q3): rejection sample inside the unit ball. Candidate j of
// a (pixel, sample, bounce) is Philox block (SCATTER, j), words 0..2.
static Vec3 random_in_unit_space() {
    for (uint32_t j = 0;; ++j) {
        uint32_t w[4];
        g_sampler.block(P_SCATTER, j, w);
        Vec3 v = 2.0 * Vec3(u01(w[0]), u01(w[1]), u01(w[2])) - Vec3::repeat(1.0);
        if (v.squared_length() < 1.0) return v;
    }
}

// ray.rs
struct Ray {
    Vec3 a, b;
    double time = 0;
    Vec3 at(double t) const { return a + t * b; }                                       // ray.rs:24-26
};

// bound.rs
struct Bound {
    Vec3 min, max;
    bool hit(const Ray& ray, double tmin, double tmax) const {                          // bound.rs:13-32 (Q19)
        for (int d = 0; d < 3; ++d) {
            double inv = 1.0 / ray.b.at(d);
            double t0 = (min.at(d) - ray.a.at(d)) * inv;
            double t1 = (max.at(d) - ray.a.at(d)) * inv;
            if (inv < 0.0) std::swap(t0, t1);
            tmin = std::fmax(t0, tmin);  // f64::max ignores NaN, like fmax
            tmax = std::fmin(t1, tmax);
            // no fragility probe here: a marginal box cull only matters if something inside is hit at that
            // same marginal t, and every primitive test below probes its own t / range comparisons
            if (tmax < tmin) return false;
        }
        return true;
    }
    Bound surrounding(const Bound& o) const {                                           // bound.rs:34-46
        Bound r;
        r.min = Vec3(std::fmin(min.x, o.min.x), std::fmin(min.y, o.min.y), std::fmin(min.z, o.min.z));
        r.max = Vec3(std::fmax(max.x, o.max.x), std::fmax(max.y, o.max.y), std::fmax(max.z, o.max.z));
        return r;
    }
};

// ---------------------------------------------------------------------------
// texture.rs / noise.rs
// ---------------------------------------------------------------------------
struct Texture {
    virtual ~Texture() {}
    virtual Vec3 value(double u, double v, const Vec3& p) const = 0;
};
struct Solid : Texture {                                                                // texture.rs:9-13
    Vec3 c;
    explicit Solid(Vec3 c_) : c(c_) {}
    Vec3 value(double, double, const Vec3&) const override { return c; }
};
struct Checker : Texture {                                                              // texture.rs:15-30 (Q21)
    std::shared_ptr<Texture> odd, even;
    Checker(std::shared_ptr<Texture> o, std::shared_ptr<Texture> e) : odd(o), even(e) {}
    Vec3 value(double u, double v, const Vec3& p) const override {
        double sines = std::sin(10.0 * p.x) * std::sin(10.0 * p.y) * std::sin(10.0 * p.z);
        return sines < 0.0 ? odd->value(u, v, p) : even->value(u, v, p);
    }
};

struct Perlin {                                                                         // noise.rs:5-109 (Q22)
    Vec3 random_points[256];
    int px[256], py[256], pz[256];
    // noise.rs:15-29,40-47: 256 x Vec3f::random(-1..1), then three shuffled permutations.
    // rand's SliceRandom::shuffle is a Fisher-Yates from the back: for i in (1..n).rev()
    // swap(i, gen_range(0..i+1)).
    static void generate(SceneRng& rng, Perlin& out) {
        for (int i = 0; i < 256; ++i) {
            double x = rng.range(-1.0, 1.0), y = rng.range(-1.0, 1.0), z = rng.range(-1.0, 1.0);
            out.random_points[i] = Vec3(x, y, z);
        }
        int* perms[3] = {out.px, out.py, out.pz};
        for (int t = 0; t < 3; ++t) {
            int* p = perms[t];
            for (int i = 0; i < 256; ++i) p[i] = i;
            for (int i = 255; i >= 1; --i) std::swap(p[i], p[rng.below((uint32_t)i + 1)]);
        }
    }
    static Perlin from_table(const rtx_perlin& t) {
        Perlin p;
        for (int i = 0; i < 256; ++i) {
            p.random_points[i] = Vec3(t.ranvec[i][0], t.ranvec[i][1], t.ranvec[i][2]);
            p.px[i] = t.perm_x[i]; p.py[i] = t.perm_y[i]; p.pz[i] = t.perm_z[i];
        }
        return p;
    }
    void to_table(rtx_perlin& t) const {
        for (int i = 0; i < 256; ++i) {
            t.ranvec[i][0] = random_points[i].x; t.ranvec[i][1] = random_points[i].y; t.ranvec[i][2] = random_points[i].z;
            t.perm_x[i] = px[i]; t.perm_y[i] = py[i]; t.perm_z[i] = pz[i];
        }
    }
    double noise(const Vec3& p) const {                                                 // noise.rs:49-75
        double u = p.x - std::floor(p.x), v = p.y - std::floor(p.y), w = p.z - std::floor(p.z);
        int i = (int)std::floor(p.x), j = (int)std::floor(p.y), k = (int)std::floor(p.z);
        Vec3 c[2][2][2];
        for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj)
                for (int dk = 0; dk < 2; ++dk)
                    c[di][dj][dk] = random_points[px[(i + di) & 255] ^ py[(j + dj) & 255] ^ pz[(k + dk) & 255]];
        // noise.rs:77-94: Hermite-smoothed trilinear blend of gradient dot products
        double uu = u * u * (3. - 2. * u), vv = v * v * (3. - 2. * v), ww = w * w * (3. - 2. * w);
        double acc = 0.0;
        for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b)
                for (int d = 0; d < 2; ++d) {
                    Vec3 weight(u - a, v - b, w - d);
                    acc += (a * uu + (1 - a) * (1. - uu)) * (b * vv + (1 - b) * (1. - vv)) *
                           (d * ww + (1 - d) * (1. - ww)) * c[a][b][d].dot(weight);
                }
        return acc;
    }
    double turbulence(Vec3 p, int depth) const {                                        // noise.rs:96-108 (no final abs)
        double acc = 0.0, weight = 1.0;
        for (int i = 0; i < depth; ++i) {
            acc += weight * noise(p);
            weight *= 0.5;
            p = p * 2.0;
        }
        return acc;
    }
};
struct Noise : Texture {                                                                // texture.rs:32-59
    Perlin perlin;
    double scale;
    Vec3 value(double, double, const Vec3& p) const override {
        return Vec3::repeat(1.0) * 0.5 * (1. + std::sin(scale * p.z + 10. * perlin.turbulence(p, 7)));
    }
};
// Rust `x as u32` for f64: saturating, NaN -> 0.
static inline uint32_t as_u32(double x) {
    if (!(x > 0.0)) return 0;
    if (x >= 4294967295.0) return 4294967295u;
    return (uint32_t)x;
}
struct Image : Texture {                                                                // texture.rs:61-107 (Q23)
    std::vector<uint8_t> data;  // empty: image failed to load
    uint32_t w = 0, h = 0;
    Vec3 value(double u, double v, const Vec3&) const override {
        if (data.empty()) return Vec3(0., 1., 1.);
        u = std::fmin(std::fmax(u, 0.), 1.);
        v = 1. - std::fmin(std::fmax(v, 0.), 1.);
        uint32_t i = as_u32(u * (double)w), j = as_u32(v * (double)h);
        if (i >= w) i = w - 1;
        if (j >= h) j = h - 1;
        const uint8_t* px = &data[4 * ((size_t)j * w + i)];
        double s = 1.0 / 255.0;
        return Vec3(px[0] * s, px[1] * s, px[2] * s);                                   // vec3.rs:44-51
    }
};

// ---------------------------------------------------------------------------
// material.rs
// ---------------------------------------------------------------------------
struct Material;
struct HitRecord {                                                                      // hittable.rs:15-27
    double t = 0;
    Vec3 p, normal;
    const Material* material = nullptr;
    double u = 0, v = 0;
    bool front_face = false;
    int prim_id = -1;  // instrumentation only
};
static inline void face_normal(const Ray& ray, const Vec3& outward, Vec3& normal, bool& front) {  // hittable.rs:30-44
    front = ray.b.dot(outward) < 0.;
    normal = front ? outward : -outward;
}
struct Material {
    int index = -1;  // index in the description (instrumentation only)
    virtual ~Material() {}
    virtual bool scatter(const Ray& ray, const HitRecord& rec, Vec3& attenuation, Ray& scattered) const = 0;
    virtual Vec3 emitted(double, double, const Vec3&) const { return Vec3::repeat(0.); }  // material.rs:10-12
};
struct Lambertian : Material {                                                          // material.rs:89-100 (Q2)
    std::shared_ptr<Texture> albedo;
    explicit Lambertian(std::shared_ptr<Texture> a) : albedo(a) {}
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& att, Ray& sc) const override {
        Vec3 target = rec.p + rec.normal + random_in_unit_space();
        sc.a = rec.p; sc.b = target - rec.p; sc.time = ray.time;
        att = albedo->value(rec.u, rec.v, rec.p);
        return true;
    }
};
struct Metal : Material {                                                               // material.rs:134-149 (Q4)
    Vec3 albedo;
    double fuzz;
    Metal(Vec3 a, double f) : albedo(a), fuzz(std::fmin(f, 1.0)) {}                     // material.rs:126-131
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& att, Ray& sc) const override {
        Vec3 reflected = ray.b.unit().reflect(rec.normal);
        sc.a = rec.p; sc.b = reflected + fuzz * random_in_unit_space(); sc.time = ray.time;
        att = albedo;
        return sc.b.dot(rec.normal) > 0.0;
    }
};
struct Dielectric : Material {                                                          // material.rs:173-203 (Q5)
    double ir;
    explicit Dielectric(double i) : ir(i) {}
    static double schlick(double cosine, double ri) {                                   // material.rs:173-176
        double r0 = (1.0 - ri) / (1.0 + ri);
        r0 = r0 * r0;
        return r0 + (1.0 - r0) * std::pow(1.0 - cosine, 5.0);
    }
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& att, Ray& sc) const override {
        att = Vec3(1., 1., 1.);
        double ratio = rec.front_face ? 1.0 / ir : ir;
        Vec3 ud = ray.b.unit();
        double cos_theta = std::fmin((-ud).dot(rec.normal), 1.);
        double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        bool cannot_refract = ratio * sin_theta > 1.0;
        // short-circuit ||: the uniform is drawn only when refraction is possible
        Vec3 dir = (cannot_refract || schlick(cos_theta, ratio) > g_sampler.dielectric_u())
                       ? ud.reflect(rec.normal)
                       : ud.refract(rec.normal, ratio);
        sc.a = rec.p; sc.b = dir; sc.time = ray.time;
        return true;
    }
};
struct DiffuseLight : Material {                                                        // material.rs:242-250 (Q24)
    std::shared_ptr<Texture> emit;
    explicit DiffuseLight(std::shared_ptr<Texture> e) : emit(e) {}
    bool scatter(const Ray&, const HitRecord&, Vec3&, Ray&) const override { return false; }
    Vec3 emitted(double u, double v, const Vec3& p) const override { return emit->value(u, v, p); }
};
struct Isotropic : Material {                                                           // material.rs:252-266 (Q3)
    std::shared_ptr<Texture> albedo;
    bool scatter(const Ray& ray, const HitRecord& rec, Vec3& att, Ray& sc) const override {
        sc.a = rec.p; sc.b = random_in_unit_space(); sc.time = ray.time;
        att = albedo->value(rec.u, rec.v, rec.p);
        return true;
    }
};

// ---------------------------------------------------------------------------
// hittable.rs
// ---------------------------------------------------------------------------
// depth-first numbering state: leaf primitives and media are counted separately
struct IdCounter {
    int prim = 0;
    int medium = 0;
};
struct Hittable {
    virtual ~Hittable() {}
    virtual bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const = 0;
    virtual bool bounding_box(double t0, double t1, Bound& out) const = 0;
    // depth-first primitive numbering (see rttnw_b200.h); first visit only
    virtual void assign_ids(IdCounter& next) = 0;
};
using HitPtr = std::shared_ptr<Hittable>;

static inline void sphere_uv(const Vec3& p, double& u, double& v) {                     // hittable.rs:77-83
    const double PI = 3.14159265358979323846;
    double theta = std::acos(-p.y);
    double phi = std::atan2(-p.z, p.x) + PI;
    u = phi / (2.0 * PI);
    v = theta / PI;
}

struct Sphere : Hittable {                                                              // hittable.rs:86-131 (Q9)
    Vec3 center;
    double radius;
    std::shared_ptr<Material> material;
    int id = -1;
    Sphere(Vec3 c, double r, std::shared_ptr<Material> m) : center(c), radius(r), material(m) {}
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        Vec3 oc = ray.a - center;
        double a = ray.b.dot(ray.b);
        double half_b = oc.dot(ray.b);
        double c = oc.dot(oc) - radius * radius;
        probe_cmp(half_b * half_b, a * c, -half_b / a);
        double disc = half_b * half_b - a * c;
        if (disc < 0.0) return false;
        double sqrtd = std::sqrt(disc);
        double root = (-half_b - sqrtd) / a;
        probe_cmp(root, tmin, root); probe_cmp(root, tmax, root);
        if (root < tmin || tmax < root) {
            root = (-half_b + sqrtd) / a;
            probe_cmp(root, tmin, root); probe_cmp(root, tmax, root);
            if (root < tmin || tmax < root) return false;
        }
        rec.t = root;
        rec.p = ray.at(root);
        Vec3 n = (rec.p - center) / radius;
        sphere_uv(n, rec.u, rec.v);
        face_normal(ray, n, rec.normal, rec.front_face);
        rec.material = material.get();
        rec.prim_id = id;
        return true;
    }
    bool bounding_box(double, double, Bound& out) const override {
        out.min = center - Vec3::repeat(radius);
        out.max = center + Vec3::repeat(radius);
        return true;
    }
    void assign_ids(IdCounter& next) override { if (id < 0) id = next.prim++; }
};

struct MovingSphere : Hittable {                                                        // hittable.rs:179-245 (Q10)
    Vec3 c0, c1;
    double t0, t1, radius;
    std::shared_ptr<Material> material;
    int id = -1;
    Vec3 center(double time) const { return c0 + ((time - t0) / (t1 - t0)) * (c1 - c0); }  // hittable.rs:187-191
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        Vec3 oc = ray.a - center(ray.time);
        double a = ray.b.dot(ray.b);
        double half_b = oc.dot(ray.b);
        double c = oc.dot(oc) - radius * radius;
        probe_cmp(half_b * half_b, a * c, -half_b / a);
        double disc = half_b * half_b - a * c;
        if (disc < 0.0) return false;
        double sqrtd = std::sqrt(disc);
        double root = (-half_b - sqrtd) / a;
        probe_cmp(root, tmin, root); probe_cmp(root, tmax, root);
        if (root < tmin || tmax < root) {
            root = (-half_b + sqrtd) / a;
            probe_cmp(root, tmin, root); probe_cmp(root, tmax, root);
            if (root < tmin || tmax < root) return false;
        }
        rec.t = root;
        rec.p = ray.at(root);
        Vec3 n = (rec.p - center(ray.time)) / radius;
        rec.u = 0; rec.v = 0;
        face_normal(ray, n, rec.normal, rec.front_face);
        rec.material = material.get();
        rec.prim_id = id;
        return true;
    }
    bool bounding_box(double ta, double tb, Bound& out) const override {
        Bound b0{center(ta) - Vec3::repeat(radius), center(ta) + Vec3::repeat(radius)};
        Bound b1{center(tb) - Vec3::repeat(radius), center(tb) + Vec3::repeat(radius)};
        out = b0.surrounding(b1);
        return true;
    }
    void assign_ids(IdCounter& next) override { if (id < 0) id = next.prim++; }
};

struct List : Hittable {                                                                // hittable.rs:133-177
    std::vector<HitPtr> list;
    void push(HitPtr h) { list.push_back(h); }
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        bool any = false;
        double closest = tmax;
        HitRecord tmp;
        for (const auto& it : list) {
            if (it->hit(ray, tmin, closest, tmp)) {
                closest = tmp.t;
                rec = tmp;
                any = true;
            }
        }
        return any;
    }
    bool bounding_box(double t0, double t1, Bound& out) const override {
        if (list.empty()) return false;
        Bound acc;
        if (!list[0]->bounding_box(t0, t1, acc)) return false;
        for (size_t i = 1; i < list.size(); ++i) {
            Bound b;
            if (!list[i]->bounding_box(t0, t1, b)) return false;
            acc = b.surrounding(acc);
        }
        out = acc;
        return true;
    }
    void assign_ids(IdCounter& next) override { for (auto& it : list) it->assign_ids(next); }
};

struct BvhTree : Hittable {                                                             // hittable.rs:247-373 (Q17, Q18)
    HitPtr left, right;
    Bound bound;
    static double key(const Hittable& h, int axis) {                                    // hittable.rs:323-333
        Bound b;
        if (!h.bounding_box(0.0, 0.0, b)) { fprintf(stderr, "No bounding box in BvhTree constructor\n"); b = Bound(); }
        return b.min.at(axis);
    }
    // hittable.rs:265-321: random axis per node; the WHOLE remaining vector is re-sorted;
    // leaves are taken with remove(0); span 1 => left == right.
    static std::shared_ptr<BvhTree> build(std::vector<HitPtr>& objects, size_t start, size_t end, double t0, double t1,
                                          SceneRng& rng) {
        int axis = (int)rng.below(3);
        auto node = std::make_shared<BvhTree>();
        size_t span = end - start;
        if (span == 1) {
            HitPtr first = objects.front(); objects.erase(objects.begin());
            node->left = first; node->right = first;
        } else if (span == 2) {
            HitPtr first = objects.front(); objects.erase(objects.begin());
            HitPtr second = objects.front(); objects.erase(objects.begin());
            if (key(*first, axis) < key(*second, axis)) { node->left = first; node->right = second; }
            else { node->left = second; node->right = first; }
        } else {
            std::stable_sort(objects.begin(), objects.end(),
                             [axis](const HitPtr& a, const HitPtr& b) { return key(*a, axis) < key(*b, axis); });
            size_t mid = start + span / 2;
            node->left = build(objects, start, mid, t0, t1, rng);
            node->right = build(objects, mid, end, t0, t1, rng);
        }
        Bound bl, br;
        if (!node->left->bounding_box(t0, t1, bl)) { fprintf(stderr, "No bounding box in BvhTree constructor\n"); bl = Bound(); }
        if (!node->right->bounding_box(t0, t1, br)) { fprintf(stderr, "No bounding box in BvhTree constructor\n"); br = Bound(); }
        node->bound = bl.surrounding(br);
        return node;
    }
    static HitPtr from(const List& list, SceneRng& rng) {                               // hittable.rs:254-264
        std::vector<HitPtr> objects = list.list;
        return build(objects, 0, objects.size(), 0., 1., rng);
    }
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override { // hittable.rs:355-368
        if (!bound.hit(ray, tmin, tmax)) {
            // Fragility probe only: a box culled by a whisker (its entry distance within a few 1e-9 of
            // t_max — 1/d rounding decides) may hide a primitive that ties with the closest hit so far,
            // e.g. the coplanar side faces adjacent ground boxes share. Look inside with the interval
            // widened; whatever would have been accepted there is a marginal candidate at its own t.
            if (g_probe.on && std::isfinite(tmax)) {
                double wide = tmax + 4e-9 * std::fmax(1.0, std::fabs(tmax));
                if (bound.hit(ray, tmin, wide)) {
                    HitRecord shadow;
                    if (left->hit(ray, tmin, wide, shadow)) probe_mark(shadow.t);
                    if (left.get() != right.get() && right->hit(ray, tmin, wide, shadow)) probe_mark(shadow.t);
                }
            }
            return false;
        }
        HitRecord lrec, rrec;
        bool hl = left->hit(ray, tmin, tmax, lrec);
        // span-1 nodes have left == right (hittable.rs:282-285): the reference tests the same
        // object again with t_max = its own t, which (inclusive bounds) returns the same record.
        // Skipped here only so that the fragility probe does not see an object tie with itself.
        if (left.get() == right.get()) { if (hl) rec = lrec; return hl; }
        double t = hl ? lrec.t : tmax;
        bool hr = right->hit(ray, tmin, t, rrec);
        if (hr) { rec = rrec; return true; }
        if (hl) { rec = lrec; return true; }
        return false;
    }
    bool bounding_box(double, double, Bound& out) const override { out = bound; return true; }
    void assign_ids(IdCounter&) override {}  // ids are assigned through the List the tree was built from
};
// A BVH node of the description: keeps the original list for id numbering.
struct BvhOfList : Hittable {
    List source;
    HitPtr tree;
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override { return tree->hit(ray, tmin, tmax, rec); }
    bool bounding_box(double a, double b, Bound& out) const override { return tree->bounding_box(a, b, out); }
    void assign_ids(IdCounter& next) override { source.assign_ids(next); }
};

struct Rect : Hittable {                                                                // hittable.rs:502-547 (Q11)
    int axis0, axis1, k_axis;  // Xy: 0,1,2  Xz: 0,2,1  Yz: 1,2,0  (hittable.rs:450-488)
    double a0, a1, b0, b1, k;
    std::shared_ptr<Material> material;
    int id = -1;
    Rect(int plane, double a0_, double a1_, double b0_, double b1_, double k_, std::shared_ptr<Material> m)
        : a0(a0_), a1(a1_), b0(b0_), b1(b1_), k(k_), material(m) {
        if (plane == 0) { axis0 = 0; axis1 = 1; k_axis = 2; }
        else if (plane == 1) { axis0 = 0; axis1 = 2; k_axis = 1; }
        else { axis0 = 1; axis1 = 2; k_axis = 0; }
    }
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        double t = (k - ray.a.at(k_axis)) / ray.b.at(k_axis);
        if (g_probe.on) {
            // The decision is (t in range) AND (point in rectangle): it is fragile only if one
            // factor is marginal while the other holds (or is marginal too) — a plane that ties
            // in t but is hit far outside its extent (coplanar faces of other boxes) is not.
            auto nearly = [](double a, double b) { return !std::isinf(a) && !std::isinf(b) && std::fabs(a - b) <= 1e-9 * std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b))); };
            double q0 = ray.a.at(axis0) + t * ray.b.at(axis0), q1 = ray.a.at(axis1) + t * ray.b.at(axis1);
            bool t_ok = !(t < tmin || t > tmax), t_near = nearly(t, tmin) || nearly(t, tmax);
            bool p_ok = (a0 <= q0 && q0 < a1) && (b0 <= q1 && q1 < b1);
            bool p_near = (nearly(q0, a0) || nearly(q0, a1)) && ((b0 <= q1 && q1 < b1) || nearly(q1, b0) || nearly(q1, b1));
            p_near = p_near || ((nearly(q1, b0) || nearly(q1, b1)) && ((a0 <= q0 && q0 < a1) || nearly(q0, a0) || nearly(q0, a1)));
            if ((t_near && (p_ok || p_near)) || (p_near && (t_ok || t_near))) probe_mark(t);
        }
        if (t < tmin || t > tmax) return false;
        double p0 = ray.a.at(axis0) + t * ray.b.at(axis0);
        double p1 = ray.a.at(axis1) + t * ray.b.at(axis1);
        // Range::contains is half-open: start <= x < end (NaN is never contained)
        if (!(a0 <= p0 && p0 < a1) || !(b0 <= p1 && p1 < b1)) return false;
        rec.u = (p0 - a0) / (a1 - a0);
        rec.v = (p1 - b0) / (b1 - b0);
        Vec3 outward;
        outward.ref(k_axis) = 1.;
        face_normal(ray, outward, rec.normal, rec.front_face);
        rec.t = t;
        rec.p = ray.at(t);
        rec.material = material.get();
        rec.prim_id = id;
        return true;
    }
    bool bounding_box(double, double, Bound& out) const override {
        Vec3 mn, mx;
        mn.ref(axis0) = a0; mn.ref(axis1) = b0; mn.ref(k_axis) = k - 0.0001;
        mx.ref(axis0) = a1; mx.ref(axis1) = b1; mx.ref(k_axis) = k + 0.0001;
        out.min = mn; out.max = mx;
        return true;
    }
    void assign_ids(IdCounter& next) override { if (id < 0) id = next.prim++; }
};

struct Cube : Hittable {                                                                // hittable.rs:549-592 (Q12)
    Vec3 bmin, bmax;
    List sides;
    Cube(Vec3 p0, Vec3 p1, std::shared_ptr<Material> m) : bmin(p0), bmax(p1) {
        // Plane::rectangles + points(), hittable.rs:414-426,451-479
        sides.push(std::make_shared<Rect>(0, p0.x, p1.x, p0.y, p1.y, p0.z, m));
        sides.push(std::make_shared<Rect>(0, p0.x, p1.x, p0.y, p1.y, p1.z, m));
        sides.push(std::make_shared<Rect>(1, p0.x, p1.x, p0.z, p1.z, p0.y, m));
        sides.push(std::make_shared<Rect>(1, p0.x, p1.x, p0.z, p1.z, p1.y, m));
        sides.push(std::make_shared<Rect>(2, p0.y, p1.y, p0.z, p1.z, p0.x, m));
        sides.push(std::make_shared<Rect>(2, p0.y, p1.y, p0.z, p1.z, p1.x, m));
    }
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override { return sides.hit(ray, tmin, tmax, rec); }
    bool bounding_box(double, double, Bound& out) const override { out.min = bmin; out.max = bmax; return true; }
    void assign_ids(IdCounter& next) override { sides.assign_ids(next); }
};

struct Translate : Hittable {                                                           // hittable.rs:594-629 (Q13)
    HitPtr item;
    Vec3 offset;
    Translate(HitPtr i, Vec3 o) : item(i), offset(o) {}
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        Ray moved{ray.a - offset, ray.b, ray.time};
        if (!item->hit(moved, tmin, tmax, rec)) return false;
        Vec3 n; bool ff;
        face_normal(moved, rec.normal, n, ff);  // re-run on the already face-flipped normal
        rec.normal = n; rec.front_face = ff;
        rec.p = rec.p + offset;
        return true;
    }
    bool bounding_box(double t0, double t1, Bound& out) const override {
        Bound b;
        if (!item->bounding_box(t0, t1, b)) return false;
        out.min = b.min + offset; out.max = b.max + offset;
        return true;
    }
    void assign_ids(IdCounter& next) override { item->assign_ids(next); }
};

struct YRotate : Hittable {                                                             // hittable.rs:631-722 (Q14, Q15)
    HitPtr item;
    double sin_theta, cos_theta;
    bool has_bound;
    Bound bound;
    YRotate(HitPtr i, double angle) : item(i) {
        const double PI = 3.14159265358979323846;
        double radians = angle * (PI / 180.0);  // f64::to_radians
        sin_theta = std::sin(radians);
        cos_theta = std::cos(radians);
        Bound b;
        has_bound = item->bounding_box(0., 1., b);
        if (!has_bound) b = Bound();
        Vec3 mn = Vec3::repeat(INFINITY), mx = Vec3::repeat(-INFINITY);
        for (int a = 0; a < 2; ++a)
            for (int c = 0; c < 2; ++c)
                for (int d = 0; d < 2; ++d) {
                    double x = a * b.max.x + (1 - a) * b.min.x;
                    double y = c * b.max.y + (1 - c) * b.min.y;
                    double z = d * b.max.z + (1 - d) * b.min.z;
                    double nx = cos_theta * x + sin_theta * z;
                    double nz = -sin_theta * nx + cos_theta * z;  // Q15: shadowed x (hittable.rs:661-662)
                    Vec3 tmp(nx, y, nz);
                    for (int e = 0; e < 3; ++e) {
                        mn.ref(e) = std::fmin(mn.at(e), tmp.at(e));
                        mx.ref(e) = std::fmax(mx.at(e), tmp.at(e));
                    }
                }
        bound.min = mn; bound.max = mx;
    }
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        Vec3 o = ray.a, d = ray.b;
        o.x = cos_theta * ray.a.x - sin_theta * ray.a.z;
        o.z = sin_theta * ray.a.x + cos_theta * ray.a.z;
        d.x = cos_theta * ray.b.x - sin_theta * ray.b.z;
        d.z = sin_theta * ray.b.x + cos_theta * ray.b.z;
        Ray rot{o, d, ray.time};
        if (!item->hit(rot, tmin, tmax, rec)) return false;
        // Q14 (hittable.rs:700-705): sequential update — the new [0] feeds [2]
        rec.p.x = cos_theta * rec.p.x + sin_theta * rec.p.z;
        rec.p.z = -sin_theta * rec.p.x + cos_theta * rec.p.z;
        rec.normal.x = cos_theta * rec.normal.x + sin_theta * rec.normal.z;
        rec.normal.z = -sin_theta * rec.normal.x + cos_theta * rec.normal.z;
        Vec3 n; bool ff;
        face_normal(rot, rec.normal, n, ff);  // object-space ray vs the mangled normal (hittable.rs:706)
        rec.normal = n; rec.front_face = ff;
        return true;
    }
    bool bounding_box(double, double, Bound& out) const override { out = bound; return true; }
    void assign_ids(IdCounter& next) override { item->assign_ids(next); }
};

struct ConstantMedium : Hittable {                                                      // hittable.rs:724-801 (Q16)
    HitPtr boundary;
    Isotropic phase;
    double neg_inv_density;
    int id = -1;
    int medium_key = 0;  // which Philox word this medium draws: its ordinal in depth-first order
    ConstantMedium(HitPtr b, double density, std::shared_ptr<Texture> tex) : boundary(b), neg_inv_density(-1. / density) {
        phase.albedo = tex;
    }
    bool hit(const Ray& ray, double tmin, double tmax, HitRecord& rec) const override {
        HitRecord r1, r2;
        g_probe.nested++;
        bool in = boundary->hit(ray, -INFINITY, INFINITY, r1) && boundary->hit(ray, r1.t + 0.0001, INFINITY, r2);
        g_probe.nested--;
        if (!in) return false;
        r1.t = std::fmax(r1.t, tmin);
        r2.t = std::fmin(r2.t, tmax);
        if (g_probe.on) {
            // the candidate this comparison is about is the scatter point, not the entry point: when the far
            // end was clipped by a surface hit at the entry distance (the glass sphere that is also this
            // medium's boundary, scenes.rs:282-292), the surface wins whichever is tested first
            double entry = std::fmax(r1.t, 0.);
            probe_cmp(r1.t, r2.t, entry + neg_inv_density * std::log(g_sampler.medium_u(medium_key)) / ray.b.magnitude());
        }
        if (r1.t >= r2.t) return false;
        r1.t = std::fmax(r1.t, 0.);
        double ray_length = ray.b.magnitude();
        double distance_inside = (r2.t - r1.t) * ray_length;
        double hit_distance = neg_inv_density * std::log(g_sampler.medium_u(medium_key));
        probe_cmp(hit_distance, distance_inside, r1.t);
        if (hit_distance > distance_inside) return false;
        rec.t = r1.t + hit_distance / ray_length;
        rec.p = ray.at(rec.t);
        rec.normal = Vec3(1., 0., 0.);
        rec.front_face = true;
        rec.material = &phase;
        rec.u = 0.0; rec.v = 0.0;
        rec.prim_id = id;
        return true;
    }
    bool bounding_box(double t0, double t1, Bound& out) const override { return boundary->bounding_box(t0, t1, out); }
    void assign_ids(IdCounter& next) override {
        boundary->assign_ids(next);
        if (id < 0) { id = next.prim++; medium_key = next.medium++; }
    }
};

// ---------------------------------------------------------------------------
// camera.rs
// ---------------------------------------------------------------------------
struct Camera {
    Vec3 origin, lower_left_corner, horizontal, vertical, u, v, w;
    double lens_radius = 0, open_time = 0, close_time = 1;
    explicit Camera(const rtx_camera& d) {                                              // camera.rs:32-61
        const double PI = 3.14159265358979323846;
        lens_radius = d.aperture / 2.0;
        double theta = d.vertical_fov * PI / 180.0;
        double half_height = std::tan(theta / 2.0);
        double half_width = d.aspect_ratio * half_height;
        Vec3 lookfrom(d.lookfrom[0], d.lookfrom[1], d.lookfrom[2]);
        Vec3 lookat(d.lookat[0], d.lookat[1], d.lookat[2]);
        Vec3 vup(d.view_up[0], d.view_up[1], d.view_up[2]);
        origin = lookfrom;
        w = (lookfrom - lookat).unit();
        u = vup.cross(w).unit();
        v = w.cross(u);
        lower_left_corner = origin - half_width * d.focus_distance * u - half_height * d.focus_distance * v -
                            d.focus_distance * w;
        horizontal = 2.0 * half_width * d.focus_distance * u;
        vertical = 2.0 * half_height * d.focus_distance * v;
        open_time = d.open_time;
        close_time = d.close_time;
    }
    // camera.rs:63-84 (Q25). Lens candidates: Philox blocks (LENS, j): (w0,w1), (w2,w3).
    // Shutter time: block (CAMERA, 0) word 2.
    Ray ray(double s, double t) const {
        Vec3 p;
        for (uint32_t j = 0;; ++j) {
            uint32_t wd[4];
            g_sampler.block(P_LENS, j, wd);
            p = 2.0 * Vec3(u01(wd[0]), u01(wd[1]), 0.0) - Vec3(1.0, 1.0, 0.0);
            if (p.dot(p) < 1.0) break;
            p = 2.0 * Vec3(u01(wd[2]), u01(wd[3]), 0.0) - Vec3(1.0, 1.0, 0.0);
            if (p.dot(p) < 1.0) break;
        }
        Vec3 rd = lens_radius * p;
        Vec3 offset = u * rd.x + v * rd.y;
        uint32_t wd[4];
        g_sampler.block(P_CAMERA, 0, wd);
        Ray r;
        r.a = origin + offset;
        r.b = lower_left_corner + s * horizontal + t * vertical - origin - offset;
        r.time = open_time + (close_time - open_time) * u01(wd[2]);  // gen_range(open..close)
        return r;
    }
};

// ---------------------------------------------------------------------------
// main.rs: color + scene table
// ---------------------------------------------------------------------------
static Vec3 color(const Ray& ray, const Vec3& background, const Hittable& world, int depth, int max_depth) {  // main.rs:26-45 (Q7, Q8)
    if (depth <= 0) return Vec3::repeat(0.);
    g_sampler.bounce = (uint32_t)(max_depth - depth);
    g_sampler.rays++;
    HitRecord rec;
    if (world.hit(ray, 0.001, DBL_MAX, rec)) {
        Vec3 emitted = rec.material->emitted(rec.u, rec.v, rec.p);
        Vec3 att;
        Ray scattered;
        if (rec.material->scatter(ray, rec, att, scattered))
            return emitted + att * color(scattered, background, world, depth - 1, max_depth);
        return emitted;
    }
    return background;
}

}  // namespace orc

using namespace orc;

struct orc_scene {
    std::shared_ptr<List> world;
    rtx_camera camera{};
    Vec3 background;
    int prim_count = 0;
    std::vector<std::shared_ptr<Texture>> textures;  // from-desc scenes only
};

// ---------------------------------------------------------------------------
// scenes.rs (restated; geometry randomness from SplitMix64)
// ---------------------------------------------------------------------------
namespace {
std::shared_ptr<Texture> solid(double r, double g, double b) { return std::make_shared<Solid>(Vec3(r, g, b)); }
std::shared_ptr<Material> lambert(std::shared_ptr<Texture> t) { return std::make_shared<Lambertian>(t); }
std::shared_ptr<Material> lambert(double r, double g, double b) { return lambert(solid(r, g, b)); }
HitPtr sphere(Vec3 c, double r, std::shared_ptr<Material> m) { return std::make_shared<Sphere>(c, r, m); }
HitPtr rect(int plane, std::shared_ptr<Material> m, double a0, double a1, double b0, double b1, double k) {
    return std::make_shared<Rect>(plane, a0, a1, b0, b1, k, m);
}
std::shared_ptr<Texture> noise_scaled(double scale, SceneRng& rng) {                    // texture.rs:45-50
    auto n = std::make_shared<Noise>();
    Perlin::generate(rng, n->perlin);
    n->scale = scale;
    return n;
}
std::shared_ptr<Texture> image_tex(const uint8_t* rgba, int w, int h) {
    auto t = std::make_shared<Image>();
    if (rgba && w > 0 && h > 0) { t->data.assign(rgba, rgba + (size_t)4 * w * h); t->w = (uint32_t)w; t->h = (uint32_t)h; }
    return t;
}

void random_scene(List& list, SceneRng& rng) {                                          // scenes.rs:11-88
    auto checker = std::make_shared<Checker>(solid(0.2, 0.3, 0.1), solid(0.9, 0.9, 0.9));
    list.push(sphere(Vec3(0.0, -1000.0, 0.0), 1000.0, lambert(checker)));
    for (int a = -11; a < 11; ++a)
        for (int b = -11; b < 11; ++b) {
            double choose_mat = rng.gen();
            double cx = (double)a + 0.9 + rng.gen();
            double cz = (double)b + 0.9 + rng.gen();
            Vec3 center(cx, 0.2, cz);
            if ((center - Vec3(4.0, 0.2, 0.0)).magnitude() > 0.9) {
                if (choose_mat < 0.8) {
                    Vec3 final_center = center + Vec3(0.0, rng.range(0.0, 0.5), 0.0);
                    auto ms = std::make_shared<MovingSphere>();
                    ms->c0 = center; ms->c1 = final_center; ms->t0 = 0.; ms->t1 = 1.; ms->radius = 0.2;
                    double r = rng.gen() * rng.gen();
                    double g = rng.gen() * rng.gen();
                    double bl = rng.gen() * rng.gen();
                    ms->material = lambert(r, g, bl);
                    list.push(ms);
                } else if (choose_mat < 0.95) {
                    double r = 0.5 * (1.0 - rng.gen());
                    double g = 0.5 * (1.0 - rng.gen());
                    double bl = 0.5 * (1.0 - rng.gen());
                    double fuzz = 0.5 * rng.gen();
                    list.push(sphere(center, 0.2, std::make_shared<Metal>(Vec3(r, g, bl), fuzz)));
                } else {
                    list.push(sphere(center, 0.2, std::make_shared<Dielectric>(1.5)));
                }
            }
        }
    list.push(sphere(Vec3(0.0, 1.0, 0.0), 1.0, std::make_shared<Dielectric>(1.5)));
    list.push(sphere(Vec3(-4.0, 1.0, 0.0), 1.0, lambert(0.4, 0.2, 0.1)));
    list.push(sphere(Vec3(4.0, 1.0, 0.0), 1.0, std::make_shared<Metal>(Vec3(0.7, 0.6, 0.5), 0.0)));
}
void two_spheres(List& world) {                                                         // scenes.rs:90-108
    auto checker = std::make_shared<Checker>(solid(0.2, 0.3, 0.1), solid(0.9, 0.9, 0.9));
    world.push(sphere(Vec3(0.0, -10.0, 0.0), 10.0, lambert(checker)));
    world.push(sphere(Vec3(0.0, 10.0, 0.0), 10.0, lambert(checker)));
}
void two_perlin_spheres(List& world, SceneRng& rng) {                                   // scenes.rs:110-125
    auto perlin = noise_scaled(4., rng);
    world.push(sphere(Vec3(0.0, -1000.0, 0.0), 1000.0, lambert(perlin)));
    world.push(sphere(Vec3(0.0, 2.0, 0.0), 2.0, lambert(perlin)));
}
void earth(List& world, const uint8_t* rgba, int w, int h) {                            // scenes.rs:127-136
    world.push(sphere(Vec3::repeat(0.0), 2., lambert(image_tex(rgba, w, h))));
}
void simple_light(List& world, SceneRng& rng) {                                         // scenes.rs:138-155
    auto perlin = noise_scaled(4., rng);
    world.push(sphere(Vec3(0.0, -1000.0, 0.0), 1000.0, lambert(perlin)));
    world.push(sphere(Vec3(0.0, 2.0, 0.0), 2.0, lambert(perlin)));
    auto light = std::make_shared<DiffuseLight>(solid(4., 4., 4.));
    world.push(rect(0, light, 3., 5., 1., 3., -2.0));
}
void empty_cornell_box(List& world) {                                                   // scenes.rs:157-173
    auto red = lambert(0.65, 0.05, 0.05);
    auto white = lambert(0.73, 0.73, 0.73);
    auto green = lambert(0.12, 0.45, 0.15);
    auto light = std::make_shared<DiffuseLight>(solid(15., 15., 15.));
    world.push(rect(2, green, 0., 555., 0., 555., 555.));
    world.push(rect(2, red, 0., 555., 0., 555., 0.));
    world.push(rect(1, light, 213., 343., 227., 332., 554.));
    world.push(rect(1, white, 0., 555., 0., 555., 555.));
    world.push(rect(1, white, 0., 555., 0., 555., 0.));
    world.push(rect(0, white, 0., 555., 0., 555., 555.));
}
HitPtr rotate_translate(HitPtr item, double angle, Vec3 offset) {
    return std::make_shared<Translate>(std::make_shared<YRotate>(item, angle), offset);
}
void cornell_box(List& world) {                                                         // scenes.rs:175-196
    empty_cornell_box(world);
    auto white = lambert(0.73, 0.73, 0.73);
    world.push(rotate_translate(std::make_shared<Cube>(Vec3(0., 0., 0.), Vec3(165., 330., 165.), white), 15., Vec3(265., 0., 295.)));
    world.push(rotate_translate(std::make_shared<Cube>(Vec3(0., 0., 0.), Vec3::repeat(165.), white), -18., Vec3(130., 0., 65.)));
}
void smoke_cornell_box(List& world) {                                                   // scenes.rs:198-236
    auto red = lambert(0.65, 0.05, 0.05);
    auto white = lambert(0.73, 0.73, 0.73);
    auto green = lambert(0.12, 0.45, 0.15);
    auto light = std::make_shared<DiffuseLight>(solid(7., 7., 7.));
    world.push(rect(2, green, 0., 555., 0., 555., 555.));
    world.push(rect(2, red, 0., 555., 0., 555., 0.));
    world.push(rect(1, light, 113., 443., 127., 432., 554.));
    world.push(rect(1, white, 0., 555., 0., 555., 555.));
    world.push(rect(1, white, 0., 555., 0., 555., 0.));
    world.push(rect(0, white, 0., 555., 0., 555., 555.));
    auto c1 = rotate_translate(std::make_shared<Cube>(Vec3(0., 0., 0.), Vec3(165., 330., 165.), white), 15., Vec3(265., 0., 295.));
    auto c2 = rotate_translate(std::make_shared<Cube>(Vec3(0., 0., 0.), Vec3::repeat(165.), white), -18., Vec3(130., 0., 65.));
    auto m1 = std::make_shared<ConstantMedium>(c1, 0.01, solid(0., 0., 0.));
    auto m2 = std::make_shared<ConstantMedium>(c2, 0.01, solid(1., 1., 1.));
    world.push(m1);
    world.push(m2);
}
void final_scene(List& world, SceneRng& rng, SceneRng& bvh_rng, const uint8_t* rgba, int w, int h) {  // scenes.rs:238-334
    List boxes;
    auto ground = lambert(0.48, 0.83, 0.53);
    const int boxes_per_side = 20;
    for (int i = 0; i < boxes_per_side; ++i)
        for (int j = 0; j < boxes_per_side; ++j) {
            double wd = 100.;
            Vec3 v0(-1000. + i * wd, 0., -1000. + j * wd);
            Vec3 v1(v0.x + wd, rng.range(1., 101.), v0.z + wd);
            boxes.push(std::make_shared<Cube>(v0, v1, ground));
        }
    auto b1 = std::make_shared<BvhOfList>();
    b1->source = boxes;
    b1->tree = BvhTree::from(boxes, bvh_rng);
    world.push(b1);
    auto light = std::make_shared<DiffuseLight>(solid(7., 7., 7.));
    world.push(rect(1, light, 123., 423., 147., 412., 554.));
    Vec3 center1 = Vec3::repeat(400.);
    Vec3 center2 = center1 + Vec3(30., 0., 0.);
    auto ms = std::make_shared<MovingSphere>();
    ms->c0 = center1; ms->c1 = center2; ms->t0 = 0.; ms->t1 = 1.; ms->radius = 50.;
    ms->material = lambert(0.7, 0.3, 0.1);
    world.push(ms);
    world.push(sphere(Vec3(260., 150., 45.), 50.0, std::make_shared<Dielectric>(1.5)));
    world.push(sphere(Vec3(0., 150., 45.), 50.0, std::make_shared<Metal>(Vec3(0.8, 0.8, 0.9), 1.)));
    auto boundary = std::make_shared<Sphere>(Vec3(360., 150., 145.), 70., std::make_shared<Dielectric>(1.5));
    world.push(std::make_shared<Sphere>(*boundary));  // boundary.clone()
    auto m1 = std::make_shared<ConstantMedium>(boundary, 0.2, solid(0.2, 0.4, 0.9));
    world.push(m1);
    auto m2 = std::make_shared<ConstantMedium>(sphere(Vec3::repeat(0.), 5000., std::make_shared<Dielectric>(1.5)), 0.0001,
                                               solid(1., 1., 1.));
    world.push(m2);
    world.push(sphere(Vec3(400., 200., 400.), 100., lambert(image_tex(rgba, w, h))));
    world.push(sphere(Vec3(220., 280., 300.), 80.0, lambert(noise_scaled(0.1, rng))));
    List spheres;
    auto white = lambert(0.73, 0.73, 0.73);
    const int ns = 1000;
    for (int i = 0; i < ns; ++i) {
        double x = rng.range(0., 165.), y = rng.range(0., 165.), z = rng.range(0., 165.);
        spheres.push(sphere(Vec3(x, y, z), 10., white));
    }
    auto b2 = std::make_shared<BvhOfList>();
    b2->source = spheres;
    b2->tree = BvhTree::from(spheres, bvh_rng);
    world.push(rotate_translate(b2, 15., Vec3(-100., 270., 395.)));
}

void set_camera(rtx_camera& c, Vec3 from, Vec3 at, double vfov, double aspect, double aperture) {  // main.rs:184-197
    c.lookfrom[0] = from.x; c.lookfrom[1] = from.y; c.lookfrom[2] = from.z;
    c.lookat[0] = at.x; c.lookat[1] = at.y; c.lookat[2] = at.z;
    c.view_up[0] = 0.; c.view_up[1] = 1.; c.view_up[2] = 0.;
    c.vertical_fov = vfov; c.aspect_ratio = aspect; c.aperture = aperture;
    c.focus_distance = 10.0; c.open_time = 0.0; c.close_time = 1.0;
}
}  // namespace

extern "C" {

orc_scene* orc_scene_builtin(int scene_number, uint64_t seed, const uint8_t* earth_rgba, int earth_w, int earth_h) {
    auto s = new orc_scene();
    s->world = std::make_shared<List>();
    SceneRng rng(seed), bvh_rng(seed ^ 0xB5Dull);
    double wide = 16.0 / 9.0;
    Vec3 sky(0.7, 0.8, 1.), black(0., 0., 0.);
    switch (scene_number) {  // main.rs:66-183
        case 1: random_scene(*s->world, rng); s->background = sky; set_camera(s->camera, Vec3(13., 2., 3.), Vec3(0, 0, 0), 20., wide, 0.1); break;
        case 2: two_spheres(*s->world); s->background = sky; set_camera(s->camera, Vec3(13., 2., 3.), Vec3(0, 0, 0), 20., wide, 0.); break;
        case 3: two_perlin_spheres(*s->world, rng); s->background = sky; set_camera(s->camera, Vec3(13., 2., 3.), Vec3(0, 0, 0), 20., wide, 0.); break;
        case 4: earth(*s->world, earth_rgba, earth_w, earth_h); s->background = sky; set_camera(s->camera, Vec3(13., 2., 3.), Vec3(0, 0, 0), 20., wide, 0.); break;
        case 5: simple_light(*s->world, rng); s->background = black; set_camera(s->camera, Vec3(26., 3., 6.), Vec3(0., 2., 0.), 20., wide, 0.); break;
        case 6: empty_cornell_box(*s->world); s->background = black; set_camera(s->camera, Vec3(278., 278., -800.), Vec3(278., 278., 0.), 40., 1.0, 0.); break;
        case 7: cornell_box(*s->world); s->background = black; set_camera(s->camera, Vec3(278., 278., -800.), Vec3(278., 278., 0.), 40., 1.0, 0.); break;
        case 8: smoke_cornell_box(*s->world); s->background = black; set_camera(s->camera, Vec3(278., 278., -800.), Vec3(278., 278., 0.), 40., 1.0, 0.); break;
        case 9: final_scene(*s->world, rng, bvh_rng, earth_rgba, earth_w, earth_h); s->background = black; set_camera(s->camera, Vec3(478., 278., -600.), Vec3(278., 278., 0.), 40., 1.0, 0.); break;
        default: delete s; return nullptr;
    }
    IdCounter next;
    s->world->assign_ids(next);
    s->prim_count = next.prim;
    return s;
}

orc_scene* orc_scene_from_desc(const rtx_scene_desc* d, uint64_t bvh_seed) {
    auto s = new orc_scene();
    SceneRng bvh_rng(bvh_seed);
    // textures (children may reference any index: build lazily, memoised)
    std::vector<std::shared_ptr<Texture>> tex((size_t)d->n_textures);
    std::function<std::shared_ptr<Texture>(int)> get_tex = [&](int i) -> std::shared_ptr<Texture> {
        if (i < 0 || i >= d->n_textures) return solid(0, 0, 0);
        if (tex[(size_t)i]) return tex[(size_t)i];
        const rtx_texture& t = d->textures[i];
        std::shared_ptr<Texture> r;
        switch (t.kind) {
            case RTX_TEX_SOLID: r = solid(t.f[0], t.f[1], t.f[2]); break;
            case RTX_TEX_CHECKER: r = std::make_shared<Checker>(get_tex(t.a), get_tex(t.b)); break;
            case RTX_TEX_NOISE: {
                auto n = std::make_shared<Noise>();
                n->perlin = Perlin::from_table(d->perlins[t.a]);
                n->scale = t.f[0];
                r = n;
                break;
            }
            case RTX_TEX_IMAGE: r = image_tex(d->images[t.a].rgba, d->images[t.a].width, d->images[t.a].height); break;
            default: r = solid(0, 0, 0);
        }
        tex[(size_t)i] = r;
        return r;
    };
    std::vector<std::shared_ptr<Material>> mats((size_t)d->n_materials);
    for (int i = 0; i < d->n_materials; ++i) {
        const rtx_material& m = d->materials[i];
        std::shared_ptr<Material> r;
        switch (m.kind) {
            case RTX_MAT_LAMBERTIAN: r = std::make_shared<Lambertian>(get_tex(m.texture)); break;
            case RTX_MAT_METAL: r = std::make_shared<Metal>(Vec3(m.albedo[0], m.albedo[1], m.albedo[2]), m.param); break;
            case RTX_MAT_DIELECTRIC: r = std::make_shared<Dielectric>(m.param); break;
            case RTX_MAT_DIFFUSE_LIGHT: r = std::make_shared<DiffuseLight>(get_tex(m.texture)); break;
            case RTX_MAT_ISOTROPIC: { auto iso = std::make_shared<Isotropic>(); iso->albedo = get_tex(m.texture); r = iso; break; }
            default: r = std::make_shared<Lambertian>(solid(0, 0, 0));
        }
        r->index = i;
        mats[(size_t)i] = r;
    }
    std::vector<HitPtr> built((size_t)d->n_nodes);
    std::function<HitPtr(int)> get = [&](int i) -> HitPtr {
        if (built[(size_t)i]) return built[(size_t)i];
        const rtx_node& n = d->nodes[i];
        const double* f = n.f;
        HitPtr r;
        switch (n.kind) {
            case RTX_NODE_SPHERE: r = sphere(Vec3(f[0], f[1], f[2]), f[3], mats[(size_t)n.material]); break;
            case RTX_NODE_MOVING_SPHERE: {
                auto ms = std::make_shared<MovingSphere>();
                ms->c0 = Vec3(f[0], f[1], f[2]); ms->c1 = Vec3(f[3], f[4], f[5]);
                ms->radius = f[6]; ms->t0 = f[7]; ms->t1 = f[8];
                ms->material = mats[(size_t)n.material];
                r = ms;
                break;
            }
            case RTX_NODE_RECT_XY: r = rect(0, mats[(size_t)n.material], f[0], f[1], f[2], f[3], f[4]); break;
            case RTX_NODE_RECT_XZ: r = rect(1, mats[(size_t)n.material], f[0], f[1], f[2], f[3], f[4]); break;
            case RTX_NODE_RECT_YZ: r = rect(2, mats[(size_t)n.material], f[0], f[1], f[2], f[3], f[4]); break;
            case RTX_NODE_CUBE: r = std::make_shared<Cube>(Vec3(f[0], f[1], f[2]), Vec3(f[3], f[4], f[5]), mats[(size_t)n.material]); break;
            case RTX_NODE_LIST: {
                auto l = std::make_shared<List>();
                for (int c = 0; c < n.n_children; ++c) l->push(get(d->children[n.child + c]));
                r = l;
                break;
            }
            case RTX_NODE_BVH: {
                auto b = std::make_shared<BvhOfList>();
                for (int c = 0; c < n.n_children; ++c) b->source.push(get(d->children[n.child + c]));
                b->tree = BvhTree::from(b->source, bvh_rng);
                r = b;
                break;
            }
            case RTX_NODE_TRANSLATE: r = std::make_shared<Translate>(get(n.child), Vec3(f[0], f[1], f[2])); break;
            case RTX_NODE_ROTATE_Y: r = std::make_shared<YRotate>(get(n.child), f[0]); break;
            case RTX_NODE_MEDIUM: {
                auto m = std::make_shared<ConstantMedium>(get(n.child), f[0], get_tex(n.material));
                m->phase.index = -(1 + n.material);
                r = m;
                break;
            }
            default: r = std::make_shared<List>();
        }
        built[(size_t)i] = r;
        return r;
    };
    HitPtr root = get(d->root);
    s->world = std::make_shared<List>();
    s->world->push(root);
    s->camera = d->camera;
    s->background = Vec3(d->background[0], d->background[1], d->background[2]);
    for (int i = 0; i < d->n_textures; ++i) get_tex(i);
    s->textures = tex;
    IdCounter next;
    s->world->assign_ids(next);
    s->prim_count = next.prim;
    return s;
}

void orc_scene_free(orc_scene* s) { delete s; }
int orc_scene_prim_count(const orc_scene* s) { return s->prim_count; }
void orc_scene_camera(const orc_scene* s, rtx_camera* cam, double background[3]) {
    *cam = s->camera;
    background[0] = s->background.x; background[1] = s->background.y; background[2] = s->background.z;
}

static void run_parallel(int n_threads, int64_t n_items, int64_t chunk, const std::function<void(int64_t, int64_t)>& body) {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(chunk);
            if (b >= n_items) break;
            body(b, std::min(n_items, b + chunk));
        }
    };
    if (n_threads == 1) { worker(); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
    for (auto& t : th) t.join();
}

void orc_trace_rays(const orc_scene* s, int64_t n, const rtx_ray* rays, rtx_hit* hits, uint8_t* fragile, int n_threads) {
    run_parallel(n_threads, n, 4096, [&](int64_t b, int64_t e) {
        g_sampler = Sampler();
        g_sampler.fixed = true;
        g_probe.on = fragile != nullptr;
        for (int64_t i = b; i < e; ++i) {
            const rtx_ray& r = rays[i];
            Ray ray{Vec3(r.origin[0], r.origin[1], r.origin[2]), Vec3(r.direction[0], r.direction[1], r.direction[2]), r.time};
            g_sampler.xi = r.xi;
            g_probe.fragile = false;
            g_probe.nested = 0;
            g_probe.min_t = INFINITY;
            HitRecord rec;
            rtx_hit& h = hits[i];
            std::memset(&h, 0, sizeof(h));
            bool got = s->world->hit(ray, r.t_min, r.t_max, rec);
            bool relevant = g_probe.fragile;  // a miss: every marginal decision counts
            if (got) {
                relevant = g_probe.fragile && !(g_probe.min_t > rec.t + 1e-9 * std::fmax(1.0, std::fabs(rec.t)));
                h.prim_id = rec.prim_id;
                h.material = rec.material ? rec.material->index : -1;
                h.front_face = rec.front_face ? 1 : 0;
                h.t = rec.t;
                h.p[0] = rec.p.x; h.p[1] = rec.p.y; h.p[2] = rec.p.z;
                h.normal[0] = rec.normal.x; h.normal[1] = rec.normal.y; h.normal[2] = rec.normal.z;
                h.u = rec.u; h.v = rec.v;
            } else {
                h.prim_id = RTX_MISS;
                h.material = -1;
            }
            if (fragile) fragile[i] = relevant ? 1 : 0;
        }
        g_probe.on = false;
    });
}

uint64_t orc_render(const orc_scene* s, int width, int height, int spp_begin, int spp_count, int max_depth, uint64_t seed,
                    int row_begin, int row_end, int row_stride, double* rgb_sum, int n_threads) {
    Camera camera(s->camera);
    std::atomic<uint64_t> total_rays{0};
    row_begin = std::max(0, row_begin);
    row_end = std::min(height, row_end);
    if (row_stride < 1) row_stride = 1;
    int64_t n_rows = row_end > row_begin ? (row_end - row_begin + row_stride - 1) / row_stride : 0;
    run_parallel(n_threads, n_rows, 1, [&](int64_t b, int64_t e) {
        g_sampler = Sampler();
        g_sampler.fixed = false;
        g_sampler.key[0] = (uint32_t)seed;
        g_sampler.key[1] = (uint32_t)(seed >> 32);
        for (int64_t rr = b; rr < e; ++rr) {
            int r = row_begin + (int)rr * row_stride;  // row from the top (main.rs:202-204: rows are emitted top first)
            int j = height - 1 - r;
            for (int i = 0; i < width; ++i) {
                Vec3 acc;
                g_sampler.pixel = (uint32_t)(r * width + i);
                for (int sidx = spp_begin; sidx < spp_begin + spp_count; ++sidx) {      // main.rs:211-217
                    g_sampler.sample = (uint32_t)sidx;
                    g_sampler.bounce = 0;
                    uint32_t w[4];
                    g_sampler.block(P_CAMERA, 0, w);
                    double u = ((double)i + u01(w[0])) / (double)width;
                    double v = ((double)j + u01(w[1])) / (double)height;
                    Ray ray = camera.ray(u, v);
                    acc = acc + color(ray, s->background, *s->world, max_depth, max_depth);
                }
                double* px = rgb_sum + 3 * ((size_t)r * width + i);
                px[0] += acc.x; px[1] += acc.y; px[2] += acc.z;
            }
        }
        total_rays += g_sampler.rays;
    });
    return total_rays.load();
}

void orc_tonemap(const double* rgb_sum, int n_pixels, double samples, uint8_t* rgba) {  // main.rs:217-225 (Q26)
    for (int i = 0; i < n_pixels; ++i) {
        for (int c = 0; c < 3; ++c) {
            double x = rgb_sum[3 * i + c] / samples;
            x = std::sqrt(x);
            if (x < 0.0) x = 0.0;          // f64::clamp keeps NaN; `as u8` then maps NaN to 0
            if (x > 0.999) x = 0.999;
            x *= 256.;
            uint8_t q = 0;
            if (x == x) q = (uint8_t)(x >= 255.0 ? 255 : (x <= 0.0 ? 0 : (int)x));
            rgba[4 * i + c] = q;
        }
        rgba[4 * i + 3] = 255;
    }
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
void orc_sphere_uv(const double p[3], double uv[2]) { sphere_uv(Vec3(p[0], p[1], p[2]), uv[0], uv[1]); }
int orc_bound_hit(const double bmin[3], const double bmax[3], const rtx_ray* r) {
    Bound b{Vec3(bmin[0], bmin[1], bmin[2]), Vec3(bmax[0], bmax[1], bmax[2])};
    Ray ray{Vec3(r->origin[0], r->origin[1], r->origin[2]), Vec3(r->direction[0], r->direction[1], r->direction[2]), r->time};
    return b.hit(ray, r->t_min, r->t_max) ? 1 : 0;
}
double orc_perlin_noise(const rtx_perlin* tab, const double p[3]) { return Perlin::from_table(*tab).noise(Vec3(p[0], p[1], p[2])); }
double orc_perlin_turbulence(const rtx_perlin* tab, const double p[3], int depth) {
    return Perlin::from_table(*tab).turbulence(Vec3(p[0], p[1], p[2]), depth);
}
void orc_texture_value(const orc_scene* s, int texture_index, double u, double v, const double p[3], double rgb[3]) {
    Vec3 c = s->textures[(size_t)texture_index]->value(u, v, Vec3(p[0], p[1], p[2]));
    rgb[0] = c.x; rgb[1] = c.y; rgb[2] = c.z;
}
void orc_perlin_generate(uint64_t seed, rtx_perlin* out) {
    SceneRng rng(seed);
    Perlin p;
    Perlin::generate(rng, p);
    p.to_table(*out);
}
int orc_hardware_threads(void) { return (int)std::max(1u, std::thread::hardware_concurrency()); }

}  // extern "C"
