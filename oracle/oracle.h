/* oracle.h — C interface of the CPU oracle. TEST INFRASTRUCTURE ONLY.
 *
 * The oracle is a C++17 / f64 restatement of the reference's algorithm
 * (luliic2/rttnw: src/main.rs:26-45,184-229, every file of src/math/, src/scenes.rs),
 * quirks included. It is the checker the CUDA path is compared
 * against. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product (rttnw_b200/) never does.
 *
 * PARITY PIN: the reference ships no tests, no golden vectors, no seeded
 * output and cannot be compiled here (no Rust toolchain; HEAD does not
 * type-check). What the oracle is pinned against: (1) analytic known-answers
 * derived from the cited source lines (tests/test_oracle_kat.py), (2) the
 * published Philox4x32-10 known-answer vectors, (3) the reference's own
 * shipped render of the deterministic Cornell-box scene (cornel_box.png,
 * committed 8x8 box-downsampled as tests/golden/cornell_ref_75.npy), which the
 * oracle's render must match region by region — including the dark rotated-box
 * faces that only the YRotate sequential-update behaviour (hittable.rs:700-705)
 * produces. Everything beyond that is "parity unpinned" by the reference.
 */
#ifndef RTTNW_ORACLE_H
#define RTTNW_ORACLE_H

#include "../include/rttnw_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* scenes.rs constructors 1..9 + the scene table of main.rs:66-183. Geometry /
 * Perlin randomness from SplitMix64(seed) (the reference uses thread_rng()).
 * earth_rgba may be NULL (cyan texture, texture.rs:96-99). */
orc_scene* orc_scene_builtin(int scene_number, uint64_t seed, const uint8_t* earth_rgba,
                             int earth_w, int earth_h);
/* Builds the reference's object tree 1:1 from a description (LIST -> List,
 * BVH -> BvhTree with the reference's construction, hittable.rs:260-321). */
orc_scene* orc_scene_from_desc(const rtx_scene_desc* desc, uint64_t bvh_seed);
void orc_scene_free(orc_scene* s);
int orc_scene_prim_count(const orc_scene* s);
void orc_scene_camera(const orc_scene* s, rtx_camera* cam, double background[3]);

/* world.hit(ray, t_min, t_max) for n rays; ConstantMedium draws ray.xi.
 * fragile[i] (may be NULL) is set to 1 when some comparison the reference makes
 * on the way was within 1e-9 relative of flipping (grazing / tie). */
void orc_trace_rays(const orc_scene* s, int64_t n, const rtx_ray* rays, rtx_hit* hits,
                    uint8_t* fragile, int n_threads);

/* The pixel loop of main.rs:202-217 (without /samples): adds spp_count samples
 * per pixel to rgb_sum (width*height*3 doubles, row 0 = top). Randomness:
 * Philox4x32-10 keyed by seed, counter (pixel, sample, bounce|purpose, block) —
 * the same streams the CUDA kernels use. Returns the number of world.hit
 * queries. Only rows row_begin, row_begin + row_stride, ... < row_end are rendered (bounded
 * CPU baselines sample every k-th row of the frame). */
uint64_t orc_render(const orc_scene* s, int width, int height, int spp_begin, int spp_count,
                    int max_depth, uint64_t seed, int row_begin, int row_end, int row_stride,
                    double* rgb_sum, int n_threads);
/* main.rs:217-225: mean, sqrt, clamp(0,0.999)*256 as u8, alpha 255. */
void orc_tonemap(const double* rgb_sum, int n_pixels, double samples, uint8_t* rgba);

/* unit helpers exposed for known-answer tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void orc_sphere_uv(const double p[3], double uv[2]);
int orc_bound_hit(const double bmin[3], const double bmax[3], const rtx_ray* ray);
double orc_perlin_noise(const rtx_perlin* tab, const double p[3]);
double orc_perlin_turbulence(const rtx_perlin* tab, const double p[3], int depth);
void orc_texture_value(const orc_scene* s, int texture_index, double u, double v,
                       const double p[3], double rgb[3]);
void orc_perlin_generate(uint64_t seed, rtx_perlin* out);
int orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
