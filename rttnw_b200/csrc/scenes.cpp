// scenes.cpp — scene_api.hpp implementation and the nine scene constructors of
// src/scenes.rs, parameter-exact, plus the scene table of src/main.rs:66-183.
// Host only: this describes scenes; it never intersects or shades anything.
#include <cmath>
#include <cstring>

#include "scene_api.hpp"

namespace rttnw {

double Vec3f::magnitude() const { return std::sqrt(x * x + y * y + z * z); }

uint64_t SceneRng::next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double SceneRng::gen() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
double SceneRng::range(double a, double b) { return a + (b - a) * gen(); }
uint32_t SceneRng::below(uint32_t n) { return (uint32_t)(gen() * n); }

static thread_local SceneBuilder* g_builder = nullptr;
SceneBuilder& current_builder() { return *g_builder; }
SceneBuilder::Scope::Scope(SceneBuilder& b) : prev(g_builder) { g_builder = &b; }
SceneBuilder::Scope::~Scope() { g_builder = prev; }

static rtx_node blank_node(int kind) {
    rtx_node n;
    std::memset(&n, 0, sizeof(n));
    n.kind = kind;
    n.material = -1;
    n.child = -1;
    return n;
}

Texture solid(const Vec3f& c) {
    rtx_texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = RTX_TEX_SOLID;
    t.f[0] = c.x; t.f[1] = c.y; t.f[2] = c.z;
    auto& b = current_builder();
    b.textures.push_back(t);
    return Texture{(int)b.textures.size() - 1};
}
Texture CheckerTexture::make(Texture odd, Texture even) {
    rtx_texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = RTX_TEX_CHECKER;
    t.a = odd.id; t.b = even.id;
    auto& b = current_builder();
    b.textures.push_back(t);
    return Texture{(int)b.textures.size() - 1};
}
// NoiseTexture::scaled (texture.rs:45-50) -> Perlin::new (noise.rs:40-47): 256 x Vec3f::random(-1..1),
// then three shuffled permutations (SliceRandom::shuffle = Fisher-Yates from the back).
Texture NoiseTexture::scaled(double scale, SceneRng& rng) {
    auto& b = current_builder();
    rtx_perlin p;
    for (int i = 0; i < 256; ++i)
        for (int c = 0; c < 3; ++c) p.ranvec[i][c] = rng.range(-1.0, 1.0);
    int32_t* perms[3] = {p.perm_x, p.perm_y, p.perm_z};
    for (int t = 0; t < 3; ++t) {
        for (int i = 0; i < 256; ++i) perms[t][i] = i;
        for (int i = 255; i >= 1; --i) {
            uint32_t j = rng.below((uint32_t)i + 1);
            int32_t tmp = perms[t][i]; perms[t][i] = perms[t][j]; perms[t][j] = tmp;
        }
    }
    b.perlins.push_back(p);
    rtx_texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = RTX_TEX_NOISE;
    t.a = (int)b.perlins.size() - 1;
    t.f[0] = scale;
    b.textures.push_back(t);
    return Texture{(int)b.textures.size() - 1};
}
// ImageTexture::new (texture.rs:69-75): a failed load is not an error, it is the cyan texture.
Texture ImageTexture::new_(const char* path) {
    auto& b = current_builder();
    rtx_image img;
    std::memset(&img, 0, sizeof(img));
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    std::string err;
    if (path && png_read_rgba8(path, w, h, px, err)) {
        b.image_pixels.push_back(std::move(px));
        img.width = w; img.height = h;  // pointer patched in finish() (vector may move)
    } else {
        b.image_pixels.emplace_back();
    }
    b.images.push_back(img);
    rtx_texture t;
    std::memset(&t, 0, sizeof(t));
    t.kind = RTX_TEX_IMAGE;
    t.a = (int)b.images.size() - 1;
    b.textures.push_back(t);
    return Texture{(int)b.textures.size() - 1};
}

static Material add_material(int kind, int texture, const Vec3f& albedo, double param) {
    rtx_material m;
    std::memset(&m, 0, sizeof(m));
    m.kind = kind; m.texture = texture;
    m.albedo[0] = albedo.x; m.albedo[1] = albedo.y; m.albedo[2] = albedo.z;
    m.param = param;
    auto& b = current_builder();
    b.materials.push_back(m);
    return Material{(int)b.materials.size() - 1};
}
Material Lambertian::arc(Texture albedo) { return add_material(RTX_MAT_LAMBERTIAN, albedo.id, Vec3f(), 0); }
Material Metal::arc(const Vec3f& albedo, double fuzz) { return add_material(RTX_MAT_METAL, -1, albedo, std::fmin(fuzz, 1.0)); }
Material Dielectric::arc(double ir) { return add_material(RTX_MAT_DIELECTRIC, -1, Vec3f(), ir); }
Material DiffuseLight::arc(Texture emit) { return add_material(RTX_MAT_DIFFUSE_LIGHT, emit.id, Vec3f(), 0); }

Hittable Sphere(const Vec3f& c, double r, Material m) {
    rtx_node n = blank_node(RTX_NODE_SPHERE);
    n.material = m.id;
    n.f[0] = c.x; n.f[1] = c.y; n.f[2] = c.z; n.f[3] = r;
    return Hittable{current_builder().add_node(n)};
}
Hittable MovingSphere(const Vec3f& c0, const Vec3f& c1, Range time, double r, Material m) {
    rtx_node n = blank_node(RTX_NODE_MOVING_SPHERE);
    n.material = m.id;
    n.f[0] = c0.x; n.f[1] = c0.y; n.f[2] = c0.z; n.f[3] = c1.x; n.f[4] = c1.y; n.f[5] = c1.z;
    n.f[6] = r; n.f[7] = time.start; n.f[8] = time.end;
    return Hittable{current_builder().add_node(n)};
}
static Hittable rect(int kind, Material m, Range p0, Range p1, double k) {
    rtx_node n = blank_node(kind);
    n.material = m.id;
    n.f[0] = p0.start; n.f[1] = p0.end; n.f[2] = p1.start; n.f[3] = p1.end; n.f[4] = k;
    return Hittable{current_builder().add_node(n)};
}
Hittable XY::rectangle(Material m, Range p0, Range p1, double k) { return rect(RTX_NODE_RECT_XY, m, p0, p1, k); }
Hittable XZ::rectangle(Material m, Range p0, Range p1, double k) { return rect(RTX_NODE_RECT_XZ, m, p0, p1, k); }
Hittable YZ::rectangle(Material m, Range p0, Range p1, double k) { return rect(RTX_NODE_RECT_YZ, m, p0, p1, k); }
Hittable Cube::new_(const Vec3f& a, const Vec3f& b, Material m) {
    rtx_node n = blank_node(RTX_NODE_CUBE);
    n.material = m.id;
    n.f[0] = a.x; n.f[1] = a.y; n.f[2] = a.z; n.f[3] = b.x; n.f[4] = b.y; n.f[5] = b.z;
    return Hittable{current_builder().add_node(n)};
}
static Hittable list_node(int kind, const std::vector<int>& items) {
    auto& b = current_builder();
    rtx_node n = blank_node(kind);
    n.child = (int)b.children.size();
    n.n_children = (int)items.size();
    for (int id : items) b.children.push_back(id);
    return Hittable{b.add_node(n)};
}
Hittable List::into_hittable() const { return list_node(RTX_NODE_LIST, items); }
Hittable BvhTree::from(const List& list) { return list_node(RTX_NODE_BVH, list.items); }
Hittable Hittable::translate(const Vec3f& o) const {
    rtx_node n = blank_node(RTX_NODE_TRANSLATE);
    n.child = id;
    n.f[0] = o.x; n.f[1] = o.y; n.f[2] = o.z;
    return Hittable{current_builder().add_node(n)};
}
Hittable Hittable::rotate_y(double angle) const {
    rtx_node n = blank_node(RTX_NODE_ROTATE_Y);
    n.child = id;
    n.f[0] = angle;
    return Hittable{current_builder().add_node(n)};
}
Hittable ConstantMedium::new_(Hittable boundary, double density, Texture phase) {
    rtx_node n = blank_node(RTX_NODE_MEDIUM);
    n.child = boundary.id;
    n.material = phase.id;
    n.f[0] = density;
    return Hittable{current_builder().add_node(n)};
}

const rtx_scene_desc& SceneBuilder::finish(Hittable world, const CameraDescriptor& cam, const Vec3f& background) {
    for (size_t i = 0; i < images.size(); ++i) images[i].rgba = image_pixels[i].empty() ? nullptr : image_pixels[i].data();
    std::memset(&desc, 0, sizeof(desc));
    desc.nodes = nodes.data(); desc.n_nodes = (int)nodes.size(); desc.root = world.id;
    desc.children = children.data(); desc.n_children = (int)children.size();
    desc.materials = materials.data(); desc.n_materials = (int)materials.size();
    desc.textures = textures.data(); desc.n_textures = (int)textures.size();
    desc.perlins = perlins.data(); desc.n_perlins = (int)perlins.size();
    desc.images = images.data(); desc.n_images = (int)images.size();
    desc.background[0] = background.x; desc.background[1] = background.y; desc.background[2] = background.z;
    rtx_camera& c = desc.camera;
    c.lookfrom[0] = cam.lookfrom.x; c.lookfrom[1] = cam.lookfrom.y; c.lookfrom[2] = cam.lookfrom.z;
    c.lookat[0] = cam.lookat.x; c.lookat[1] = cam.lookat.y; c.lookat[2] = cam.lookat.z;
    c.view_up[0] = cam.view_up.x; c.view_up[1] = cam.view_up.y; c.view_up[2] = cam.view_up.z;
    c.vertical_fov = cam.vertical_fov; c.aspect_ratio = cam.aspect_ratio; c.aperture = cam.aperture;
    c.focus_distance = cam.focus_distance; c.open_time = cam.open_time; c.close_time = cam.close_time;
    return desc;
}

// ---------------------------------------------------------------------------
// src/scenes.rs
// ---------------------------------------------------------------------------
static List random_scene(SceneRng& rng) {  // scenes.rs:11-88
    List list;
    Texture checker = CheckerTexture::make(solid(Vec3f(0.2, 0.3, 0.1)), solid(Vec3f(0.9, 0.9, 0.9)));
    list.push(Sphere(Vec3f(0.0, -1000.0, 0.0), 1000.0, Lambertian::arc(checker)));
    for (int a = -11; a < 11; ++a) {
        for (int b = -11; b < 11; ++b) {
            double choose_mat = rng.gen();
            double cx = (double)a + 0.9 + rng.gen();
            double cz = (double)b + 0.9 + rng.gen();
            Vec3f center(cx, 0.2, cz);
            if ((center - Vec3f(4.0, 0.2, 0.0)).magnitude() > 0.9) {
                if (choose_mat < 0.8) {  // diffuse, moving
                    Vec3f final_center = center + Vec3f(0.0, rng.range(0.0, 0.5), 0.0);
                    double r = rng.gen() * rng.gen();
                    double g = rng.gen() * rng.gen();
                    double bl = rng.gen() * rng.gen();
                    list.push(MovingSphere(center, final_center, Range{0., 1.}, 0.2, Lambertian::boxed(Vec3f(r, g, bl))));
                } else if (choose_mat < 0.95) {  // metal
                    double r = 0.5 * (1.0 - rng.gen());
                    double g = 0.5 * (1.0 - rng.gen());
                    double bl = 0.5 * (1.0 - rng.gen());
                    double fuzz = 0.5 * rng.gen();
                    list.push(Sphere(center, 0.2, Metal::arc(Vec3f(r, g, bl), fuzz)));
                } else {  // glass
                    list.push(Sphere(center, 0.2, Dielectric::arc(1.5)));
                }
            }
        }
    }
    list.push(Sphere(Vec3f(0.0, 1.0, 0.0), 1.0, Dielectric::arc(1.5)));
    list.push(Sphere(Vec3f(-4.0, 1.0, 0.0), 1.0, Lambertian::arc(Vec3f(0.4, 0.2, 0.1))));
    list.push(Sphere(Vec3f(4.0, 1.0, 0.0), 1.0, Metal::arc(Vec3f(0.7, 0.6, 0.5), 0.0)));
    return list;
}

static List two_spheres() {  // scenes.rs:90-108
    List world;
    Texture checker = CheckerTexture::make(solid(Vec3f(0.2, 0.3, 0.1)), solid(Vec3f(0.9, 0.9, 0.9)));
    Material m = Lambertian::arc(checker);
    world.push(Sphere(Vec3f(0.0, -10.0, 0.0), 10.0, m));
    world.push(Sphere(Vec3f(0.0, 10.0, 0.0), 10.0, m));
    return world;
}

static List two_perlin_spheres(SceneRng& rng) {  // scenes.rs:110-125
    List world;
    Material m = Lambertian::arc(NoiseTexture::scaled(4., rng));
    world.push(Sphere(Vec3f(0.0, -1000.0, 0.0), 1000.0, m));
    world.push(Sphere(Vec3f(0.0, 2.0, 0.0), 2.0, m));
    return world;
}

static List earth(const char* png) {  // scenes.rs:127-136
    List world;
    world.push(Sphere(Vec3f::repeat(0.0), 2., Lambertian::arc(ImageTexture::new_(png))));
    return world;
}

static List simple_light(SceneRng& rng) {  // scenes.rs:138-155
    List world;
    Material m = Lambertian::arc(NoiseTexture::scaled(4., rng));
    world.push(Sphere(Vec3f(0.0, -1000.0, 0.0), 1000.0, m));
    world.push(Sphere(Vec3f(0.0, 2.0, 0.0), 2.0, m));
    Material light = DiffuseLight::arc(Vec3f::repeat(4.));
    world.push(XY::rectangle(light, Range{3., 5.}, Range{1., 3.}, -2.0));
    return world;
}

static List empty_cornell_box() {  // scenes.rs:157-173
    List world;
    Material red = Lambertian::arc(Vec3f(0.65, 0.05, 0.05));
    Material white = Lambertian::arc(Vec3f::repeat(0.73));
    Material green = Lambertian::arc(Vec3f(0.12, 0.45, 0.15));
    Material light = DiffuseLight::arc(Vec3f::repeat(15.));
    world.push(YZ::rectangle(green, Range{0., 555.}, Range{0., 555.}, 555.));
    world.push(YZ::rectangle(red, Range{0., 555.}, Range{0., 555.}, 0.));
    world.push(XZ::rectangle(light, Range{213., 343.}, Range{227., 332.}, 554.));
    world.push(XZ::rectangle(white, Range{0., 555.}, Range{0., 555.}, 555.));
    world.push(XZ::rectangle(white, Range{0., 555.}, Range{0., 555.}, 0.));
    world.push(XY::rectangle(white, Range{0., 555.}, Range{0., 555.}, 555.));
    return world;
}

static List cornell_box() {  // scenes.rs:175-196
    List world = empty_cornell_box();
    Material white = Lambertian::arc(Vec3f::repeat(0.73));
    world.push(Cube::new_(Vec3f(0., 0., 0.), Vec3f(165., 330., 165.), white).rotate_y(15.).translate(Vec3f(265., 0., 295.)));
    world.push(Cube::new_(Vec3f(0., 0., 0.), Vec3f::repeat(165.), white).rotate_y(-18.).translate(Vec3f(130., 0., 65.)));
    return world;
}

static List smoke_cornell_box() {  // scenes.rs:198-236
    List world;
    Material red = Lambertian::arc(Vec3f(0.65, 0.05, 0.05));
    Material white = Lambertian::arc(Vec3f::repeat(0.73));
    Material green = Lambertian::arc(Vec3f(0.12, 0.45, 0.15));
    Material light = DiffuseLight::arc(Vec3f::repeat(7.));
    world.push(YZ::rectangle(green, Range{0., 555.}, Range{0., 555.}, 555.));
    world.push(YZ::rectangle(red, Range{0., 555.}, Range{0., 555.}, 0.));
    world.push(XZ::rectangle(light, Range{113., 443.}, Range{127., 432.}, 554.));
    world.push(XZ::rectangle(white, Range{0., 555.}, Range{0., 555.}, 555.));
    world.push(XZ::rectangle(white, Range{0., 555.}, Range{0., 555.}, 0.));
    world.push(XY::rectangle(white, Range{0., 555.}, Range{0., 555.}, 555.));
    Hittable c1 = Cube::new_(Vec3f(0., 0., 0.), Vec3f(165., 330., 165.), white).rotate_y(15.).translate(Vec3f(265., 0., 295.));
    Hittable c2 = Cube::new_(Vec3f(0., 0., 0.), Vec3f::repeat(165.), white).rotate_y(-18.).translate(Vec3f(130., 0., 65.));
    world.push(ConstantMedium::new_(c1, 0.01, solid(Vec3f::repeat(0.))));
    world.push(ConstantMedium::new_(c2, 0.01, solid(Vec3f::repeat(1.))));
    return world;
}

static List final_scene(SceneRng& rng, const char* png) {  // scenes.rs:238-334
    List boxes;
    Material ground = Lambertian::arc(Vec3f(0.48, 0.83, 0.53));
    const int boxes_per_side = 20;
    for (int i = 0; i < boxes_per_side; ++i) {
        for (int j = 0; j < boxes_per_side; ++j) {
            double w = 100.;
            Vec3f v0(-1000. + i * w, 0., -1000. + j * w);
            Vec3f v1(v0.x + w, rng.range(1., 101.), v0.z + w);
            boxes.push(Cube::new_(v0, v1, ground));
        }
    }
    List world;
    world.push(BvhTree::from(boxes));
    Material light = DiffuseLight::arc(Vec3f::repeat(7.));
    world.push(XZ::rectangle(light, Range{123., 423.}, Range{147., 412.}, 554.));
    Vec3f center1 = Vec3f::repeat(400.);
    Vec3f center2 = center1 + Vec3f(30., 0., 0.);
    world.push(MovingSphere(center1, center2, Range{0., 1.}, 50., Lambertian::boxed(Vec3f(0.7, 0.3, 0.1))));
    world.push(Sphere(Vec3f(260., 150., 45.), 50.0, Dielectric::arc(1.5)));
    world.push(Sphere(Vec3f(0., 150., 45.), 50.0, Metal::arc(Vec3f(0.8, 0.8, 0.9), 1.)));
    Material glass = Dielectric::arc(1.5);
    world.push(Sphere(Vec3f(360., 150., 145.), 70., glass));                       // boundary.clone()
    Hittable boundary = Sphere(Vec3f(360., 150., 145.), 70., glass);
    world.push(ConstantMedium::new_(boundary, 0.2, solid(Vec3f(0.2, 0.4, 0.9))));
    world.push(ConstantMedium::new_(Sphere(Vec3f::repeat(0.), 5000., Dielectric::arc(1.5)), 0.0001, solid(Vec3f::repeat(1.))));
    world.push(Sphere(Vec3f(400., 200., 400.), 100., Lambertian::arc(ImageTexture::new_(png))));
    world.push(Sphere(Vec3f(220., 280., 300.), 80.0, Lambertian::arc(NoiseTexture::scaled(0.1, rng))));
    List spheres;
    Material white = Lambertian::arc(Vec3f::repeat(0.73));
    const int ns = 1000;
    for (int i = 0; i < ns; ++i) {
        double x = rng.range(0., 165.), y = rng.range(0., 165.), z = rng.range(0., 165.);  // Vec3f::random(0..165)
        spheres.push(Sphere(Vec3f(x, y, z), 10., white));
    }
    world.push(BvhTree::from(spheres).rotate_y(15.).translate(Vec3f(-100., 270., 395.)));
    return world;
}

// main.rs:58-197,255: render(400, 16/9, 100, scene) with per-scene overrides; depth 50 (:216).
bool builtin_scene_defaults(int scene, SceneDefaults& o) {
    static const char* names[10] = {"", "random_scene", "two_spheres", "two_perlin_spheres", "earth", "simple_light",
                                    "empty_cornell_box", "cornell_box", "smoke_cornell_box", "final_scene"};
    if (scene < 1 || scene > 9) return false;
    int width = 400, samples = 100;
    double aspect = 16.0 / 9.0;
    if (scene == 5) samples = 400;
    if (scene >= 6 && scene <= 8) { samples = 200; aspect = 1.0; width = 600; }
    if (scene == 9) { samples = 10000; aspect = 1.0; width = 800; }
    o.width = width;
    o.height = (int)(uint32_t)((double)width / aspect);  // main.rs:184
    o.samples = samples;
    o.max_depth = 50;
    o.name = names[scene];
    return true;
}

bool builtin_scene(int scene, uint64_t seed, const char* earth_png_path, SceneBuilder& b) {
    if (scene < 1 || scene > 9) return false;
    SceneBuilder::Scope scope(b);
    SceneRng rng(seed);
    const char* png = earth_png_path ? earth_png_path : "assets/earth.png";  // scenes.rs:129,303
    CameraDescriptor cam;
    cam.view_up = Vec3f(0., 1., 0.);
    cam.focus_distance = 10.0;
    cam.open_time = 0.0;
    cam.close_time = 1.0;
    cam.aspect_ratio = (scene >= 6) ? 1.0 : 16.0 / 9.0;
    Vec3f background(0.7, 0.8, 1.);
    List world;
    switch (scene) {
        case 1: world = random_scene(rng); cam.lookfrom = Vec3f(13., 2., 3.); cam.lookat = Vec3f(); cam.vertical_fov = 20.; cam.aperture = 0.1; break;
        case 2: world = two_spheres(); cam.lookfrom = Vec3f(13., 2., 3.); cam.lookat = Vec3f(); cam.vertical_fov = 20.; break;
        case 3: world = two_perlin_spheres(rng); cam.lookfrom = Vec3f(13., 2., 3.); cam.lookat = Vec3f(); cam.vertical_fov = 20.; break;
        case 4: world = earth(png); cam.lookfrom = Vec3f(13., 2., 3.); cam.lookat = Vec3f(); cam.vertical_fov = 20.; break;
        case 5: world = simple_light(rng); background = Vec3f(); cam.lookfrom = Vec3f(26., 3., 6.); cam.lookat = Vec3f(0., 2., 0.); cam.vertical_fov = 20.; break;
        case 6: world = empty_cornell_box(); background = Vec3f(); cam.lookfrom = Vec3f(278., 278., -800.); cam.lookat = Vec3f(278., 278., 0.); cam.vertical_fov = 40.; break;
        case 7: world = cornell_box(); background = Vec3f(); cam.lookfrom = Vec3f(278., 278., -800.); cam.lookat = Vec3f(278., 278., 0.); cam.vertical_fov = 40.; break;
        case 8: world = smoke_cornell_box(); background = Vec3f(); cam.lookfrom = Vec3f(278., 278., -800.); cam.lookat = Vec3f(278., 278., 0.); cam.vertical_fov = 40.; break;
        case 9: world = final_scene(rng, png); background = Vec3f(); cam.lookfrom = Vec3f(478., 278., -600.); cam.lookat = Vec3f(278., 278., 0.); cam.vertical_fov = 40.; break;
    }
    b.finish(world.into_hittable(), cam, background);
    return true;
}

}  // namespace rttnw
