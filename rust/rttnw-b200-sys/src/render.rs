//! `render()` — the reference's entry point (`fn render(width, aspect_ratio, samples, scene) -> Option<()>`,
//! src/main.rs:58-233) with the pixel loop (main.rs:199-229) replaced by the CUDA library: same scene table, same
//! camera, same `image.png` (RGBA8, top row first), same stdout lines. The reference's `main()` (main.rs:236-258)
//! can call this function instead of its own and stay as it is.
use std::ptr;

use crate::ffi;
use crate::math::{CameraDescriptor, Color, Hittable, Position, Vec3f};
use crate::{check, Gpu, RtxError, SceneBuilder};

/// One row of the scene table of src/main.rs:66-183: what `match scene { ... }` sets besides the world.
pub struct SceneSetup {
    pub name: &'static str,
    pub background: Vec3f<Color>,
    pub lookfrom: Vec3f<Position>,
    pub lookat: Vec3f<Position>,
    pub vertical_fov: f64,
    pub aperture: f64,
    /// `samples = ...; aspect_ratio = ...; width = ...;` overrides of the arm (None: the caller's values stand)
    pub samples: Option<usize>,
    pub square_width: Option<u32>,
}

pub fn scene_setup(scene: usize) -> Option<SceneSetup> {
    let sky = Vec3f::new(0.7, 0.8, 1.0);
    let black = Vec3f::repeat(0.0);
    let book1 = |name, aperture| SceneSetup {
        name, background: sky, lookfrom: Vec3f::new(13.0, 2.0, 3.0), lookat: Vec3f::repeat(0.0), vertical_fov: 20.0, aperture,
        samples: None, square_width: None,
    };
    let cornell = |name| SceneSetup {
        name, background: black, lookfrom: Vec3f::new(278.0, 278.0, -800.0), lookat: Vec3f::new(278.0, 278.0, 0.0),
        vertical_fov: 40.0, aperture: 0.0, samples: Some(200), square_width: Some(600),
    };
    Some(match scene {
        1 => book1("random_scene", 0.1),
        2 => book1("two_spheres", 0.0),
        3 => book1("two_perlin_spheres", 0.0),
        4 => book1("earth", 0.0),
        5 => SceneSetup {
            name: "simple_light", background: black, lookfrom: Vec3f::new(26.0, 3.0, 6.0), lookat: Vec3f::new(0.0, 2.0, 0.0),
            vertical_fov: 20.0, aperture: 0.0, samples: Some(400), square_width: None,
        },
        6 => cornell("empty_cornell_box"),
        7 => cornell("cornell_box"),
        8 => cornell("smoke_cornell_box"),
        9 => SceneSetup {
            name: "final_scene", background: black, lookfrom: Vec3f::new(478.0, 278.0, -600.0), lookat: Vec3f::new(278.0, 278.0, 0.0),
            vertical_fov: 40.0, aperture: 0.0, samples: Some(10000), square_width: Some(800),
        },
        _ => return None,
    })
}

/// The world of scene `scene`, described into `b`. With the `reference-scenes` feature the trait-object tree is built
/// by the reference's own constructors (scenes.rs) and walked by `Hittable::describe`.
#[cfg(feature = "reference-scenes")]
fn describe_world(scene: usize, b: &mut SceneBuilder) -> Option<i32> {
    use crate::scenes;
    let world = match scene {
        1 => scenes::random_scene(),
        2 => scenes::two_spheres(),
        3 => scenes::two_perlin_spheres(),
        4 => scenes::earth(),
        5 => scenes::simple_light(),
        6 => scenes::empty_cornell_box(),
        7 => scenes::cornell_box(),
        8 => scenes::smoke_cornell_box(),
        9 => scenes::final_scene(),
        _ => return None,
    };
    Some(world.describe(b))
}

/// Renders any world built from this crate's `math` types: the drop-in for the closure of main.rs:199-229.
#[allow(clippy::too_many_arguments)]
pub fn render_world(world: &dyn Hittable, camera: &CameraDescriptor, background: Vec3f<Color>, width: u32, height: u32,
                    samples: usize, seed: u64, device: i32) -> Result<Vec<u8>, RtxError> {
    let mut b = SceneBuilder::new();
    let root = world.describe(&mut b);
    Gpu::new(device)?.render(&b, root, camera.to_ffi(), background.to_array(), width, height, samples, seed)
}

/// `fn render(width, aspect_ratio, samples, scene) -> Option<()>` of src/main.rs:58: prints "Running scene <name>",
/// renders on GPU 0 and saves `image.png`. `None` for an unknown scene (main.rs:179-182) or a failed render.
pub fn render(mut width: u32, mut aspect_ratio: f64, mut samples: usize, scene: usize) -> Option<()> {
    let setup = match scene_setup(scene) {
        Some(s) => s,
        None => {
            eprintln!("There is no scene {}", scene);
            return None;
        }
    };
    println!("Running scene {}", setup.name);
    if let Some(s) = setup.samples {
        samples = s;
    }
    if let Some(w) = setup.square_width {
        aspect_ratio = 1.0;
        width = w;
    }
    let height = (width as f64 / aspect_ratio) as u32; // main.rs:184
    let camera = CameraDescriptor {
        lookfrom: setup.lookfrom, lookat: setup.lookat, view_up: Vec3f::new(0.0, 1.0, 0.0), vertical_fov: setup.vertical_fov,
        aspect_ratio, aperture: setup.aperture, focus_distance: 10.0, open_time: 0.0, close_time: 1.0,
    };
    let seed: u64 = rand::random(); // the reference is unseeded (thread_rng); pass a constant for reproducible frames
    let pixels = render_scene_number(scene, &camera, setup.background, width, height, samples, seed).map_err(|e| {
        eprintln!("render failed ({}): {}", e.status, e.message);
        e
    }).ok()?;
    image::save_buffer("image.png", &pixels, width, height, image::ColorType::Rgba8).unwrap(); // main.rs:231
    Some(())
}

#[cfg(feature = "reference-scenes")]
fn render_scene_number(scene: usize, camera: &CameraDescriptor, background: Vec3f<Color>, width: u32, height: u32, samples: usize,
                       seed: u64) -> Result<Vec<u8>, RtxError> {
    let mut b = SceneBuilder::new();
    let root = describe_world(scene, &mut b).ok_or(RtxError { status: -1, message: format!("There is no scene {}", scene) })?;
    Gpu::new(0)?.render(&b, root, camera.to_ffi(), background.to_array(), width, height, samples, seed)
}

/// Without the reference's scenes.rs: the library's own seeded constructors (`rtx_builtin_scene`), which also carry the
/// camera and background of the scene table; width / height / samples are the caller's.
#[cfg(not(feature = "reference-scenes"))]
fn render_scene_number(scene: usize, _camera: &CameraDescriptor, _background: Vec3f<Color>, width: u32, height: u32, samples: usize,
                       seed: u64) -> Result<Vec<u8>, RtxError> {
    let gpu = Gpu::new(0)?;
    let mut desc: *mut ffi::rtx_scene_desc = ptr::null_mut();
    check(unsafe { ffi::rtx_builtin_scene(scene as i32, 0x5254_544E_57u64 + scene as u64, ptr::null(), &mut desc) })?;
    let pixels = gpu.render_desc(desc, width, height, samples, seed);
    unsafe { ffi::rtx_scene_desc_free(desc) };
    pixels
}

/// Samples sharded over `n` GPUs by global sample index (they are i.i.d., main.rs:211-217): rank r renders
/// `[r * samples / n, (r + 1) * samples / n)` of every pixel on its own device, one host thread per device; rank 0
/// then sums the other accumulators over NVLink and tonemaps in one kernel (`rtx_reduce_tonemap_peers`).
/// (One PROCESS per device brings the CUDA contexts up in parallel — what `rttnw --gpus N` does, rttnw_b200/csrc/cli.cpp —
/// with `rtx_ipc_export` / `rtx_ipc_open` for the accumulators, or `rtx_comm_*` + `rtx_accum_reduce` for NCCL.)
pub fn shard_spp(samples: usize, rank: usize, n: usize) -> (usize, usize) {
    let begin = rank * samples / n;
    (begin, (rank + 1) * samples / n - begin)
}
