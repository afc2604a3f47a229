//! The nine scene constructors of the reference — THE REFERENCE'S OWN FILE, compiled inside this crate.
//!
//! `src/scenes.rs` of luliic2/rttnw starts with `use crate::math::{BvhTree, CheckerTexture, ..., XY, XZ, YZ}`; included
//! here, `crate::math` is this crate's description-only mirror (math.rs), which exports every name that list asks
//! for (also `XY / XZ / YZ`, which the reference's own `math/mod.rs:13` forgets — Q28). Nothing of the file is copied
//! into this repository: build.rs points `RTTNW_REFERENCE_SRC` at a checkout of the reference
//! (`RTTNW_REFERENCE_DIR=/path/to/rttnw cargo build --features reference-scenes`).
//!
//! Without the feature the scenes come from the library's own seeded constructors (`rtx_builtin_scene`, the C++
//! restatement in rttnw_b200/csrc/scenes.cpp that the parity tests compare with the oracle's): see `render::render`.
#[cfg(feature = "reference-scenes")]
include!(concat!(env!("RTTNW_REFERENCE_SRC"), "/scenes.rs"));
