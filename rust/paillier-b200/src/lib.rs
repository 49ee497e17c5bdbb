//! `PaillierChip` of aerius-labs/paillier-halo2 with the arithmetic moved to the GPU.
//!
//! Same public surface as the reference crate (`src/lib.rs:1-2` there: `pub mod bench; pub mod paillier;`), same names and
//! signatures in `paillier` (`EncryptionPublicKeyAssigned`, `PaillierChip::{construct, get_biguint, encrypt, add}`,
//! `paillier_enc_native`, `paillier_add_native`).  What changes is where the values come from: every `(q, rem)` that
//! `BigUintChip::mul_mod` would compute with num-bigint arrives from `libpaillier_b200` through `gpu::GpuWitness`.
//!
//! STATUS: source only.  This repository's build environment has no Rust toolchain, so nothing in this crate has met a
//! compiler, and the halo2 / biguint-halo2 calls are written against those crates' APIs as the reference uses them
//! (`/root/reference/src/paillier.rs:1-2,39-57`) plus the `BigUintChip` methods listed in SURVEY.md Appendix A
//! [UPSTREAM-RECALL].  `tests/mockprover_cells.rs` is the first thing to run on a box that has `cargo`: it is both the
//! MockProver acceptance test of the GPU-fed circuit and the pin of this repository's oracle.
pub mod bench;
pub mod gpu;
pub mod paillier;
