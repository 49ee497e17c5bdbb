//! The acceptance test the reference runs (`/root/reference/src/paillier.rs:113-182`: `base_test().k(16).lookup_bits(15)
//! .expect_satisfied(true).run(..)`) on the GPU-fed chip, AND the pin of this repository's oracle:
//!
//!   1. MockProver must accept the circuit whose every (q, rem) came from the GPU;
//!   2. every advice cell `ctx` holds after synthesis is dumped and hashed; for the seeded inputs of
//!      `tests/golden/cells.json` the SHA-256 over the 32-byte little-endian cells must equal the `sha256` recorded there by the
//!      Python restatement (`oracle/paillier_oracle.py`).  Until this has run once, the cell ORDER and the
//!      set of cells halo2-base's `range_check` / `inner_product` assign are "parity unpinned" (DESIGN.md §6): a mismatch here
//!      is expected to show exactly which gate-internal cells the restatement does not model, and the dump
//!      (`target/advice_cells_<tag>.hex`) is what to diff.
//!
//! Run on a box with a Rust toolchain and a B200:   PB200_GOLDEN=../../tests/golden/cells.json cargo test --release -- --nocapture
use biguint_halo2::big_uint::chip::BigUintChip;
use halo2_base::{
    halo2_proofs::{circuit::Value, halo2curves::bn256::Fr},
    utils::{fe_to_biguint, testing::base_test},
};
use num_bigint::{BigUint, RandBigInt};
use paillier_b200::paillier::{paillier_add_native, paillier_enc_native, EncryptionPublicKeyAssigned, PaillierChip};
use rand::thread_rng;
use sha2::{Digest, Sha256};

fn dump_cells(tag: &str, cells: &[BigUint]) -> String {
    let mut h = Sha256::new();
    let mut text = String::new();
    for c in cells {
        let mut b = c.to_bytes_le();
        b.resize(32, 0);
        h.update(&b);
        text.push_str(&hex::encode(&b));
        text.push('\n');
    }
    std::fs::create_dir_all("target").ok();
    std::fs::write(format!("target/advice_cells_{tag}.hex"), text).unwrap();
    hex::encode(h.finalize())
}

fn run_encrypt(enc_bits: usize, limb_bits: usize, n: BigUint, g: BigUint, m: BigUint, r: BigUint, tag: &str) -> String {
    let expected = paillier_enc_native(&n, &g, &m, &r);
    let mut digest = String::new();
    base_test().k(16).lookup_bits(15).expect_satisfied(true).run(|ctx, range| {
        let biguint_chip = BigUintChip::<Fr>::construct(range, limb_bits);
        let paillier_chip = PaillierChip::construct(&biguint_chip, enc_bits);
        let n_assigned = biguint_chip.assign_integer(ctx, Value::known(n.clone()), enc_bits).unwrap();
        let g_assigned = biguint_chip.assign_integer(ctx, Value::known(g.clone()), enc_bits).unwrap();
        let pk_enc = EncryptionPublicKeyAssigned { n: n_assigned, g: g_assigned };
        let m_assigned = biguint_chip.assign_integer(ctx, Value::known(m.clone()), enc_bits).unwrap();
        let r_assigned = biguint_chip.assign_integer(ctx, Value::known(r.clone()), enc_bits).unwrap();
        let c_assigned = paillier_chip.encrypt(ctx, &pk_enc, &m_assigned, &r_assigned).unwrap();
        let res_assigned = biguint_chip.assign_integer(ctx, Value::known(expected.clone()), enc_bits * 2).unwrap();
        c_assigned.value().zip(res_assigned.value()).map(|(a, b)| assert_eq!(a, b));
        biguint_chip.assert_equal_fresh(ctx, &c_assigned, &res_assigned).unwrap();
        // every advice value of the context, in assignment order
        let cells: Vec<BigUint> = ctx.advice.iter().map(|a| fe_to_biguint(&a.evaluate())).collect();
        digest = dump_cells(tag, &cells);
    });
    digest
}

#[test]
fn test_paillier_encryption_gpu_fed() {
    // the reference's own distribution (`/root/reference/src/paillier.rs:173-176`): unseeded, n may be even or short
    const ENC_BIT_LEN: usize = 128;
    const LIMB_BIT_LEN: usize = 64;
    let mut rng = thread_rng();
    let (n, g, m, r) = (rng.gen_biguint(128), rng.gen_biguint(128), rng.gen_biguint(128), rng.gen_biguint(128));
    if n.bits() == 0 {
        return; // num-bigint would panic on a zero modulus
    }
    run_encrypt(ENC_BIT_LEN, LIMB_BIT_LEN, n, g, m, r, "random");
}

#[test]
fn pin_oracle_cell_streams() {
    let path = std::env::var("PB200_GOLDEN").unwrap_or_else(|_| "../../tests/golden/cells.json".into());
    let golden: serde_json::Value = serde_json::from_str(&std::fs::read_to_string(path).unwrap()).unwrap();
    let hexint = |v: &serde_json::Value| BigUint::parse_bytes(v.as_str().unwrap().trim_start_matches("0x").as_bytes(), 16).unwrap();
    let mut mismatches = Vec::new();
    for (idx, case) in golden["flows"].as_array().unwrap().iter().enumerate() {
        if case["lookup_bits"].as_u64() != Some(15) {
            continue; // base_test().lookup_bits(15); the cells of ctx are canonical integers, as in the fixture
        }
        let tag = format!("flow{idx}");
        let tag = tag.as_str();
        let got = run_encrypt(case["enc_bits"].as_u64().unwrap() as usize, case["limb_bits"].as_u64().unwrap() as usize,
                              hexint(&case["n"]), hexint(&case["g"]), hexint(&case["m"]), hexint(&case["r"]), tag);
        if got != case["sha256"].as_str().unwrap() {
            mismatches.push(format!("{tag}: cells differ from the oracle's stream (diff target/advice_cells_{tag}.hex against `python tools/gen_golden_cells.py --dump {idx}`)"));
        }
    }
    assert!(mismatches.is_empty(), "oracle NOT pinned:\n{}", mismatches.join("\n"));
}

#[test]
fn test_encryption_addition_gpu_fed() {
    // `/root/reference/src/paillier.rs:184-259`: 264-bit values on 88-bit limbs, c1 and c2 assigned with enc_bits
    const ENC_BIT_LEN: usize = 264;
    const LIMB_BIT_LEN: usize = 88;
    let mut rng = thread_rng();
    let (n, g, c1, c2) = (rng.gen_biguint(264), rng.gen_biguint(264), rng.gen_biguint(264), rng.gen_biguint(264));
    if n.bits() == 0 {
        return;
    }
    let expected = paillier_add_native(&n, &c1, &c2);
    base_test().k(16).lookup_bits(15).expect_satisfied(true).run(|ctx, range| {
        let biguint_chip = BigUintChip::<Fr>::construct(range, LIMB_BIT_LEN);
        let paillier_chip = PaillierChip::construct(&biguint_chip, ENC_BIT_LEN);
        let n_assigned = biguint_chip.assign_integer(ctx, Value::known(n.clone()), ENC_BIT_LEN).unwrap();
        let g_assigned = biguint_chip.assign_integer(ctx, Value::known(g.clone()), ENC_BIT_LEN).unwrap();
        let pk_enc = EncryptionPublicKeyAssigned { n: n_assigned, g: g_assigned };
        let c1_assigned = biguint_chip.assign_integer(ctx, Value::known(c1.clone()), ENC_BIT_LEN).unwrap();
        let c2_assigned = biguint_chip.assign_integer(ctx, Value::known(c2.clone()), ENC_BIT_LEN).unwrap();
        let result = paillier_chip.add(ctx, &pk_enc, &c1_assigned, &c2_assigned).unwrap();
        let res_assigned = biguint_chip.assign_integer(ctx, Value::known(expected.clone()), ENC_BIT_LEN * 2).unwrap();
        result.value().zip(res_assigned.value()).map(|(a, b)| assert_eq!(a, b));
        biguint_chip.assert_equal_fresh(ctx, &result, &res_assigned).unwrap();
    });
}
