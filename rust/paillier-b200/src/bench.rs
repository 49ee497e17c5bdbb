//! Circuit drivers of `/root/reference/src/bench.rs:11-117`, unchanged except for the chip they construct.
use crate::paillier::{EncryptionPublicKeyAssigned, PaillierChip};
use biguint_halo2::big_uint::chip::BigUintChip;
use halo2_base::{
    gates::{circuit::builder::BaseCircuitBuilder, RangeChip},
    halo2_proofs::circuit::Value,
    utils::BigPrimeField,
};
use num_bigint::BigUint;

#[derive(Clone, Debug)]
pub struct PaillierEncryptionInput {
    pub enc_bits: usize,
    pub limb_bits: usize,
    pub n: BigUint,
    pub g: BigUint,
    pub m: BigUint,
    pub r: BigUint,
    pub res: BigUint,
}

#[derive(Clone, Debug)]
pub struct PaillierAddCipherInput {
    pub enc_bits: usize,
    pub limb_bits: usize,
    pub n: BigUint,
    pub g: BigUint,
    pub c1: BigUint,
    pub c2: BigUint,
    pub res: BigUint,
}

/// `/root/reference/src/bench.rs:33-75`
pub fn paillier_enc_test<F: BigPrimeField>(pool: &mut BaseCircuitBuilder<F>, range: &RangeChip<F>, input: PaillierEncryptionInput) {
    let ctx = pool.main(0);
    let biguint_chip = BigUintChip::construct(range, input.limb_bits);
    let paillier_chip = PaillierChip::construct(&biguint_chip, input.enc_bits);
    let n_assigned = biguint_chip.assign_integer(ctx, Value::known(input.n.clone()), input.enc_bits).unwrap();
    let g_assigned = biguint_chip.assign_integer(ctx, Value::known(input.g.clone()), input.enc_bits).unwrap();
    let pk_enc = EncryptionPublicKeyAssigned { n: n_assigned, g: g_assigned };
    let m_assigned = biguint_chip.assign_integer(ctx, Value::known(input.m.clone()), input.enc_bits).unwrap();
    let r_assigned = biguint_chip.assign_integer(ctx, Value::known(input.r.clone()), input.enc_bits).unwrap();
    let c_assigned = paillier_chip.encrypt(ctx, &pk_enc, &m_assigned, &r_assigned).unwrap();
    let res_assigned = biguint_chip.assign_integer(ctx, Value::known(input.res.clone()), input.enc_bits * 2).unwrap();
    c_assigned.value().zip(res_assigned.value()).map(|(a, b)| assert_eq!(a, b));
    biguint_chip.assert_equal_fresh(ctx, &c_assigned, &res_assigned).unwrap();
}

/// `/root/reference/src/bench.rs:77-117`
pub fn paillier_enc_add_test<F: BigPrimeField>(pool: &mut BaseCircuitBuilder<F>, range: &RangeChip<F>, input: PaillierAddCipherInput) {
    let ctx = pool.main(0);
    let biguint_chip = BigUintChip::construct(range, input.limb_bits);
    let paillier_chip = PaillierChip::construct(&biguint_chip, input.enc_bits);
    let n_assigned = biguint_chip.assign_integer(ctx, Value::known(input.n.clone()), input.enc_bits).unwrap();
    let g_assigned = biguint_chip.assign_integer(ctx, Value::known(input.g.clone()), input.enc_bits).unwrap();
    let pk_enc = EncryptionPublicKeyAssigned { n: n_assigned, g: g_assigned };
    let c1_assigned = biguint_chip.assign_integer(ctx, Value::known(input.c1.clone()), input.enc_bits).unwrap();
    let c2_assigned = biguint_chip.assign_integer(ctx, Value::known(input.c2.clone()), input.enc_bits).unwrap();
    let result = paillier_chip.add(ctx, &pk_enc, &c1_assigned, &c2_assigned).unwrap();
    let res_assigned = biguint_chip.assign_integer(ctx, Value::known(input.res.clone()), input.enc_bits * 2).unwrap();
    result.value().zip(res_assigned.value()).map(|(a, b)| assert_eq!(a, b));
    biguint_chip.assert_equal_fresh(ctx, &result, &res_assigned).unwrap();
}
