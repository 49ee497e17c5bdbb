//! Drop-in for `/root/reference/src/paillier.rs`: same items, same signatures.  The body of `encrypt` / `add` is the
//! reference's call sequence line for line; the only difference is `mul_mod_with_witness` / `pow_mod_fixed_exp_with_witness`
//! in place of `mul_mod` / `pow_mod_fixed_exp`, which take `(q, rem)` from the GPU stream instead of calling
//! `BigUint::div_rem` (SURVEY.md Appendix A.4, A.5 [UPSTREAM-RECALL]: everything else `BigUintChip` does is unchanged, so
//! the constraint system — and therefore the proving / verifying keys — are the reference's).
use crate::gpu::{GpuError, GpuWitness, Record};
use biguint_halo2::big_uint::{chip::BigUintChip, AssignedBigUint, Fresh, RefreshAux};
use halo2_base::{
    halo2_proofs::plonk::Error,
    utils::{fe_to_biguint, BigPrimeField},
    Context,
};
use num_bigint::BigUint;
use num_traits::Zero;
use std::cell::RefCell;

pub struct EncryptionPublicKeyAssigned<F: BigPrimeField> {
    pub n: AssignedBigUint<F, Fresh>,
    pub g: AssignedBigUint<F, Fresh>,
}

/// Where the next `(q, rem)` comes from.  `prefetch` is filled by `PaillierChip::prefetch` (one GPU call for a batch of
/// circuits); without it `encrypt` / `add` fetch their own unit on first use.
#[derive(Default)]
pub struct WitnessQueue {
    records: std::collections::VecDeque<Record>,
}

#[derive(Clone)]
pub struct PaillierChip<'a, F: BigPrimeField> {
    pub biguint: &'a BigUintChip<'a, F>,
    pub enc_bits: usize,
    /// shared with clones of the chip, like the `&BigUintChip` borrow (`/root/reference/src/paillier.rs:11-15`)
    gpu: std::rc::Rc<RefCell<Option<GpuWitness>>>,
    queue: std::rc::Rc<RefCell<WitnessQueue>>,
    device: i32,
}

impl<'a, F: BigPrimeField> PaillierChip<'a, F> {
    /// `/root/reference/src/paillier.rs:18-20`.  The device key is created lazily on the first `encrypt` / `add` (it needs n, g).
    pub fn construct(biguint: &'a BigUintChip<'a, F>, enc_bits: usize) -> Self {
        Self { biguint, enc_bits, gpu: Default::default(), queue: Default::default(), device: 0 }
    }
    pub fn on_device(mut self, device: i32) -> Self {
        self.device = device;
        self
    }

    /// `/root/reference/src/paillier.rs:22-30`, unchanged.
    pub fn get_biguint(&self, assigned: &AssignedBigUint<F, Fresh>) -> BigUint {
        assigned
            .limbs()
            .iter()
            .rev()
            .fold(BigUint::zero(), |acc, acell| (acc << assigned.int_ref().max_limb_bits) + fe_to_biguint(acell.value()))
    }

    fn with_gpu<T>(&self, n: &BigUint, g: &BigUint, f: impl FnOnce(&mut GpuWitness) -> Result<T, GpuError>) -> Result<T, Error> {
        let mut slot = self.gpu.borrow_mut();
        let stale = match slot.as_ref() {
            Some(k) => k.n() != n,
            None => true,
        };
        if stale {
            *slot = Some(GpuWitness::new(self.device, n, g, self.enc_bits, self.biguint.limb_bits)?);
        }
        Ok(f(slot.as_mut().unwrap())?)
    }

    /// Batch entry: the witnesses of `encrypt(m_i, r_i)`, i < count, in ONE GPU call; the following `encrypt` calls (one per
    /// circuit, in the same order) consume them.  This is what makes the GPU pay: the reference builds one circuit per ciphertext.
    pub fn prefetch(&self, n: &BigUint, g: &BigUint, m: &[BigUint], r: &[BigUint]) -> Result<Vec<BigUint>, Error> {
        let units = self.with_gpu(n, g, |k| k.encrypt_witness(m, r))?;
        let mut q = self.queue.borrow_mut();
        let mut cs = Vec::with_capacity(units.len());
        for u in units {
            cs.push(u.ciphertext);
            q.records.extend(u.records);
        }
        Ok(cs)
    }

    fn next_record(&self) -> Result<Record, Error> {
        self.queue.borrow_mut().records.pop_front().ok_or(Error::Synthesis)
    }

    /// `BigUintChip::mul_mod` with the witness supplied: assigns q and rem (range-checked), the no-carry products
    /// `a*b` and `q*n`, the sums, and constrains `is_equal_muled` — the same cells in the same order; only the `div_rem` is gone.
    fn mul_mod_with_witness(
        &self,
        ctx: &mut Context<F>,
        a: &AssignedBigUint<F, Fresh>,
        b: &AssignedBigUint<F, Fresh>,
        n: &AssignedBigUint<F, Fresh>,
        w: &Record,
    ) -> Result<AssignedBigUint<F, Fresh>, Error> {
        let big = self.biguint;
        let limb_bits = big.limb_bits;
        let n1 = a.num_limbs();
        assert_eq!(n1, n.num_limbs());
        let bits = n1 * limb_bits;
        let assign_q = big.assign_integer(ctx, halo2_base::halo2_proofs::circuit::Value::known(w.q.clone()), bits)?;
        let assign_rem = big.assign_integer(ctx, halo2_base::halo2_proofs::circuit::Value::known(w.rem.clone()), bits)?;
        let ab = big.mul(ctx, a, b)?;
        let qn = big.mul(ctx, &assign_q, n)?;
        let qn_rem = big.add_muled_fresh(ctx, &qn, &assign_rem)?; // limb-wise qn_i + rem_i for i < n1 (A.4)
        let eq = big.is_equal_muled(ctx, &ab, &qn_rem, n1, n1)?;
        big.gate().assert_is_const(ctx, &eq, &F::ONE);
        Ok(assign_rem)
    }

    /// `BigUintChip::pow_mod_fixed_exp`: LSB-first, the last squaring is assigned although unused.
    fn pow_mod_fixed_exp_with_witness(
        &self,
        ctx: &mut Context<F>,
        a: &AssignedBigUint<F, Fresh>,
        e: &BigUint,
        n: &AssignedBigUint<F, Fresh>,
    ) -> Result<AssignedBigUint<F, Fresh>, Error> {
        let num_limbs = a.num_limbs();
        let zero = ctx.load_zero();
        let mut acc = self.biguint.assign_constant(ctx, BigUint::from(1u32))?.extend_limbs(num_limbs - 1, zero);
        let mut squared = a.clone();
        for i in 0..e.bits() {
            let cur = squared.clone();
            let w = self.next_record()?;
            squared = self.mul_mod_with_witness(ctx, &cur, &cur, n, &w)?;
            if !e.bit(i) {
                continue;
            }
            let w = self.next_record()?;
            acc = self.mul_mod_with_witness(ctx, &acc, &cur, n, &w)?;
        }
        Ok(acc)
    }

    /// `/root/reference/src/paillier.rs:32-60`.
    pub fn encrypt(
        &self,
        ctx: &mut Context<F>,
        pk_enc: &EncryptionPublicKeyAssigned<F>,
        m: &AssignedBigUint<F, Fresh>,
        r: &AssignedBigUint<F, Fresh>,
    ) -> Result<AssignedBigUint<F, Fresh>, Error> {
        let n2 = self.biguint.square(ctx, &pk_enc.n)?;
        let aux = RefreshAux::new(self.biguint.limb_bits, pk_enc.n.num_limbs(), pk_enc.n.num_limbs());
        let n2 = self.biguint.refresh(ctx, &n2, &aux)?;

        let zero_value = ctx.load_zero();

        let g_extended = pk_enc.g.extend_limbs(n2.num_limbs() - pk_enc.g.num_limbs(), zero_value);
        let m_biguint = self.get_biguint(m);
        let n_biguint = self.get_biguint(&pk_enc.n);
        if self.queue.borrow().records.is_empty() {
            // not prefetched: fetch this unit's witnesses now (a batch of one)
            let (g_big, r_big) = (self.get_biguint(&pk_enc.g), self.get_biguint(r));
            self.prefetch(&n_biguint, &g_big, &[m_biguint.clone()], &[r_big])?;
        }
        let gm = self.pow_mod_fixed_exp_with_witness(ctx, &g_extended, &m_biguint, &n2)?;

        let r_extended = r.extend_limbs(n2.num_limbs() - r.num_limbs(), zero_value);
        let rn = self.pow_mod_fixed_exp_with_witness(ctx, &r_extended, &n_biguint, &n2)?;

        let w = self.next_record()?;
        let c = self.mul_mod_with_witness(ctx, &gm, &rn, &n2, &w)?;

        Ok(c)
    }

    /// `/root/reference/src/paillier.rs:62-85`.
    pub fn add(
        &self,
        ctx: &mut Context<F>,
        pk_enc: &EncryptionPublicKeyAssigned<F>,
        c1: &AssignedBigUint<F, Fresh>,
        c2: &AssignedBigUint<F, Fresh>,
    ) -> Result<AssignedBigUint<F, Fresh>, Error> {
        let n2 = self.biguint.square(ctx, &pk_enc.n)?;
        let aux = RefreshAux::new(self.biguint.limb_bits, pk_enc.n.num_limbs(), pk_enc.n.num_limbs());
        let n2 = self.biguint.refresh(ctx, &n2, &aux)?;

        let zero_value = ctx.load_zero();

        let c1_extended = c1.extend_limbs(n2.num_limbs() - c1.num_limbs(), zero_value);
        let c2_extended = c2.extend_limbs(n2.num_limbs() - c2.num_limbs(), zero_value);
        let (n_big, g_big) = (self.get_biguint(&pk_enc.n), self.get_biguint(&pk_enc.g));
        let (a, b) = (self.get_biguint(c1), self.get_biguint(c2));
        let w = self.with_gpu(&n_big, &g_big, |k| k.paillier_add_native(&[a], &[b]))?.remove(0);
        let result = self.mul_mod_with_witness(ctx, &c1_extended, &c2_extended, &n2, &w)?;

        Ok(result)
    }
}

/// `/root/reference/src/paillier.rs:87-92` on the GPU (a batch of one; use `GpuWitness::paillier_enc_native` for batches).
pub fn paillier_enc_native(n: &BigUint, g: &BigUint, m: &BigUint, r: &BigUint) -> BigUint {
    let bits = n.bits().max(g.bits()).max(m.bits()).max(r.bits()).max(1) as usize;
    let enc_bits = (bits + 63) / 64 * 64;
    let mut k = GpuWitness::new(0, n, g, enc_bits, 64).expect("pb200_key_create (num-bigint panics on a zero modulus too)");
    k.paillier_enc_native(&[m.clone()], &[r.clone()]).expect("pb200_encrypt_batch").remove(0)
}

/// `/root/reference/src/paillier.rs:94-97` on the GPU.
pub fn paillier_add_native(n: &BigUint, c1: &BigUint, c2: &BigUint) -> BigUint {
    let bits = (2 * n.bits()).max(c1.bits()).max(c2.bits()).max(1) as usize;
    let enc_bits = (bits + 127) / 128 * 64;
    let mut k = GpuWitness::new(0, n, &BigUint::from(1u32), enc_bits, 64).expect("pb200_key_create");
    k.paillier_add_native(&[c1.clone()], &[c2.clone()]).expect("pb200_add_batch").remove(0).rem
}
