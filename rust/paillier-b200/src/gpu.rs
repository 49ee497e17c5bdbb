//! Host side of the witness stream: one `pb200_key` per (n, g), the per-key g-chain, and per-unit record streams.
use num_bigint::BigUint;
use paillier_b200_sys as sys;
use std::os::raw::{c_int, c_void};

/// Library status mapped to the error type the chip returns (`halo2_proofs::plonk::Error`, `/root/reference/src/paillier.rs:38`).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub struct GpuError(pub i32);

impl From<GpuError> for halo2_base::halo2_proofs::plonk::Error {
    fn from(_: GpuError) -> Self {
        halo2_base::halo2_proofs::plonk::Error::Synthesis
    }
}

fn check(rc: c_int) -> Result<(), GpuError> {
    if rc == sys::PB200_OK { Ok(()) } else { Err(GpuError(rc)) }
}

/// `BigUint::to_u64_digits()` is little-endian 64-bit words: exactly the ABI's layout (include/paillier_b200.h:11-12) and the
/// order `PaillierChip::get_biguint` folds (`/root/reference/src/paillier.rs:22-30`).
pub fn to_words(v: &BigUint, words: usize) -> Result<Vec<u64>, GpuError> {
    let mut w = v.to_u64_digits();
    if w.len() > words {
        return Err(GpuError(sys::PB200_ERR_RANGE));
    }
    w.resize(words, 0);
    Ok(w)
}
pub fn from_words(w: &[u64]) -> BigUint {
    let mut bytes = Vec::with_capacity(w.len() * 8);
    for x in w {
        bytes.extend_from_slice(&x.to_le_bytes());
    }
    BigUint::from_bytes_le(&bytes)
}

/// One `(q, rem)` of a `mul_mod(a, b, n^2)`: `q = floor(a b / n^2)`, `rem = a b mod n^2`.
#[derive(Clone, Debug, PartialEq, Eq)]
pub struct Record {
    pub q: BigUint,
    pub rem: BigUint,
}

/// The records of ONE `PaillierChip::encrypt` call in the order the chip issues its `mul_mod`s
/// (`/root/reference/src/paillier.rs:51,55,57`; `paillier_halo2_b200.api.chip_order` is the tested Python statement of this walk).
pub struct UnitWitness {
    pub ciphertext: BigUint,
    pub records: Vec<Record>,
}

/// Owns the device key of one public key.  Not `Sync`: calls on one key are serialised, like `&mut Context<F>`.
pub struct GpuWitness {
    key: sys::Key,
    n: BigUint,
    enc_bits: usize,
    words_in: usize,
    words_out: usize,
    g_chain: Vec<Record>, // record i = square of g^(2^i), i < enc_bits (per key, shared by every unit)
}

struct Collect {
    words_out: usize,
    units: Vec<Vec<Record>>,
}

extern "C" fn collect_sink(user: *mut c_void, chunk: *const sys::pb200_witness_chunk) -> c_int {
    // SAFETY: `user` is the `Collect` passed by `unit_streams` below and outlives the call; the chunk's pointers are valid
    // for the duration of the callback (include/paillier_b200.h, pb200_witness_sink_fn).
    let (c, ch) = unsafe { (&mut *(user as *mut Collect), &*chunk) };
    let wo = c.words_out;
    let offs = unsafe { std::slice::from_raw_parts(ch.offsets, ch.n_units + 1) };
    let recs = unsafe { std::slice::from_raw_parts(ch.records, offs[ch.n_units] as usize * 2 * wo) };
    for u in 0..ch.n_units {
        let mut v = Vec::with_capacity((offs[u + 1] - offs[u]) as usize);
        for r in offs[u] as usize..offs[u + 1] as usize {
            let base = r * 2 * wo;
            v.push(Record { q: from_words(&recs[base..base + wo]), rem: from_words(&recs[base + wo..base + 2 * wo]) });
        }
        c.units.push(v);
    }
    0
}

impl GpuWitness {
    /// Replaces `EncryptionPublicKeyAssigned{n, g}` + the per-call `square(n)` / `refresh` VALUES
    /// (`/root/reference/src/paillier.rs:6-9,39-45`).
    pub fn new(device: i32, n: &BigUint, g: &BigUint, enc_bits: usize, limb_bits: usize) -> Result<Self, GpuError> {
        let words_in = (enc_bits + 63) / 64;
        let key = sys::Key::new(device, enc_bits as u32, limb_bits as u32, &to_words(n, words_in)?, &to_words(g, words_in)?)
            .map_err(GpuError)?;
        let words_out = key.words_out();
        let mut me = GpuWitness { key, n: n.clone(), enc_bits, words_in, words_out, g_chain: Vec::new() };
        let mut raw = vec![0u64; enc_bits * 2 * words_out];
        check(unsafe { sys::pb200_key_g_chain(me.key.as_ptr(), raw.as_mut_ptr()) })?;
        me.g_chain = (0..enc_bits)
            .map(|i| Record {
                q: from_words(&raw[i * 2 * words_out..i * 2 * words_out + words_out]),
                rem: from_words(&raw[i * 2 * words_out + words_out..(i + 1) * 2 * words_out]),
            })
            .collect();
        Ok(me)
    }

    /// Batched `paillier_enc_native` (`/root/reference/src/paillier.rs:87-92`).
    pub fn paillier_enc_native(&mut self, m: &[BigUint], r: &[BigUint]) -> Result<Vec<BigUint>, GpuError> {
        let (mw, rw) = (self.pack(m)?, self.pack(r)?);
        let out = self.key.encrypt_batch(&mw, &rw).map_err(GpuError)?;
        Ok(out.chunks(self.words_out).map(from_words).collect())
    }

    /// Batched `paillier_add_native` (`/root/reference/src/paillier.rs:94-97`) with the `mul_mod` quotient.
    pub fn paillier_add_native(&mut self, c1: &[BigUint], c2: &[BigUint]) -> Result<Vec<Record>, GpuError> {
        let wo = self.words_out;
        let pack = |v: &[BigUint]| -> Result<Vec<u64>, GpuError> {
            let mut w = Vec::with_capacity(v.len() * wo);
            for x in v {
                w.extend(to_words(x, wo)?);
            }
            Ok(w)
        };
        let (rem, q) = self.key.add_batch(&pack(c1)?, &pack(c2)?).map_err(GpuError)?;
        Ok(rem.chunks(wo).zip(q.chunks(wo)).map(|(r, q)| Record { q: from_words(q), rem: from_words(r) }).collect())
    }

    fn pack(&self, v: &[BigUint]) -> Result<Vec<u64>, GpuError> {
        let mut w = Vec::with_capacity(v.len() * self.words_in);
        for x in v {
            w.extend(to_words(x, self.words_in)?);
        }
        Ok(w)
    }

    /// The witnesses of `encrypt(m_i, r_i)` for a whole batch: ONE GPU call, then the per-unit streams are interleaved with the
    /// per-key g-chain squarings into chip order:
    ///   g-chain, bit i of m low to high:  sqr_i (per key) [, mul (unit stream)];   then the unit's r-chain records;   then the final one.
    pub fn encrypt_witness(&mut self, m: &[BigUint], r: &[BigUint]) -> Result<Vec<UnitWitness>, GpuError> {
        let (mw, rw) = (self.pack(m)?, self.pack(r)?);
        let mut c_out = vec![0u64; m.len() * self.words_out];
        let mut sink = Collect { words_out: self.words_out, units: Vec::with_capacity(m.len()) };
        check(unsafe {
            sys::pb200_encrypt_witness_batch(self.key.as_ptr(), mw.as_ptr(), rw.as_ptr(), m.len(), c_out.as_mut_ptr(), 0, collect_sink,
                                             &mut sink as *mut Collect as *mut c_void)
        })?;
        let mut out = Vec::with_capacity(m.len());
        for (u, unit) in sink.units.into_iter().enumerate() {
            let mut it = unit.into_iter();
            let mut records = Vec::new();
            for i in 0..m[u].bits() as usize {
                records.push(self.g_chain[i].clone());
                if m[u].bit(i as u64) {
                    records.push(it.next().expect("g-chain mul record"));
                }
            }
            records.extend(it);
            out.push(UnitWitness { ciphertext: from_words(&c_out[u * self.words_out..(u + 1) * self.words_out]), records });
        }
        Ok(out)
    }

    pub fn n(&self) -> &BigUint { &self.n }
    pub fn enc_bits(&self) -> usize { self.enc_bits }
}
