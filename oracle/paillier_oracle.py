"""CPU oracle for the paillier-halo2 hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, with exact Python integers, the algorithm of the reference's batched-Paillier
path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it; the product path (`paillier_halo2_b200/`) never does and fails loudly when the
CUDA extension is missing.

What is restated, and from where (citations into /root/reference):

  * `paillier_enc_native`, `paillier_add_native`      src/paillier.rs:87-92, :94-97
  * `PaillierChip::get_biguint` limb order             src/paillier.rs:22-30  (little-endian limbs)
  * `PaillierChip::encrypt` / `add` op sequence        src/paillier.rs:32-60, :62-85
  * driver order (assign n,g,m,r -> encrypt -> assign res -> equality)   src/bench.rs:33-75, :77-117

The arithmetic below the chip lives in an un-vendored, un-pinned git dependency
(`biguint-halo2`, Cargo.toml:11, default-branch HEAD, no Cargo.lock) plus `halo2-base`
(Cargo.toml:9) and `num-bigint 0.4.4` (Cargo.toml:12).  None of their sources exist in this
container, so the witness semantics follow the published algorithm of
`biguint-halo2::big_uint::chip::BigUintChip` as recorded in SURVEY.md Appendix A
(assign_integer, mul, refresh/RefreshAux, mul_mod, pow_mod_fixed_exp, is_equal_muled).

PARITY PINNING
  * ciphertext VALUES: pinned.  Integers are integers: any exact bignum equals num-bigint.  The
    oracle is cross-checked against OpenSSL BN (oracle/paillier_ref.cpp) and GMP (ctypes) in
    tests/test_oracle.py, and against the mathematical identities the reference README states
    (README.md:10,22: c = g^m r^n mod n^2, c1*c2 mod n^2; (n+1)^m = 1 + m n mod n^2).
  * per-step WITNESS limbs and their ORDER: "parity unpinned" — the reference holds no golden
    vectors or KATs for them and its tests only assert the end value plus
    MockProver satisfiability.  q and rem of every mul_mod are nevertheless unique given
    (a, b, n^2) and the range checks, so the values are forced; the constraint re-checker
    (`check_constraints`) is the MockProver stand-in.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------
# value semantics  (src/paillier.rs:87-97)
# --------------------------------------------------------------------------------------


def paillier_enc_native(n: int, g: int, m: int, r: int) -> int:
    """src/paillier.rs:87-92 — n2 = n*n; gm = g.modpow(m,n2); rn = r.modpow(n,n2); (gm*rn) % n2.

    num-bigint's modpow panics on a zero modulus; mirrored as ZeroDivisionError.
    """
    n2 = n * n
    if n2 == 0:
        raise ZeroDivisionError("modpow with zero modulus (num-bigint panics)")
    gm = pow(g, m, n2)
    rn = pow(r, n, n2)
    return (gm * rn) % n2


def paillier_add_native(n: int, c1: int, c2: int) -> int:
    """src/paillier.rs:94-97 — n2 = n*n; (c1*c2) % n2."""
    n2 = n * n
    if n2 == 0:
        raise ZeroDivisionError("remainder by zero modulus (num-bigint panics)")
    return (c1 * c2) % n2


def tally_native(n: int, cs: Sequence[int]) -> int:
    """N-ary fold of paillier_add_native (BASELINE.json config 3); empty product is 1 mod n^2."""
    n2 = n * n
    acc = 1 % n2
    for c in cs:
        acc = (acc * c) % n2
    return acc


# --------------------------------------------------------------------------------------
# limb helpers  (src/paillier.rs:22-30 fixes the order: limb i has weight 2^(i*limb_bits))
# --------------------------------------------------------------------------------------


def decompose(v: int, num_limbs: int, limb_bits: int) -> List[int]:
    """`decompose_biguint`: i-th limb = (v >> i*limb_bits) & (B-1)."""
    mask = (1 << limb_bits) - 1
    out = [(v >> (i * limb_bits)) & mask for i in range(num_limbs)]
    if v >> (num_limbs * limb_bits):
        raise ValueError("value does not fit the requested limbs (range check would fail)")
    return out


def get_biguint(limbs: Sequence[int], limb_bits: int) -> int:
    """src/paillier.rs:22-30 — fold MSB->LSB: (acc << max_limb_bits) + limb."""
    acc = 0
    for l in reversed(limbs):
        acc = (acc << limb_bits) + l
    return acc


def to_u64_le(v: int, n64: int) -> List[int]:
    return decompose(v, n64, 64)


def from_u64_le(limbs: Sequence[int]) -> int:
    return get_biguint(limbs, 64)


# --------------------------------------------------------------------------------------
# witness semantics of biguint-halo2
# --------------------------------------------------------------------------------------


@dataclass
class MulModStep:
    """One `mul_mod` witness group: q = floor(a*b / n2), rem = a*b mod n2."""

    kind: str  # "sqr" (square_mod inside pow_mod_fixed_exp), "mul" (acc*cur), "final", "add"
    a: int
    b: int
    q: int
    rem: int


@dataclass
class Context:
    """Stand-in for halo2-base `Context<F>`: an append-only list of advice cells (ints that never
    wrap the BN254 scalar field: max value < 2^135) plus a list of constraint
    closures re-evaluated by `check_constraints`."""

    cells: List[int] = field(default_factory=list)
    checks: List[Tuple[str, bool]] = field(default_factory=list)
    steps: List[MulModStep] = field(default_factory=list)

    def load(self, v: int) -> int:
        assert v >= 0
        self.cells.append(v)
        return v

    def constrain(self, what: str, ok: bool) -> None:
        self.checks.append((what, bool(ok)))


BN254_FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617


@dataclass
class Assigned:
    """AssignedBigUint<F, Fresh|Muled>: limbs (little-endian) + the integer they represent."""

    limbs: List[int]
    value: int
    limb_bits: int
    fresh: bool = True

    def num_limbs(self) -> int:
        return len(self.limbs)

    def extend_limbs(self, k: int) -> "Assigned":
        """AssignedBigUint::extend_limbs (src/paillier.rs:49,53,79-80): zero-pad by k limbs."""
        return Assigned(self.limbs + [0] * k, self.value, self.limb_bits, self.fresh)


class RefreshAux:
    """RefreshAux::new(limb_bits, nl, nr): simulate all-(B-1) operands to find how far
    each column sum of a product can spill."""

    def __init__(self, limb_bits: int, nl: int, nr: int):
        self.limb_bits = limb_bits
        B = 1 << limb_bits
        n_cols = nl + nr - 1
        # worst-case column sums of a*b with all limbs B-1
        cols = [0] * n_cols
        for i in range(nl):
            for j in range(nr):
                cols[i + j] += (B - 1) * (B - 1)
        inc: List[int] = []
        vals = list(cols)
        i = 0
        while i < len(vals):
            v = vals[i]
            cnt = 0
            carry = v >> limb_bits
            k = 1
            while carry:
                if i + k >= len(vals):
                    vals.append(0)
                vals[i + k] += carry & (B - 1)
                carry >>= limb_bits
                cnt += 1
                k += 1
            vals[i] = v & (B - 1)
            inc.append(cnt)
            i += 1
        self.increased_limbs_vec = inc
        self.num_limbs_out = len(inc)


class BigUintChip:
    """Restatement of biguint-halo2 `BigUintChip` witness generation.

    `witness_source`: optional iterator of externally produced (q, rem) pairs, consumed by `mul_mod` in call
    order INSTEAD of computing div_rem here — this is how the tests feed the GPU-produced witness stream
    through the chip and let the constraints (not the oracle's own division) decide whether it is accepted."""

    def __init__(self, limb_bits: int, lookup_bits: Optional[int] = None, witness_source=None):
        self.limb_bits = limb_bits
        self.lookup_bits = lookup_bits
        self.B = 1 << limb_bits
        self.witness_source = iter(witness_source) if witness_source is not None else None

    # A.1 ------------------------------------------------------------------------------
    def range_check(self, ctx: Context, v: int, bits: int) -> None:
        ctx.constrain(f"range_check({bits})", 0 <= v < (1 << bits))
        if self.lookup_bits:
            lb = self.lookup_bits
            k = -(-bits // lb)
            chunks = decompose(v, k, lb) if v < (1 << (k * lb)) else [v]
            for c in chunks:
                ctx.load(c)
            if bits % lb:
                ctx.load(chunks[-1] << (lb - bits % lb))

    def assign_integer(self, ctx: Context, v: int, bit_len: int) -> Assigned:
        if bit_len % self.limb_bits != 0:
            raise AssertionError("assign_integer: bit_len % limb_bits != 0")
        nl = bit_len // self.limb_bits
        if v >> bit_len:
            ctx.constrain("assign_integer: value fits bit_len", False)
            v &= (1 << bit_len) - 1
        limbs = decompose(v, nl, self.limb_bits)
        for l in limbs:
            ctx.load(l)
            self.range_check(ctx, l, self.limb_bits)
        return Assigned(limbs, v, self.limb_bits, True)

    def assign_constant(self, ctx: Context, v: int, num_limbs: int = 1) -> Assigned:
        limbs = decompose(v, num_limbs, self.limb_bits)
        for l in limbs:
            ctx.load(l)
        return Assigned(limbs, v, self.limb_bits, True)

    # A.2 ------------------------------------------------------------------------------
    def mul(self, ctx: Context, a: Assigned, b: Assigned) -> Assigned:
        n1, n2 = a.num_limbs(), b.num_limbs()
        n = n1 + n2 - 1
        al = a.limbs + [0] * (n - n1)
        bl = b.limbs + [0] * (n - n2)
        out = []
        for i in range(n):
            c = 0
            for j in range(i + 1):
                c += al[j] * bl[i - j]
            ctx.load(c)
            out.append(c)
        ctx.constrain("mul columns below Fr", all(c < BN254_FR for c in out))
        return Assigned(out, a.value * b.value, self.limb_bits, False)

    def square(self, ctx: Context, a: Assigned) -> Assigned:
        return self.mul(ctx, a, a)

    # A.3 ------------------------------------------------------------------------------
    def div_mod_unsafe(self, ctx: Context, a: int, B: int) -> Tuple[int, int]:
        q, r = divmod(a, B)
        ctx.load(q)
        ctx.load(r)
        ctx.constrain("div_mod_unsafe r == a - q*B", r == a - q * B)
        return q, r

    def refresh(self, ctx: Context, a: Assigned, aux: RefreshAux) -> Assigned:
        assert aux.limb_bits == self.limb_bits
        inc = aux.increased_limbs_vec
        n_out = aux.num_limbs_out
        x = list(a.limbs) + [0] * (n_out - a.num_limbs())
        for i in range(a.num_limbs()):
            limb = x[i]
            for j in range(inc[i] + 1):
                q, r = self.div_mod_unsafe(ctx, limb, self.B)
                if j == 0:
                    x[i] = r
                else:
                    x[i + j] += r
                limb = q
            ctx.constrain("refresh: final carry zero", limb == 0)
        for l in x:
            self.range_check(ctx, l, self.limb_bits)
        out = Assigned(x, get_biguint(x, self.limb_bits), self.limb_bits, True)
        ctx.constrain("refresh preserves value", out.value == a.value)
        return out

    # A.6 ------------------------------------------------------------------------------
    def is_equal_muled(self, ctx: Context, a: Sequence[int], b: Sequence[int], nl: int, nr: int) -> int:
        B = self.B
        min_n = min(nl, nr)
        word_max = min_n * (B - 1) * (B - 1) + (B - 1)
        carry_bits = (2 * word_max).bit_length() - self.limb_bits
        n = len(a)
        assert len(b) == n
        eq = 1
        carry = 0
        acc_extra = 0
        for i in range(n):
            s = a[i] - b[i] + carry + word_max
            ctx.constrain("is_equal_muled: sum non-negative", s >= 0)
            if s < 0:
                s = 0
            carry_next, cs = self.div_mod_unsafe(ctx, s, B)
            acc_extra += word_max
            q_acc, mod_acc = self.div_mod_unsafe(ctx, acc_extra, B)
            eq &= int(cs == mod_acc)
            acc_extra = q_acc
            if i < n - 1:
                self.range_check(ctx, carry_next, carry_bits)
            else:
                eq &= int(carry_next == acc_extra)
            carry = carry_next
        ctx.load(eq)
        return eq

    # A.4 ------------------------------------------------------------------------------
    def mul_mod(self, ctx: Context, a: Assigned, b: Assigned, n: Assigned, kind: str = "mul") -> Assigned:
        L = n.num_limbs()
        assert a.num_limbs() == L and b.num_limbs() == L, "mul_mod: limb counts differ"
        full = a.value * b.value
        if n.value == 0:
            raise ZeroDivisionError("mul_mod by zero modulus")
        if self.witness_source is not None:
            q, rem = next(self.witness_source)      # GPU-fed witness: accepted or rejected by the constraints below
        else:
            q, rem = divmod(full, n.value)
        ctx.steps.append(MulModStep(kind, a.value, b.value, q, rem))
        q_as = self.assign_integer(ctx, q, L * self.limb_bits)
        rem_as = self.assign_integer(ctx, rem, L * self.limb_bits)
        ab = self.mul(ctx, a, b)
        qn = self.mul(ctx, q_as, n)
        qn_prod = []
        for i in range(2 * L - 1):
            v = qn.limbs[i] + rem_as.limbs[i] if i < L else qn.limbs[i]
            ctx.load(v)
            qn_prod.append(v)
        eq = self.is_equal_muled(ctx, ab.limbs, qn_prod, L, L)
        ctx.constrain("mul_mod: ab == q*n + rem", eq == 1)
        return rem_as

    def square_mod(self, ctx: Context, a: Assigned, n: Assigned) -> Assigned:
        return self.mul_mod(ctx, a, a, n, kind="sqr")

    # A.5 ------------------------------------------------------------------------------
    def pow_mod_fixed_exp(self, ctx: Context, a: Assigned, e: int, n: Assigned) -> Assigned:
        L = n.num_limbs()
        assert a.num_limbs() == L
        acc = self.assign_constant(ctx, 1).extend_limbs(L - 1)
        squared = a
        for i in range(e.bit_length()):
            cur = squared
            squared = self.square_mod(ctx, cur, n)
            if not (e >> i) & 1:
                continue
            acc = self.mul_mod(ctx, acc, cur, n, kind="mul")
        return acc

    def assert_equal_fresh(self, ctx: Context, a: Assigned, b: Assigned) -> None:
        ctx.constrain("assert_equal_fresh: limb counts", a.num_limbs() == b.num_limbs())
        ctx.constrain("assert_equal_fresh: limbs", list(a.limbs) == list(b.limbs))


@dataclass
class EncryptionPublicKeyAssigned:
    """src/paillier.rs:6-9."""

    n: Assigned
    g: Assigned


class PaillierChip:
    """Restatement of `PaillierChip` (src/paillier.rs:11-85) over the BigUintChip restatement."""

    def __init__(self, biguint: BigUintChip, enc_bits: int):
        # src/paillier.rs:18-20 — enc_bits is stored and never read by encrypt/add.
        self.biguint = biguint
        self.enc_bits = enc_bits

    @classmethod
    def construct(cls, biguint: BigUintChip, enc_bits: int) -> "PaillierChip":
        return cls(biguint, enc_bits)

    def get_biguint(self, a: Assigned) -> int:
        return get_biguint(a.limbs, a.limb_bits)

    def _n2(self, ctx: Context, pk: EncryptionPublicKeyAssigned) -> Assigned:
        # src/paillier.rs:39-45 / :69-75
        n2 = self.biguint.square(ctx, pk.n)
        aux = RefreshAux(self.biguint.limb_bits, pk.n.num_limbs(), pk.n.num_limbs())
        return self.biguint.refresh(ctx, n2, aux)

    def encrypt(self, ctx: Context, pk: EncryptionPublicKeyAssigned, m: Assigned, r: Assigned) -> Assigned:
        """src/paillier.rs:32-60."""
        n2 = self._n2(ctx, pk)
        ctx.load(0)  # ctx.load_zero()  :47
        g_ext = pk.g.extend_limbs(n2.num_limbs() - pk.g.num_limbs())  # :49
        m_big = self.get_biguint(m)  # :50
        gm = self.biguint.pow_mod_fixed_exp(ctx, g_ext, m_big, n2)  # :51
        r_ext = r.extend_limbs(n2.num_limbs() - r.num_limbs())  # :53
        n_big = self.get_biguint(pk.n)  # :54
        rn = self.biguint.pow_mod_fixed_exp(ctx, r_ext, n_big, n2)  # :55
        return self.biguint.mul_mod(ctx, gm, rn, n2, kind="final")  # :57

    def add(self, ctx: Context, pk: EncryptionPublicKeyAssigned, c1: Assigned, c2: Assigned) -> Assigned:
        """src/paillier.rs:62-85."""
        n2 = self._n2(ctx, pk)
        ctx.load(0)  # :77
        c1e = c1.extend_limbs(n2.num_limbs() - c1.num_limbs())  # :79
        c2e = c2.extend_limbs(n2.num_limbs() - c2.num_limbs())  # :80
        return self.biguint.mul_mod(ctx, c1e, c2e, n2, kind="add")  # :81


def check_constraints(ctx: Context) -> None:
    """MockProver stand-in (`expect_satisfied(true)`, src/paillier.rs:167-170): every recorded
    constraint must hold and every cell must be a canonical field element."""
    for what, ok in ctx.checks:
        if not ok:
            raise AssertionError(f"constraint not satisfied: {what}")
    for c in ctx.cells:
        if not (0 <= c < BN254_FR):
            raise AssertionError("cell value out of field")


# --------------------------------------------------------------------------------------
# drivers mirroring src/bench.rs:33-75 and :77-117
# --------------------------------------------------------------------------------------


def paillier_enc_test(enc_bits: int, limb_bits: int, n: int, g: int, m: int, r: int, res: int,
                      lookup_bits: Optional[int] = None, witness_source=None) -> Context:
    """src/bench.rs:33-75 — returns the filled Context (cells, steps) after all assertions.
    witness_source: (q, rem) pairs in mul_mod call order (see BigUintChip)."""
    ctx = Context()
    big = BigUintChip(limb_bits, lookup_bits, witness_source)
    chip = PaillierChip.construct(big, enc_bits)
    n_as = big.assign_integer(ctx, n, enc_bits)
    g_as = big.assign_integer(ctx, g, enc_bits)
    pk = EncryptionPublicKeyAssigned(n_as, g_as)
    m_as = big.assign_integer(ctx, m, enc_bits)
    r_as = big.assign_integer(ctx, r, enc_bits)
    c_as = chip.encrypt(ctx, pk, m_as, r_as)
    res_as = big.assign_integer(ctx, res, enc_bits * 2)
    assert c_as.value == res_as.value, "assert_eq!(c.value(), res.value()) failed"
    big.assert_equal_fresh(ctx, c_as, res_as)
    check_constraints(ctx)
    return ctx


def paillier_enc_add_test(enc_bits: int, limb_bits: int, n: int, g: int, c1: int, c2: int, res: int,
                          lookup_bits: Optional[int] = None, c_bits: Optional[int] = None, witness_source=None) -> Context:
    """src/bench.rs:77-117 — c1, c2 are assigned with enc_bits in the reference (half-width,
    src/paillier.rs:216-221); pass c_bits=2*enc_bits for real ciphertexts."""
    ctx = Context()
    big = BigUintChip(limb_bits, lookup_bits, witness_source)
    chip = PaillierChip.construct(big, enc_bits)
    n_as = big.assign_integer(ctx, n, enc_bits)
    g_as = big.assign_integer(ctx, g, enc_bits)
    pk = EncryptionPublicKeyAssigned(n_as, g_as)
    cb = c_bits or enc_bits
    c1_as = big.assign_integer(ctx, c1, cb)
    c2_as = big.assign_integer(ctx, c2, cb)
    out = chip.add(ctx, pk, c1_as, c2_as)
    res_as = big.assign_integer(ctx, res, enc_bits * 2)
    assert out.value == res_as.value
    big.assert_equal_fresh(ctx, out, res_as)
    check_constraints(ctx)
    return ctx


# --------------------------------------------------------------------------------------
# compact step record (what the GPU emits: q, rem per mul_mod, chain order of A.5)
# --------------------------------------------------------------------------------------


def pow_chain_steps(a: int, e: int, n2: int) -> Tuple[int, List[MulModStep]]:
    """The (q, rem) sequence of `pow_mod_fixed_exp(a, e, n2)` without cell bookkeeping."""
    steps: List[MulModStep] = []
    acc = 1
    sq = a
    for i in range(e.bit_length()):
        cur = sq
        q, sq = divmod(cur * cur, n2)
        steps.append(MulModStep("sqr", cur, cur, q, sq))
        if (e >> i) & 1:
            q, rem = divmod(acc * cur, n2)
            steps.append(MulModStep("mul", acc, cur, q, rem))
            acc = rem
    return acc, steps


def encrypt_steps(n: int, g: int, m: int, r: int) -> Tuple[int, List[MulModStep]]:
    """All mul_mod groups of `PaillierChip::encrypt` in assignment order (src/paillier.rs:51,55,57)."""
    n2 = n * n
    gm, s1 = pow_chain_steps(g, m, n2)
    rn, s2 = pow_chain_steps(r, n, n2)
    q, c = divmod(gm * rn, n2)
    return c, s1 + s2 + [MulModStep("final", gm, rn, q, c)]


def paillier_dec_native(n: int, lam: int, mu: int, c: int) -> int:
    """Decryption as the reference's README states it (README.md:17-22 there; its code has no decryption):
    m = L(c^lambda mod n^2) * mu mod n with L(x) = (x - 1) / n.  Raises if c is not a valid ciphertext (x != 1 mod n)."""
    x = pow(c, lam, n * n)
    if (x - 1) % n:
        raise ValueError("c^lambda mod n^2 is not 1 modulo n")
    return ((x - 1) // n) * mu % n
