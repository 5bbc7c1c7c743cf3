"""Gate circuits on the batched bootstrap: `FhewBool` / `FhewU8` mirrors against an eager,
gate-at-a-time evaluation with the oracle in the reference's order (scheme/fhew/src/fhew/boolean.rs:241-386 truth tables,
uint8.rs:292-339 arithmetic).  Ciphertexts must be bit-identical (a gate is a deterministic function of its inputs) and
decrypt to the plain result."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TABLES = {"and": ([0, 0, 0, 1], "add"), "nand": ([1, 1, 1, 0], "add"), "or": ([0, 1, 1, 1], "add"), "nor": ([1, 0, 0, 0], "add"),
          "xor": ([0, 1, 1, 1], "sub2"), "xnor": ([1, 0, 0, 0], "sub2"), "majority": ([0, 0, 0, 1], "add3")}


class RefBool:
    """Eager reference: one oracle bootstrap per gate, in program order (fhew.rs:27-67)."""

    def __init__(self, K, ct):
        self.K, self.ct = K, ct

    def _gate(self, name, *others):
        q = np.uint64(self.K.param.big_q)
        table, lin = TABLES[name]
        cts = [self.ct] + [o.ct for o in others]
        if lin == "add":
            x = (cts[0] + cts[1]) % q
        elif lin == "add3":
            x = (cts[0] + cts[1] + cts[2]) % q
        else:
            d = (cts[0] + (q - cts[1])) % q
            x = (d + d) % q
        return RefBool(self.K, self.K.op(table, x, threads=4))

    def __invert__(self):
        q = np.uint64(self.K.param.big_q)
        out = (q - self.ct) % q
        out[:, -1] = (out[:, -1] + np.uint64(int(round(self.K.param.big_q / 4.0)))) % q
        return RefBool(self.K, out)

    def __and__(self, o):
        return self._gate("and", o)

    def __or__(self, o):
        return self._gate("or", o)

    def __xor__(self, o):
        return self._gate("xor", o)

    def select(self, f, t):
        return (~self & f) | (self & t)

    def overflowing_add(self, rhs):
        return self ^ rhs, self & rhs

    def carrying_add(self, rhs, carry):
        t = self ^ rhs
        return t ^ carry, (self & rhs) | (t & carry)

    def overflowing_sub(self, rhs):
        return self ^ rhs, ~self & rhs

    def borrowing_sub(self, rhs, borrow):
        t = self ^ rhs
        return t ^ borrow, (~self & rhs) | (~t & borrow)


@pytest.fixture(scope="module")
def setup(pkg, ctx, fhew_setup):
    from learn_fhe_b200 import circuits, fhew
    P, K, ex = fhew_setup
    K.param = P
    param = fhew.single_key_testing_param(P.big_q)
    bk = fhew.BootstrappingKey(ctx, param, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])
    yield K, bk, circuits
    bk.free()


def enc_u8(K, vals, seed):
    """little-endian bit ciphertexts [8, B, N+1]"""
    vals = np.asarray(vals, dtype=np.uint8)
    return np.stack([K.encrypt(((vals >> i) & 1).astype(np.int32), seed + i) for i in range(8)])


def dec_u8(K, cts):
    bits = np.stack([K.decrypt(cts[i]) for i in range(8)])
    return sum((bits[i].astype(np.int64) << i) for i in range(8)).astype(np.uint8)


def test_gate_truth_tables(setup):
    """boolean.rs:254-318: not/and/nand/or/nor/xor/xnor/majority over all inputs, one level-batched launch per table."""
    K, bk, circuits = setup
    eng = circuits.GateEngine(bk)
    m = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)], dtype=np.int32)
    cts = [K.encrypt(m[:, i], 40 + i) for i in range(3)]
    a, b, c = (eng.input(x) for x in cts)
    ra, rb, rc = (RefBool(K, x) for x in cts)
    outs = {"not": ~a, "and": a.bitand(b), "nand": a.bitnand(b), "or": a.bitor(b), "nor": a.bitnor(b), "xor": a.bitxor(b),
            "xnor": a.bitxnor(b), "majority": a.bitmajority(b, c)}
    refs = {"not": ~ra, "and": ra._gate("and", rb), "nand": ra._gate("nand", rb), "or": ra._gate("or", rb), "nor": ra._gate("nor", rb),
            "xor": ra._gate("xor", rb), "xnor": ra._gate("xnor", rb), "majority": ra._gate("majority", rb, rc)}
    plain = {"not": 1 - m[:, 0], "and": m[:, 0] & m[:, 1], "nand": 1 - (m[:, 0] & m[:, 1]), "or": m[:, 0] | m[:, 1],
             "nor": 1 - (m[:, 0] | m[:, 1]), "xor": m[:, 0] ^ m[:, 1], "xnor": 1 - (m[:, 0] ^ m[:, 1]),
             "majority": ((m.sum(axis=1)) >= 2).astype(np.int32)}
    got = eng.ciphertexts(list(outs.values()))
    assert eng.launches == 4  # seven gates of one level, four distinct truth tables
    for (name, _), g in zip(outs.items(), got):
        assert (g == refs[name].ct).all(), name
        assert (K.decrypt(g) == plain[name]).all(), name


def test_u8_add_sub_neg_not(setup):
    K, bk, circuits = setup
    eng = circuits.GateEngine(bk)
    m0, m1 = np.array([200, 17, 255], dtype=np.uint8), np.array([100, 250, 1], dtype=np.uint8)
    c0, c1 = enc_u8(K, m0, 100), enc_u8(K, m1, 200)
    x, y = circuits.FhewU8.from_ciphertexts(eng, c0), circuits.FhewU8.from_ciphertexts(eng, c1)
    s, carry = x.overflowing_add(y)
    d, borrow = x.overflowing_sub(y)
    neg, inv = x.wrapping_neg(), ~x
    assert (dec_u8(K, s.ciphertexts()) == (m0 + m1)).all()
    assert (K.decrypt(carry.ciphertexts()) == ((m0.astype(int) + m1) > 255)).all()
    assert (dec_u8(K, d.ciphertexts()) == (m0 - m1)).all()
    assert (K.decrypt(borrow.ciphertexts()) == (m0 < m1)).all()
    assert (dec_u8(K, neg.ciphertexts()) == (np.uint8(0) - m0)).all()
    assert (dec_u8(K, inv.ciphertexts()) == ~m0).all()
    # bit-identical to the eager reference for the ripple-carry adder
    rx, ry = [RefBool(K, c0[i]) for i in range(8)], [RefBool(K, c1[i]) for i in range(8)]
    rc, rs = None, []
    for i in range(8):
        t, rc = rx[i].overflowing_add(ry[i]) if rc is None else rx[i].carrying_add(ry[i], rc)
        rs.append(t.ct)
    assert (s.ciphertexts() == np.stack(rs)).all()
    assert (carry.ciphertexts() == rc.ct).all()


def test_u8_mul_div_rem(setup):
    """uint8.rs:123-157: array multiplier (36 ANDs in the first level) and restoring division, on a vector of bytes."""
    K, bk, circuits = setup
    eng = circuits.GateEngine(bk)
    m0, m1 = np.array([13, 250], dtype=np.uint8), np.array([11, 7], dtype=np.uint8)
    c0, c1 = enc_u8(K, m0, 300), enc_u8(K, m1, 400)
    x, y = circuits.FhewU8.from_ciphertexts(eng, c0), circuits.FhewU8.from_ciphertexts(eng, c1)
    prod = x * y
    q, r = x.div_rem(y)
    before = eng.launches
    assert (dec_u8(K, prod.ciphertexts()) == (m0 * m1)).all()
    levels_mul = eng.launches - before
    assert (dec_u8(K, q.ciphertexts()) == m0 // m1).all()
    assert (dec_u8(K, r.ciphertexts()) == m0 % m1).all()
    # the DAG batches: far fewer launches than gates (155 gates in the multiplier)
    assert levels_mul < 80
    # eager reference of the multiplier, gate by gate, must give the same ciphertexts
    lhs, rhs = [RefBool(K, c0[i]) for i in range(8)], [RefBool(K, c1[i]) for i in range(8)]
    carries, out = [None] * 7, []
    for i in range(8):
        t = [lhs[j] & rhs[i - j] for j in range(i + 1)]
        s = t[0]
        for k, tj in enumerate(t[1:]):
            if carries[k] is not None:
                s, carries[k] = s.carrying_add(tj, carries[k])
            else:
                s, carries[k] = s.overflowing_add(tj)
        out.append(s.ct)
    assert (prod.ciphertexts() == np.stack(out)).all()
