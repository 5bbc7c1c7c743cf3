"""CPU tier: the gate-DAG builder and level scheduler of learn-fhe_b200/circuits.py with each bootstrapped gate replaced by
its plain boolean function (no GPU, no ciphertexts): checks the circuits of scheme/fhew/src/fhew/boolean.rs:134-164 and
uint8.rs:53-157 against integer arithmetic, and that evaluation batches by (level, truth table)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def plain_engine(pkg):
    from learn_fhe_b200 import circuits
    from learn_fhe_b200.fhew import GATES

    PLAIN = {"and": lambda a, b: a & b, "nand": lambda a, b: 1 - (a & b), "or": lambda a, b: a | b, "nor": lambda a, b: 1 - (a | b),
             "xor": lambda a, b: a ^ b, "xnor": lambda a, b: 1 - (a ^ b), "majority": lambda a, b, c: ((a + b + c) >= 2).astype(np.int64)}

    class PlainEngine(circuits.GateEngine):
        """Values are numpy 0/1 vectors; a 'bootstrap batch' applies the gate's boolean function (fhew.rs:58-67 table names)."""

        def __init__(self):  # no key, no device
            self.nodes, self.values, self.launches, self.gates = [], {}, 0, 0
            self.batches = []

        def input(self, bits):
            self.nodes.append(("input", None, (), 0))
            self.values[len(self.nodes) - 1] = np.asarray(bits, dtype=np.int64)
            return circuits.FhewBool(self, len(self.nodes) - 1)

        def evaluate(self, targets):
            need, stack = set(), [t for t in targets if t not in self.values]
            while stack:
                i = stack.pop()
                if i in need or i in self.values:
                    continue
                need.add(i)
                stack.extend(a for a in self.nodes[i][2] if a not in self.values)
            for level in sorted({self.nodes[i][3] for i in need}):
                todo = sorted(i for i in need if self.nodes[i][3] == level)
                groups = {}
                for i in todo:
                    if self.nodes[i][0] != "not":
                        groups.setdefault(tuple(GATES[self.nodes[i][0]][0]), []).append(i)
                for table, ids in groups.items():
                    self.launches += 1
                    self.batches.append((level, table, len(ids)))
                    for i in ids:
                        kind, _, args, _ = self.nodes[i]
                        xs = [self.values[a] for a in args]
                        self.values[i] = PLAIN[kind](*xs)
                        self.gates += len(xs[0])
                for i in todo:
                    if self.nodes[i][0] == "not":
                        self.values[i] = 1 - self.values[self.nodes[i][2][0]]

        def bits(self, nodes):
            self.evaluate([b.node for b in nodes])
            return np.stack([self.values[b.node] for b in nodes])

    return PlainEngine, circuits


def u8_in(eng, circuits, vals):
    vals = np.asarray(vals, dtype=np.int64)
    return circuits.FhewU8([eng.input((vals >> i) & 1) for i in range(8)])


def u8_out(eng, x):
    b = eng.bits(x.bits)
    return sum(b[i] << i for i in range(8))


def test_bool_gates_and_adders(plain_engine):
    PlainEngine, circuits = plain_engine
    eng = PlainEngine()
    m = np.array([[a, b, c] for a in (0, 1) for b in (0, 1) for c in (0, 1)])
    a, b, c = (eng.input(m[:, i]) for i in range(3))
    got = eng.bits([a & b, a | b, a ^ b, a.bitnand(b), a.bitnor(b), a.bitxnor(b), a.bitmajority(b, c), ~a, a.select(b, c)])
    exp = [m[:, 0] & m[:, 1], m[:, 0] | m[:, 1], m[:, 0] ^ m[:, 1], 1 - (m[:, 0] & m[:, 1]), 1 - (m[:, 0] | m[:, 1]), 1 - (m[:, 0] ^ m[:, 1]),
           (m.sum(axis=1) >= 2).astype(int), 1 - m[:, 0], np.where(m[:, 0] == 1, m[:, 2], m[:, 1])]
    for g, e in zip(got, exp):
        assert (g == e).all()
    s, cy = a.carrying_add(b, c)
    d, bw = a.borrowing_sub(b, c)
    assert (eng.bits([s])[0] == m.sum(axis=1) % 2).all() and (eng.bits([cy])[0] == (m.sum(axis=1) >= 2)).all()
    diff = m[:, 0] - m[:, 1] - m[:, 2]
    assert (eng.bits([d])[0] == diff % 2).all() and (eng.bits([bw])[0] == (diff < 0)).all()


def test_u8_arithmetic_against_integers(plain_engine):
    PlainEngine, circuits = plain_engine
    rng = np.random.default_rng(5)
    m0 = np.concatenate([rng.integers(0, 256, 60), [0, 255, 1, 128, 200]])
    m1 = np.concatenate([rng.integers(1, 256, 60), [255, 255, 1, 3, 200]])
    eng = PlainEngine()
    x, y = u8_in(eng, circuits, m0), u8_in(eng, circuits, m1)
    assert (u8_out(eng, x + y) == (m0 + m1) % 256).all()
    assert (u8_out(eng, x - y) == (m0 - m1) % 256).all()
    assert (u8_out(eng, x * y) == (m0 * m1) % 256).all()
    assert (u8_out(eng, x.wrapping_neg()) == (-m0) % 256).all()
    assert (u8_out(eng, ~x) == 255 - m0).all()
    q, r = x.div_rem(y)
    assert (u8_out(eng, q) == m0 // m1).all() and (u8_out(eng, r) == m0 % m1).all()
    s, c = x.carrying_add(y, eng.input(np.ones_like(m0)))
    assert (u8_out(eng, s) == (m0 + m1 + 1) % 256).all() and (eng.bits([c])[0] == (m0 + m1 + 1 > 255)).all()


def test_level_batching(plain_engine):
    """The multiplier's 36 partial-product ANDs are independent: they must arrive as one level-1 batch."""
    PlainEngine, circuits = plain_engine
    eng = PlainEngine()
    x, y = u8_in(eng, circuits, [7]), u8_in(eng, circuits, [9])
    p = x * y
    assert (u8_out(eng, p) == 63).all()
    lvl1 = [b for b in eng.batches if b[0] == 1]
    assert lvl1 == [(1, (0, 0, 0, 1), 36)]
    n_gates = sum(b[2] for b in eng.batches)
    # the reference evaluates 36 + 7*2 + 21*5 = 155 gates; the lazy DAG never evaluates the 19 whose carries are discarded
    assert n_gates == 136 and eng.launches < n_gates / 2
