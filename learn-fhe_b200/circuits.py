"""Gate-circuit callers of the batched FHEW bootstrap: mirrors of `FhewBool` (scheme/fhew/src/fhew/boolean.rs:10-176) and
`FhewU8` (scheme/fhew/src/fhew/uint8.rs:13-163) — SURVEY.md §8(f) rank 1.

The reference evaluates one gate (= one bootstrap) at a time.  Here every `FhewBool` is a *vector* of encrypted bits and a
node of a lazily built gate DAG: nothing runs until a value is requested; then the DAG is walked level by level and all
gates of one level that share a truth table go to the GPU as ONE `fhe_fhew_bootstrap_batch` call (the 8x8 `wrapping_mul`
has 36 independent ANDs in its first level; a vector of B bytes multiplies that by B).  Every gate is a deterministic
function of its input ciphertexts, so the ciphertexts are bit-identical to the reference's sequential evaluation.

Linear pre-/post-processing (ct0 + ct1, (ct0 - ct1).double(), NOT; fhew.rs:27-29,58-67) runs on the device through the
library's own element-wise kernels.  Nothing here computes on the CPU; encryption / decryption stay with the caller.
"""
import numpy as np

from . import dptr, to_dev, to_host
from .fhew import GATES, big_q_by_8, gate_poly


class GateEngine:
    """Owns the DAG, the device-resident ciphertext batches ([B, N+1] int64 tensors) and the level-batched evaluation."""

    def __init__(self, bk):
        import torch
        self.bk, self.ctx, self.param = bk, bk.ctx, bk.param
        self.torch = torch
        self.dev = "cuda:%d" % self.ctx.device
        self.ctx.use_torch_stream()  # torch glue (cat / slicing) and the library's kernels must share one stream
        self.nodes = []  # (kind, payload, args, level); kind in {"input", "not", gate name}
        self.values = {}
        self.post = big_q_by_8(self.param)
        self.q4 = int(round(self.param.big_q / 4.0)) % self.param.big_q
        self._f = {}
        self._q4row = None
        self.launches = 0  # bootstrap batches issued (for tests / reporting)
        self.gates = 0     # gate bootstraps evaluated (x batch width)

    # ---- DAG construction ----------------------------------------------------------------------------------------------
    def input(self, cts):
        """cts: [B, N+1] uint64 numpy array (or int64 CUDA tensor) of LWE ciphertexts mod Q."""
        t = cts if self.torch.is_tensor(cts) else to_dev(np.ascontiguousarray(cts, dtype=np.uint64), self.ctx.device)
        assert t.dim() == 2 and t.shape[1] == self.param.n + 1
        self.nodes.append(("input", None, (), 0))
        self.values[len(self.nodes) - 1] = t
        return FhewBool(self, len(self.nodes) - 1)

    def _add(self, kind, args):
        level = max(self.nodes[a][3] for a in args) + (0 if kind == "not" else 1)
        self.nodes.append((kind, None, tuple(args), level))
        return FhewBool(self, len(self.nodes) - 1)

    # ---- device-side linear algebra mod Q (library kernels) ---------------------------------------------------------------
    def _ew(self, name, a, b):
        out = self.torch.empty_like(a)
        self.ctx.call(name, self.param.big_q, a.numel(), dptr(a), dptr(b), dptr(out))
        return out

    def _linear(self, lin, xs):
        if lin == "add":
            return self._ew("fhe_vec_add_u64", xs[0], xs[1])
        if lin == "add3":
            return self._ew("fhe_vec_add_u64", self._ew("fhe_vec_add_u64", xs[0], xs[1]), xs[2])
        if lin == "sub2":  # (ct0 - ct1).double()
            d = self._ew("fhe_vec_sub_u64", xs[0], xs[1])
            return self._ew("fhe_vec_add_u64", d, d)
        raise ValueError(lin)

    def _not(self, x):
        """Fhew::not (fhew.rs:27-29): (-a, -b + Q/4)."""
        out = self.torch.empty_like(x)
        self.ctx.call("fhe_vec_neg_u64", self.param.big_q, x.numel(), dptr(x), dptr(out))
        if self._q4row is None or self._q4row.shape[0] < x.shape[0]:
            row = np.zeros((x.shape[0], self.param.n + 1), dtype=np.uint64)
            row[:, -1] = self.q4
            self._q4row = to_dev(row, self.ctx.device)
        return self._ew("fhe_vec_add_u64", out, self._q4row[:x.shape[0]].contiguous())

    def _table(self, table):
        key = tuple(table)
        if key not in self._f:
            self._f[key] = to_dev(gate_poly(self.param, table), self.ctx.device)
        return self._f[key]

    # ---- evaluation -------------------------------------------------------------------------------------------------------
    def evaluate(self, targets):
        """Materialise the nodes in `targets` (and everything they depend on)."""
        need, stack = set(), [t for t in targets if t not in self.values]
        while stack:
            i = stack.pop()
            if i in need or i in self.values:
                continue
            need.add(i)
            stack.extend(a for a in self.nodes[i][2] if a not in self.values)
        for level in sorted({self.nodes[i][3] for i in need}):
            todo = sorted(i for i in need if self.nodes[i][3] == level)
            # gates of this level, grouped by truth table: one bootstrap batch per table
            groups = {}
            for i in todo:
                kind = self.nodes[i][0]
                if kind != "not":
                    groups.setdefault(tuple(GATES[kind][0]), []).append(i)
            for table, ids in groups.items():
                lins = [self._linear(GATES[self.nodes[i][0]][1], [self.values[a] for a in self.nodes[i][2]]) for i in ids]
                batch = lins[0] if len(lins) == 1 else self.torch.cat(lins, dim=0)
                out = self.torch.empty_like(batch)
                self.ctx.call("fhe_fhew_bootstrap_batch", self.bk.h, dptr(self._table(table)), self.post, batch.shape[0], dptr(batch), dptr(out))
                self.launches += 1
                self.gates += batch.shape[0]
                off = 0
                for i, l in zip(ids, lins):
                    self.values[i] = out[off:off + l.shape[0]]
                    off += l.shape[0]
            # NOT is linear: same level as its argument, evaluated once that is available (chains resolve in index order)
            for i in todo:
                if self.nodes[i][0] == "not":
                    self.values[i] = self._not(self.values[self.nodes[i][2][0]].contiguous())

    def ciphertexts(self, bits):
        """Evaluate and download: list of FhewBool -> numpy uint64 [len(bits), B, N+1]."""
        self.evaluate([b.node for b in bits])
        self.ctx.sync()
        return np.stack([to_host(self.values[b.node].contiguous()) for b in bits])


class FhewBool:
    """A vector of encrypted bits (boolean.rs:10-14) as a node of the engine's gate DAG."""
    __slots__ = ("eng", "node")

    def __init__(self, eng, node):
        self.eng, self.node = eng, node

    def _gate(self, name, *others):
        return self.eng._add(name, [self.node] + [o.node for o in others])

    # boolean.rs:45-75 (impl_op!): bitnot / bitand / bitnand / bitor / bitnor / bitxor / bitxnor / bitmajority
    def bitnot(self):
        return self.eng._add("not", [self.node])

    def bitand(self, o):
        return self._gate("and", o)

    def bitnand(self, o):
        return self._gate("nand", o)

    def bitor(self, o):
        return self._gate("or", o)

    def bitnor(self, o):
        return self._gate("nor", o)

    def bitxor(self, o):
        return self._gate("xor", o)

    def bitxnor(self, o):
        return self._gate("xnor", o)

    def bitmajority(self, a, b):
        return self._gate("majority", a, b)

    __invert__ = bitnot
    __and__ = bitand
    __or__ = bitor
    __xor__ = bitxor

    # boolean.rs:134-164
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

    def ciphertexts(self):
        return self.eng.ciphertexts([self])[0]


class FhewU8:
    """Eight little-endian FhewBool vectors (uint8.rs:13-14)."""

    def __init__(self, bits):
        assert len(bits) == 8
        self.bits = list(bits)

    @classmethod
    def from_ciphertexts(cls, eng, cts):
        """cts [8, B, N+1]: little-endian bit ciphertexts (uint8.rs:17-20)."""
        return cls([eng.input(cts[i]) for i in range(8)])

    def ciphertexts(self):
        return self.bits[0].eng.ciphertexts(self.bits)

    def __invert__(self):  # uint8.rs:35-50
        return FhewU8([~b for b in self.bits])

    def wrapping_neg(self):  # uint8.rs:53-65
        v = self.bits
        carry = ~v[0]
        out = []
        for i in range(8):
            if i == 0:
                out.append(v[0])
            else:
                s, carry = (~v[i]).overflowing_add(carry)
                out.append(s)
        return FhewU8(out)

    def overflowing_add(self, rhs):  # uint8.rs:67-79
        carry, out = None, []
        for i in range(8):
            if carry is None:
                s, carry = self.bits[i].overflowing_add(rhs.bits[i])
            else:
                s, carry = self.bits[i].carrying_add(rhs.bits[i], carry)
            out.append(s)
        return FhewU8(out), carry

    def carrying_add(self, rhs, carry):  # uint8.rs:81-89
        out = []
        for i in range(8):
            s, carry = self.bits[i].carrying_add(rhs.bits[i], carry)
            out.append(s)
        return FhewU8(out), carry

    def wrapping_add(self, rhs):
        return self.overflowing_add(rhs)[0]

    def overflowing_sub(self, rhs):  # uint8.rs:95-107
        borrow, out = None, []
        for i in range(8):
            if borrow is None:
                s, borrow = self.bits[i].overflowing_sub(rhs.bits[i])
            else:
                s, borrow = self.bits[i].borrowing_sub(rhs.bits[i], borrow)
            out.append(s)
        return FhewU8(out), borrow

    def borrowing_sub(self, rhs, borrow):  # uint8.rs:109-117
        out = []
        for i in range(8):
            s, borrow = self.bits[i].borrowing_sub(rhs.bits[i], borrow)
            out.append(s)
        return FhewU8(out), borrow

    def wrapping_sub(self, rhs):
        return self.overflowing_sub(rhs)[0]

    def wrapping_mul(self, rhs):  # uint8.rs:123-135
        lhs, rb = self.bits, rhs.bits
        carries = [None] * 7
        out = []
        for i in range(8):
            t = [lhs[j] & rb[i - j] for j in range(i + 1)]
            s = t[0]
            for k, tj in enumerate(t[1:]):
                if carries[k] is not None:
                    s, carries[k] = s.carrying_add(tj, carries[k])
                else:
                    s, carries[k] = s.overflowing_add(tj)
            out.append(s)
        return FhewU8(out)

    def div_rem(self, rhs):  # uint8.rs:137-157 (restoring division on VecDeques)
        lhs, neg = self.bits, rhs.wrapping_neg().bits
        q, r = [], []
        for i in range(8):
            r.insert(0, lhs[7 - i])
            d = list(r)
            d[0], carry = d[0].overflowing_add(neg[0])
            for j in range(1, 8):
                if j < len(d):
                    d[j], carry = d[j].carrying_add(neg[j], carry)
                else:
                    carry = carry & neg[j]
            r = [carry.select(rj, dj) for rj, dj in zip(r, d)]
            q.insert(0, carry)
        return FhewU8(q), FhewU8(r)

    def wrapping_div(self, rhs):
        return self.div_rem(rhs)[0]

    def wrapping_rem(self, rhs):
        return self.div_rem(rhs)[1]

    __add__ = wrapping_add
    __sub__ = wrapping_sub
    __mul__ = wrapping_mul
    __floordiv__ = wrapping_div
    __mod__ = wrapping_rem
