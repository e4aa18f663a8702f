"""TEST INFRASTRUCTURE. A second, fully independent restatement of the reference protocol in pure Python: big-int field
arithmetic and the official `blake3` package only -- no ctypes, no C++, nothing shared with multi_stark_b200/host/ or
oracle/ (which share one challenger, one graph compiler and one proof parser between them). It restates, from the
reference sources directly:

  PyChallenger          DeterministicPow<SerializingChallenger64<Goldilocks, HashChallenger<u8, Blake3, 32>>>
                        (src/types.rs:28-81; p3-challenger 0.5.1 published semantics), seeded as
                        GoldilocksBlake3Config::new does (src/types.rs:118-130)
  Expr / compile        the frontend operators with constant folding (src/expr.rs:151-214) and `compile()` with its hash-consing
                        interner and commutative normalisation (src/graph.rs:120-188, 213-330)
  u32_add / byte_table  the benchmark circuits (benches/multi_stark.rs:73-165)
  verify                System::verify_multiple_claims (src/verifier.rs:208-532): transcript replay, opening rounds, the OOD
                        check with sweep (src/eval.rs:67-106), logup_constraint_values (src/lookup.rs:152-208), selectors at a
                        point, reversed alpha fold, quotient recombination
  pcs_verify            TwoAdicFriPcs::verify / verify_fri / verify_query / fold_row and MerkleTreeMmcs::verify_batch of
                        p3-fri / p3-merkle-tree 0.5.1 (not vendored in the reference; published semantics)

A proof made by the device prover (or the C++ oracle) that this verifier accepts has the transcript order, sample-pop direction,
node numbering, alpha ordering, FRI folding rule and query-index derivation that THIS file derives from the reference text; a
wrong-but-self-consistent host layer cannot pass it. The proof bytes are parsed with tests/_proof.py (also pure Python)."""
import blake3 as _b3

P = 2**64 - 2**32 + 1
GENERATOR = 7
W = 7  # X^2 = 7
TWO_ADIC_ROOT = 1753635133440165772  # = 7^((p-1)/2^32): p3 Goldilocks two_adic_generator(32)


def two_adic_generator(bits):
    assert 0 <= bits <= 32
    return pow(TWO_ADIC_ROOT, 1 << (32 - bits), P)


def inv(a):
    return pow(a % P, P - 2, P)


def rev_bits(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


class E:
    """BinomialExtensionField<Goldilocks, 2>, basis [1, X], X^2 = 7."""
    __slots__ = ("a", "b")

    def __init__(self, a=0, b=0):
        self.a, self.b = a % P, b % P

    @staticmethod
    def of(x):
        return x if isinstance(x, E) else E(x, 0)

    def __add__(self, o):
        o = E.of(o)
        return E(self.a + o.a, self.b + o.b)

    __radd__ = __add__

    def __sub__(self, o):
        o = E.of(o)
        return E(self.a - o.a, self.b - o.b)

    def __rsub__(self, o):
        return E.of(o) - self

    def __neg__(self):
        return E(-self.a, -self.b)

    def __mul__(self, o):
        o = E.of(o)
        return E(self.a * o.a + W * self.b * o.b, self.a * o.b + self.b * o.a)

    __rmul__ = __mul__

    def __eq__(self, o):
        o = E.of(o)
        return self.a == o.a and self.b == o.b

    def __hash__(self):
        return hash((self.a, self.b))

    def is_zero(self):
        return self.a == 0 and self.b == 0

    def inverse(self):
        n = inv(self.a * self.a - W * self.b * self.b)
        return E(self.a * n, -self.b * n)

    def pow(self, e):
        r, b = E(1), self
        while e:
            if e & 1:
                r = r * b
            b = b * b
            e >>= 1
        return r

    def __repr__(self):
        return "E(%d,%d)" % (self.a, self.b)


# ------------------------------------------------------------------------------------------------------------------
# transcript
# ------------------------------------------------------------------------------------------------------------------
class PyChallenger:
    def __init__(self, seed):
        self.inp = bytearray(seed)
        self.out = bytearray()

    @staticmethod
    def for_config(log_blowup, cap_height, log_final_poly_len, max_log_arity, num_queries, commit_pow_bits, query_pow_bits):
        seed = bytearray(b"multi-stark/v0")
        for p in (log_blowup, cap_height, log_final_poly_len, max_log_arity, num_queries, commit_pow_bits, query_pow_bits):
            seed += int(p).to_bytes(8, "little")
        return PyChallenger(seed)

    def clone(self):
        c = PyChallenger(self.inp)
        c.out = bytearray(self.out)
        return c

    # HashChallenger<u8>
    def observe_bytes(self, bs):
        for b in bs:          # observe(byte): output buffer cleared, byte appended
            self.out.clear()
            self.inp.append(b)

    def _flush(self):
        d = _b3.blake3(bytes(self.inp)).digest()
        self.inp = bytearray(d)      # chaining
        self.out = bytearray(d)

    def sample_byte(self):
        if not self.out:
            self._flush()
        return self.out.pop()        # from the END

    # SerializingChallenger64
    def observe(self, v):
        if isinstance(v, E):
            self.observe(v.a)
            self.observe(v.b)
        elif isinstance(v, (bytes, bytearray)):
            self.observe_bytes(v)
        else:
            assert 0 <= v < P
            self.observe_bytes(int(v).to_bytes(8, "little"))

    def sample_u64(self):
        return int.from_bytes(bytes(self.sample_byte() for _ in range(8)), "little")

    def sample_base(self):
        while True:
            v = self.sample_u64()
            if v < P:
                return v

    def sample_ext(self):
        a = self.sample_base()
        b = self.sample_base()
        return E(a, b)

    def sample_bits(self, bits):
        return self.sample_u64() & ((1 << bits) - 1)

    # GrindingChallenger under DeterministicPow: zero bits observe nothing
    def check_witness(self, bits, w):
        if bits == 0:
            return True
        self.observe(w)
        return self.sample_bits(bits) == 0


# ------------------------------------------------------------------------------------------------------------------
# frontend expressions (src/expr.rs) and the compiler (src/graph.rs)
# ------------------------------------------------------------------------------------------------------------------
PRE, MAIN, STAGE2 = 0, 1, 2


class Expr:
    __slots__ = ("k", "x")  # kind, payload

    def __init__(self, k, x=None):
        self.k, self.x = k, x

    @staticmethod
    def const(v):
        return Expr("const", v % P)

    @staticmethod
    def var(source, offset, index):
        return Expr("var", (source, offset, index))

    @staticmethod
    def main(i):
        return Expr.var(MAIN, 0, i)

    @staticmethod
    def main_next(i):
        return Expr.var(MAIN, 1, i)

    @staticmethod
    def preprocessed(i):
        return Expr.var(PRE, 0, i)

    def _c(self):
        return self.x if self.k == "const" else None

    def __add__(self, o):  # src/expr.rs:151-161
        a, b = self._c(), o._c()
        if a is not None and b is not None:
            return Expr.const(a + b)
        if a == 0:
            return o
        if b == 0:
            return self
        return Expr("add", (self, o))

    def __sub__(self, o):  # :163-174
        a, b = self._c(), o._c()
        if a is not None and b is not None:
            return Expr.const(a - b)
        if b == 0:
            return self
        if a == 0:
            return -o
        return Expr("sub", (self, o))

    def __mul__(self, o):  # :176-187
        a, b = self._c(), o._c()
        if a is not None and b is not None:
            return Expr.const(a * b)
        if a == 0 or b == 0:
            return Expr.const(0)
        if a == 1:
            return o
        if b == 1:
            return self
        return Expr("mul", (self, o))

    def __neg__(self):  # :189-199
        if self.k == "const":
            return Expr.const(-self.x)
        if self.k == "neg":
            return self.x
        return Expr("neg", self)


class Interner:
    """src/graph.rs:213-330. A node is a tuple: ("const", v) ("var", source, offset, index) ("public", i) ("first",) ("last",)
    ("trans",) ("add", a, b) ("sub", a, b) ("mul", a, b) ("neg", a)."""

    def __init__(self):
        self.nodes, self.degrees, self.map = [], [], {}

    def intern(self, node):
        if node in self.map:
            return self.map[node]
        i = len(self.nodes)
        self.degrees.append(self.degree_of(node))
        self.nodes.append(node)
        self.map[node] = i
        return i

    def degree_of(self, n):
        k = n[0]
        if k in ("const", "public", "trans"):
            return 0
        if k in ("var", "first", "last"):
            return 1
        if k in ("add", "sub"):
            return max(self.degrees[n[1]], self.degrees[n[2]])
        if k == "mul":
            return self.degrees[n[1]] + self.degrees[n[2]]
        return self.degrees[n[1]]

    def as_const(self, i):
        return self.nodes[i][1] if self.nodes[i][0] == "const" else None

    def constant(self, v):
        return self.intern(("const", v % P))

    def add(self, a, b):
        x, y = self.as_const(a), self.as_const(b)
        if x is not None and y is not None:
            return self.constant(x + y)
        if x == 0:
            return b
        if y == 0:
            return a
        a, b = (a, b) if a <= b else (b, a)
        return self.intern(("add", a, b))

    def sub(self, a, b):
        if a == b:
            return self.constant(0)
        x, y = self.as_const(a), self.as_const(b)
        if x is not None and y is not None:
            return self.constant(x - y)
        if y == 0:
            return a
        if x == 0:
            return self.neg(b)
        return self.intern(("sub", a, b))

    def mul(self, a, b):
        x, y = self.as_const(a), self.as_const(b)
        if x is not None and y is not None:
            return self.constant(x * y)
        if x is not None:
            if x == 0:
                return a
            if x == 1:
                return b
        if y is not None:
            if y == 0:
                return b
            if y == 1:
                return a
        a, b = (a, b) if a <= b else (b, a)
        return self.intern(("mul", a, b))

    def neg(self, a):
        x = self.as_const(a)
        if x is not None:
            return self.constant(-x)
        if self.nodes[a][0] == "neg":
            return self.nodes[a][1]
        return self.intern(("neg", a))

    def compile_expr(self, e):
        k = e.k
        if k == "const":
            return self.constant(e.x)
        if k == "var":
            return self.intern(("var",) + e.x)
        if k in ("first", "last", "trans"):
            return self.intern((k,))
        if k == "public":
            return self.intern(("public", e.x))
        if k == "neg":
            return self.neg(self.compile_expr(e.x))
        a = self.compile_expr(e.x[0])
        b = self.compile_expr(e.x[1])
        return getattr(self, k)(a, b)


class Graph:
    pass


def compile_circuit(lookups, constraints):
    """src/graph.rs:120-188 for circuits without extension constraints (none of the benchmark circuits has any):
    lookups first (multiplicity, then the arguments in order), then the base constraints; roots that fold to the zero constant
    are dropped, the rest sorted by node id and deduplicated."""
    it = Interner()
    out_lookups = []
    for mult, args in lookups:
        m = it.compile_expr(mult)
        out_lookups.append((m, [it.compile_expr(a) for a in args]))
    g = Graph()
    g.lookup_prefix_len = len(it.nodes)
    zeros = []
    for c in constraints:
        r = it.compile_expr(c)
        cv = it.as_const(r)
        if cv is None:
            zeros.append(r)
        else:
            assert cv == 0, "unsatisfiable constant constraint"
    g.zeros = sorted(set(zeros))
    g.nodes, g.degrees, g.lookups = it.nodes, it.degrees, out_lookups
    g.max_constraint_degree = max([it.degrees[z] for z in g.zeros], default=0)
    return g


def logup_max_degree(g):  # src/lookup.rs:262-278
    best = None
    for mult, args in g.lookups:
        md = max([g.degrees[a] for a in args], default=0)
        d = max(md + 1, g.degrees[mult])
        best = d if best is None else max(best, d)
    return 1 if best is None else best


class Circuit:
    def __init__(self, main_width, lookups, constraints, pre_width=0, pre_height=0):
        self.main_width, self.pre_width, self.pre_height = main_width, pre_width, pre_height
        self.graph = compile_circuit(lookups, constraints)
        self.num_lookups = len(lookups)
        self.stage2_width = max(self.num_lookups, 1) * 2                       # src/lookup.rs:90-95
        self.constraint_count = len(self.graph.zeros) + max(self.num_lookups, 1) * 2   # src/system.rs:151
        self.max_constraint_degree = max(self.graph.max_constraint_degree, logup_max_degree(self.graph))

    def quotient_degree(self):  # src/system.rs:85-87
        x = max(self.max_constraint_degree, 2) - 1
        p = 1
        while p < x:
            p <<= 1
        return p


def pull(mult, args):  # src/lookup.rs:58-66: the multiplicity enters negated
    return (-mult, args)


def push(mult, args):
    return (mult, args)


def byte_table():  # benches/multi_stark.rs:85-90,136-140
    return Circuit(1, [pull(Expr.main(0), [Expr.const(0), Expr.preprocessed(0)])], [], pre_width=1, pre_height=256)


def _weighted(c0):
    return (Expr.main(c0) + Expr.main(c0 + 1) * Expr.const(256) + Expr.main(c0 + 2) * Expr.const(256 ** 2)
            + Expr.main(c0 + 3) * Expr.const(256 ** 3))


def u32_add():
    """benches/multi_stark.rs:101-165. p3-air: assert_bool(x) = assert_zero(x.bool_check()), assert_eq(a, b) = assert_zero(a - b);
    p3-field 0.5.1: bool_check(x) = x.andn(x) = (ONE - x) * x ("x * (1 - x) instead of x * (x - 1) as this lets us delegate to
    the andn function")."""
    carry = Expr.main(12)
    m = Expr.main
    c = Expr.const
    bool_check = (c(1) - carry) * carry
    expr1 = (m(0) + m(1) * c(256) + m(2) * c(256 ** 2) + m(3) * c(256 ** 3)
             + m(4) + m(5) * c(256) + m(6) * c(256 ** 2) + m(7) * c(256 ** 3))
    expr2 = m(8) + m(9) * c(256) + m(10) * c(256 ** 2) + m(11) * c(256 ** 3) + carry * c(256 ** 4)
    lookups = [pull(m(13), [c(1), _weighted(0), _weighted(4), _weighted(8)])]
    lookups += [push(c(1), [c(0), m(i)]) for i in range(12)]
    return Circuit(14, lookups, [bool_check, expr1 - expr2])


def named_system(kind):
    if kind == "u32_add":
        return [byte_table(), u32_add()]
    if kind.startswith("multi:"):
        return [byte_table()] + [u32_add() for _ in range(int(kind[6:]))]
    raise ValueError(kind)


# ------------------------------------------------------------------------------------------------------------------
# MMCS (p3-merkle-tree verify_batch with the reference's hashes, src/types.rs:82-84,199-207)
# ------------------------------------------------------------------------------------------------------------------
def hash_rows(rows):
    h = _b3.blake3()
    for row in rows:
        for v in row:
            h.update(int(v).to_bytes(8, "little"))
    return h.digest()


def compress(l, r):
    return _b3.blake3(l + r).digest()


def _np2(x):
    p = 1
    while p < x:
        p <<= 1
    return p


def verify_batch(commit, heights, index, opened_values, proof):
    if len(heights) != len(opened_values):
        return False
    order = sorted(range(len(heights)), key=lambda i: -heights[i])  # stable: ties keep matrix order
    pos = 0
    cur = _np2(heights[order[0]])
    group = []
    while pos < len(order) and _np2(heights[order[pos]]) == cur:
        group.append(opened_values[order[pos]])
        pos += 1
    root = hash_rows(group)
    if len(proof) != cur.bit_length() - 1:
        return False
    for sib in proof:
        root = compress(root, sib) if index & 1 == 0 else compress(sib, root)
        index >>= 1
        cur >>= 1
        if pos < len(order) and _np2(heights[order[pos]]) == cur:
            group = []
            while pos < len(order) and _np2(heights[order[pos]]) == cur:
                group.append(opened_values[order[pos]])
                pos += 1
            root = compress(root, hash_rows(group))
    return pos == len(order) and root == commit


# ------------------------------------------------------------------------------------------------------------------
# FRI PCS verification
# ------------------------------------------------------------------------------------------------------------------
def fold_row(index, log_height, beta, e0, e1):
    x0 = pow(two_adic_generator(log_height + 1), rev_bits(index, log_height), P)
    x1 = (-x0) % P
    return e0 + (beta - x0) * (e1 - e0) * inv(x1 - x0)


def pcs_verify(rounds, fri, params, ch, out=None):
    """rounds: [(commit, [(log_degree, [(z, [values])...])...])...]. Returns None on success, else a reason string.
    `out` (dict) receives the derived challenges."""
    lb = params["log_blowup"]
    for _, mats in rounds:
        for _, pts in mats:
            for _, vals in pts:
                for y in vals:
                    ch.observe(y)
    alpha = ch.sample_ext()
    commits = fri["commit_phase_commits"]
    if len(fri["commit_pow_witnesses"]) != len(commits):
        return "pow witness count"
    betas = []
    for c, wv in zip(commits, fri["commit_pow_witnesses"]):
        ch.observe(c)
        if not ch.check_witness(params["commit_pow_bits"], wv):
            return "commit pow"
        betas.append(ch.sample_ext())
    final_poly = [E(*c) for c in fri["final_poly"]]
    if len(final_poly) != 1 << params["log_final_poly_len"]:
        return "final poly length"
    for c in final_poly:
        ch.observe(c)
    if len(fri["query_proofs"]) != params["num_queries"]:
        return "query count"
    if not ch.check_witness(params["query_pow_bits"], fri["query_pow_witness"]):
        return "query pow"
    log_max = len(commits) + lb + params["log_final_poly_len"]
    indices = []
    for qp in fri["query_proofs"]:
        index = ch.sample_bits(log_max)
        indices.append(index)
        if len(qp["input_proof"]) != len(rounds):
            return "input proof count"
        reduced = {}
        for bo, (commit, mats) in zip(qp["input_proof"], rounds):
            if len(bo["opened_values"]) != len(mats):
                return "opened matrices"
            if mats:
                heights = [1 << (ld + lb) for ld, _ in mats]
                lbm = max(ld + lb for ld, _ in mats)
                if lbm > log_max:
                    return "height"
                if not verify_batch(commit, heights, index >> (log_max - lbm), bo["opened_values"], bo["opening_proof"]):
                    return "input merkle path"
            for row, (ld, pts) in zip(bo["opened_values"], mats):
                lh = ld + lb
                x = GENERATOR * pow(two_adic_generator(lh), rev_bits(index >> (log_max - lh), lh), P) % P
                ap, ro = reduced.get(lh, (E(1), E(0)))
                for z, vals in pts:
                    if len(vals) != len(row):
                        return "opened width"
                    q = (z - x).inverse()
                    for px, pz in zip(row, vals):
                        ro = ro + ap * (pz - px) * q
                        ap = ap * alpha
                reduced[lh] = (ap, ro)
        if lb in reduced:
            if not reduced[lb][1].is_zero():
                return "constant matrix"
            del reduced[lb]
        ro_desc = sorted(((lh, ro) for lh, (_, ro) in reduced.items()), reverse=True)
        steps = qp["commit_phase_openings"]
        if len(steps) != len(commits) or not ro_desc or ro_desc[0][0] != log_max:
            return "query shape"
        folded = ro_desc[0][1]
        rp = 1
        di = index
        for k, (step, commit) in enumerate(zip(steps, commits)):
            lfh = log_max - 1 - k
            if step["log_arity"] != 1 or len(step["sibling_values"]) != 1:
                return "arity"
            evals = [folded, folded]
            evals[(di ^ 1) & 1] = E(*step["sibling_values"][0])
            di >>= 1
            row = [evals[0].a, evals[0].b, evals[1].a, evals[1].b]  # ExtensionMmcs: flattened to base
            if not verify_batch(commit, [1 << lfh], di, [row], step["opening_proof"]):
                return "layer merkle path"
            folded = fold_row(di, lfh, betas[k], evals[0], evals[1])
            if rp < len(ro_desc) and ro_desc[rp][0] == lfh:
                folded = folded + betas[k] * betas[k] * ro_desc[rp][1]
                rp += 1
        if rp != len(ro_desc):
            return "input never rolled in"
        x = pow(two_adic_generator(log_max), rev_bits(di, log_max), P)
        ev = E(0)
        for c in reversed(final_poly):
            ev = ev * x + c
        if ev != folded:
            return "final polynomial mismatch"
    if out is not None:
        out.update(alpha_pcs=alpha, fri_betas=betas, query_indices=indices)
    return None


# ------------------------------------------------------------------------------------------------------------------
# System::verify_multiple_claims
# ------------------------------------------------------------------------------------------------------------------
def sweep(graph, pre, main, stage2, publics, first, last, trans):
    """src/eval.rs:67-106 over extension values (the verifier's instantiation)."""
    buf = []
    views = {PRE: pre, MAIN: main, STAGE2: stage2}
    for n in graph.nodes:
        k = n[0]
        if k == "const":
            v = E(n[1])
        elif k == "var":
            v = views[n[1]][n[2]][n[3]]
        elif k == "public":
            v = publics[n[1]]
        elif k == "first":
            v = first
        elif k == "last":
            v = last
        elif k == "trans":
            v = trans
        elif k == "add":
            v = buf[n[1]] + buf[n[2]]
        elif k == "sub":
            v = buf[n[1]] - buf[n[2]]
        elif k == "mul":
            v = buf[n[1]] * buf[n[2]]
        else:
            v = -buf[n[1]]
        buf.append(v)
    return buf


def _mul2(a, b):  # src/lookup.rs:123-128 over "coordinates" that are themselves extension values
    v0, v1 = a[0] * b[0], a[1] * b[1]
    cross = (a[0] + a[1]) * (b[0] + b[1]) - v0 - v1
    return (v0 + v1 * W, cross)


def logup_constraint_values(lookups, buf, s2, s2n, publics, delta_scaled, is_last, out):  # src/lookup.rs:152-208, D = 2
    beta, gamma = (publics[0], publics[1]), (publics[2], publics[3])
    inj = (is_last * delta_scaled[0], is_last * delta_scaled[1])
    if not lookups:
        out.append(s2n[0] - s2[0] + inj[0])
        out.append(s2n[1] - s2[1] + inj[1])
        return
    last = len(lookups) - 1
    for j, (mult, args) in enumerate(lookups):
        source = (s2[2 * j], s2[2 * j + 1])
        target = (s2[2 * j + 2], s2[2 * j + 3]) if j < last else (s2n[0] + inj[0], s2n[1] + inj[1])
        f = (E(0), E(0))
        for a in reversed(args):
            f = _mul2(f, gamma)
            f = (f[0] + buf[a], f[1])
        c = _mul2((f[0] + beta[0], f[1] + beta[1]), (target[0] - source[0], target[1] - source[1]))
        out.append(c[0] - buf[mult])
        out.append(c[1])


def selectors_at_point(log_n, z):
    """p3-commit TwoAdicMultiplicativeCoset::selectors_at_point on the natural domain H_n (shift 1), unnormalised:
    is_first = Z_H / (z - 1), is_last = Z_H / (z - g^-1), is_transition = z - g^-1, inv_vanishing = 1 / Z_H."""
    zh = z.pow(1 << log_n) - 1
    ginv = inv(two_adic_generator(log_n))
    return dict(first=zh * (z - 1).inverse(), last=zh * (z - ginv).inverse(), trans=z - ginv, inv_vanishing=zh.inverse())


def verify(circuits, params, pre_commit, claims, proof, out=None):
    """Returns "Ok" or the name of the reference's VerificationError variant (src/verifier.rs:176-192)."""
    active = [bool(a) for a in proof["active"]]
    pre_idx, n_pre = [], 0
    for c in circuits:
        if c.pre_width:
            pre_idx.append(n_pre)
            n_pre += 1
        else:
            pre_idx.append(None)
    if (n_pre == 0) != (pre_commit is None) or not circuits:
        return "InvalidSystem"
    # verify_shape (src/verifier.rs:536-695)
    if len(active) != len(circuits):
        return "InvalidProofShape"
    act = [i for i, a in enumerate(active) if a]
    ld = list(proof["log_degrees"])
    pov = proof["preprocessed_opened_values"]
    if not act or len(ld) != len(act) or (len(pov) if pov is not None else 0) != n_pre:
        return "InvalidProofShape"
    s1, s2, qv = proof["stage_1_opened_values"], proof["stage_2_opened_values"], proof["quotient_opened_values"]
    if len(s1) != len(act) or len(s2) != len(act) or len(qv) != len(act) or len(proof["intermediate_accumulators"]) != len(act):
        return "InvalidProofShape"
    for ci, c in enumerate(circuits):
        if pre_idx[ci] is not None and not active[ci] and len(pov[pre_idx[ci]]) != 0:
            return "InvalidProofShape"
    qdeg = []
    for pos, ci in enumerate(act):
        c = circuits[ci]
        if len(s1[pos]) != 2 or len(s2[pos]) != 2 or len(qv[pos]) != 1:
            return "InvalidProofShape"
        slot = pre_idx[ci]
        if slot is not None and len(pov[slot]) != 2:
            return "InvalidProofShape"
        for j in range(2):
            if slot is not None and len(pov[slot][j]) != c.pre_width:
                return "InvalidProofShape"
            if len(s1[pos][j]) != c.main_width or len(s2[pos][j]) != c.stage2_width:
                return "InvalidProofShape"
        q = c.quotient_degree()
        if ld[pos] + (q.bit_length() - 1) > 32 - params["log_blowup"] or len(qv[pos][0]) != q * 2:
            return "InvalidProofShape"
        qdeg.append(q)
    accs = [E(*a) for a in proof["intermediate_accumulators"]]
    if not accs[-1].is_zero():
        return "UnbalancedChannel"

    ch = PyChallenger.for_config(params["log_blowup"], 0, params["log_final_poly_len"], params["max_log_arity"],
                                 params["num_queries"], params["commit_pow_bits"], params["query_pow_bits"])
    ch.observe(len(circuits))                       # observe_shape, src/system.rs:211-222
    for c in circuits:
        for x in (c.constraint_count, c.max_constraint_degree, c.pre_height, c.pre_width, c.main_width, c.stage2_width):
            ch.observe(x)
    for a in active:
        ch.observe(1 if a else 0)
    if pre_commit is not None:
        ch.observe(pre_commit)
    ch.observe(proof["stage_1_trace"])
    for l in ld:
        ch.observe(l)
    ch.observe(len(claims))
    for cl in claims:
        ch.observe(len(cl))
        for v in cl:
            ch.observe(int(v))
    beta = ch.sample_ext()
    ch.observe(beta)
    gamma = ch.sample_ext()
    ch.observe(gamma)
    ch.observe(proof["stage_2_trace"])
    for a in accs:
        ch.observe(a)
    acc = E(0)
    for cl in claims:                               # fingerprint = sum_i claim_i gamma^i (Horner over the reversed claim)
        f = E(0)
        for v in reversed(cl):
            f = f * gamma + int(v)
        acc = acc + (beta + f).inverse()
    alpha = ch.sample_ext()
    ch.observe(proof["quotient_chunks"])
    zeta = ch.sample_ext()
    if out is not None:
        out.update(beta=beta, gamma=gamma, alpha=alpha, zeta=zeta)

    def ev(rows):
        return [[E(*v) for v in row] for row in rows]
    r1, r2, r3 = [], [], []
    for pos in range(len(act)):
        zn = zeta * two_adic_generator(ld[pos])
        a, b = ev(s1[pos]), ev(s2[pos])
        r1.append((ld[pos], [(zeta, a[0]), (zn, a[1])]))
        r2.append((ld[pos], [(zeta, b[0]), (zn, b[1])]))
        r3.append((ld[pos], [(zeta, ev(qv[pos])[0])]))
    rounds = [(proof["stage_1_trace"], r1), (proof["stage_2_trace"], r2), (proof["quotient_chunks"], r3)]
    if pre_commit is not None:
        apos = {ci: pos for pos, ci in enumerate(act)}
        r0 = []
        for ci, c in enumerate(circuits):
            if pre_idx[ci] is None:
                continue
            if ci in apos:
                l = ld[apos[ci]]
                pv = ev(pov[pre_idx[ci]])
                r0.append((l, [(zeta, pv[0]), (zeta * two_adic_generator(l), pv[1])]))
            else:
                r0.append((c.pre_height.bit_length() - 1, []))
        rounds.append((pre_commit, r0))
    why = pcs_verify(rounds, proof["opening_proof"], params, ch, out)
    if why is not None:
        if out is not None:
            out["pcs_error"] = why
        return "InvalidOpeningArgument"

    for pos, ci in enumerate(act):
        c = circuits[ci]
        n = 1 << ld[pos]
        next_acc = accs[pos]
        sels = selectors_at_point(ld[pos], zeta)
        inj_norm = inv(n * two_adic_generator(ld[pos]))
        publics = []
        for ef in (beta, gamma, acc, next_acc):
            publics += [E(ef.a), E(ef.b)]
        slot = pre_idx[ci]
        pre = ev(pov[slot]) if slot is not None else [[], []]
        main, st2 = ev(s1[pos]), ev(s2[pos])
        buf = sweep(c.graph, pre, main, st2, publics, sels["first"], sels["last"], sels["trans"])
        cv = [buf[z] for z in c.graph.zeros]
        delta = [(publics[6] - publics[4]) * inj_norm, (publics[7] - publics[5]) * inj_norm]
        logup_constraint_values(c.graph.lookups, buf, st2[0], st2[1], publics, delta, sels["last"], cv)
        assert len(cv) == c.constraint_count
        comp = E(0)
        for v in cv:
            comp = comp * alpha + v
        qrow = ev(qv[pos])[0]
        zpn = zeta.pow(n)
        zp, quot = E(1), E(0)
        for k in range(0, len(qrow), 2):
            quot = quot + zp * (qrow[k] + qrow[k + 1] * E(0, 1))
            zp = zp * zpn
        if comp * sels["inv_vanishing"] != quot:
            return "OodEvaluationMismatch"
        acc = next_acc
    return "Ok"


def graph_dict(c):
    """A compiled circuit in the form multi_stark_b200.System.from_graphs / msgpu_graph_desc take."""
    return dict(nodes=list(c.graph.nodes), zeros=list(c.graph.zeros), lookups=[(m, list(a)) for m, a in c.graph.lookups],
                lookup_prefix_len=c.graph.lookup_prefix_len, main_width=c.main_width, pre_width=c.pre_width)
