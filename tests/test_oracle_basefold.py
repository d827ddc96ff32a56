"""CPU tests: the Basefold oracle against the invariants the reference's own mpcs tests pin
(SURVEY.md 8c items 5-6): NTT == naive Horner over the coset (rs.rs:558-624), folding coefficients
w*(x1-x0) = 1 (arithmetic.rs:126-129), fold(encode(m)) == encode(fold(m)) (encoding.rs:174-238), the
sanity-check asserts of commit_phase.rs:148-169 / :326-345 (raised as exceptions by the oracle), and the
Merkle shapes of merkle_tree.rs:261-420."""
import numpy as np
import oracle_py as O

P = O.P
GEN = 7
ROOT32 = 1753635133440165772


def two_adic(bits):
    return pow(ROOT32, 1 << (32 - bits), P)


def bitrev_perm(n_log):
    n = 1 << n_log
    return np.array([int(format(i, "0%db" % n_log)[::-1], 2) if n_log else 0 for i in range(n)])


def test_rs_encode_matches_naive_horner():
    """encode_internal == naive_fft (rs.rs:540-556): codeword[i] = poly(shift * w^i)"""
    for lg_m, full_log in [(3, 3), (4, 6), (5, 5)]:
        m = 1 << lg_m
        c = O.splitmix_f(lg_m, m)
        cw = O.rs_encode(c, False, full_log)
        shift = pow(GEN, 1 << (full_log - lg_m), P)
        w = two_adic(lg_m + 1)
        for i in range(2 * m):
            x = shift * pow(w, i, P) % P
            assert int(cw[i]) == sum(int(c[j]) * pow(x, j, P) for j in range(m)) % P
    # Ext coefficients: componentwise
    ce = O.splitmix_e(9, 8)
    cwe = O.rs_encode(ce, True, 4)
    assert (cwe[:, 0] == O.rs_encode(ce[:, 0].copy(), False, 4)).all() and (cwe[:, 1] == O.rs_encode(ce[:, 1].copy(), False, 4)).all()


def test_interpolate_is_moebius_and_commutes_with_bitrev():
    v = O.splitmix_f(1, 32)
    c = O.interpolate_hc(v, False)
    # evaluating the multilinear coefficients at every hypercube point gives back the evaluations
    for x in range(32):
        assert sum(int(c[s]) for s in range(32) if s & x == s) % P == int(v[x])
    perm = bitrev_perm(5)
    assert (O.interpolate_hc(v[perm], False) == c[perm]).all()


def test_folding_coeffs_weight_identity():
    """interpolate2_weights sanity: w * (x1 - x0) == 1, x1 = -x0 (arithmetic.rs:126-129, rs.rs:377-410)"""
    for full_log, level in [(10, 3), (10, 10), (12, 0), (20, 7)]:
        for idx in [0, 1, (1 << level) - 1, (1 << level) // 2]:
            if idx >= (1 << level) and level > 0:
                continue
            x0, w = O.folding_coeffs(full_log, level, idx)
            assert w * ((P - x0 - x0) % P) % P == 1
            assert pow(x0, 1 << (level + 1), P) == pow(GEN, 1 << (full_log + 1), P)  # x0 in the coset of size 2^(level+1)


def test_rs_codeword_folding():
    """encoding.rs:174-238: folding the bit-reversed codeword == encoding the folded (even/odd) message"""
    full_log = 8
    nv = 8
    msg = O.splitmix_e(3, 1 << nv)                     # coefficients, natural order
    perm = bitrev_perm(nv)
    cw = O.rs_encode(msg[perm], True, full_log)         # the reference encodes the bit-reversed message
    cw = cw[bitrev_perm(nv + 1)]
    for rnd in range(5):
        r = O.splitmix_e(50 + rnd, 1)[0]
        cw = O.fri_fold(cw, full_log, r)
        lo, hi = msg[: msg.shape[0] // 2], msg[msg.shape[0] // 2:]   # left-right fold of the coefficient vector
        msg = np.array([O.pe_add(lo[i], O.pe_mul(r, hi[i])) for i in range(lo.shape[0])], dtype=np.uint64)
        k = msg.shape[0].bit_length() - 1
        exp = O.rs_encode(msg[bitrev_perm(k)], True, full_log)[bitrev_perm(k + 1)]
        assert (cw == exp).all(), "round %d" % rnd


def test_merkle_shapes():
    # 2 base leaves: root is the zero-padded pair (hash_or_noop, no permutation)
    assert (O.merkle_root([5, 6], False) == np.array([5, 6, 0, 0], dtype=np.uint64)).all()
    # 4 base leaves: compress of the two padded pairs
    assert (O.merkle_root([1, 2, 3, 4], False) == O.compress([1, 2, 0, 0], [3, 4, 0, 0])).all()
    # 4 ext leaves: leaf pairs are packed (4 limbs), then compressed
    e = O.splitmix_e(1, 4)
    assert (O.merkle_root(e, True) == O.compress(e[:2].reshape(-1), e[2:].reshape(-1))).all()
    b = O.splitmix_f(2, 8)
    l1 = [O.compress([b[0], b[1], 0, 0], [b[2], b[3], 0, 0]), O.compress([b[4], b[5], 0, 0], [b[6], b[7], 0, 0])]
    assert (O.merkle_root(b, False) == O.compress(l1[0], l1[1])).all()


def test_commit_layout():
    nv, full_log = 8, 10
    ev = O.splitmix_f(4, 1 << nv)
    root, cw, bh = O.pcs_commit(ev, False, full_log)
    perm = bitrev_perm(nv)
    assert (bh == ev[perm]).all()
    coeffs = O.interpolate_hc(ev, False)[perm]
    assert (cw == O.rs_encode(coeffs, False, full_log)[bitrev_perm(nv + 1)]).all()
    assert (root == O.merkle_root(cw, False)).all()
    # trivial commitment: Merkle tree over the raw evaluations (basefold.rs:102-104)
    ev7 = O.splitmix_f(5, 1 << 7)
    root7, cw7, bh7 = O.pcs_commit(ev7, False, full_log)
    assert (cw7 == ev7).all() and (root7 == O.merkle_root(ev7, False)).all()


def parse_flat(f):
    """decode flatten_proof for structural checks"""
    f = [int(x) for x in f]
    pos = 0

    def take(n):
        nonlocal pos
        v = f[pos:pos + n]
        pos += n
        return v
    out = {}
    n = take(1)[0]; out["sumcheck_proof"] = [take(6) for _ in range(n)]
    n = take(1)[0]; out["sumcheck_messages"] = [take(6) for _ in range(n)]
    n = take(1)[0]; out["roots"] = [take(4) for _ in range(n)]
    n = take(1)[0]; out["final_message"] = [take(2) for _ in range(n)]

    def q():
        idx, is_base = take(2)
        vals = take(2) if is_base else take(4)
        k = take(1)[0]
        return {"index": idx, "is_base": is_base, "vals": vals, "path": [take(4) for _ in range(k)]}
    n = take(1)[0]; out["single"] = []
    for _ in range(n):
        x = take(1)[0]; cq = q(); k = take(1)[0]
        out["single"].append({"x": x, "commitment": cq, "oracle": [q() for _ in range(k)]})
    n = take(1)[0]; out["batched"] = []
    for _ in range(n):
        x = take(1)[0]; k = take(1)[0]; oq = [q() for _ in range(k)]; k = take(1)[0]
        out["batched"].append({"x": x, "oracle": oq, "commitments": [q() for _ in range(k)]})
    assert pos == len(f)
    return out


def authenticate(qr, root, is_base_leaf):
    """authenticate_merkle_path_root (merkle_tree.rs:424-...): leaf pair -> packed digest -> compress up"""
    v = qr["vals"]
    cur = np.array([v[0], v[1], 0, 0] if qr["is_base"] else v, dtype=np.uint64)
    idx = qr["index"] >> 1
    for d in qr["path"]:
        d = np.array(d, dtype=np.uint64)
        cur = O.compress(cur, d) if idx & 1 == 0 else O.compress(d, cur)
        idx >>= 1
    return [int(x) for x in cur] == list(root)


def test_open_structure_and_merkle_paths():
    nv, full_log = 9, 9
    ev = O.splitmix_f(6, 1 << nv)
    pt = O.splitmix_e(7, nv)
    root, _, _ = O.pcs_commit(ev, False, full_log)
    pr = parse_flat(O.pcs_open(ev, False, full_log, pt))
    assert len(pr["sumcheck_messages"]) == nv - 7 and len(pr["roots"]) == nv - 8 and len(pr["final_message"]) == 128
    assert len(pr["single"]) == 200 and not pr["batched"] and not pr["sumcheck_proof"]
    # first sumcheck message: p(0)+p(1) = 2 c0 + c1 + c2 = poly(point)
    m = pr["sumcheck_messages"][0]
    c0, c1, c2 = (m[0], m[1]), (m[2], m[3]), (m[4], m[5])
    s = O.pe_add(O.pe_add(O.pe_add(c0, c0), c1), c2)
    assert s == tuple(int(x) for x in O.evaluate(ev, False, pt))
    for qr in pr["single"][:40]:
        assert authenticate(qr["commitment"], [int(x) for x in root], True)
        for k, oq in enumerate(qr["oracle"]):
            assert authenticate(oq, pr["roots"][k], False)
            assert oq["index"] == ((qr["x"] >> (k + 1)) | 1) - 1


def test_batch_open_structure():
    full_log = 10
    polys = [(O.splitmix_f(1, 1 << 10), False), (O.splitmix_f(2, 1 << 8), False), (O.splitmix_e(3, 1 << 9), True)]
    pts = [O.splitmix_e(10 + i, (p[0].reshape(-1).size // (2 if p[1] else 1)).bit_length() - 1) for i, p in enumerate(polys)]
    pr = parse_flat(O.pcs_batch_open(polys, full_log, pts))
    assert len(pr["sumcheck_proof"]) == 10 and len(pr["sumcheck_messages"]) == 3 and len(pr["roots"]) == 2
    assert len(pr["batched"]) == 200 and len(pr["batched"][0]["commitments"]) == 3 and len(pr["batched"][0]["oracle"]) == 2
    roots = [O.pcs_commit(p[0], p[1], full_log, want_codeword=False)[0] for p in polys]
    for qr in pr["batched"][:20]:
        for k, cq in enumerate(qr["commitments"]):
            assert authenticate(cq, [int(x) for x in roots[k]], not polys[k][1])
