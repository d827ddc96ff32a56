"""CPU tests (-m "not gpu"): the oracle against every known-answer / invariant the reference's own tests
hold for the field, MLE, eq and sumcheck layers."""
import numpy as np
import oracle_py as O

P = O.P


def E(v):
    return np.array([v % P, 0], dtype=np.uint64)


def negE(v):
    return np.array([(P - v) % P, 0], dtype=np.uint64)


def test_field_constants():
    # SURVEY.md "facts verified": p-1 = 2^32 * 3*5*17*257*65537; 7 generates F*; X^2-7 irreducible
    assert P - 1 == 2**32 * 3 * 5 * 17 * 257 * 65537
    assert pow(7, (P - 1) // 2, P) == P - 1
    assert pow(7, (P - 1) // 2**32, P) == 1753635133440165772
    a = O.splitmix_e(11, 64)
    inv = O.e_inv(a)
    one = O.e_binop(2, a, inv)
    assert (one[:, 0] == 1).all() and (one[:, 1] == 0).all()


def test_ext_mul_matches_python_bigint():
    a, b = O.splitmix_e(5, 500), O.splitmix_e(6, 500)
    got = O.e_binop(2, a, b)
    for i in range(500):
        assert tuple(int(x) for x in got[i]) == O.pe_mul(a[i], b[i])
    # non-canonical inputs are reduced first
    big = np.array([[P + 5, 2**64 - 1]], dtype=np.uint64)
    got = O.e_binop(2, big, big)[0]
    assert tuple(int(x) for x in got) == O.pe_mul((5, (2**64 - 1) % P), (5, (2**64 - 1) % P))


def test_fix_high_variables_kat():
    """reference KAT: multilinear_extensions/src/test.rs:47-82 (exact integers)"""
    evals = np.array([13, 97, 11, 101, 7, 103, 5, 107], dtype=np.uint64)
    pt = np.stack([E(3), E(5)])
    r1 = O.fix_high(evals, False, pt[1:])
    exp1 = np.stack([negE(17), E(127), negE(19), E(131)])
    assert (r1 == exp1).all()
    r2 = O.fix_high(evals, False, pt)
    exp2 = np.stack([negE(23), E(139)])
    assert (r2 == exp2).all()


def naive_eq(r):
    nv = len(r)
    out = []
    for x in range(1 << nv):
        cur = (1, 0)
        for i in range(nv):
            ri = (int(r[i][0]), int(r[i][1]))
            cur = O.pe_mul(cur, ri if (x >> i) & 1 else O.pe_sub((1, 0), ri))
        out.append(cur)
    return np.array(out, dtype=np.uint64)


def test_eq_xr_matches_naive():
    """multilinear_extensions/src/test.rs:36-44 (build_eq_x_r vs build_eq_x_r_for_test)"""
    for nv in range(0, 10):
        r = O.splitmix_e(100 + nv, nv) if nv else np.zeros((0, 2), dtype=np.uint64)
        assert (O.build_eq(r) == naive_eq(r)).all()


def test_eq_eval_and_evaluate_consistency():
    nv = 6
    r, y = O.splitmix_e(1, nv), O.splitmix_e(2, nv)
    eq = O.build_eq(r)
    # eq(x, r) as an MLE evaluated at y == eq_eval(r, y)
    assert (O.evaluate(eq, True, y) == O.eq_eval(r, y)).all()
    # fix_low then fix_high commute with evaluate
    f = O.splitmix_f(3, 1 << nv)
    full = O.evaluate(f, False, y)
    lo = O.fix_low(f, False, y[:2])
    assert (O.evaluate(lo, True, y[2:]) == full).all()
    hi = O.fix_high(f, False, y[4:])
    assert (O.evaluate(hi, True, y[:4]) == full).all()


def rand_vp(seed, nv, shape):
    """shape: list of products, each a list of 'b'/'e' operand kinds; returns (mles, products)"""
    mles, products = [], []
    s = seed
    for kinds in shape:
        idx = []
        for k in kinds:
            s += 1
            if k == "b":
                mles.append((O.splitmix_f(s, 1 << nv), False))
            else:
                mles.append((O.splitmix_e(s, 1 << nv), True))
            idx.append(len(mles) - 1)
        s += 1
        products.append((tuple(int(v) for v in O.splitmix_e(s, 1)[0]), idx))
    return mles, products


def vp_sum(mles, products, nv):
    tot = (0, 0)
    for coef, idx in products:
        acc = (0, 0)
        n = mles[idx[0]][0].reshape(-1).size // (2 if mles[idx[0]][1] else 1)
        for x in range(n):
            t = (1, 0)
            for i in idx:
                arr, ext = mles[i]
                t = O.pe_mul(t, tuple(arr.reshape(-1, 2)[x]) if ext else (int(arr[x]), 0))
            acc = O.pe_add(acc, t)
        scale = 1 << (nv - (n.bit_length() - 1))
        acc = O.pe_mul(acc, (scale % P, 0))
        tot = O.pe_add(tot, O.pe_mul(acc, coef))
    return tot


def test_sumcheck_prove_verify_roundtrip():
    """sumcheck/src/test.rs:23-56,170-177: extract_sum == true sum; verify accepts; subclaim == vp(point)"""
    for nv, shape in [(1, [["b"]]), (3, [["b", "b", "b"]]), (4, [["e", "b"], ["e"]]), (5, [["e", "e", "e"], ["b", "e"]]),
                      (4, [["b", "b", "b", "b", "b"], ["e", "e", "e", "e"]])]:
        mles, products = rand_vp(1000 + nv, nv, shape)
        point, msgs, fin = O.sumcheck_prove(mles, products, nv)
        true_sum = vp_sum(mles, products, nv)
        s = O.pe_add(msgs[0][0], msgs[0][1])
        assert s == true_sum
        max_deg = max(len(p[1]) for p in products)
        vpoint, expected = O.sumcheck_verify(np.array(true_sum, dtype=np.uint64), nv, max_deg, msgs)
        assert (vpoint == point).all()
        # subclaim: sum_i c_i prod_j f_ij(point) using the prover's final evaluations AND mle.evaluate
        tot = (0, 0)
        for coef, idx in products:
            t = coef
            for i in idx:
                ev = O.evaluate(mles[i][0], mles[i][1], point)
                assert (ev == fin[i]).all()
                t = O.pe_mul(t, ev)
            tot = O.pe_add(tot, t)
        assert tot == tuple(int(v) for v in expected)


def test_sumcheck_mixed_num_vars_multiplicity():
    """products with fewer variables than max are scaled by 2^(missing) (sumcheck_macro/src/lib.rs:242-247)"""
    nv = 4
    big = (O.splitmix_f(1, 1 << nv), False)
    small_a = (O.splitmix_e(2, 1 << 2), True)
    small_b = (O.splitmix_f(3, 1 << 2), False)
    mles = [big, small_a, small_b]
    products = [((3, 0), [0]), ((5, 1), [1, 2])]
    point, msgs, fin = O.sumcheck_prove(mles, products, nv)
    assert O.pe_add(msgs[0][0], msgs[0][1]) == vp_sum(mles, products, nv)
    O.sumcheck_verify(np.array(vp_sum(mles, products, nv), dtype=np.uint64), nv, 2, msgs)
    # small MLEs are fully folded after 2 rounds and then stay constant
    assert (O.evaluate(small_a[0], True, point[:2]) == fin[1]).all()


def test_sumcheck_rejects_bad_claim():
    mles, products = rand_vp(7, 3, [["b", "e"]])
    _, msgs, _ = O.sumcheck_prove(mles, products, 3)
    bad = np.array([1, 2], dtype=np.uint64)
    try:
        O.sumcheck_verify(bad, 3, 2, msgs)
        assert False, "verifier accepted a wrong sum"
    except RuntimeError as e:
        assert "claim" in str(e)


def test_fixed_challenge_rounds_match_transcript_run():
    mles, products = rand_vp(42, 4, [["e", "b", "b"], ["b"]])
    point, msgs, fin = O.sumcheck_prove(mles, products, 4)
    msgs2, fin2 = O.sumcheck_rounds_fixed(mles, products, 4, point)
    assert (msgs == msgs2).all() and (fin == fin2).all()


def test_prove_batch_polys_equals_prove_parallel():
    """zkml/src/model/mod.rs:987-993 (test_model_sequential): the devirgo split produces the same proof"""
    for nv, shape, T in [(5, [["b", "e"]], 2), (6, [["e", "e", "e"], ["b", "e"]], 4), (4, [["b", "b", "b"]], 4), (3, [["e"]], 1)]:
        mles, products = rand_vp(77 + nv, nv, shape)
        p1, m1, f1 = O.sumcheck_prove(mles, products, nv)
        p2, m2, f2 = O.sumcheck_prove_batch(T, mles, products, nv)
        assert (p1 == p2).all() and (m1 == m2).all() and (f1 == f2).all()
