"""Property tests (hypothesis) for the CPU checker's algebra: the identities every later layer relies on.
Field: Goldilocks p = 2^64 - 2^32 + 1 and F[X]/(X^2 - 7) against Python integers; MLE: fix_low / fix_high / evaluate / eq agree."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle_py as O

P = 0xFFFFFFFF00000001
felt = st.integers(min_value=0, max_value=P - 1)
edge = st.sampled_from([0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, (P - 1) // 2])
fe = st.one_of(felt, edge)


def e_mul_py(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


@settings(max_examples=200, deadline=None)
@given(fe, fe)
def test_base_field_ops_match_python_integers(a, b):
    A, B = np.array([a], dtype=np.uint64), np.array([b], dtype=np.uint64)
    assert int(O.f_binop(0, A, B)[0]) == (a + b) % P
    assert int(O.f_binop(1, A, B)[0]) == (a - b) % P
    assert int(O.f_binop(2, A, B)[0]) == (a * b) % P


@settings(max_examples=200, deadline=None)
@given(fe, fe, fe, fe)
def test_extension_field_ops_match_the_definition(a0, a1, b0, b1):
    A, B = np.array([[a0, a1]], dtype=np.uint64), np.array([[b0, b1]], dtype=np.uint64)
    got = O.e_binop(2, A, B)[0]
    assert (int(got[0]), int(got[1])) == e_mul_py((a0, a1), (b0, b1))
    if (a0, a1) != (0, 0):
        inv = O.e_inv(A)[0]
        assert e_mul_py((a0, a1), (int(inv[0]), int(inv[1]))) == (1, 0)


@settings(max_examples=25, deadline=None)
@given(st.integers(min_value=1, max_value=7), st.integers(min_value=0, max_value=2**31), st.booleans())
def test_mle_folds_and_evaluation_agree(nv, seed, ext):
    ev = O.splitmix_e(seed, 1 << nv) if ext else O.splitmix_f(seed, 1 << nv)
    pt = O.splitmix_e(seed + 1, nv)
    full = O.evaluate(ev, ext, pt)
    k = (seed % nv) + 1
    # fixing the low k variables then evaluating the rest == evaluating at the whole point (mle.rs:454-525, 607-623)
    low = O.fix_low(ev, ext, pt[:k])
    assert (O.evaluate(low, True, pt[k:]) == full).all() if nv > k else (low.reshape(-1, 2)[0] == full).all()
    # fixing the high k variables (mle.rs:562-603) then evaluating the low ones
    high = O.fix_high(ev, ext, pt[nv - k:])
    assert (O.evaluate(high, True, pt[:nv - k]) == full).all() if nv > k else (high.reshape(-1, 2)[0] == full).all()
    # evaluate == <eq(point, .), evals>  (virtual_poly.rs:346-453)
    eq = O.build_eq(pt).reshape(-1, 2)
    vals = ev.reshape(-1, 2) if ext else np.stack([ev, np.zeros_like(ev)], axis=1)
    acc = (0, 0)
    for i in range(1 << nv):
        t = e_mul_py((int(eq[i, 0]), int(eq[i, 1])), (int(vals[i, 0]), int(vals[i, 1])))
        acc = ((acc[0] + t[0]) % P, (acc[1] + t[1]) % P)
    assert (int(full[0]), int(full[1])) == acc
