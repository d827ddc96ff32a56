"""ctypes view of oracle/libdp_oracle.so -- TEST INFRASTRUCTURE.  Imported only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
import ctypes as C
import os
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(_ROOT, "oracle", "libdp_oracle.so")
P = 0xFFFFFFFF00000001
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_PATH):
            raise ImportError("oracle not built: run `make oracle`")
        _lib = C.CDLL(ORACLE_PATH)
        _lib.dpo_last_error.restype = C.c_char_p
        _lib.dpo_transcript_new.restype = C.c_void_p
        _lib.dpo_transcript_new.argtypes = [C.c_char_p]
        for name in ("dpo_transcript_free",):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.dpo_transcript_append_f.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.dpo_transcript_append_msg.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        _lib.dpo_transcript_append_e.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.dpo_transcript_challenge.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        _lib.dpo_transcript_read_challenge.argtypes = [C.c_void_p, C.c_void_p]
    return _lib


def u64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint64))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def set_hash_mode(kind):
    """0 = Poseidon2 + BasicTranscript (default), 1 = BLAKE3 + BlakeTranscript (the reference's `blake` feature); process-wide"""
    lib().dpo_set_hash_mode(int(kind))


def splitmix_f(seed, n):
    out = np.empty(n, dtype=np.uint64)
    lib().dpo_splitmix_f(C.c_uint64(seed), C.c_uint64(n), ptr(out))
    return out


def splitmix_e(seed, n):
    return splitmix_f(seed, 2 * n).reshape(n, 2)


def f_binop(op, a, b):
    a, b = u64(a), u64(b)
    out = np.empty_like(a)
    lib().dpo_f_binop(op, ptr(a), ptr(b), C.c_uint64(a.size), ptr(out))
    return out


def e_binop(op, a, b):
    a, b = u64(a).reshape(-1, 2), u64(b).reshape(-1, 2)
    out = np.empty_like(a)
    lib().dpo_e_binop(op, ptr(a), ptr(b), C.c_uint64(a.shape[0]), ptr(out))
    return out


def e_inv(a):
    a = u64(a).reshape(-1, 2)
    out = np.empty_like(a)
    lib().dpo_e_inv(ptr(a), C.c_uint64(a.shape[0]), ptr(out))
    return out


def _mle_call(fn, evals, is_ext, point, out_len):
    ev = u64(evals).reshape(-1)
    n = ev.size // (2 if is_ext else 1)
    pt = u64(point).reshape(-1)
    out = np.zeros((out_len, 2), dtype=np.uint64)
    rc = fn(ptr(ev), C.c_uint64(n), int(bool(is_ext)), ptr(pt), C.c_uint32(pt.size // 2), ptr(out))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return out


def fix_high(evals, is_ext, point):
    ev = u64(evals).reshape(-1)
    n = ev.size // (2 if is_ext else 1)
    k = u64(point).size // 2
    return _mle_call(lib().dpo_fix_high, evals, is_ext, point, n >> k)


def fix_low(evals, is_ext, point):
    ev = u64(evals).reshape(-1)
    n = ev.size // (2 if is_ext else 1)
    k = u64(point).size // 2
    return _mle_call(lib().dpo_fix_low, evals, is_ext, point, n >> k)


def evaluate(evals, is_ext, point):
    return _mle_call(lib().dpo_evaluate, evals, is_ext, point, 1)[0]


def build_eq(point):
    pt = u64(point).reshape(-1)
    nv = pt.size // 2
    out = np.zeros((1 << nv, 2), dtype=np.uint64)
    lib().dpo_build_eq(ptr(pt), C.c_uint32(nv), ptr(out))
    return out


def eq_eval(x, y):
    x, y = u64(x).reshape(-1), u64(y).reshape(-1)
    out = np.zeros(2, dtype=np.uint64)
    lib().dpo_eq_eval(ptr(x), ptr(y), C.c_uint32(x.size // 2), ptr(out))
    return out


def poseidon2_permute(state):
    s = u64(state).copy()
    lib().dpo_poseidon2_permute(ptr(s))
    return s


def compress(x, y):
    x, y = u64(x), u64(y)
    out = np.zeros(4, dtype=np.uint64)
    lib().dpo_compress(ptr(x), ptr(y), ptr(out))
    return out


def hash_or_noop(v):
    v = u64(v)
    out = np.zeros(4, dtype=np.uint64)
    lib().dpo_hash_or_noop(ptr(v), C.c_uint64(v.size), ptr(out))
    return out


class Transcript:
    def __init__(self, label=b"m2vec"):
        self.h = C.c_void_p(lib().dpo_transcript_new(label))

    def append_f(self, f):
        f = u64(f)
        lib().dpo_transcript_append_f(self.h, ptr(f), C.c_uint64(f.size))

    def append_msg(self, m):
        lib().dpo_transcript_append_msg(self.h, m, C.c_uint64(len(m)))

    def append_e(self, e):
        e = u64(e).reshape(-1, 2)
        lib().dpo_transcript_append_e(self.h, ptr(e), C.c_uint64(e.shape[0]))

    def challenge(self, label):
        out = np.zeros(2, dtype=np.uint64)
        lib().dpo_transcript_challenge(self.h, label, ptr(out))
        return out

    def read_challenge(self):
        out = np.zeros(2, dtype=np.uint64)
        lib().dpo_transcript_read_challenge(self.h, ptr(out))
        return out

    def __del__(self):
        try:
            lib().dpo_transcript_free(self.h)
        except Exception:
            pass


def _vp_args(mles, products):
    """mles: list of (np array, is_ext); products: list of (coef(c0,c1), [idx])"""
    arrs = [u64(m[0]).reshape(-1) for m in mles]
    n = len(arrs)
    data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = u64([a.size // (2 if m[1] else 1) for a, m in zip(arrs, mles)])
    is_ext = np.ascontiguousarray(np.asarray([int(bool(m[1])) for m in mles], dtype=np.int32))
    coefs = u64([c for p in products for c in p[0]])
    deg = np.ascontiguousarray(np.asarray([len(p[1]) for p in products], dtype=np.uint32))
    idx = np.ascontiguousarray(np.asarray([i for p in products for i in p[1]], dtype=np.uint32))
    return arrs, data, lens, is_ext, coefs, deg, idx


def sumcheck_prove(mles, products, max_nv, label=b"m2vec"):
    arrs, data, lens, is_ext, coefs, deg, idx = _vp_args(mles, products)
    max_deg = int(deg.max())
    point = np.zeros((max_nv, 2), dtype=np.uint64)
    msgs = np.zeros((max_nv, max_deg + 1, 2), dtype=np.uint64)
    fin = np.zeros((len(mles), 2), dtype=np.uint64)
    md = C.c_uint32()
    rc = lib().dpo_sumcheck_prove(C.c_uint32(len(mles)), data, ptr(lens), ptr(is_ext), C.c_uint32(len(products)), ptr(coefs),
                                  ptr(deg), ptr(idx), C.c_uint32(max_nv), label, ptr(point), ptr(msgs), ptr(fin), C.byref(md))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return point, msgs, fin


def sumcheck_rounds_fixed(mles, products, max_nv, challenges):
    arrs, data, lens, is_ext, coefs, deg, idx = _vp_args(mles, products)
    max_deg = int(deg.max())
    ch = u64(challenges).reshape(-1)
    msgs = np.zeros((max_nv, max_deg + 1, 2), dtype=np.uint64)
    fin = np.zeros((len(mles), 2), dtype=np.uint64)
    rc = lib().dpo_sumcheck_rounds_fixed(C.c_uint32(len(mles)), data, ptr(lens), ptr(is_ext), C.c_uint32(len(products)),
                                         ptr(coefs), ptr(deg), ptr(idx), C.c_uint32(max_nv), ptr(ch), ptr(msgs), ptr(fin))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return msgs, fin


def sumcheck_verify(claimed_sum, nv, max_deg, msgs, label=b"m2vec"):
    cs = u64(claimed_sum)
    m = u64(msgs).reshape(-1)
    point = np.zeros((nv, 2), dtype=np.uint64)
    exp = np.zeros(2, dtype=np.uint64)
    rc = lib().dpo_sumcheck_verify(ptr(cs), C.c_uint32(nv), C.c_uint32(max_deg), ptr(m), label, ptr(point), ptr(exp))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return point, exp


# ---- tiny pure-Python field helpers for naive cross-checks ----
def pf_mul(a, b):
    return (int(a) * int(b)) % P


def pe_mul(a, b):
    a0, a1, b0, b1 = int(a[0]), int(a[1]), int(b[0]), int(b[1])
    return ((a0 * b0 + 7 * a1 * b1) % P, (a0 * b1 + a1 * b0) % P)


def pe_add(a, b):
    return ((int(a[0]) + int(b[0])) % P, (int(a[1]) + int(b[1])) % P)


def pe_sub(a, b):
    return ((int(a[0]) - int(b[0])) % P, (int(a[1]) - int(b[1])) % P)


# ---- Basefold ----
def _fv(a, is_ext):
    a = u64(a).reshape(-1)
    return a, a.size // (2 if is_ext else 1)


def rs_encode(coeffs, is_ext, full_log):
    a, n = _fv(coeffs, is_ext)
    out = np.zeros(2 * a.size, dtype=np.uint64)
    lib().dpo_rs_encode(ptr(a), C.c_uint64(n), int(is_ext), C.c_uint32(full_log), ptr(out))
    return out.reshape(-1, 2) if is_ext else out


def interpolate_hc(evals, is_ext):
    a, n = _fv(evals, is_ext)
    out = np.zeros_like(a)
    lib().dpo_interpolate_hc(ptr(a), C.c_uint64(n), int(is_ext), ptr(out))
    return out.reshape(-1, 2) if is_ext else out


def merkle_root(leaves, is_ext):
    a, n = _fv(leaves, is_ext)
    out = np.zeros(4, dtype=np.uint64)
    lib().dpo_merkle_root(ptr(a), C.c_uint64(n), int(is_ext), ptr(out))
    return out


def folding_coeffs(full_log, level, index):
    x0, w = C.c_uint64(), C.c_uint64()
    lib().dpo_folding_coeffs(C.c_uint32(full_log), C.c_uint32(level), C.c_uint64(index), C.byref(x0), C.byref(w))
    return x0.value, w.value


def fri_fold(vals, full_log, r):
    a = u64(vals).reshape(-1)
    n = a.size // 2
    rr = u64(r)
    out = np.zeros((n // 2, 2), dtype=np.uint64)
    lib().dpo_fri_fold(ptr(a), C.c_uint64(n), C.c_uint32(full_log), ptr(rr), ptr(out))
    return out


def pcs_commit(evals, is_ext, full_log, want_codeword=True):
    a, n = _fv(evals, is_ext)
    root = np.zeros(4, dtype=np.uint64)
    nv = n.bit_length() - 1
    trivial = nv <= 7
    lim = 2 if is_ext else 1
    cw = np.zeros((n if trivial else 2 * n) * lim, dtype=np.uint64)
    bh = np.zeros(n * lim, dtype=np.uint64)
    rc = lib().dpo_pcs_commit(ptr(a), C.c_uint64(n), int(is_ext), C.c_uint32(full_log), ptr(root), ptr(cw) if want_codeword else None,
                              ptr(bh) if want_codeword else None)
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    if is_ext:
        cw, bh = cw.reshape(-1, 2), bh.reshape(-1, 2)
    return root, cw, bh


def pcs_open(evals, is_ext, full_log, point, label=b"m2vec", cap=1 << 24):
    a, n = _fv(evals, is_ext)
    p = u64(point).reshape(-1)
    out = np.zeros(cap, dtype=np.uint64)
    ln = C.c_uint64()
    rc = lib().dpo_pcs_open(ptr(a), C.c_uint64(n), int(is_ext), C.c_uint32(full_log), ptr(p), label, ptr(out), C.c_uint64(cap), C.byref(ln))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return out[: ln.value].copy()


def pcs_batch_open(polys, full_log, points, label=b"m2vec", cap=1 << 25):
    """polys: list of (array, is_ext); points: list of (nv_i, 2) arrays"""
    arrs = [u64(p[0]).reshape(-1) for p in polys]
    n = len(arrs)
    data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = u64([a.size // (2 if p[1] else 1) for a, p in zip(arrs, polys)])
    is_ext = np.ascontiguousarray(np.asarray([int(bool(p[1])) for p in polys], dtype=np.int32))
    pts = np.concatenate([u64(x).reshape(-1) for x in points])
    out = np.zeros(cap, dtype=np.uint64)
    ln = C.c_uint64()
    rc = lib().dpo_pcs_batch_open(C.c_uint32(n), data, ptr(lens), ptr(is_ext), C.c_uint32(full_log), ptr(pts), label, ptr(out),
                                  C.c_uint64(cap), C.byref(ln))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return out[: ln.value].copy()


# ---- zkml synthetic MLP ----
def synthetic_mlp(n_layers, width, seed):
    w = np.zeros(n_layers * width * width, dtype=np.int64)
    b = np.zeros(n_layers * width, dtype=np.int64)
    rq = np.zeros(n_layers * 4, dtype=np.int64)
    lib().dpo_synthetic_mlp(C.c_uint32(n_layers), C.c_uint32(width), C.c_uint64(seed), ptr(w), ptr(b), ptr(rq))
    return w, b, rq.reshape(n_layers, 4)


def synthetic_input(width, seed):
    x = np.zeros(width, dtype=np.int64)
    lib().dpo_synthetic_input(C.c_uint32(width), C.c_uint64(seed), ptr(x))
    return x


def zkml_prove(n_layers, width, seed_model, seed_input, label=b"m2vec", cap=1 << 22, want_proof=True):
    out = np.zeros(cap if want_proof else 1, dtype=np.uint64)
    n = C.c_uint64()
    ms = (C.c_double * 2)()
    rc = lib().dpo_zkml_prove(C.c_uint32(n_layers), C.c_uint32(width), C.c_uint64(seed_model), C.c_uint64(seed_input), label,
                              ptr(out) if want_proof else None, C.c_uint64(cap), C.byref(n), ms)
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return (out[: n.value].copy() if want_proof else None), (ms[0], ms[1])


def sumcheck_prove_batch(T, mles, products, max_nv, label=b"m2vec"):
    """prove_batch_polys over T contiguous slices of the given full MLEs"""
    arrs, data, lens, is_ext, coefs, deg, idx = _vp_args(mles, products)
    max_deg = int(deg.max())
    point = np.zeros((max_nv, 2), dtype=np.uint64)
    msgs = np.zeros((max_nv, max_deg + 1, 2), dtype=np.uint64)
    fin = np.zeros((len(mles), 2), dtype=np.uint64)
    rc = lib().dpo_sumcheck_prove_batch(C.c_uint32(T), C.c_uint32(len(mles)), data, ptr(lens), ptr(is_ext), C.c_uint32(len(products)),
                                        ptr(coefs), ptr(deg), ptr(idx), C.c_uint32(max_nv), label, ptr(point), ptr(msgs), ptr(fin))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return point, msgs, fin


# ---- FFT convolution layer (oracle/conv.hpp) ----
def i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def synthetic_conv(kw, kx, n_x, real_nw, kw_u, k_u, kx_u, n_x_u, seed_model, seed_input):
    filt = np.zeros((kw, kx, real_nw, real_nw), dtype=np.int64)
    bias = np.zeros(kw, dtype=np.int64)
    x = np.zeros((kx, n_x, n_x), dtype=np.int64)
    lib().dpo_synthetic_conv(C.c_uint32(kw), C.c_uint32(kx), C.c_uint32(n_x), C.c_uint32(real_nw), C.c_uint32(kw_u), C.c_uint32(k_u), C.c_uint32(kx_u),
                             C.c_uint32(n_x_u), C.c_uint64(seed_model), C.c_uint64(seed_input), ptr(filt), ptr(bias), ptr(x))
    unpadded_out = np.asarray([kw_u, n_x_u - k_u + 1, n_x_u - k_u + 1], dtype=np.uint32)
    return filt, bias, x, unpadded_out


def conv_op(filt, bias, unpadded_out, x):
    kw, kx, rn, _ = filt.shape
    n_x = x.shape[1]
    after = np.zeros((kw, n_x, n_x), dtype=np.int64)
    cleared = np.zeros((kw, n_x, n_x), dtype=np.int64)
    rc = lib().dpo_conv_op(C.c_uint32(kw), C.c_uint32(kx), C.c_uint32(n_x), C.c_uint32(rn), ptr(i64(filt)), ptr(i64(bias)), ptr(np.ascontiguousarray(unpadded_out, dtype=np.uint32)),
                           ptr(i64(x)), ptr(after), ptr(cleared))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return after, cleared


def conv_prove(filt, bias, unpadded_out, x, label=b"m2vec", cap=1 << 20):
    kw, kx, rn, _ = filt.shape
    n_x = x.shape[1]
    out = np.zeros(cap, dtype=np.uint64)
    n = C.c_uint64()
    rc = lib().dpo_conv_prove(C.c_uint32(kw), C.c_uint32(kx), C.c_uint32(n_x), C.c_uint32(rn), ptr(i64(filt)), ptr(i64(bias)), ptr(np.ascontiguousarray(unpadded_out, dtype=np.uint32)),
                              ptr(i64(x)), label, ptr(out), C.c_uint64(cap), C.byref(n))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return out[:n.value].copy()


def fft_ext(rows, inverse=False):
    a = u64(rows).copy()
    r, n = a.shape[0], a.shape[1]
    lib().dpo_fft_ext(ptr(a), C.c_uint64(r), C.c_uint64(n), C.c_int(int(inverse)))
    return a


def cnn_prove(small, seed_model, seed_input, label=b"m2vec", cap=1 << 23, want_proof=True):
    out = np.zeros(cap if want_proof else 1, dtype=np.uint64)
    n = C.c_uint64()
    ms = (C.c_double * 2)()
    rc = lib().dpo_cnn_prove(C.c_int(int(small)), C.c_uint64(seed_model), C.c_uint64(seed_input), label, ptr(out) if want_proof else None, C.c_uint64(cap), C.byref(n), ms)
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return (out[:n.value].copy() if want_proof else None), (ms[0], ms[1])


def synthetic_cnn(small, seed_model, seed_input):
    dl, wl, il = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib().dpo_synthetic_cnn(C.c_int(int(small)), C.c_uint64(seed_model), C.c_uint64(seed_input), None, C.c_uint64(0), C.byref(dl), None, C.c_uint64(0), C.byref(wl), None, C.byref(il))
    desc = np.zeros(dl.value, dtype=np.int64); data = np.zeros(wl.value, dtype=np.int64); inp = np.zeros(il.value, dtype=np.int64)
    rc = lib().dpo_synthetic_cnn(C.c_int(int(small)), C.c_uint64(seed_model), C.c_uint64(seed_input), ptr(desc), C.c_uint64(desc.size), C.byref(dl), ptr(data), C.c_uint64(data.size), C.byref(wl), ptr(inp), C.byref(il))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return desc.reshape(-1, 9), data, inp


def model_prove(desc, data, x, label=b"m2vec", cap=1 << 23, want_proof=True):
    """Prover::prove of a model given as a layer descriptor (deep-prove_b200/models.py); returns (flat | None, (ctx_ms, prove_ms))"""
    d = i64(desc).reshape(-1, 9); w = i64(data); xi = i64(x)
    out = np.zeros(cap if want_proof else 1, dtype=np.uint64)
    n = C.c_uint64(); ms = (C.c_double * 2)()
    rc = lib().dpo_model_prove(ptr(d), C.c_uint32(d.shape[0]), ptr(w), ptr(xi), C.c_uint64(xi.size), label, ptr(out) if want_proof else None, C.c_uint64(cap), C.byref(n), ms)
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return (out[:n.value].copy() if want_proof else None), (ms[0], ms[1])


def pcs_verify(flat, root, num_vars, is_base, full_log, point, eval_, label=b"m2vec"):
    """Basefold::verify on a flat proof image; returns None when accepted, else the rejection reason"""
    f = u64(flat); r = u64(root); pt = u64(point).reshape(-1); ev = u64(eval_).reshape(-1)
    rc = lib().dpo_pcs_verify(ptr(f), C.c_uint64(f.size), ptr(r), C.c_uint32(num_vars), C.c_int(int(is_base)), C.c_uint32(full_log), ptr(pt), ptr(ev), label)
    return None if rc == 0 else lib().dpo_last_error().decode()


def pcs_batch_verify(flat, roots, num_vars, is_base, full_log, points, evals, label=b"m2vec"):
    f = u64(flat); r = u64(roots).reshape(-1); nv = np.ascontiguousarray(num_vars, dtype=np.uint32); ib = np.ascontiguousarray(is_base, dtype=np.int32)
    pt = u64(np.concatenate([u64(p).reshape(-1) for p in points])); ev = u64(evals).reshape(-1)
    rc = lib().dpo_pcs_batch_verify(ptr(f), C.c_uint64(f.size), C.c_uint32(len(nv)), ptr(r), ptr(nv), ptr(ib), C.c_uint32(full_log), ptr(pt), ptr(ev), label)
    return None if rc == 0 else lib().dpo_last_error().decode()


def zkml_prove_verify(n_layers, width, seed_model, seed_input, label=b"m2vec", tamper=0):
    """prove on the checker, verify on a fresh transcript with the restated model verifier; None = accepted, else the reason"""
    rc = lib().dpo_zkml_prove_verify(C.c_uint32(n_layers), C.c_uint32(width), C.c_uint64(seed_model), C.c_uint64(seed_input), label, C.c_int(tamper))
    return None if rc == 0 else lib().dpo_last_error().decode()


def model_prove_verify(desc, data, x, label=b"m2vec", tamper=0):
    d = i64(desc).reshape(-1, 9); w = i64(data); xi = i64(x)
    rc = lib().dpo_model_prove_verify(ptr(d), C.c_uint32(d.shape[0]), ptr(w), ptr(xi), C.c_uint64(xi.size), label, C.c_int(tamper))
    return None if rc == 0 else lib().dpo_last_error().decode()


def pcs_simple_batch(polys, is_ext, full_log, point=None, label=b"m2vec", cap=1 << 24):
    """batch_commit of same-size polynomials (+ simple_batch_open at `point`): returns (root, evals | None, flat | None)"""
    arrs = [u64(p).reshape(-1) for p in polys]
    n = len(arrs); ln = arrs[0].size // (2 if is_ext else 1)
    data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    root = np.zeros(4, dtype=np.uint64); evals = np.zeros((n, 2), dtype=np.uint64)
    out = np.zeros(cap if point is not None else 1, dtype=np.uint64); ol = C.c_uint64()
    pt = u64(point).reshape(-1) if point is not None else None
    rc = lib().dpo_pcs_simple_batch(C.c_uint32(n), data, C.c_uint64(ln), C.c_int(int(is_ext)), C.c_uint32(full_log), ptr(pt) if pt is not None else None, label,
                                    ptr(root), ptr(evals), ptr(out), C.c_uint64(cap), C.byref(ol))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return root, (evals if point is not None else None), (out[:ol.value].copy() if point is not None else None)


def pcs_simple_batch_verify(flat, root, num_vars, is_base, n_polys, full_log, point, evals, label=b"m2vec"):
    f = u64(flat); r = u64(root); pt = u64(point).reshape(-1); ev = u64(evals).reshape(-1)
    rc = lib().dpo_pcs_simple_batch_verify(ptr(f), C.c_uint64(f.size), ptr(r), C.c_uint32(num_vars), C.c_int(int(is_base)), C.c_uint32(n_polys), C.c_uint32(full_log), ptr(pt), ptr(ev), label)
    return None if rc == 0 else lib().dpo_last_error().decode()


def pcs_batch_open_evals(polys, full_log, points, eval_poly, eval_point, label=b"m2vec", cap=1 << 25, verify=True):
    """batch_open with an explicit (poly, point) evaluation list, verified by the restated batch_verify when `verify`"""
    arrs = [u64(p[0]).reshape(-1) for p in polys]
    n = len(arrs)
    data = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = u64([a.size // (2 if p[1] else 1) for a, p in zip(arrs, polys)])
    ie = np.ascontiguousarray(np.asarray([int(bool(p[1])) for p in polys], dtype=np.int32))
    pts = u64(np.concatenate([u64(x).reshape(-1) for x in points]))
    pnv = np.ascontiguousarray([u64(x).reshape(-1, 2).shape[0] for x in points], dtype=np.uint32)
    ep = np.ascontiguousarray(eval_poly, dtype=np.uint32); eq = np.ascontiguousarray(eval_point, dtype=np.uint32)
    roots = np.zeros((n, 4), dtype=np.uint64); vals = np.zeros((len(ep), 2), dtype=np.uint64)
    out = np.zeros(cap, dtype=np.uint64); ol = C.c_uint64()
    rc = lib().dpo_pcs_batch_open_evals(C.c_uint32(n), data, ptr(lens), ptr(ie), C.c_uint32(full_log), ptr(pts), ptr(pnv), C.c_uint32(len(pnv)), ptr(ep), ptr(eq), C.c_uint32(len(ep)),
                                        label, ptr(roots), ptr(vals), ptr(out), C.c_uint64(cap), C.byref(ol), C.c_int(int(verify)))
    if rc:
        raise RuntimeError(lib().dpo_last_error().decode())
    return out[:ol.value].copy(), roots, vals
