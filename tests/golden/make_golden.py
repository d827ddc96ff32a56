#!/usr/bin/env python3
"""Regenerates tests/golden/checker_regression.json: SHA-256 digests of the CPU checker's outputs on fixed seeded inputs.

These are REGRESSION vectors for the checker itself (the reference is Rust and cannot run here, so no vector in this file comes
from deep-prove): they freeze what rounds of this repo agreed on bit for bit -- device == checker is tested elsewhere, so a
change in any digest means the restated algorithm changed and must be explained.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "deep-prove_b200"))
import oracle_py as O  # noqa: E402
import models  # noqa: E402


def h(*arrays):
    m = hashlib.sha256()
    for a in arrays:
        m.update(np.ascontiguousarray(a, dtype=np.uint64).tobytes())
    return m.hexdigest()


def vectors():
    out = {}
    out["poseidon2_permute_zeros"] = [int(x) for x in O.poseidon2_permute(np.zeros(8, dtype=np.uint64))]
    out["poseidon2_permute_0_to_7"] = [int(x) for x in O.poseidon2_permute(np.arange(8, dtype=np.uint64))]
    out["compress_1234_5678"] = [int(x) for x in O.compress(np.array([1, 2, 3, 4], dtype=np.uint64), np.array([5, 6, 7, 8], dtype=np.uint64))]
    t = O.Transcript(b"m2vec"); t.append_e(np.array([[1, 2], [3, 4]], dtype=np.uint64))
    out["transcript_challenge"] = [int(x) for x in t.challenge(b"Internal round")]
    nv = 8
    mles = [(O.splitmix_f(1, 1 << nv), False), (O.splitmix_e(2, 1 << nv), True), (O.splitmix_f(3, 1 << nv), False)]
    out["sumcheck_nv8_deg3"] = h(*O.sumcheck_prove(mles, [((1, 0), [0, 1, 2]), ((5, 7), [1, 2])], nv))
    ev = O.splitmix_f(7, 1 << 10)
    root, cw, bh = O.pcs_commit(ev, False, 10)
    out["basefold_commit_root_nv10"] = [int(x) for x in root]
    out["basefold_open_nv10"] = h(O.pcs_open(ev, False, 10, O.splitmix_e(8, 10)))
    polys = [(O.splitmix_f(400 + i, 1 << nvp), False) for i, nvp in enumerate((10, 8, 9))]
    out["basefold_batch_open_10_8_9"] = h(O.pcs_batch_open(polys, 10, [O.splitmix_e(500 + i, nvp) for i, nvp in enumerate((10, 8, 9))]))
    sb = [O.splitmix_f(900 + i, 1 << 9) for i in range(3)]
    r, evs, flat = O.pcs_simple_batch(sb, False, 10, O.splitmix_e(950, 9))
    out["simple_batch_root_3x_nv9"] = [int(x) for x in r]
    out["simple_batch_open_3x_nv9"] = h(flat)
    out["zkml_mlp_2x64"] = h(O.zkml_prove(2, 64, 5, 6)[0])
    filt, bias, x, uo = O.synthetic_conv(2, 2, 8, 4, 2, 3, 2, 7, 15, 16)
    out["conv_layer_proof_2x2x8"] = h(O.conv_prove(filt, bias, uo, x))
    d, w, xin, _ = models.cnn_small(seed=3)
    out["cnn_small_proof"] = h(O.model_prove(d, w, xin)[0])
    return out


if __name__ == "__main__":
    v = vectors()
    with open(os.path.join(HERE, "checker_regression.json"), "w") as f:
        json.dump(v, f, indent=1, sort_keys=True)
    print("wrote", len(v), "vectors")
