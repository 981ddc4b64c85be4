"""Mirror of packages/fri/src/fri.rs::prove_low_degree on the GPU backend."""
import ctypes as C

import numpy as np

from .fft import _as_elems, _root_limbs
from ._lib import _ptr, default_context
from .merkle import Proof


def _read(ptr, n):
    return C.string_at(ptr, n) if n else b""


def prove_low_degree(values, root_of_unity, max_deg_plus_1, exclude_multiples_of, ctx=None, as_json=False):
    """fri.rs:46-62.  Returns Vec<FriProof> as a list of
    {"Middle": {"root2", "column_branches", "poly_branches"}} / {"Last": {"last"}} (fri.rs:16-26), or the
    serde_json text when as_json=True."""
    ctx = ctx or default_context()
    v = _as_elems(values)
    root = _root_limbs(root_of_unity)
    h = C.c_void_p()
    ctx.check(ctx.lib.sb_fri_prove(ctx.h, _ptr(v), v.shape[0], _ptr(root), max_deg_plus_1, exclude_multiples_of, C.byref(h)))
    try:
        return fri_json(ctx, h) if as_json else unpack_proof(ctx, h)
    finally:
        ctx.lib.sb_fri_proof_free(h)


def fri_json(ctx, h):
    s = ctx.lib.sb_fri_proof_json(h)
    try:
        return C.string_at(s).decode()
    finally:
        ctx.lib.sb_free_string(s)


def layer_roots(ctx, h):
    out = []
    for i in range(ctx.lib.sb_fri_n_layers(h)):
        if not ctx.lib.sb_fri_layer_is_last(h, i):
            r = np.empty(32, dtype=np.uint8)
            ctx.check(ctx.lib.sb_fri_layer_root(h, i, _ptr(r)))
            out.append(r.tobytes())
    return out


def unpack_proof(ctx, h):
    lib = ctx.lib
    out = []
    for i in range(lib.sb_fri_n_layers(h)):
        if lib.sb_fri_layer_is_last(h, i):
            p, n = C.c_void_p(), C.c_size_t()
            ctx.check(lib.sb_fri_last(h, i, C.byref(p), C.byref(n)))
            raw = _read(p, n.value * 32)
            out.append({"Last": {"last": [raw[k * 32:(k + 1) * 32] for k in range(n.value)]}})
            continue
        root2, cl, cn, pl, pn = (C.c_void_p() for _ in range(5))
        nc, dc, npo, dp = (C.c_size_t() for _ in range(4))
        ctx.check(lib.sb_fri_middle(h, i, C.byref(root2), C.byref(nc), C.byref(dc), C.byref(cl), C.byref(cn),
                                    C.byref(npo), C.byref(dp), C.byref(pl), C.byref(pn)))

        def branches(leaves, nodes, count, depth):
            lv, nd = _read(leaves, count * 32), _read(nodes, count * depth * 32)
            return [Proof(lv[q * 32:(q + 1) * 32], [nd[(q * depth + l) * 32:(q * depth + l + 1) * 32] for l in range(depth)])
                    for q in range(count)]

        out.append({"Middle": {"root2": _read(root2, 32),
                               "column_branches": branches(cl, cn, nc.value, dc.value),
                               "poly_branches": branches(pl, pn, npo.value, dp.value)}})
    return out


def verify_low_degree_proof(merkle_root, root_of_unity, proof_json, max_deg_plus_1, exclude_multiples_of, n, ctx=None):
    """fri.rs:226-404 on the serde text of Vec<FriProof>.  Host-side check (no GPU needed when ctx is None).  Returns True,
    or raises StarkB200Error(SB_ERR_VERIFY) where the reference asserts."""
    from ._lib import StarkB200Error, load
    lib = ctx.lib if ctx is not None else load()
    text = proof_json.encode() if isinstance(proof_json, str) else bytes(proof_json)
    root = _root_limbs(root_of_unity)
    mr = np.frombuffer(bytes(merkle_root), dtype=np.uint8).copy()
    rc = lib.sb_fri_verify_json(ctx.h if ctx is not None else None, text, len(text), _ptr(mr), _ptr(root), n, max_deg_plus_1, exclude_multiples_of)
    if rc != 0:
        raise StarkB200Error(rc, lib.sb_last_error(ctx.h).decode(errors="replace") if ctx is not None else "FRI proof rejected")
    return True
