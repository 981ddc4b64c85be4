"""ctypes binding of include/stark_b200.h.  Fails loudly when the CUDA library is missing or no
sm_100 device is present -- there is deliberately no other execution path."""
import ctypes as C
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "libstark_b200.so")
_HEADER = os.path.join(HERE, "..", "include", "stark_b200.h")

SB_OK = 0
ERRORS = {-1: "SB_ERR_NO_DEVICE", -2: "SB_ERR_CUDA", -3: "SB_ERR_ARG", -4: "SB_ERR_ROOT", -5: "SB_ERR_OOM", -6: "SB_ERR_VERIFY"}


class StarkB200Error(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__("%s (%d)%s" % (ERRORS.get(code, "error"), code, (": " + msg) if msg else ""))


def library_path():
    return _LIB_PATH


def header_symbols():
    """every function name declared in include/stark_b200.h"""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


_lib = None


def load():
    """dlopen libstark_b200.so (no device needed for that) and declare the prototypes"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise StarkB200Error(-1, "libstark_b200.so is not built: run `python -m stark_pure_rust_b200.build` "
                                 "(there is no CPU fallback)")
    L = C.CDLL(_LIB_PATH)
    vp, u64p, u8p, sz, u32, i32 = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_size_t, C.c_uint32, C.c_int
    szp = C.POINTER(C.c_size_t)
    proto = {
        "sb_init": (i32, [i32, C.POINTER(vp)]),
        "sb_init_multi": (i32, [C.POINTER(i32), i32, C.POINTER(vp)]),
        "sb_device_count": (i32, [vp]),
        "sb_ext_create": (i32, [vp, sz, u32, u32, C.POINTER(vp)]),
        "sb_ext_free": (None, [vp, vp]),
        "sb_ext_devices": (i32, [vp]),
        "sb_ext_load": (i32, [vp, vp, sz, sz, vp, sz]),
        "sb_ext_extend": (i32, [vp, vp, sz, sz]),
        "sb_ext_commit": (i32, [vp, vp, szp, sz, vp, C.POINTER(vp)]),
        "sb_ext_fri_prove": (i32, [vp, vp, sz, vp, sz, u32, C.POINTER(vp)]),
        "sb_ext_read": (i32, [vp, vp, sz, vp]),
        "sb_destroy": (None, [vp]),
        "sb_last_error": (C.c_char_p, [vp]),
        "sb_set_stream": (i32, [vp, vp]),
        "sb_sync": (i32, [vp]),
        "sb_timer_start": (i32, [vp]),
        "sb_timer_stop": (i32, [vp, C.POINTER(C.c_float)]),
        "sb_launch_count": (C.c_uint64, [vp]),
        "sb_dev_alloc": (i32, [vp, sz, C.POINTER(vp)]),
        "sb_dev_free": (i32, [vp, vp]),
        "sb_h2d": (i32, [vp, vp, vp, sz]),
        "sb_d2h": (i32, [vp, vp, vp, sz]),
        "sb_host_alloc_pinned": (i32, [vp, sz, C.POINTER(vp)]),
        "sb_host_free_pinned": (i32, [vp, vp]),
        "sb_ntt": (i32, [vp, vp, sz, vp, u32, i32]),
        "sb_ntt_dev": (i32, [vp, vp, sz, sz, vp, sz, sz, vp, u32, i32]),
        "sb_lde_batch": (i32, [vp, vp, sz, sz, vp, u32, u32, vp]),
        "sb_lde_batch_dev": (i32, [vp, vp, sz, sz, sz, vp, u32, u32, vp]),
        "sb_powers": (i32, [vp, vp, sz, vp]),
        "sb_powers_dev": (i32, [vp, vp, sz, vp]),
        "sb_batch_inverse": (i32, [vp, vp, sz]),
        "sb_batch_inverse_dev": (i32, [vp, vp, sz]),
        "sb_merkle_commit": (i32, [vp, vp, sz, sz, vp, C.POINTER(vp)]),
        "sb_ntt_multi_dev": (i32, [vp, C.POINTER(vp), vp, u32, i32]),
        "sb_dev_alloc_on": (i32, [vp, i32, sz, C.POINTER(vp)]),
        "sb_merkle_commit_poseidon": (i32, [vp, vp, sz, sz, vp, C.POINTER(vp)]),
        "sb_poseidon_hash": (i32, [vp, vp, sz, sz, vp]),
        "sb_poseidon_hash_host": (i32, [vp, sz, vp]),
        "sb_merkle_commit_cols_dev": (i32, [vp, C.POINTER(vp), sz, sz, vp, C.POINTER(vp)]),
        "sb_merkle_open": (i32, [vp, vp, szp, sz, vp, vp]),
        "sb_tree_width": (sz, [vp]),
        "sb_tree_leaf_bytes": (sz, [vp]),
        "sb_tree_root": (i32, [vp, vp]),
        "sb_tree_free": (None, [vp, vp]),
        "sb_fri_prove": (i32, [vp, vp, sz, vp, sz, u32, C.POINTER(vp)]),
        "sb_fri_prove_dev": (i32, [vp, vp, sz, vp, sz, u32, vp, C.POINTER(vp)]),
        "sb_twiddle_mul_dev": (i32, [vp, vp, sz, sz, sz, vp, u32, i32]),
        "sb_fri_fold_dev": (i32, [vp, vp, sz, vp, vp, vp]),
        "sb_fri_n_layers": (sz, [vp]),
        "sb_fri_layer_is_last": (i32, [vp, sz]),
        "sb_fri_middle": (i32, [vp, sz, C.POINTER(vp), szp, szp, C.POINTER(vp), C.POINTER(vp), szp, szp, C.POINTER(vp), C.POINTER(vp)]),
        "sb_fri_last": (i32, [vp, sz, C.POINTER(vp), szp]),
        "sb_fri_layer_root": (i32, [vp, sz, vp]),
        "sb_fri_proof_json": (vp, [vp]),
        "sb_free_string": (None, [vp]),
        "sb_fri_proof_free": (None, [vp]),
        "sb_pseudorandom_indices": (i32, [vp, sz, u32, sz, u32, vp]),
        "sb_pseudorandom_indices_ctx": (i32, [vp, vp, sz, u32, sz, u32, vp]),
        "sb_blake2s": (None, [vp, sz, vp]),
        "sb_fp_vec_op": (i32, [vp, i32, vp, vp, vp, sz]),
        "sb_prove_r1cs": (i32, [vp, vp, C.POINTER(vp)]),
        "sb_stark_proof_roots": (i32, [vp, vp, vp, vp]),
        "sb_stark_proof_stage_ms": (i32, [vp, C.POINTER(C.c_double)]),
        "sb_stark_proof_json": (vp, [vp, szp]),
        "sb_stark_proof_free": (None, [vp]),
        "sb_verify_r1cs": (i32, [vp, vp, vp]),
        "sb_stark_proof_from_json": (i32, [C.c_char_p, sz, C.POINTER(vp)]),
        "sb_fri_verify_json": (i32, [vp, C.c_char_p, sz, vp, vp, sz, sz, u32]),
        "sb_verify_files": (i32, [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]),
        "sb_trace_from_files": (i32, [C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(vp)]),
        "sb_host_trace_free": (None, [vp]),
        "sb_prove_files": (i32, [vp, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]),
        "sb_set_extended_domain": (i32, [vp, i32]),
        "sb_pipe_peak": (i32, [vp, i32, C.POINTER(C.c_double)]),
        "sb_profile": (i32, [vp, i32]),
        "sb_profile_read": (i32, [vp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._proto = proto
    _lib = L
    return L


def _ptr(a):
    """raw pointer of a C-contiguous numpy array / bytes-like / int"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    raise TypeError(type(a))


class Context:
    """sb_ctx: one GPU (device = ordinal) or several GPUs of the node driven by one process (devices = list of ordinals,
    1 / 2 / 4 / 8 entries, the first one is the primary; an ordinal may repeat -- logical devices on one GPU).
    Raises StarkB200Error(SB_ERR_NO_DEVICE) without a B200."""

    def __init__(self, device=0, devices=None):
        self.lib = load()
        h = C.c_void_p()
        if devices is not None:
            devices = [int(d) for d in devices]
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.sb_init_multi(arr, len(devices), C.byref(h))
            if rc != SB_OK:
                raise StarkB200Error(rc, "sb_init_multi(%s) failed: 1/2/4/8 CUDA devices of compute capability 10.x with peer access are required" % devices)
            device = devices[0]
        else:
            rc = self.lib.sb_init(device, C.byref(h))
            if rc != SB_OK:
                raise StarkB200Error(rc, "sb_init(device=%d) failed: a CUDA device of compute capability 10.x is required" % device)
        self.h = h
        self.device = device
        self.devices = devices or [device]

    def close(self):
        if getattr(self, "h", None):
            self.lib.sb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != SB_OK:
            raise StarkB200Error(rc, self.lib.sb_last_error(self.h).decode(errors="replace"))

    # ---- device memory ----
    def alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.sb_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr):
        self.check(self.lib.sb_dev_free(self.h, C.c_void_p(dptr)))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.check(self.lib.sb_h2d(self.h, C.c_void_p(dptr), _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        assert arr.flags["C_CONTIGUOUS"]
        self.check(self.lib.sb_d2h(self.h, _ptr(arr), C.c_void_p(dptr), arr.nbytes))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.alloc(arr.nbytes)
        self.h2d(d, arr)
        return d

    def sync(self):
        self.check(self.lib.sb_sync(self.h))

    def set_stream(self, cuda_stream):
        self.check(self.lib.sb_set_stream(self.h, C.c_void_p(cuda_stream)))

    def timer_start(self):
        self.check(self.lib.sb_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.lib.sb_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        self.check(self.lib.sb_profile(self.h, 1 if enable else 0))

    def profile_read(self, kind):
        """(launches, total_ms) of one kernel family since profile(True)"""
        n, ms = C.c_uint64(), C.c_double()
        self.check(self.lib.sb_profile_read(self.h, kind, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def launch_count(self):
        return int(self.lib.sb_launch_count(self.h))


_default = None


def default_context():
    global _default
    if _default is None:
        _default = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default
