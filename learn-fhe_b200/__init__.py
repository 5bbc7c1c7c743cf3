"""learn_fhe_b200 — Python binding of libfhe_b200.so (the sm_100a polynomial-ring engine behind the C ABI of
include/fhe_b200.h).  The directory is named ``learn-fhe_b200``; import it through ``_pkg.load_package()`` at the
repo root (registers it as ``learn_fhe_b200``).

The host logic lives in C++ inside the shared library (the reference is compiled Rust, so the host side above the
C ABI is C++); this package only (a) builds / loads the library, (b) declares ctypes prototypes and (c) mirrors the
reference's operator names for tests and bench.py.  There is NO CPU fallback: every call goes to the CUDA library
and raises if it is missing or no B200-class GPU is present.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
# FHE_B200_LIB: developer override used for A/B builds of the same library (learn-fhe_b200/csrc/Makefile, EXTRA=...)
LIB_PATH = os.environ.get("FHE_B200_LIB") or os.path.join(_HERE, "libfhe_b200.so")
HEADER = os.path.join(REPO, "include", "fhe_b200.h")

FHE_OK, FHE_EINVAL, FHE_ECUDA, FHE_ENOMEM, FHE_EUNSUPPORTED = 0, 1, 2, 3, 4
_STATUS = {1: "FHE_EINVAL", 2: "FHE_ECUDA", 3: "FHE_ENOMEM", 4: "FHE_EUNSUPPORTED"}


class FheError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (_STATUS.get(status, status), msg))
        self.status = status


class FhewParam(C.Structure):
    """fhe_fhew_param (mirrors BootstrappingParam, scheme/fhew/src/bootstrapping.rs:21-90)."""
    _fields_ = [("log_n", C.c_uint), ("big_q", C.c_uint64), ("p", C.c_uint64), ("rlwe_log_b", C.c_uint),
                ("rlwe_d", C.c_uint), ("rgsw_log_b", C.c_uint), ("rgsw_d", C.c_uint), ("n_s", C.c_uint),
                ("q_ks", C.c_uint64), ("ks_log_b", C.c_uint), ("ks_d", C.c_uint), ("w", C.c_uint)]

    @property
    def n(self):
        return 1 << self.log_n


class CkksRot(C.Structure):
    """fhe_ckks_rot: one rotation of a BSGS plan (t = 5^j mod 2N, 0 = none)."""
    _fields_ = [("t", C.c_int64), ("key", C.c_void_p)]


class TfheParam(C.Structure):
    """fhe_tfhe_param (mirrors tfhe BootstrappingParam, scheme/tfhe/src/bootstrapping.rs:14-38)."""
    _fields_ = [("log_p", C.c_uint), ("padding", C.c_uint), ("n", C.c_uint), ("ks_log_b", C.c_uint), ("ks_d", C.c_uint),
                ("log_big_n", C.c_uint), ("k", C.c_uint), ("bs_log_b", C.c_uint), ("bs_d", C.c_uint)]

    @property
    def big_n(self):
        return 1 << self.log_big_n


def header_symbols():
    """Every function the public header declares (used by the CPU-tier export test)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fhe_[a-z0-9_]+)\s*\(", src)))


def build(verbose=False):
    """Compile libfhe_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library; fails loudly if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError("libfhe_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i64, sz, ui = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int64, C.c_size_t, C.c_uint
    L.fhe_last_error.restype = C.c_char_p
    L.fhe_last_error.argtypes = [vp]
    L.fhe_version.restype = C.c_char_p
    L.fhe_launch_count.restype = u64
    L.fhe_launch_count.argtypes = [vp]
    L.fhe_sm_count.argtypes = [vp]
    L.fhe_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.fhe_ctx_destroy.argtypes = [vp]
    L.fhe_ctx_destroy.restype = None
    L.fhe_ctx_set_stream.argtypes = [vp, vp]
    L.fhe_sync.argtypes = [vp]
    L.fhe_prof_begin.argtypes = [vp]
    L.fhe_prof_end.argtypes = [vp, C.c_char_p, sz]
    L.fhe_two_adic_primes.argtypes = [ui, ui, sz, vp]
    L.fhe_diag_int32_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fhe_diag_fp64_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fhe_diag_butterfly_rate.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.fhe_keys_broadcast.argtypes = [vp, vp, C.c_int, vp, sz]
    L.fhe_fhew_key_bytes.argtypes = [vp]
    L.fhe_fhew_key_bytes.restype = sz
    L.fhe_fhew_key_broadcast.argtypes = [vp, vp, vp, C.c_int]
    L.fhe_malloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.fhe_free.argtypes = [vp, vp]
    L.fhe_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.fhe_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    for nm in ("fhe_ntt_fwd_u64", "fhe_ntt_inv_u64"):
        getattr(L, nm).argtypes = [vp, u64, ui, sz, vp]
    for nm in ("fhe_ntt_fwd_u32", "fhe_ntt_inv_u32"):
        getattr(L, nm).argtypes = [vp, u32, ui, sz, vp]
    for nm in ("fhe_ntt_fwd_rns", "fhe_ntt_inv_rns"):
        getattr(L, nm).argtypes = [vp, vp, sz, ui, sz, vp]
    for nm in ("fhe_ntt_fwd_host", "fhe_ntt_inv_host"):
        getattr(L, nm).argtypes = [vp, u64, vp, sz, sz]
    L.fhe_twiddles_host.argtypes = [vp, u64, sz, vp, vp]
    L.fhe_negacyclic_mul_u64.argtypes = [vp, u64, ui, sz, vp, vp]
    L.fhe_negacyclic_mul_host.argtypes = [vp, u64, vp, vp, sz, sz]
    for nm in ("fhe_pointwise_mul_u64", "fhe_pointwise_mac_u64", "fhe_vec_add_u64", "fhe_vec_sub_u64"):
        getattr(L, nm).argtypes = [vp, u64, sz, vp, vp, vp]
    L.fhe_vec_neg_u64.argtypes = [vp, u64, sz, vp, vp]
    L.fhe_vec_scalar_mul_u64.argtypes = [vp, u64, sz, vp, u64, vp]
    L.fhe_automorphism_u64.argtypes = [vp, u64, ui, sz, i64, vp, vp]
    L.fhe_monomial_mul_u64.argtypes = [vp, u64, ui, sz, i64, vp, vp]
    L.fhe_mod_switch_u64.argtypes = [vp, u64, u64, sz, vp, vp]
    L.fhe_mod_switch_odd_u64.argtypes = [vp, u64, u64, sz, vp, vp]
    L.fhe_decompose_zq.argtypes = [vp, u64, ui, ui, sz, vp, vp]
    L.fhe_decompose_t64.argtypes = [vp, ui, ui, sz, vp, vp]
    L.fhe_rounding_shr_t64.argtypes = [vp, ui, sz, vp, vp]
    if hasattr(L, "fhe_fft64_negacyclic_mul"):
        L.fhe_fft64_negacyclic_mul.argtypes = [vp, ui, sz, vp, vp]
        L.fhe_fft64_negacyclic_mul_host.argtypes = [vp, vp, vp, sz, sz]
    if hasattr(L, "fhe_rns_extend_bases"):
        L.fhe_rns_extend_bases.argtypes = [vp, vp, sz, vp, sz, ui, sz, vp, vp]
        L.fhe_rns_rescale_k.argtypes = [vp, vp, sz, sz, ui, sz, vp, vp]
    # FHEW
    L.fhe_fhew_key_upload.argtypes = [vp, C.POINTER(FhewParam), vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.fhe_fhew_key_free.argtypes = [vp, vp]
    L.fhe_fhew_key_free.restype = None
    L.fhe_fhew_bootstrap_batch.argtypes = [vp, vp, vp, u64, sz, vp, vp]
    L.fhe_fhew_bootstrap_batch_host.argtypes = [vp, vp, vp, u64, sz, vp, vp]
    L.fhe_fhew_prologue_batch.argtypes = [vp, vp, sz, vp, vp]
    L.fhe_lwe_key_switch_batch.argtypes = [vp, vp, sz, vp, vp]
    L.fhe_fhew_external_product.argtypes = [vp, vp, sz, vp, vp, vp]
    L.fhe_fhew_automorphism.argtypes = [vp, vp, sz, vp, vp, vp]
    L.fhe_fhew_blind_rotate_batch.argtypes = [vp, vp, vp, sz, vp, vp]
    L.fhe_fhew_key_check_error.argtypes = [vp, vp]
    L.fhe_fhew_keygen.argtypes = [vp, C.POINTER(FhewParam), u64, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.fhe_tfhe_key_serialized_size.argtypes = [vp]
    L.fhe_tfhe_key_serialized_size.restype = sz
    L.fhe_tfhe_key_serialize.argtypes = [vp, vp, vp, sz]
    L.fhe_tfhe_key_deserialize.argtypes = [vp, vp, sz, C.POINTER(vp)]
    L.fhe_fhew_key_serialized_size.argtypes = [vp]
    L.fhe_fhew_key_serialized_size.restype = sz
    L.fhe_fhew_key_serialize.argtypes = [vp, vp, vp, sz]
    L.fhe_fhew_key_deserialize.argtypes = [vp, vp, sz, C.POINTER(vp)]
    L.fhe_rgsw_internal_product.argtypes = [vp, u64, ui, ui, ui, sz, vp, vp, vp]
    # TFHE
    if hasattr(L, "fhe_tfhe_key_upload"):
        L.fhe_tfhe_key_upload.argtypes = [vp, C.POINTER(TfheParam), vp, vp, vp, C.POINTER(vp)]
        L.fhe_tfhe_key_free.argtypes = [vp, vp]
        L.fhe_tfhe_keygen.argtypes = [vp, C.POINTER(TfheParam), C.c_double, C.c_double, u64, vp, vp, vp, vp, vp, C.POINTER(vp)]
        L.fhe_tfhe_key_free.restype = None
        L.fhe_tfhe_key_set_mode.argtypes = [vp, vp, C.c_int]
        L.fhe_tfhe_key_bytes.argtypes = [vp]
        L.fhe_tfhe_key_bytes.restype = sz
        L.fhe_tfhe_key_broadcast.argtypes = [vp, vp, vp, C.c_int]
        L.fhe_tfhe_pbs_batch.argtypes = [vp, vp, vp, sz, vp, vp]
        L.fhe_tfhe_pbs_batch_host.argtypes = [vp, vp, vp, sz, vp, vp]
        L.fhe_tfhe_external_product.argtypes = [vp, vp, sz, vp, vp, vp]
        L.fhe_tfhe_cmux.argtypes = [vp, vp, sz, vp, vp, vp, vp]
        L.fhe_tfhe_blind_rotate_extract_batch.argtypes = [vp, vp, vp, sz, vp, vp]
        L.fhe_tlwe_key_switch_batch.argtypes = [vp, vp, sz, vp, vp]
    # CKKS
    if hasattr(L, "fhe_ckks_create"):
        L.fhe_ckks_create.argtypes = [vp, ui, vp, vp, sz, C.POINTER(vp)]
        L.fhe_ckks_destroy.argtypes = [vp, vp]
        L.fhe_ckks_destroy.restype = None
        L.fhe_ckks_ksk_upload.argtypes = [vp, vp, vp, C.POINTER(vp)]
        L.fhe_ckks_ksk_free.argtypes = [vp, vp]
        L.fhe_ckks_ksk_free.restype = None
        L.fhe_ckks_ksk_bytes.argtypes = [vp]
        L.fhe_ckks_ksk_bytes.restype = sz
        L.fhe_ckks_ksk_broadcast.argtypes = [vp, vp, vp, C.c_int]
        L.fhe_ckks_keygen.argtypes = [vp, vp, u64, sz, vp, vp, vp, vp]
        L.fhe_ckks_ksk_serialized_size.argtypes = [vp]
        L.fhe_ckks_ksk_serialized_size.restype = sz
        L.fhe_ckks_ksk_serialize.argtypes = [vp, vp, vp, vp, sz]
        L.fhe_ckks_ksk_deserialize.argtypes = [vp, vp, vp, sz, C.POINTER(vp)]
        L.fhe_ckks_mul_relin_rescale_batch.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp]
        L.fhe_ckks_mul_relin_rescale_batch_host.argtypes = [vp, vp, vp, sz, sz, vp, vp, vp]
        L.fhe_ckks_key_switch.argtypes = [vp, vp, vp, i64, sz, sz, vp, vp]
        L.fhe_ckks_rescale.argtypes = [vp, vp, sz, sz, vp, vp]
        L.fhe_ckks_mul_plain_rescale_batch.argtypes = [vp, vp, sz, sz, sz, vp, vp, vp]
        L.fhe_ckks_mul_mat.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp, vp, vp, vp, vp]
    _lib = L
    return L


class Context:
    """fhe_ctx wrapper: one per process/device (the reference's global twiddle caches live here)."""

    def __init__(self, device=0):
        L = lib()
        h = C.c_void_p()
        st = L.fhe_ctx_create(device, C.byref(h))
        if st != FHE_OK:
            raise FheError(st, "fhe_ctx_create(device=%d) failed: no usable sm_100-class GPU (no CPU fallback)" % device)
        self.h = h
        self.L = L
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.fhe_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ck(self, st):
        if st != FHE_OK:
            raise FheError(st, self.L.fhe_last_error(self.h).decode())

    def call(self, name, *args):
        self.ck(getattr(self.L, name)(self.h, *args))

    def sync(self):
        self.ck(self.L.fhe_sync(self.h))

    def use_torch_stream(self):
        import torch
        h = torch.cuda.current_stream(self.device).cuda_stream
        # torch's default stream has handle 0, which fhe_ctx_set_stream reads as "restore the own stream": name the
        # legacy default stream explicitly (cudaStreamLegacy == (cudaStream_t)0x1)
        self.ck(self.L.fhe_ctx_set_stream(self.h, C.c_void_p(h if h else 1)))

    def prof_begin(self):
        self.ck(self.L.fhe_prof_begin(self.h))

    def prof_end(self):
        """{"kernel": {"ms": total device ms, "launches": n}} since prof_begin (CUDA events on the context's stream)."""
        import json
        buf = C.create_string_buffer(1 << 16)
        self.ck(self.L.fhe_prof_end(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def int32_peak(self):
        """Measured IMAD / IMAD.HI / IMAD.WIDE peaks of this device, 10^12 thread-ops per second."""
        if getattr(self, "_int32_peak", None) is None:
            a, b, c = C.c_double(), C.c_double(), C.c_double()
            self.ck(self.L.fhe_diag_int32_peak(self.h, C.byref(a), C.byref(b), C.byref(c)))
            self._int32_peak = {"imad": a.value, "imad_hi": b.value, "imad_wide": c.value}
        return dict(self._int32_peak)

    def fp64_peak(self):
        """Measured DADD / DMUL / DFMA rates of this device, 10^12 thread-instructions per second."""
        if getattr(self, "_fp64_peak", None) is None:
            a, b, c = C.c_double(), C.c_double(), C.c_double()
            self.ck(self.L.fhe_diag_fp64_peak(self.h, C.byref(a), C.byref(b), C.byref(c)))
            self._fp64_peak = {"dadd": a.value, "dmul": b.value, "dfma": c.value}
        return dict(self._fp64_peak)

    def butterfly_rate(self):
        """u32 radix-16 register pass, 10^12 butterflies/s: Shoup quotient by IMAD.HI vs by DFMA (FP64 pipe); `same` = identical words."""
        a, b, c = C.c_double(), C.c_double(), C.c_int()
        self.ck(self.L.fhe_diag_butterfly_rate(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"imad_hi": a.value, "dfma": b.value, "same": bool(c.value)}

    @property
    def launches(self):
        return int(self.L.fhe_launch_count(self.h))

    @property
    def sm_count(self):
        return int(self.L.fhe_sm_count(self.h))


def two_adic_primes(bits, log_n, count=1):
    """util/src/zq.rs:325-329 (host-side setup helper, computed by the library)."""
    import numpy as np
    out = np.zeros(count, dtype=np.uint64)
    st = lib().fhe_two_adic_primes(bits, log_n, count, C.c_void_p(out.ctypes.data))
    if st != FHE_OK:
        raise FheError(st, "two_adic_primes(%d, %d): not enough primes" % (bits, log_n))
    return [int(x) for x in out]


def first_two_adic_prime(bits, log_n):
    return two_adic_primes(bits, log_n, 1)[0]


def nccl_comm_ptr(dist, device):
    """ncclComm_t of torch.distributed's default NCCL process group (for fhe_keys_broadcast)."""
    import torch
    pg = dist.distributed_c10d._get_default_group()
    return C.c_void_p(pg._get_backend(torch.device(device))._comm_ptr())


def dptr(t):
    """Device pointer of a torch CUDA tensor (int64/int32 storage reinterpreted as u64/u32 words)."""
    assert t.is_cuda and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def hptr(a):
    """Host pointer of a C-contiguous numpy array."""
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def to_dev(a, device=0):
    """numpy uint64/uint32 array -> torch CUDA tensor holding the same bits."""
    import numpy as np
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        return torch.from_numpy(a.view(np.int64)).to("cuda:%d" % device)
    if a.dtype == np.uint32:
        return torch.from_numpy(a.view(np.int32)).to("cuda:%d" % device)
    return torch.from_numpy(a).to("cuda:%d" % device)


def to_host(t, dtype=None):
    """torch tensor -> numpy array of unsigned words."""
    import numpy as np
    import torch
    a = t.detach().cpu().numpy()
    if dtype is None:
        dtype = {torch.int64: np.uint64, torch.int32: np.uint32}.get(t.dtype)
    return a.view(dtype) if dtype is not None else a
