"""ctypes binding of libmchrg_b200.so (include/mchrg_b200.h).

The library is the product; this module only loads it and declares prototypes.  There is no
Python/NumPy fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCHRG_B200_LIB: alternative build of the same library (A/B timing of kernel variants); default: the in-tree one
LIB_PATH = os.environ.get("MCHRG_B200_LIB") or os.path.join(_HERE, "libmchrg_b200.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_vp = C.c_void_p

ERR_ARG, ERR_CUDA, ERR_ALLOC, ERR_MODEL = -1, -2, -3, -4


class Structure(C.Structure):
    """mchrg_b200_structure"""
    _fields_ = [("nat", C.c_int32), ("nid", C.c_int32), ("id", c_vp), ("xyz", c_vp), ("charge", C.c_double),
                ("periodic", C.c_int32 * 3), ("lattice", C.c_double * 9)]


# name -> (restype, argtypes); every symbol include/mchrg_b200.h declares
PROTOTYPES = {
    "mchrg_b200_context_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "mchrg_b200_context_destroy": (None, [c_vp]),
    "mchrg_b200_last_error": (C.c_char_p, []),
    "mchrg_b200_version": (C.c_int, []),
    "mchrg_b200_synchronize": (C.c_int, [c_vp]),
    "mchrg_b200_stream": (c_vp, [c_vp]),
    "mchrg_b200_launch_count": (C.c_uint64, [c_vp]),
    "mchrg_b200_side_stream": (c_vp, [c_vp]),
    "mchrg_b200_batch_trace": (C.c_int, [c_vp, c_vp, c_vp]),
    "mchrg_b200_comm_unique_id": (C.c_int, [c_vp]),
    "mchrg_b200_comm_init_nccl": (C.c_int, [c_vp, C.c_int32, C.c_int32, c_vp]),
    "mchrg_b200_comm_init_local": (C.c_int, [C.POINTER(c_vp), C.c_int32]),
    "mchrg_b200_comm_destroy": (C.c_int, [c_vp]),
    "mchrg_b200_comm_rank": (C.c_int32, [c_vp]),
    "mchrg_b200_comm_size": (C.c_int32, [c_vp]),
    "mchrg_b200_comm_bcast": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int32]),
    "mchrg_b200_comm_allreduce_sum": (C.c_int, [c_vp, c_vp, C.c_int64]),
    "mchrg_b200_comm_allgather": (C.c_int, [c_vp, c_vp, c_vp, C.c_int64]),
    "mchrg_b200_new_eeq_model": (C.c_int, [c_vp, C.c_int32, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_double, C.c_double,
                                           C.c_double, C.POINTER(c_vp)]),
    "mchrg_b200_new_eeq2019_model": (C.c_int, [c_vp, C.c_int32, c_vp, C.POINTER(c_vp)]),
    "mchrg_b200_new_eeqbc_model": (C.c_int, [c_vp, C.c_int32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_double, c_vp,
                                             c_vp, c_vp, C.c_double, C.c_double, C.c_double, c_vp, c_vp, C.c_double,
                                             C.POINTER(c_vp)]),
    "mchrg_b200_new_eeqbc2025_model": (C.c_int, [c_vp, C.c_int32, c_vp, c_vp, c_vp, C.POINTER(c_vp)]),
    "mchrg_b200_model_free": (None, [c_vp]),
    "mchrg_b200_model_kind": (C.c_int, [c_vp]),
    "mchrg_b200_get_coordination_number": (C.c_int, [c_vp, C.POINTER(Structure), c_vp, c_vp, c_vp]),
    "mchrg_b200_local_charge": (C.c_int, [c_vp, C.POINTER(Structure), c_vp, c_vp, c_vp]),
    "mchrg_b200_get_xvec": (C.c_int, [c_vp, C.POINTER(Structure), c_vp, c_vp, c_vp]),
    "mchrg_b200_get_xvec_derivs": (C.c_int, [c_vp, C.POINTER(Structure)] + [c_vp] * 8),
    "mchrg_b200_get_coulomb_matrix": (C.c_int, [c_vp, C.POINTER(Structure), c_vp, c_vp, c_vp]),
    "mchrg_b200_get_coulomb_derivs": (C.c_int, [c_vp, C.POINTER(Structure)] + [c_vp] * 10),
    "mchrg_b200_solve": (C.c_int, [c_vp, C.POINTER(Structure)] + [c_vp] * 12),
    "mchrg_b200_get_charges": (C.c_int, [c_vp, C.POINTER(Structure), c_vp, c_vp, c_vp]),
    "mchrg_b200_singlepoint": (C.c_int, [c_vp, C.POINTER(Structure)] + [c_vp] * 8),
    "mchrg_b200_singlepoint_rows": (C.c_int, [c_vp, C.POINTER(Structure), C.c_int32, C.c_int32] + [c_vp] * 7),
    "mchrg_b200_singlepoint_batch": (C.c_int, [c_vp, C.c_int32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mchrg_b200_dpotrf": (C.c_int, [c_vp, C.c_int64, c_vp, C.c_int64]),
    "mchrg_b200_dgemm": (C.c_int, [c_vp, C.c_char, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_vp, C.c_int64, c_vp,
                                   C.c_int64, C.c_double, c_vp, C.c_int64]),
    "mchrg_b200_dpotrs_right": (C.c_int, [c_vp, C.c_int64, C.c_int64, c_vp, C.c_int64, c_vp, C.c_int64]),
}

COMM_ID_BYTES = 128

_lib = None


def load():
    """Load the shared library (no GPU needed for loading; compute calls need one)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m multicharge_b200.build` "
                "(multicharge_b200 has no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class MchrgError(RuntimeError):
    """Raised for negative status codes (argument / CUDA / allocation / model errors)."""

    def __init__(self, code, message):
        super().__init__(f"mchrg_b200 error {code}: {message}")
        self.code = code


def check(status):
    if status < 0:
        raise MchrgError(status, load().mchrg_b200_last_error().decode(errors="replace"))
    return status
