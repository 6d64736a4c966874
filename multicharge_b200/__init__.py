"""multicharge_b200 -- B200-native (sm_100a CUDA) drop-in for the EEQ charge hot path of
grimme-lab/multicharge.  Public names mirror ``module multicharge`` (src/multicharge.f90:16-27).

Everything computes inside libmchrg_b200.so (C ABI: include/mchrg_b200.h); there is no CPU
fallback.  See DESIGN.md / INTEGRATION.md.
"""
from ._lib import MchrgError  # noqa: F401
from .api import (Context, FactorizationError, MchrgModel, Structure, default_context,  # noqa: F401
                  get_charges, get_eeq_charges, get_eeqbc_charges, new_eeq2019_batch_model, new_eeqbc2025_model,
                  new_eeqbc_model, new_eeq2019_model, new_eeq_model,
                  singlepoint_batch, singlepoint_rows)
from .dist import init_comm_from_torch  # noqa: F401

mchrg_model_type = MchrgModel  # reference type name


class mchrg_model:  # param.f90:34-42
    eeq2019 = 1
    eeqbc2025 = 2


def get_multicharge_version():
    from . import _lib
    v = _lib.load().mchrg_b200_version()
    return f"{v // 10000}.{(v // 100) % 100}.{v % 100}"
