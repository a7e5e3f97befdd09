"""galcamb-b200: sm_100a implementation of the scalar CMB transfer path of the Galileon fork of CAMB.

Host-side mirror of the reference's call sequence (camb/cmbmain.f90:131-291) over the C ABI of
include/galcamb.h.  The product is libgalcamb_b200.so; this package only loads it, marshals numpy arrays and
shards batches across ranks.  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from .camb_gpu import CParams, GalCamb, ModelTables, lib_path, load_library, make_params, prestage  # noqa: F401
from .cl_files import write_lensed_cls, write_scalar_cls  # noqa: F401

__all__ = ["GalCamb", "ModelTables", "CParams", "make_params", "prestage", "lib_path", "load_library", "write_scalar_cls", "write_lensed_cls"]
