"""wps_b200 -- B200-native horizontal-interpolation path of WPS geogrid/metgrid.

The compute path is libwpsgpu.so (hand-written CUDA for sm_100a behind the C ABI in
include/wpsgpu.h).  This package is the Python host-side mirror of the reference's
interface for that path; it contains no CPU implementation of any operator.
"""
from . import capi  # noqa: F401
from .capi import Handle, WpsGpuError  # noqa: F401

__all__ = ["capi", "Handle", "WpsGpuError"]
