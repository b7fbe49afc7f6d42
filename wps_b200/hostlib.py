"""ctypes access to libwpshost.so -- the C++ host side of the metgrid path (include/wpshost.h), used by tests and
drivers.  All array work behind it goes through libwpsgpu.so; nothing is computed in Python."""
import ctypes as C
import os

import numpy as np

from . import capi

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libwpshost.so")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise capi.WpsGpuError("%s not found: build it with wps_b200/csrc/build.sh" % LIB_PATH)
        capi.lib()                       # libwpsgpu.so first (same directory, also reachable through the rpath)
        _LIB = C.CDLL(LIB_PATH)
        _LIB.wpshost_last_error.restype = C.c_char_p
    return _LIB


def _check(rc):
    if rc < 0:
        raise capi.WpsGpuError("wpshost error %d: %s" % (rc, lib().wpshost_last_error().decode(errors="replace")))
    return rc


class HostDomain:
    """wpshost_domain: one model domain driven by the C++ record loop."""

    def __init__(self, device, xlat, xlon, xlat_u, xlon_u, xlat_v, xlon_v, landmask=None, process_bdy_width=0):
        ny, nx = xlat.shape
        self.nx, self.ny = nx, ny
        self._d = C.c_void_p()
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32).ctypes.data_as(C.c_void_p)
        _check(lib().wpshost_domain_create(int(device), nx, ny, f(xlat), f(xlon), f(xlat_u), f(xlon_u), f(xlat_v), f(xlon_v),
                                           f(landmask), int(process_bdy_width), C.byref(self._d)))

    def close(self):
        if self._d:
            lib().wpshost_domain_destroy(self._d)
            self._d = C.c_void_p()

    def set_table(self, text):
        return _check(lib().wpshost_set_table(self._d, text.encode()))

    def process_file(self, path, input_name):
        return _check(lib().wpshost_process_file(self._d, path.encode(), input_name.encode()))

    def create_derived_fields(self):
        _check(lib().wpshost_create_derived_fields(self._d))

    def fill_missing_levels(self):
        _check(lib().wpshost_fill_missing_levels(self._d))

    def fields(self):
        out = []
        for k in range(lib().wpshost_num_fields(self._d)):
            name = C.create_string_buffer(10)
            lev, stag, nx, ny = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
            _check(lib().wpshost_field_info(self._d, k, name, C.byref(lev), C.byref(stag), C.byref(nx), C.byref(ny)))
            out.append((name.value.decode(), lev.value, stag.value, nx.value, ny.value))
        return out

    def get_field(self, name, level):
        info = [f for f in self.fields() if f[0] == name and f[1] == level]
        if not info:
            raise KeyError((name, level))
        _, _, _, nx, ny = info[0]
        r = np.zeros((ny, nx), np.float32)
        v = np.zeros((ny, capi.nwords(nx)), np.int32)
        rc = lib().wpshost_get_field(self._d, name.encode(), int(level), r.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise KeyError((name, level))
        return r, v
