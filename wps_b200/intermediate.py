"""WPS intermediate format, version 5 (host I/O mirror of metgrid/src/read_met_module.F:259-386 and
write_met_module.F:264-382): Fortran sequential unformatted records, big-endian data and 4-byte record markers.
GRIB decoding and file I/O stay on the host; this module exists to manufacture and read synthetic inputs."""
import struct
from dataclasses import dataclass

import numpy as np

PROJ_LATLON, PROJ_MERC, PROJ_LC, PROJ_GAUSS, PROJ_PS, PROJ_CASSINI = 0, 1, 3, 4, 5, 6   # iproj codes of the FILE format


@dataclass
class MetData:
    """met_data (met_data_module.F:4-18)"""
    field: str
    xlvl: float
    slab: np.ndarray           # [ny][nx] float32
    iproj: int = 0
    hdate: str = "2026-10-19_00:00:00"
    xfcst: float = 0.0
    map_source: str = "synthetic"
    units: str = "-"
    desc: str = "-"
    startloc: str = "SWCORNER"
    startlat: float = 0.0
    startlon: float = 0.0
    deltalat: float = 0.0
    deltalon: float = 0.0
    dx: float = 0.0
    dy: float = 0.0
    xlonc: float = 0.0
    truelat1: float = 0.0
    truelat2: float = 0.0
    nlats: float = 0.0
    earth_radius: float = 6371.229   # km in the file; *1000 when the projection is pushed (process_domain_module.F:1151)
    is_wind_grid_rel: bool = False

    @property
    def nx(self):
        return self.slab.shape[1]

    @property
    def ny(self):
        return self.slab.shape[0]


def _rec(f, payload):
    f.write(struct.pack(">i", len(payload)))
    f.write(payload)
    f.write(struct.pack(">i", len(payload)))


def write_records(path, records):
    with open(path, "wb") as f:
        for r in records:
            _rec(f, struct.pack(">i", 5))
            hdr = r.hdate.ljust(24)[:24].encode() + struct.pack(">f", r.xfcst) + r.map_source.ljust(32)[:32].encode() + \
                r.field.ljust(9)[:9].encode() + r.units.ljust(25)[:25].encode() + r.desc.ljust(46)[:46].encode() + \
                struct.pack(">fiii", r.xlvl, r.nx, r.ny, r.iproj)
            _rec(f, hdr)
            sl = r.startloc.ljust(8)[:8].encode()
            if r.iproj == PROJ_LATLON:
                body = sl + struct.pack(">5f", r.startlat, r.startlon, r.deltalat, r.deltalon, r.earth_radius)
            elif r.iproj == PROJ_MERC:
                body = sl + struct.pack(">6f", r.startlat, r.startlon, r.dx, r.dy, r.truelat1, r.earth_radius)
            elif r.iproj == PROJ_LC:
                body = sl + struct.pack(">8f", r.startlat, r.startlon, r.dx, r.dy, r.xlonc, r.truelat1, r.truelat2, r.earth_radius)
            elif r.iproj == PROJ_GAUSS:
                body = sl + struct.pack(">5f", r.startlat, r.startlon, r.nlats, r.deltalon, r.earth_radius)
            elif r.iproj == PROJ_PS:
                body = sl + struct.pack(">7f", r.startlat, r.startlon, r.dx, r.dy, r.xlonc, r.truelat1, r.earth_radius)
            else:
                raise ValueError("unsupported iproj %d" % r.iproj)
            _rec(f, body)
            _rec(f, struct.pack(">i", 1 if r.is_wind_grid_rel else 0))
            _rec(f, np.ascontiguousarray(r.slab, ">f4").tobytes())


def read_records(path):
    out = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def rec():
        nonlocal pos
        n = struct.unpack_from(">i", data, pos)[0]
        payload = data[pos + 4:pos + 4 + n]
        pos += 8 + n
        return payload
    while pos < len(data):
        ver = struct.unpack(">i", rec())[0]
        if ver != 5:
            raise ValueError("only version 5 intermediate files are supported (got %d)" % ver)
        h = rec()
        hdate = h[0:24].decode().strip()
        xfcst = struct.unpack(">f", h[24:28])[0]
        map_source = h[28:60].decode().strip()
        fieldnm = h[60:69].decode().strip()
        units = h[69:94].decode().strip()
        desc = h[94:140].decode().strip()
        xlvl, nx, ny, iproj = struct.unpack(">fiii", h[140:156])
        if fieldnm == "HGT":
            fieldnm = "GHT"            # read_met_module.F:272
        b = rec()
        r = MetData(field=fieldnm, xlvl=xlvl, slab=None, iproj=iproj, hdate=hdate, xfcst=xfcst, map_source=map_source,
                    units=units, desc=desc, startloc=b[0:8].decode())
        v = struct.unpack(">%df" % ((len(b) - 8) // 4), b[8:])
        if iproj == PROJ_LATLON:
            r.startlat, r.startlon, r.deltalat, r.deltalon, r.earth_radius = v
        elif iproj == PROJ_MERC:
            r.startlat, r.startlon, r.dx, r.dy, r.truelat1, r.earth_radius = v
        elif iproj == PROJ_LC:
            r.startlat, r.startlon, r.dx, r.dy, r.xlonc, r.truelat1, r.truelat2, r.earth_radius = v
        elif iproj == PROJ_GAUSS:
            r.startlat, r.startlon, r.nlats, r.deltalon, r.earth_radius = v
        elif iproj == PROJ_PS:
            r.startlat, r.startlon, r.dx, r.dy, r.xlonc, r.truelat1, r.earth_radius = v
        r.is_wind_grid_rel = struct.unpack(">i", rec())[0] != 0
        r.slab = np.frombuffer(rec(), ">f4").astype(np.float32).reshape(ny, nx)
        out.append(r)
    return out
