"""Synthetic GFS/ERA5-shaped metgrid workloads: seeded, analytic, regenerable at any size.

This module only manufactures INPUTS (source slabs, masks, destination grids); every operator it calls goes
through libwpsgpu.so.  The field list follows ungrib/Variable_Tables/Vtable.GFS and the interp options are the
default METGRID.TBL.ARW entries (metgrid/METGRID.TBL.ARW)."""
import numpy as np

from . import capi

# name, n_levels, output stagger, METGRID.TBL.ARW entry (interp_option, missing_value, fill_missing, masked, interp_mask)
C3D = "sixteen_pt+four_pt+average_4pt"
SOIL = "sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search"


def gfs_slab_table(nlev3d=34):
    """One time of a GFS-shaped FILE: -> list of dict(field, nlev, stagger, chain, missing, fill, masked, mask, kind)."""
    t = []

    def add(field, nlev, stagger, chain, missing=1e20, fill=1e20, masked=capi.NOT_MASKED, mask=None, kind="smooth"):
        t.append(dict(field=field, nlev=nlev, stagger=stagger, chain=chain, missing=missing, fill=fill, masked=masked,
                      mask=mask, kind=kind))
    add("TT", nlev3d + 1, capi.M, C3D, fill=0.0)
    add("RH", nlev3d + 1, capi.M, C3D, fill=0.0)
    add("UU", nlev3d + 1, capi.U, C3D, fill=0.0, kind="wind")
    add("VV", nlev3d + 1, capi.V, C3D, fill=0.0, kind="wind")
    add("GHT", nlev3d, capi.M, C3D, fill=0.0)
    add("PSFC", 1, capi.M, "four_pt+average_4pt")
    add("PMSL", 1, capi.M, C3D)
    add("SOILHGT", 1, capi.M, "four_pt+average_4pt")
    add("LANDSEA", 1, capi.M, "nearest_neighbor", fill=-1.0, kind="landsea")
    add("SKINTEMP", 1, capi.M, SOIL, fill=0.0, masked=capi.MASKED_BOTH, mask=("both", 1.0, 0.0))
    add("SEAICE", 1, capi.M, "four_pt+average_4pt", missing=-1e30, fill=0.0, masked=capi.MASKED_LAND, mask=("mask", 1.0), kind="water")
    add("SNOW", 1, capi.M, "four_pt+average_4pt", fill=0.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    add("SNOWH", 1, capi.M, "four_pt+average_4pt", fill=0.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    for nm in ("ST000010", "ST010040", "ST040100", "ST100200"):
        add(nm, 1, capi.M, SOIL, missing=-1e30, fill=285.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    for nm in ("SM000010", "SM010040", "SM040100", "SM100200"):
        add(nm, 1, capi.M, SOIL, missing=-1e30, fill=1.0, masked=capi.MASKED_WATER, mask=("mask", 0.0), kind="land")
    return t


def n_slabs(table):
    return sum(e["nlev"] for e in table)


LAND_THRESHOLD = -1.2  # ~59 % land over the CONUS domain, coastlines with ~100-km structure


def landsea_fn(xp, lam, phi):
    """Synthetic land/sea pattern (xp = numpy or torch); land where the value exceeds LAND_THRESHOLD."""
    return xp.sin(5 * lam) + xp.cos(4 * phi) + 0.5 * xp.sin(19 * lam) * xp.cos(13 * phi)


def landsea_np(nx, ny):
    lam = np.linspace(0, 2 * np.pi, nx, endpoint=False)[None, :]
    phi = np.linspace(np.pi / 2, -np.pi / 2, ny)[:, None]
    return (landsea_fn(np, lam, phi) > LAND_THRESHOLD).astype(np.float32)


def slab_np(nx, ny, seed, kind, landsea=None, missing=1e20):
    """v = A + B sin(3 lam) cos(2 phi) + C lev + eps; exact zeros at 0.1 % of the cells."""
    rng = np.random.default_rng(seed)
    lam = np.linspace(0, 2 * np.pi, nx, endpoint=False)[None, :]
    phi = np.linspace(np.pi / 2, -np.pi / 2, ny)[:, None]
    if kind == "landsea":
        return landsea.copy()
    if kind == "wind":
        f = 12.0 * np.sin(3 * lam + 0.1 * seed) * np.cos(2 * phi) + rng.uniform(-0.012, 0.012, (ny, nx))
    else:
        f = 280.0 + 30.0 * np.sin(3 * lam) * np.cos(2 * phi) + 0.5 * (seed % 40) + rng.uniform(-0.03, 0.03, (ny, nx))
    f = f.astype(np.float32)
    f[rng.random((ny, nx)) < 0.001] = 0.0
    if kind == "land":
        f[landsea == 0] = missing
    elif kind == "water":
        f[landsea == 1] = missing
    return f


def add_halo(slab, bdr=3):
    """halo_slab of process_domain_module.F:1121-1135 (global lat-lon source); works for numpy and torch."""
    nx = slab.shape[-1]
    if isinstance(slab, np.ndarray):
        return np.ascontiguousarray(np.concatenate([slab[..., nx - bdr:], slab, slab[..., :bdr]], axis=-1))
    import torch
    return torch.cat([slab[..., nx - bdr:], slab, slab[..., :bdr]], dim=-1).contiguous()


def lambert_domain(nx, ny, dx, truelat1, truelat2, stand_lon, ref_lat, ref_lon):
    return capi.map_set(capi.PROJ_LC, truelat1=truelat1, truelat2=truelat2, stdlon=stand_lon, lat1=ref_lat,
                        lon1=ref_lon, knowni=(nx + 1) / 2.0, knownj=(ny + 1) / 2.0, dx=dx)


CONFIGS = {
    # BASELINE.json configs[0]: 1-degree GFS-shaped source -> 100x100 30-km Lambert (namelist.wps.all_options geometry)
    "config1": dict(src=(360, 181, -1.0, 1.0), nlev=32, dom=(100, 100, 30000.0, 30.0, 60.0, -98.0, 34.83, -81.03)),
    # BASELINE.json configs[1]: GFS 0.25 deg, 34 levels -> 3-km CONUS Lambert 1799x1059 (HRRR-like)
    "config2": dict(src=(1440, 721, -0.25, 0.25), nlev=34, dom=(1799, 1059, 3000.0, 38.5, 38.5, -97.5, 38.5, -97.5)),
}
