"""Host mirror of the field loop of process_mpas_fields (metgrid/src/process_domain_module.F:1626-1958) above the C ABI:
which table drives a field, which target grid it goes to, and which (field, level) planes it is stored as.  The remapping
and the per-level extraction run on the device (wpsgpu_mpas_remap_planes); file access (scan_input.F, netCDF) stays with the
caller, which hands over one field at a time as scan_input_next_field / scan_input_read_field would.

derive_mpas_fields (:1957-2053, TT = theta * (p / p0) ** (Rd / cp) on the stored planes) is `derive_mpas_fields` below."""
from typing import List, Sequence

import numpy as np

from . import capi

_RENAME = {"u10": "uReconstructZonal", "v10": "uReconstructMeridional", "q2": "qv", "t2m": "theta"}     # :1628-1636
_SOIL = {"tslb": ("ST000010", "ST010040", "ST040100", "ST100200"), "smois": ("SM000010", "SM010040", "SM040100", "SM100200")}
_DECOMP = {"nCells": capi.MPAS_CELLS, "nVertices": capi.MPAS_VERTICES, "nEdges": capi.MPAS_EDGES}
GRID_OF_STAGGER = {capi.M: 0, capi.U: 1, capi.V: 2}      # remap_info_m / _u / _v (:1535-1590)


def mpas_name_to_idx(entries, name: str) -> int:
    """interp_option_module.F:605-626; 1-based, 0 = no entry carries this mpas_name"""
    for i, e in enumerate(entries):
        if e.other.get("mpas_name", " ").strip() == name.strip():
            return i + 1
    return 0


def can_remap_field(dimnames: Sequence[str], dtype, time_dependent: bool) -> bool:
    """remapper.F:339-376"""
    if np.dtype(dtype) not in (np.dtype(np.int32), np.dtype(np.float32), np.dtype(np.float64)):
        return False
    ndims = len(dimnames)
    if ndims == 0 or (ndims == 1 and time_dependent):
        return False
    decomp = ndims - 1 if time_dependent else ndims
    return dimnames[decomp - 1] in _DECOMP


def process_mpas_field(h: "capi.Handle", entries, name: str, dimnames: Sequence[str], array, landmask=None, warn=None) -> List[dict]:
    """One pass of the `do while (scan_input_next_field ...)` body for a REAL field without its time dimension.
    dimnames: Fortran order, decomposed dimension last (e.g. ('nVertLevels', 'nCells')); array: [nNodes][nlev] or [nNodes]
    (C order = the file's).  Returns the storage_put_field calls as dicts(field, level, vertical_coord, stagger, r_arr)."""
    if not can_remap_field(dimnames, array.dtype, False):
        return []
    name = _RENAME.get(name, name)
    idx = mpas_name_to_idx(entries, name)
    if idx == 0:
        return []                                                          # :1648-1651
    ent = entries[idx - 1]
    stagger = ent.output_stagger                                           # mpas_output_stagger (:663-684)
    if stagger not in GRID_OF_STAGGER:
        raise capi.WpsGpuError("Cannot handle requested output stagger %d for MPAS field %s ..." % (stagger, name))
    grid = GRID_OF_STAGGER[stagger]
    masked_water = ent.masked == capi.MASKED_WATER
    use_masked_table = stagger == capi.M and masked_water                 # :1663-1669: only the mass grid has the masked table
    interm = ent.fieldname.strip() or name                                 # mpas_to_intermediate_name, else the MPAS name
    nlat, nlon = h._mpas_grids[grid]
    lead = dimnames[0] if len(dimnames) == 2 else None
    nlev = 1 if array.ndim == 1 else int(array.shape[1])
    new = (lambda: np.empty((nlat, nlon), np.float32)) if isinstance(array, np.ndarray) else None
    if new is None:
        import torch
        new = lambda: torch.empty((nlat, nlon), dtype=torch.float32, device=array.device)   # noqa: E731
    decomp = _DECOMP[dimnames[-1]]
    out = []

    def put(field, level, coord, r_arr):
        out.append(dict(field=field, level=level, vertical_coord=coord, stagger=stagger, r_arr=r_arr))
    if len(dimnames) == 2 and lead == "nVertLevels":                       # :1700-1731
        planes = [new() for _ in range(nlev)]
        h.mpas_remap_planes(grid, decomp, array, planes, mode=0, masked=use_masked_table)
        for k in range(nlev):
            put(interm, k + 1, "num_mpas_levels", planes[k])
    elif len(dimnames) == 2 and lead == "nVertLevelsP1":                   # :1733-1809
        planes = [new() for _ in range(nlev)]
        h.mpas_remap_planes(grid, decomp, array, planes, mode=1, masked=use_masked_table)
        put(interm, 200100.0, "none", planes[0])
        for k in range(1, nlev):
            put(interm, k, "num_mpas_levels", planes[k])
        if name == "zgrid":
            put("SOILHGT", 200100.0, "num_mpas_levels", planes[0])         # same values; vertical_coord as left by the loop above
    elif len(dimnames) == 2 and lead == "nSoilLevels":                     # :1811-1925
        if name not in _SOIL:
            if warn:
                warn("Skipping unknown MPAS soil field %s ..." % name)
            return []
        if nlev > 4:
            raise capi.WpsGpuError("Too many soil layers in MPAS soil field %s ..." % name)
        planes = [new() for _ in range(nlev)]
        h.mpas_remap_planes(grid, decomp, array, planes, mode=0, masked=use_masked_table, landmask=landmask if masked_water else None,
                            fill_missing=ent.fill_missing)
        for k in range(nlev):
            put(_SOIL[name][k], 200100.0, "none", planes[k])
    elif len(dimnames) == 1:                                               # :1927-1956
        planes = [new()]
        h.mpas_remap_planes(grid, decomp, array, planes, mode=0, masked=use_masked_table, landmask=landmask if masked_water else None,
                            fill_missing=ent.fill_missing)
        put(interm, 200100.0, "none", planes[0])
    return out


def derive_mpas_fields(h: "capi.Handle", stored: List[dict]) -> None:
    """process_domain_module.F:1957-2053 on the planes `process_mpas_field` returned: where TT and PRESSURE both exist with the
    same number of levels, TT (potential temperature so far) is multiplied by the Exner function, level by level, in place."""
    tt = sorted((d for d in stored if d["field"] == "TT" and d["vertical_coord"] == "num_mpas_levels"), key=lambda d: d["level"])
    pr = sorted((d for d in stored if d["field"] == "PRESSURE" and d["vertical_coord"] == "num_mpas_levels"), key=lambda d: d["level"])
    if not tt or len(tt) != len(pr):
        return
    h.mpas_derive_tt([d["r_arr"] for d in tt], [d["r_arr"] for d in pr])
