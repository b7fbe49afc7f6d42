"""Host-side mirror of metgrid's per-file record loop (metgrid/src/process_domain_module.F:999-1449), driving the
batched CUDA path: every (field, level) record of a FILE becomes one queued slab, and ONE flush replaces the
per-slab interp_met_field calls.  Storage semantics (valid_mask / modified_mask carry-over between sources of
different priority, storage_module.F) are kept so that multiple fg_name prefixes behave as in the reference."""
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import capi
from .intermediate import MetData
from .interp_option import TableEntry, find_entry

BDR_WIDTH = 3   # process_domain_module.F:1027
# iproj codes of the intermediate format -> PROJ_* (read_met_module.F:286-386)
_IPROJ = {0: capi.PROJ_LATLON, 1: capi.PROJ_MERC, 3: capi.PROJ_LC, 4: capi.PROJ_GAUSS, 5: capi.PROJ_PS}
GRID_OF_STAGGER = {capi.M: 0, capi.U: 1, capi.V: 2}


@dataclass
class FgInput:
    """fg_input (datatype_module.F:12-64), the part the interpolation path touches"""
    field: str
    level: int
    stagger: int
    r_arr: np.ndarray
    valid_mask: np.ndarray
    modified_mask: Optional[np.ndarray]
    is_wind_grid_rel: bool = False
    mask_field: bool = False


def source_projection(fg: MetData):
    """push_source_projection call of process_domain_module.F:1144-1151 for one record."""
    code = _IPROJ[fg.iproj]
    if fg.startloc.strip() == "CENTER":
        starti, startj = (fg.nx + 1) / 2.0, (fg.ny + 1) / 2.0
    else:
        starti, startj = 1.0, 1.0
    return capi.push_source_projection(code, stand_lon=fg.xlonc, truelat1=fg.truelat1, truelat2=fg.truelat2,
                                       dxkm=fg.dx * 1000.0, dykm=fg.dy * 1000.0,
                                       dlat=fg.nlats if code == capi.PROJ_GAUSS else fg.deltalat, dlon=fg.deltalon,
                                       known_x=starti, known_y=startj, known_lat=fg.startlat, known_lon=fg.startlon,
                                       earth_radius=fg.earth_radius * 1000.0)


def is_global(fg: MetData) -> bool:
    """process_domain_module.F:1121-1122"""
    return (fg.iproj == 0 and abs(np.float32(fg.nx) * np.float32(fg.deltalon) - np.float32(360.0)) < 0.0001) or fg.iproj == 4


def halo_slab(fg: MetData):
    """:1121-1140 -> (slab incl. halo, minx, bdr)"""
    s = np.ascontiguousarray(fg.slab, np.float32)
    if is_global(fg):
        b = BDR_WIDTH
        return np.ascontiguousarray(np.concatenate([s[:, fg.nx - b:], s, s[:, :b]], axis=1)), 1 - b, b
    return s, 1, 0


def output_fields(storage: Dict[Tuple[str, int], "FgInput"]):
    """The output gather of process_single_met_time (:966-983): reset_next_field + get_next_output_field
    (storage_module.F:478-552) over the stored fields.  Yields (field_name, r_array [nlevels][ny][nx]) in the order
    the reference writes met_em variables:
      * fields: storage_put_field pushes the head node of a NEW field name at the FRONT of the list (:88-104), so fields
        come out in reverse order of their first storage;
      * fields whose header says mask_field (output_this_field = no, :1186-1191) are skipped (:507-515);
      * levels: kept sorted by secondary_cmp (datatype_module.F:124-155) -- for time-dependent and constant fields
        (everything the record loop stores) by DESCENDING vertical_level: 200100 (surface) first, then 100000, 97500 ...
    `storage` is the dict process_intermediate_fields returns (insertion-ordered, as Python dicts are)."""
    first_seen: List[str] = []
    for (name, _lvl) in storage:
        if name not in first_seen:
            first_seen.append(name)
    for name in reversed(first_seen):
        levels = sorted((lvl for (n, lvl) in storage if n == name), reverse=True)
        if storage[(name, levels[0])].mask_field:
            continue
        yield name, np.stack([storage[(name, lvl)].r_arr for lvl in levels])


class MetgridDomain:
    """One model domain: its staggered lat/lon arrays (from geo_em), LANDMASK and the device handle."""

    def __init__(self, handle: capi.Handle, xlat, xlon, xlat_u, xlon_u, xlat_v, xlon_v, landmask, domain=None,
                 mem_start=(1, 1), process_bdy_width=0, dom_dx=None, dom_dy=None, dom_proj=None):
        self.h = handle
        self.grids = {capi.M: (xlat, xlon), capi.U: (xlat_u, xlon_u), capi.V: (xlat_v, xlon_v)}
        self.landmask = landmask
        self.sm = mem_start
        ny, nx = xlat.shape
        self.domain = domain or (mem_start[0], mem_start[0] + nx - 1, mem_start[1], mem_start[1] + ny - 1)
        self.process_bdy_width = process_bdy_width
        self.dom_dx, self.dom_dy = dom_dx, dom_dy
        self.dom_proj = dom_proj     # set_domain_projection's projection; needed by the average_gcell branch only
        for st, (la, lo) in self.grids.items():
            handle.metgrid_set_grid(GRID_OF_STAGGER[st], la, lo, mem_start[0], mem_start[1])
        if landmask is not None:
            handle.metgrid_set_landmask(0, landmask)
        handle.metgrid_set_domain(*self.domain)

    def process_intermediate_fields(self, records: List[MetData], table: List[TableEntry], input_name: str = "FILE",
                                    storage: Optional[Dict[Tuple[str, int], FgInput]] = None):
        """The record loop of process_intermediate_fields; returns the storage dict keyed by (field, nint(xlvl))."""
        storage = {} if storage is None else storage
        # ---- get_interp_masks (:2061-2197): stash "<NAME>.mask" slabs of this source first
        masks: Dict[str, np.ndarray] = {}
        mask_names = {e.interp_mask for e in table[:-1] if e.interp_mask != " "}
        for fg in records:
            name = self._renamed(fg.field, table, input_name)
            if name in mask_names and name + ".mask" not in masks:
                masks[name + ".mask"] = halo_slab(fg)[0]
        queued = []
        keep = []   # host buffers must stay alive until the flush
        for fg in records:
            idx = find_entry(table, fg.field, input_name)
            name = fg.field
            if table[idx].output_name != " ":
                name = table[idx].output_name[:9]
                idx = find_entry(table, name, input_name)
            e = table[idx]
            slab, minx, bdr = halo_slab(fg)
            proj = source_projection(fg)
            _, _, has_gcell, thr = capi.parse_interp_option(e.interp_method)
            do_gcell = has_gcell and self._do_gcell(fg, thr)
            if do_gcell and self.dom_proj is None:
                raise ValueError("average_gcell applies to %s but MetgridDomain was built without dom_proj" % name)
            stag = e.output_stagger if e.output_stagger in (capi.U, capi.V) else capi.M
            xlat, _ = self.grids[stag]
            gy, gx = xlat.shape
            key = (name, int(np.rint(fg.xlvl)))
            if any(k == key for k, _ in queued):
                # a second record of the same (field, level) in this file: the reference has finished the first one --
                # interpolation and bitarray_merge (:1225-1365) -- before it reads the second, so flush what is queued
                self.h.metgrid_flush()
                for _, f in queued:
                    f.valid_mask |= f.modified_mask
                queued.clear()
                keep.clear()
            if key in storage:                       # storage_query_field found it (:1225-1233): carry data + valid mask over
                fld = storage[key]
                fld.modified_mask = None
            else:
                fld = FgInput(name, key[1], stag, np.zeros((gy, gx), np.float32), np.zeros((gy, capi.nwords(gx)), np.int32), None,
                              fg.is_wind_grid_rel, not e.output_this_field)
                fld._fresh = True
            fld.modified_mask = np.zeros((gy, capi.nwords(gx)), np.int32)
            m = [masks.get(nm + ".mask") if nm != " " else None for nm in (e.interp_mask, e.interp_land_mask, e.interp_water_mask)]
            fresh = getattr(fld, "_fresh", False)
            if do_gcell:
                self.h.metgrid_enqueue_gcell_slab(proj, self.dom_proj, e.field_opts(), slab, minx, 1, bdr, GRID_OF_STAGGER[stag],
                                                  stag, fld.r_arr, fld.modified_mask,
                                                  valid_mask=None if fresh else fld.valid_mask, mask=m[0],
                                                  use_landmask=(stag == capi.M and self.landmask is not None),
                                                  process_bdy_width=self.process_bdy_width)
            else:
                self.h.metgrid_enqueue_slab(proj, e.field_opts(), slab, minx, 1, bdr, GRID_OF_STAGGER[stag], stag, fld.r_arr,
                                            fld.modified_mask, valid_mask=None if fresh else fld.valid_mask, masks=tuple(m),
                                            use_landmask=(stag == capi.M and self.landmask is not None),
                                            process_bdy_width=self.process_bdy_width)
            fld._fresh = False
            keep.append((slab, m))
            queued.append((key, fld))
            storage[key] = fld
        self.h.metgrid_flush()
        for key, fld in queued:                      # bitarray_merge(valid_mask, modified_mask) (:1365)
            fld.valid_mask |= fld.modified_mask
        return storage

    # ---- fill_missing_levels / fill_field / create_level (process_domain_module.F:2748-2933, 3037-3332) ----------
    def _dims(self, stagger):
        st = stagger if stagger in (capi.U, capi.V) else capi.M
        return self.grids[st][0].shape

    def create_level(self, storage, name, stagger, rlevel, fill_const=None, src=None):
        """create_level: no-op when (name, level) exists; constant fill, or a copy of src = (field, level) with its
        valid bits; a missing source only logs in the reference (:3322-3325)."""
        key = (name, int(np.rint(rlevel)))
        if key in storage:
            return False
        gy, gx = self._dims(stagger)
        new = FgInput(name, key[1], stagger, np.zeros((gy, gx), np.float32), np.zeros((gy, capi.nwords(gx)), np.int32),
                      np.zeros((gy, capi.nwords(gx)), np.int32))
        if src is None:
            self.h.create_level(new.r_arr, new.valid_mask, fill_const=fill_const)
        else:
            skey = (src[0], int(np.rint(src[1])))
            if skey not in storage:
                return False
            sf = storage[skey]
            if sf.r_arr.shape != new.r_arr.shape:
                raise ValueError("Dimensions for %s do not match those of %s. This is probably because the staggerings "
                                 "of the fields do not match." % (src[0], name))
            self.h.create_level(new.r_arr, new.valid_mask, src=sf.r_arr, src_valid=sf.valid_mask)
        storage[key] = new
        return True

    @staticmethod
    def _levels_of(storage, name):
        """storage_get_levels: ascending vertical_level for time-dependent data (secondary_cmp, datatype_module.F:124)."""
        return sorted(l for (n, l) in storage if n == name)

    def fill_field(self, storage, name, entry, stagger, all_level_list=None):
        """fill_field (:3037-3177): walk the entry's fill_lev list in table order."""
        from .interp_option import get_constant_fill_lev, get_fill_src_level
        filled_all = False
        for lev_s, spec in entry.fill_lev:
            if lev_s == "all":
                if filled_all:
                    continue
                filled_all = True
                ist, const = get_constant_fill_lev(spec)
                if ist == 0 or spec == "vertical_index":
                    levels = all_level_list if all_level_list is not None else self._levels_of(storage, entry.level_template)
                    for l in levels:
                        self.create_level(storage, name, stagger, float(l), fill_const=const if ist == 0 else float(l))
                else:
                    levels = all_level_list if all_level_list is not None else self._levels_of(storage, spec)
                    if all_level_list is None and not levels:
                        filled_all = False
                    for l in levels:
                        self.create_level(storage, name, stagger, float(l), src=(spec, float(l)))
            else:
                rlevel = float(lev_s)
                ist, const = get_constant_fill_lev(spec)
                if ist == 0:
                    self.create_level(storage, name, stagger, rlevel, fill_const=const)
                else:
                    self.create_level(storage, name, stagger, rlevel, src=get_fill_src_level(spec))

    def create_derived_fields(self, storage, table):
        """create_derived_fields (:2939-3029): entries with derived=yes are built from their fill_lev rules."""
        for e in table[:-1]:
            if e.is_derived_field:
                self.fill_field(storage, e.fieldname, e, e.output_stagger)

    def fill_missing_levels(self, storage, table):
        """fill_missing_levels (:2748-2933): union of the levels of all 3-d fields (201300 excluded), fill_field for
        every stored field that has a table entry, then the vertical interpolation on device per 3-d field."""
        names = list(dict.fromkeys(n for (n, _) in storage))[::-1]      # head insertion puts the newest name first
        union = sorted({l for (n, l) in storage if l != 201300})
        if not union:
            return
        for n in names:
            e = next((t for t in table[:-1] if t.fieldname == n), None)
            if e is not None:
                stag = next(f.stagger for (nn, _), f in storage.items() if nn == n)
                self.fill_field(storage, n, e, stag, all_level_list=union)
        for n in list(dict.fromkeys(nn for (nn, _) in storage))[::-1]:
            levels = self._levels_of(storage, n)
            if len(levels) > 1:
                flds = [storage[(n, l)] for l in levels]
                self.h.fill_missing_levels(levels, [f.r_arr for f in flds], [f.valid_mask for f in flds])

    def _renamed(self, name, table, input_name):
        idx = find_entry(table, name, input_name)
        if table[idx].output_name != " ":
            return table[idx].output_name[:9]
        return name

    def _do_gcell(self, fg, threshold):
        """:1198-1220 (C grid)"""
        if self.dom_dx is None:
            return False
        if fg.dx == 0.0 and fg.dy == 0.0 and fg.deltalat != 0.0 and fg.deltalon != 0.0:
            dx, dy = abs(fg.deltalon), abs(fg.deltalat)
        else:
            dx, dy = fg.dx * 1000.0 / 111000.0, fg.dy * 1000.0 / 111000.0
        return threshold * max(dx, dy) * 111.0 <= max(self.dom_dx, self.dom_dy) / 1000.0

    def rotate_winds(self, storage, proj, idir, u_name="UU", v_name="VV", only_modified=True):
        """met_to_map (idir=-1, proj = domain projection, process_domain_module.F:802-834) or map_to_met (idir=+1,
        proj = source projection, :1392-1437) over every level that holds both components."""
        levels = sorted(set(l for (n, l) in storage if n == u_name) & set(l for (n, l) in storage if n == v_name))
        if not levels:
            return
        u = np.stack([storage[(u_name, l)].r_arr for l in levels])
        v = np.stack([storage[(v_name, l)].r_arr for l in levels])
        pick = (lambda f: f.modified_mask) if only_modified else (lambda f: f.valid_mask)
        um = np.stack([pick(storage[(u_name, l)]) for l in levels])
        vm = np.stack([pick(storage[(v_name, l)]) for l in levels])
        self.h.rotate_winds(idir, proj, u, um, v, vm, self.grids[capi.U][1], self.grids[capi.V][1], us=self.sm, vs=self.sm,
                            xlat_u=self.grids[capi.U][0], xlat_v=self.grids[capi.V][0])
        for k, l in enumerate(levels):
            storage[(u_name, l)].r_arr[:] = u[k]
            storage[(v_name, l)].r_arr[:] = v[k]
