"""Host-side mirror of geogrid's calc_field driver (geogrid/src/process_tile_module.F:1208-1681) above the C ABI.

What the Fortran host keeps when it calls libwpsgpu.so restated in Python so that the parity tests read
like the reference's flow: the recursion over priority levels (highest first, :1271), the CONTINUOUS -> SP_CONTINUOUS
decision (:1296-1327), and the tile flood fill (:1347-1574 + process_neighbor :2077-2123) that decides in which ORDER
the source tiles of a level are visited.  The per-pixel / per-point work is the library's.  File access is replaced by a
`tile_source` callback (there is no geog data set here): (tile_x, tile_y) -> ndarray [nz][ny][nx] or None when the data
set has no such tile (get_data_tile's istatus /= 0)."""
from collections import deque
from dataclasses import dataclass, field as dc_field
from typing import Callable, List, Optional

import numpy as np

from . import capi


@dataclass
class SourceLevel:
    """One priority level of a GEOGRID.TBL entry (index file + table options) for a regular lat-lon data set."""
    dx_deg: float                       # index: dx = dy
    known_lat: float
    known_lon: float
    tile_x: int
    tile_y: int
    tile_bdr: int
    itype: int                          # capi.CATEGORICAL or capi.CONTINUOUS (iget_source_fieldtype)
    interp_option: str
    tile_source: Callable[[int, int], Optional[np.ndarray]]
    missing_value: float = capi.NAN
    masked: float = -1.0                # get_masked_value; -1 = none
    scale_factor: Optional[float] = None
    nz: int = 1
    raw: Optional[tuple] = None         # (wordsize, is_signed, endian, row_order): tile_source returns the file bytes


def tile_of(level: SourceLevel, rx: float, ry: float):
    """Tile that holds source point (rx, ry): lower-left index of the tile's interior (without the border), as get_tile_fname
    derives it from the index file (source_data_module.F:2914-2990)."""
    ix, iy = int(np.floor(rx + 0.5)), int(np.floor(ry + 0.5))
    tx = level.tile_x * ((ix - 1) // level.tile_x) + 1
    ty = level.tile_y * ((iy - 1) // level.tile_y) + 1
    return tx, ty


def flood_fill_tile_order(src_xy: np.ndarray, level: SourceLevel):
    """The order in which calc_field's two queues reach the source tiles of the patch (:1347-1574): a tile queue seeded with
    point (start_i, start_j); for the tile of each dequeued, unvisited point an inner queue visits every connected point
    of the patch that lies in the tile (is_point_in_tile) and hands the 8-neighbours that do not to the tile queue.
    src_xy: [ny][nx][2] source-grid coordinates of the destination points.  -> list of (tile_x, tile_y)."""
    ny, nx = src_xy.shape[:2]
    visited = np.zeros((ny, nx), bool)
    order, seen = [], set()
    tile_q = deque([(0, 0)])
    nb = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]      # process_neighbor call order
    while tile_q:
        i, j = tile_q.popleft()
        if visited[j, i]:
            continue
        t = tile_of(level, src_xy[j, i, 0], src_xy[j, i, 1])
        if t not in seen:                                   # have_processed_tile
            seen.add(t)
            order.append(t)
        visited[j, i] = True
        x0, x1 = t[0] - 0.5, t[0] + level.tile_x - 0.5      # is_point_in_tile in terms of the interior's index range
        y0, y1 = t[1] - 0.5, t[1] + level.tile_y - 0.5
        pq = deque([(i, j)])
        while pq:
            ci, cj = pq.popleft()
            for di, dj in nb:
                ni, nj = ci + di, cj + dj
                if ni < 0 or nj < 0 or ni >= nx or nj >= ny or visited[nj, ni]:
                    continue
                rx, ry = src_xy[nj, ni]
                if x0 <= rx <= x1 and y0 <= ry <= y1:         # floor(rx+0.5) >= first, ceiling(rx-0.5) <= last (:928-929)
                    visited[nj, ni] = True
                    pq.append((ni, nj))
                else:
                    tile_q.append((ni, nj))
    return order


def scan_tile_order(src_xy: np.ndarray, level: SourceLevel):
    """Every tile that holds a point of the patch, in row-major order of first appearance: enough for priority level 1 and
    continuous data, whose results do not depend on the tile order; O(npts) in numpy."""
    ix = np.floor(src_xy[..., 0] + 0.5).astype(np.int64)
    iy = np.floor(src_xy[..., 1] + 0.5).astype(np.int64)
    tx = level.tile_x * ((ix - 1) // level.tile_x) + 1
    ty = level.tile_y * ((iy - 1) // level.tile_y) + 1
    key = (ty * (1 << 32) + tx).ravel()
    _, first = np.unique(key, return_index=True)
    return [(int(tx.ravel()[k]), int(ty.ravel()[k])) for k in sorted(first)]


def effective_type(level: SourceLevel, dom_dx_m: float, gridtype: str = "C"):
    """:1296-1327: a CONTINUOUS level whose interp_option contains average_gcell(t) becomes SP_CONTINUOUS when the source
    is fine enough against the domain."""
    if level.itype != capi.CONTINUOUS or "average_gcell" not in level.interp_option:
        return level.itype
    _, _, has_gcell, thr = capi.parse_interp_option(level.interp_option)
    if not has_gcell:
        return level.itype
    src_dx = np.float32(level.dx_deg)
    if gridtype == "C":
        ok = np.float32(thr) * src_dx * np.float32(111.0) <= np.float32(dom_dx_m) / np.float32(1000.0)
    else:
        ok = src_dx >= np.float32(thr) * np.float32(dom_dx_m)
    return capi.SP_CONTINUOUS if ok else capi.CONTINUOUS


def calc_field(h, dom_proj, istagger, xlat, xlon, start_i, start_j, field, levels: List[SourceLevel], dom_dx_m, landmask=None,
               msg_fill_val=capi.NAN, start_k=1, order="reference", processed_out=None, trace=None):
    """calc_field for one field: `levels[0]` is priority level 1.  Returns the number of invalid categories met.
    trace (optional list) receives (ilevel, itype, [tiles in visiting order]) per level."""
    h.geogrid_field_begin(dom_proj, istagger, xlat, xlon, start_i, start_j, start_k, field, landmask)
    for ilevel in range(len(levels), 0, -1):                 # the recursion of :1271 bottoms out at the highest level first
        lv = levels[ilevel - 1]
        itype = effective_type(lv, dom_dx_m)
        kw = dict(dlat=lv.dx_deg, dlon=lv.dx_deg, known_x=1.0, known_y=1.0, known_lat=lv.known_lat, known_lon=lv.known_lon)
        src = capi.push_source_projection(capi.PROJ_LATLON, **kw)
        h.geogrid_level_begin(src, ilevel, itype, lv.interp_option, lv.missing_value, msg_fill_val, lv.masked, lv.scale_factor)
        lon = np.where(xlon >= 180.0, xlon - np.float32(360.0), xlon).astype(np.float32)
        rx, ry = h.lltoxy(src, xlat, lon, capi.M)
        src_xy = np.stack([rx, ry], axis=-1)
        tiles = flood_fill_tile_order(src_xy, lv) if order == "reference" else scan_tile_order(src_xy, lv)
        if trace is not None:
            trace.append((ilevel, itype, list(tiles)))
        for tx, ty in tiles:
            data = lv.tile_source(tx, ty)
            if data is None:                                 # no such tile in the data set
                continue
            b = lv.tile_bdr
            sk = 1 if itype == capi.CATEGORICAL else start_k
            if lv.raw is not None:
                ws, sg, en, ro = lv.raw
                h.geogrid_add_tile_raw(data, (lv.nz, lv.tile_y + 2 * b, lv.tile_x + 2 * b), ws, sg, en, ro, tx - b, ty - b, sk, b)
            else:
                h.geogrid_add_tile(data, tx - b, ty - b, sk, b)
        h.geogrid_level_end()
    return h.geogrid_field_finish(processed_out)
