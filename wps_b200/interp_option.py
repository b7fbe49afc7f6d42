"""METGRID.TBL parser -- host-side mirror of metgrid/src/interp_option_module.F:31-505 (read_interp_table).

The table surface is kept verbatim: blocks separated by lines containing '=====', 'key=value' specifications
separated by ';', '#' comments, whitespace ignored (despace), and one extra trailing entry holding the defaults
(interp_option_module.F:77-147).  Only the keys the horizontal-interpolation path consumes are interpreted; the
others are kept as strings."""
from dataclasses import dataclass, field
from typing import Dict, List

from . import capi

NAN = 1.0e20


@dataclass
class TableEntry:
    fieldname: str = " "
    interp_method: str = "nearest_neighbor"
    masked: float = capi.NOT_MASKED
    fill_missing: float = NAN
    missing_value: float = NAN
    interp_mask: str = " "
    interp_land_mask: str = " "
    interp_water_mask: str = " "
    interp_mask_val: float = NAN
    interp_land_mask_val: float = NAN
    interp_water_mask_val: float = NAN
    interp_mask_relational: str = " "
    interp_land_mask_relational: str = " "
    interp_water_mask_relational: str = " "
    output_stagger: int = capi.M
    output_this_field: bool = True
    is_u_field: bool = False
    is_v_field: bool = False
    output_name: str = " "
    from_input: str = "*"
    fill_lev: List = field(default_factory=list)     # [(level string | 'all', fill string)] in table order (:459-475)
    level_template: str = " "
    is_derived_field: bool = False
    flag_in_output: str = " "
    other: Dict[str, str] = field(default_factory=dict)

    def field_opts(self):
        """wpsgpu_field_opts for this entry (interp_array_from_string etc. are re-done per slab in the reference,
        process_domain_module.F:2324-2325; here once per entry)."""
        return capi.field_opts(self.interp_method, self.missing_value, self.fill_missing, self.masked,
                               (self.interp_mask_val, self.interp_land_mask_val, self.interp_water_mask_val),
                               self.interp_mask_relational + self.interp_land_mask_relational + self.interp_water_mask_relational)


def _parse_mask(value):
    """interp_mask = NAME(val) | NAME(<val) | NAME(>val)  (interp_option_module.F:311-348) -> (name, val, relational)"""
    p1, p2 = value.find("("), value.find(")")
    if p1 < 0 or p2 < 0:
        return value, 0.0, " "          # WARN 'Problem in specifying interp_mask flag. Setting masked flag to 0.'
    s1, s2 = value.find("<"), value.find(">")
    if s1 >= 0 or s2 >= 0:
        rel = "<" if s1 >= 0 else ">"
        return value[:p1], float(value[p1 + 2:p2]), rel
    return value[:p1], float(value[p1 + 1:p2]), " "


def read_interp_table(text: str, gridtype: str = "C") -> List[TableEntry]:
    """Parse METGRID.TBL text.  Returns num_entries = (number of non-empty blocks) + 1 entries; the last is the default."""
    entries: List[TableEntry] = []
    cur = TableEntry(output_stagger=capi.M if gridtype == "C" else capi.HH)
    nparams = 0
    for raw in text.splitlines():
        buf = "".join(raw.split())          # despace
        if not buf or buf[0] == "#":
            continue
        if "=====" in buf:
            if nparams > 0:
                entries.append(cur)
                cur = TableEntry(output_stagger=capi.M if gridtype == "C" else capi.HH)
            nparams = 0
            continue
        eos = buf.find("#")
        if eos >= 0:
            buf = buf[:eos]
        for spec in buf.split(";"):
            idx = spec.find("=")
            if idx < 0:
                continue
            nparams += 1
            key, val = spec[:idx], spec[idx + 1:]
            if key == "name":
                cur.fieldname = val
            elif key == "from_input":
                cur.from_input = val
            elif key == "output_stagger":
                for nm, code in (("M", capi.M), ("U", capi.U), ("V", capi.V), ("HH", capi.HH), ("VV", capi.VV)):
                    if val in nm and val:        # index('M', trim(value)) /= 0
                        cur.output_stagger = code
                        break
            elif key == "output":
                cur.output_this_field = not (val and val in "no")
            elif key == "is_u_field":
                cur.is_u_field = bool(val) and val in "yes"
            elif key == "is_v_field":
                cur.is_v_field = bool(val) and val in "yes"
            elif key == "interp_option":
                cur.interp_method = val
            elif key == "interp_mask":
                cur.interp_mask, cur.interp_mask_val, cur.interp_mask_relational = _parse_mask(val)
            elif key == "interp_land_mask":
                cur.interp_land_mask, cur.interp_land_mask_val, cur.interp_land_mask_relational = _parse_mask(val)
            elif key == "interp_water_mask":
                cur.interp_water_mask, cur.interp_water_mask_val, cur.interp_water_mask_relational = _parse_mask(val)
            elif key == "masked":
                if "water" in val:
                    cur.masked = capi.MASKED_WATER
                elif "land" in val:
                    cur.masked = capi.MASKED_LAND
                elif "both" in val:
                    cur.masked = capi.MASKED_BOTH
            elif key == "output_name":
                cur.output_name = val
            elif key == "fill_missing":
                cur.fill_missing = float(val.replace("E", "e").replace("D", "e"))
            elif key == "missing_value":
                cur.missing_value = float(val.replace("E", "e").replace("D", "e"))
            elif key == "fill_lev":          # fill_lev = [level:]spec ; no level means 'all' (:459-475)
                c = val.find(":")
                cur.fill_lev.append(("all", val) if c < 0 else (val[:c], val[c + 1:]))
            elif key == "level_template":
                cur.level_template = val
            elif key == "derived":
                cur.is_derived_field = bool(val) and val in "yes"
            elif key == "flag_in_output":
                cur.flag_in_output = val
            else:
                cur.other.setdefault(key, val)
    if nparams > 0:
        entries.append(cur)
    entries.append(TableEntry(output_stagger=capi.M if gridtype == "C" else capi.HH))   # the implicit default entry
    return entries


def find_entry(entries: List[TableEntry], fieldname: str, input_name: str) -> int:
    """Index lookup of process_domain_module.F:1080-1095 / :2253-2268: exact name match; an entry whose from_input
    occurs in input_name wins, otherwise the first '*' entry; not found -> the default (last) entry."""
    num = len(entries) - 1
    idxt = num
    found = False
    for i in range(num):
        e = entries[i]
        if e.fieldname == fieldname:
            if (e.from_input != "*" and e.from_input in input_name) or (e.from_input == "*" and not found):
                idxt = i
                found = True
    return idxt


def get_constant_fill_lev(fill_opt: str):
    """-> (istatus, value): 'const(v)' specifications (interp_option_module.F:736-772); istatus 0 = is a constant."""
    i = fill_opt.find("const")
    if i < 0:
        return 1, NAN
    p1, p2 = fill_opt.find("(", i), fill_opt.find(")", i)
    if p1 >= 0 and p2 >= 0:
        return 0, float(fill_opt[p1 + 1:p2].replace("D", "e"))
    return 0, NAN


def get_fill_src_level(fill_opt: str):
    """'FIELD(level)' -> (FIELD, level); a bare 'FIELD' means level 1 (interp_option_module.F:780-812)."""
    p1, p2 = fill_opt.find("("), fill_opt.find(")")
    if p1 >= 0 and p2 >= 0:
        return fill_opt[:p1], int(fill_opt[p1 + 1:p2])
    return fill_opt, 1
