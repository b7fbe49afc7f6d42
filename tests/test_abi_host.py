"""CPU: the C-ABI library loads without a GPU, exports every symbol include/wpsgpu.h declares, fails loudly when no
device exists, and its host-only entry points (interp_option parsing, map_set) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    from wps_b200 import capi
    lib = capi.lib()
    hdr = open(os.path.join(ROOT, "include", "wpsgpu.h")).read()
    names = sorted(set(re.findall(r"\b(wpsgpu_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_library_exports_and_fails_without_device(tmp_path):
    """libwpshost.so (C++ host side, include/wpshost.h) loads, exports its interface and reports the missing device."""
    import torch
    from wps_b200 import capi, hostlib
    lib = hostlib.lib()
    hdr = open(os.path.join(ROOT, "include", "wpshost.h")).read()
    names = sorted(set(re.findall(r"\b(wpshost_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    if not torch.cuda.is_available():
        z = np.zeros((3, 4), np.float32)
        with pytest.raises(capi.WpsGpuError):
            hostlib.HostDomain(0, z, z, np.zeros((3, 5), np.float32), np.zeros((3, 5), np.float32), np.zeros((4, 4), np.float32),
                               np.zeros((4, 4), np.float32))


def test_no_cpu_fallback_without_device():
    import torch
    from wps_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.WpsGpuError) as e:
        capi.Handle(0)
    assert "no CUDA device" in str(e.value) or "CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wps_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("CPU oracle", "").lower() or f == "__init__.py", (dirpath, f)


STRINGS = ["sixteen_pt+four_pt+average_4pt", "four_pt+average_4pt", "nearest_neighbor", "four_pt",
           "sixteen_pt+four_pt+wt_average_4pt+wt_average_16pt+search", "average_gcell(4.0)+four_pt+average_4pt",
           "average_gcell(1.5)+sixteen_pt", "search(5)", "four_pt+search(12)+average_4pt", "average_16pt+wt_average_16pt",
           "four_pt+x", "bogus+four_pt", "wt_average_4pt+search", "sixteen_pt+", "+four_pt"]


@pytest.mark.parametrize("s", STRINGS)
def test_parse_interp_option_matches_oracle(orc, s):
    from wps_b200 import capi
    codes, opts, has_gcell, thr = capi.parse_interp_option(s)
    ocodes, oopts = orc.parse_interp_option(s)
    assert codes == ocodes and opts == oopts
    assert has_gcell == ("average_gcell" in s)
    if has_gcell:
        assert thr == orc.gcell_threshold(s)


def test_parse_errors():
    from wps_b200 import capi
    with pytest.raises(capi.WpsGpuError):
        capi.parse_interp_option("search(abc)")


PROJS = [
    ("PROJ_LC", dict(truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=30000.)),
    ("PROJ_LC", dict(truelat1=38.5, truelat2=38.5, stdlon=-97.5, lat1=38.5, lon1=-97.5, knowni=900., knownj=530., dx=3000.)),
    ("PROJ_LC", dict(truelat1=-30., truelat2=-60., stdlon=140., lat1=-35., lon1=145., knowni=40., knownj=40., dx=12000.)),
    ("PROJ_PS", dict(truelat1=60., stdlon=-100., lat1=70., lon1=-110., knowni=60., knownj=60., dx=25000.)),
    ("PROJ_PS", dict(truelat1=-71., stdlon=0., lat1=-75., lon1=20., knowni=50., knownj=50., dx=30000.)),
    ("PROJ_MERC", dict(truelat1=10., lat1=5., lon1=100., knowni=30., knownj=30., dx=27000.)),
    ("PROJ_LATLON", dict(latinc=-0.25, loninc=0.25, knowni=1., knownj=1., lat1=90., lon1=0., nxmax=1440)),
    ("PROJ_LATLON", dict(latinc=0.5, loninc=0.5, knowni=1., knownj=1., lat1=-90., lon1=190., nxmax=720)),
    ("PROJ_CASSINI", dict(latinc=0.5, loninc=0.5, lat1=10., lon1=-20., lat0=60., lon0=-140., knowni=1., knownj=1., stdlon=-140., dx=55000., dy=55000.)),
    ("PROJ_PS_WGS84", dict(truelat1=70., stdlon=-45., lat1=65., lon1=-50., knowni=100., knownj=100., dx=5000.)),
    ("PROJ_ALBERS_NAD83", dict(truelat1=29.5, truelat2=45.5, stdlon=-96., lat1=23., lon1=-96., knowni=1., knownj=1., dx=1000.)),
    ("PROJ_GAUSS", dict(nlat=47, lat1=88.542, lon1=0., loninc=1.875, nxmax=192)),
    ("PROJ_ROTLL", dict(ixdim=120, jydim=241, phi=12.0, lat1=38.0, lon1=-97.0, stagger=5, **{"lambda": 10.0})),
]


@pytest.mark.parametrize("name,kw", PROJS)
def test_map_set_matches_oracle_bit_for_bit(orc, name, kw):
    """wpsgpu_map_set runs on the host with the FP64-then-round policy: identical to the oracle in transc mode 1"""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    po = orc.map_set(getattr(orc, name), **kw)
    pg = capi.map_set(getattr(capi, name), **kw)
    for f in ("lat1", "lon1", "lat0", "lon0", "dx", "dy", "latinc", "loninc", "dlon", "stdlon", "truelat1", "truelat2", "hemi",
              "cone", "polei", "polej", "rsw", "rebydx", "knowni", "knownj", "re_m", "rho0", "nc", "bigc"):
        a, b = np.float32(getattr(po, f)), np.float32(getattr(pg, f))
        assert a.tobytes() == b.tobytes(), (name, f, a, b)
    assert (po.nxmin, po.nxmax, po.nlat) == (pg.nxmin, pg.nxmax, pg.nlat)
    assert (po.ixdim, po.jydim, po.stagger, po.phi, po.lambda_) == (pg.ixdim, pg.jydim, pg.stagger, pg.phi, getattr(pg, "lambda"))
    if name == "PROJ_GAUSS":
        assert np.array_equal(np.array(po.gauss_lat[:94], np.float32), np.array(pg.gauss_lat[:94], np.float32))


NEST_FIELDS = ("lat1", "lon1", "dx", "stdlon", "truelat1", "truelat2", "hemi", "cone", "polei", "polej", "rsw", "rebydx", "knowni", "knownj",
               "re_m", "latinc", "loninc")


@pytest.mark.parametrize("name,kw", [
    ("PROJ_LC", dict(truelat1=30., truelat2=60., stdlon=-98., lat1=38.0, lon1=-98.0, knowni=75.0, knownj=65.0, dx=27000.)),
    ("PROJ_PS", dict(truelat1=60., stdlon=-100., lat1=62.0, lon1=-100.0, knowni=75.0, knownj=65.0, dx=27000.)),
    ("PROJ_MERC", dict(truelat1=10., lat1=5.0, lon1=120.0, knowni=75.0, knownj=65.0, dx=27000.)),
])
def test_compute_nest_locations_matches_oracle(orc, name, kw):
    """compute_nest_locations / find_known_latlon (llxy_module.F:331-593) for a 27/9/3-km three-domain nest plus a second
    nest on level 2: host arithmetic of the product vs the oracle's recursive restatement, bit for bit (transc mode 1)."""
    from wps_b200 import capi
    orc.set_transc_mode(1)
    nests = [(1, 1, 1, 1, 150, 130), (1, 3, 40, 35, 160, 151), (2, 3, 50, 45, 172, 163), (1, 5, 90, 20, 101, 96)]
    po = orc.compute_nest_locations(getattr(orc, name), nests, **kw)
    pg = capi.compute_nest_locations(getattr(capi, name), nests, **kw)
    for k, (a, b) in enumerate(zip(po, pg)):
        for f in NEST_FIELDS:
            x, y = np.float32(getattr(a, f)), np.float32(getattr(b, f))
            assert x.tobytes() == y.tobytes(), (name, "domain %d" % (k + 1), f, x, y)
    if True:
        # known answers: grid distances divide by the ratios, the known point is the nest's centre (ixdim/2, jydim/2)
        assert np.float32(pg[1].dx) == np.float32(27000.0 / 3) and np.float32(pg[3].dx) == np.float32(27000.0 / 5)
        assert np.float32(pg[2].dx) == np.float32(np.float32(27000.0) / np.float32(3) / np.float32(3))
        assert (pg[2].knowni, pg[2].knownj) == (86.0, 81.5)
        # the centre of nest 2 seen from the coarse domain: (80 - 2) / 3 + 40, (75.5 - 2) / 3 + 35
        lat, lon = orc.xytoll(po[0], np.array([(80.0 - 2.0) / 3.0 + 40.0], np.float32), np.array([(75.5 - 2.0) / 3.0 + 35.0], np.float32))
        assert np.float32(pg[1].lat1) == np.float32(lat[0]) and np.float32(pg[1].lon1) == np.float32(lon[0])


def test_map_set_rejects_what_the_reference_rejects(orc):
    from wps_b200 import capi
    with pytest.raises(capi.WpsGpuError):
        capi.map_set(capi.PROJ_LC, truelat1=30., truelat2=60., stdlon=-98., lat1=34.83, lon1=-81.03, knowni=50.5, knownj=50.5)  # dx missing
    with pytest.raises(capi.WpsGpuError):
        capi.map_set(capi.PROJ_LC, truelat1=30., truelat2=60., stdlon=-98., lat1=94.83, lon1=-81.03, knowni=50.5, knownj=50.5, dx=3000.)
    with pytest.raises(capi.WpsGpuError):
        capi.map_set(capi.PROJ_ROTLL, lat1=1.0, lon1=1.0)


def _header_entry_points():
    import re
    txt = open(os.path.join(ROOT, "include", "wpsgpu.h")).read()
    return sorted(set(re.findall(r"^(?:int|int64_t|void \*|const char \*)\s*(wpsgpu_\w+)\s*\(", txt, re.M)))


def test_fortran_module_binds_every_entry_point():
    """fortran/wpsgpu_mod.F90 has a bind(C) interface for every function include/wpsgpu.h declares, and nothing else."""
    import re
    f90 = open(os.path.join(ROOT, "fortran", "wpsgpu_mod.F90")).read()
    bound = sorted(set(re.findall(r"bind\(C,\s*name='(wpsgpu_\w+)'\)", f90)))
    assert bound == _header_entry_points(), (sorted(set(_header_entry_points()) - set(bound)), sorted(set(bound) - set(_header_entry_points())))


def _f90_type_members(name):
    """member names of a bind(C) derived type of fortran/wpsgpu_mod.F90, in declaration order"""
    import re
    f90 = open(os.path.join(ROOT, "fortran", "wpsgpu_mod.F90")).read()
    body = re.search(r"type, bind\(C\) :: %s\n(.*?)end type %s" % (name, name), f90, re.S).group(1)
    body = re.sub(r"&\s*\n", " ", body)
    out = []
    for line in body.split("\n"):
        line = line.split("!")[0]
        if "::" not in line:
            continue
        for item in line.split("::")[1].split(","):
            item = item.strip()
            if item and not item[0].isdigit() and not item.endswith(")") or "(" in item:
                out.append(re.sub(r"\(.*|\s*=.*", "", item).strip())
    return [m for m in out if m]


def test_struct_layouts_agree_between_c_ctypes_and_fortran(tmp_path):
    """A C program that includes include/wpsgpu.h and is LINKED with -lwpsgpu prints sizeof / offsetof of every struct; the
    ctypes mirrors of wps_b200/capi.py must have exactly those layouts, and the bind(C) derived types of the Fortran module
    must list the same members in the same order (bind(C) then guarantees the same layout)."""
    import json
    import subprocess
    from wps_b200 import capi
    exe = str(tmp_path / "abi_layout")
    libdir = os.path.join(ROOT, "wps_b200")
    cuda_lib = "/usr/local/cuda/lib64"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "abi_layout.c"),
                           "-o", exe, "-L", libdir, "-lwpsgpu", "-L", cuda_lib, "-lcudart", "-Wl,-rpath," + libdir, "-Wl,-rpath," + cuda_lib])
    lay = json.loads(subprocess.check_output([exe]).decode())
    assert lay["map_set_ok"] == 0 and lay["map_set_without_dx"] != 0 and lay["error_text_length"] > 0
    assert abs(lay["cone"] - 0.715567) < 1e-5                   # cone factor of a 30 / 60 Lambert projection
    pairs = {"wpsgpu_proj": capi.Proj, "wpsgpu_map_args": capi.MapArgs, "wpsgpu_field_opts": capi.FieldOpts, "wpsgpu_slab": capi.Slab,
             "wpsgpu_geo_field": capi.GeoField, "wpsgpu_geo_level": capi.GeoLevel, "wpsgpu_mpas_mesh": capi.MpasMesh}
    for cname, ct in pairs.items():
        assert C.sizeof(ct) == lay["sizeof(%s)" % cname], cname
        cfields = [k.split(".")[1] for k in lay if k.startswith(cname + ".")]
        assert cfields == [f[0] for f in ct._fields_], cname
        for fname in cfields:
            assert getattr(ct, fname).offset == lay["%s.%s" % (cname, fname)], (cname, fname)
        assert _f90_type_members(cname) == cfields, (cname, _f90_type_members(cname), cfields)
