// wpshost.cpp -- C++ host side of the metgrid path above the libwpsgpu C ABI (see include/wpshost.h).
// Mirrors, procedure by procedure, the parts of metgrid.exe that stay on the CPU around interp_met_field:
// interp_option_module.F (METGRID.TBL), read_met_module.F (intermediate format v5), the record loop of
// process_domain_module.F:999-1449 with storage_module.F's valid/modified mask semantics, and
// fill_missing_levels / fill_field / create_level / create_derived_fields (:2748-3332).
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/wpsgpu.h"
#include "../../../include/wpshost.h"

namespace wpshost {

static thread_local char g_err[1024] = "";
static void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
#define HOST_CHECK(cond, code, ...) do { if (!(cond)) { set_error(__VA_ARGS__); return (code); } } while (0)
#define GPU_TRY(call) do { int rc_ = (call); if (rc_ != 0) { set_error("%s: %s", #call, wpsgpu_last_error()); return rc_; } } while (0)

constexpr float NAN_WPS = 1.0e20f;   // misc_definitions_module.F:12
constexpr int BDR_WIDTH = 3;         // process_domain_module.F:1027

// ---- METGRID.TBL (interp_option_module.F:11-22 per-entry arrays) ---------------------------------------------
struct TableEntry {
    std::string fieldname = " ", interp_method = "nearest_neighbor", output_name = " ", from_input = "*";
    std::string interp_mask = " ", interp_land_mask = " ", interp_water_mask = " ", level_template = " ", flag_in_output = " ";
    float masked = WPSGPU_NOT_MASKED, fill_missing = NAN_WPS, missing_value = NAN_WPS;
    float interp_mask_val = NAN_WPS, interp_land_mask_val = NAN_WPS, interp_water_mask_val = NAN_WPS;
    char interp_mask_relational = ' ', interp_land_mask_relational = ' ', interp_water_mask_relational = ' ';
    int output_stagger = WPSGPU_M;
    bool output_this_field = true, is_u_field = false, is_v_field = false, is_derived_field = false;
    std::vector<std::pair<std::string, std::string>> fill_lev;   // (level string | "all", fill string), table order
};

// 'index(keyword, trim(value)) /= 0' of the Fortran parser: the value occurs inside the keyword
static bool value_in(const std::string& val, const char* keyword) { return !val.empty() && std::string(keyword).find(val) != std::string::npos; }

static float parse_real(const std::string& s) { std::string t = s; for (char& c : t) if (c == 'D' || c == 'd') c = 'e'; return strtof(t.c_str(), nullptr); }

// interp_mask = NAME(val) | NAME(<val) | NAME(>val)   (interp_option_module.F:311-348)
static void parse_mask(const std::string& value, std::string& name, float& val, char& rel)
{
    size_t p1 = value.find('('), p2 = value.find(')');
    if (p1 == std::string::npos || p2 == std::string::npos) { name = value; val = 0.0f; rel = ' '; return; }
    size_t s1 = value.find('<'), s2 = value.find('>');
    name = value.substr(0, p1);
    if (s1 != std::string::npos || s2 != std::string::npos) {
        rel = s1 != std::string::npos ? '<' : '>';
        val = parse_real(value.substr(p1 + 2, p2 - p1 - 2));
    } else {
        rel = ' ';
        val = parse_real(value.substr(p1 + 1, p2 - p1 - 1));
    }
}

// read_interp_table (interp_option_module.F:31-505): blocks separated by lines containing '=====', 'key=value'
// specifications separated by ';', '#' comments, blanks ignored; one extra trailing entry holds the defaults.
static std::vector<TableEntry> read_interp_table(const std::string& text, char gridtype = 'C')
{
    std::vector<TableEntry> entries;
    auto fresh = [&]() { TableEntry e; e.output_stagger = gridtype == 'C' ? WPSGPU_M : WPSGPU_HH; return e; };
    TableEntry cur = fresh();
    int nparams = 0;
    std::istringstream in(text);
    std::string raw;
    while (std::getline(in, raw)) {
        std::string buf;
        for (char c : raw) if (c != ' ' && c != '\t' && c != '\r') buf.push_back(c);   // despace
        if (buf.empty() || buf[0] == '#') continue;
        if (buf.find("=====") != std::string::npos) {
            if (nparams > 0) { entries.push_back(cur); cur = fresh(); }
            nparams = 0;
            continue;
        }
        size_t eos = buf.find('#');
        if (eos != std::string::npos) buf = buf.substr(0, eos);
        std::istringstream specs(buf);
        std::string spec;
        while (std::getline(specs, spec, ';')) {
            size_t idx = spec.find('=');
            if (idx == std::string::npos) continue;
            ++nparams;
            std::string key = spec.substr(0, idx), val = spec.substr(idx + 1);
            if (key == "name") cur.fieldname = val;
            else if (key == "from_input") cur.from_input = val;
            else if (key == "output_stagger") {
                const std::pair<const char*, int> st[] = { {"M", WPSGPU_M}, {"U", WPSGPU_U}, {"V", WPSGPU_V}, {"HH", WPSGPU_HH}, {"VV", WPSGPU_VV} };
                for (const auto& s : st) if (value_in(val, s.first)) { cur.output_stagger = s.second; break; }
            } else if (key == "output") cur.output_this_field = !value_in(val, "no");
            else if (key == "is_u_field") cur.is_u_field = value_in(val, "yes");
            else if (key == "is_v_field") cur.is_v_field = value_in(val, "yes");
            else if (key == "interp_option") cur.interp_method = val;
            else if (key == "interp_mask") parse_mask(val, cur.interp_mask, cur.interp_mask_val, cur.interp_mask_relational);
            else if (key == "interp_land_mask") parse_mask(val, cur.interp_land_mask, cur.interp_land_mask_val, cur.interp_land_mask_relational);
            else if (key == "interp_water_mask") parse_mask(val, cur.interp_water_mask, cur.interp_water_mask_val, cur.interp_water_mask_relational);
            else if (key == "masked") {
                if (val.find("water") != std::string::npos) cur.masked = WPSGPU_MASKED_WATER;
                else if (val.find("land") != std::string::npos) cur.masked = WPSGPU_MASKED_LAND;
                else if (val.find("both") != std::string::npos) cur.masked = WPSGPU_MASKED_BOTH;
            } else if (key == "output_name") cur.output_name = val;
            else if (key == "fill_missing") cur.fill_missing = parse_real(val);
            else if (key == "missing_value") cur.missing_value = parse_real(val);
            else if (key == "fill_lev") {          // fill_lev = [level:]spec ; no level means 'all' (:459-475)
                size_t c = val.find(':');
                if (c == std::string::npos) cur.fill_lev.emplace_back("all", val);
                else cur.fill_lev.emplace_back(val.substr(0, c), val.substr(c + 1));
            } else if (key == "level_template") cur.level_template = val;
            else if (key == "derived") cur.is_derived_field = value_in(val, "yes");
            else if (key == "flag_in_output") cur.flag_in_output = val;
            // other keys (z_dim_name, mandatory, vertical_interp_option, ...) do not touch the interpolation path
        }
    }
    if (nparams > 0) entries.push_back(cur);
    entries.push_back(fresh());   // the implicit default entry
    return entries;
}

// Index lookup of process_domain_module.F:1080-1095 / :2253-2268
static int find_entry(const std::vector<TableEntry>& t, const std::string& fieldname, const std::string& input_name)
{
    const int num = (int)t.size() - 1;
    int idxt = num;
    bool found = false;
    for (int i = 0; i < num; ++i)
        if (t[i].fieldname == fieldname) {
            if ((t[i].from_input != "*" && input_name.find(t[i].from_input) != std::string::npos) || (t[i].from_input == "*" && !found)) {
                idxt = i;
                found = true;
            }
        }
    return idxt;
}

// get_constant_fill_lev / get_fill_src_level (interp_option_module.F:736-812)
static bool constant_fill_lev(const std::string& opt, float& value)
{
    size_t i = opt.find("const");
    if (i == std::string::npos) return false;
    size_t p1 = opt.find('(', i), p2 = opt.find(')', i);
    value = (p1 != std::string::npos && p2 != std::string::npos) ? parse_real(opt.substr(p1 + 1, p2 - p1 - 1)) : NAN_WPS;
    return true;
}
static void fill_src_level(const std::string& opt, std::string& src, int& level)
{
    size_t p1 = opt.find('('), p2 = opt.find(')');
    if (p1 != std::string::npos && p2 != std::string::npos) { src = opt.substr(0, p1); level = atoi(opt.substr(p1 + 1, p2 - p1 - 1).c_str()); }
    else { src = opt; level = 1; }
}

// ---- intermediate format, version 5 (read_met_module.F:259-386) ----------------------------------------------
struct MetRecord {
    std::string field, hdate, map_source, units, desc, startloc;
    float xfcst = 0, xlvl = 0, startlat = 0, startlon = 0, deltalat = 0, deltalon = 0, dx = 0, dy = 0, xlonc = 0,
          truelat1 = 0, truelat2 = 0, nlats = 0, earth_radius = 6371.229f;
    int nx = 0, ny = 0, iproj = 0;
    bool is_wind_grid_rel = false;
    std::vector<uint32_t> raw;   // slab(1:nx,1:ny) in FILE byte order (big-endian words): decoded on the device
};

static uint32_t be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static float be_float(const unsigned char* p) { uint32_t w = be32(p); float f; memcpy(&f, &w, 4); return f; }
static std::string trimmed(const unsigned char* p, size_t n)
{
    std::string s((const char*)p, n);
    size_t a = s.find_first_not_of(' '), b = s.find_last_not_of(' ');
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

static int read_intermediate(const std::string& path, std::vector<MetRecord>& out)
{
    std::ifstream f(path, std::ios::binary);
    HOST_CHECK(f.good(), -2, "cannot open %s", path.c_str());
    std::vector<unsigned char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t pos = 0;
    auto rec = [&](const unsigned char*& p, size_t& n) -> bool {   // one Fortran sequential record
        if (pos + 4 > data.size()) return false;
        n = be32(&data[pos]);
        if (pos + 8 + n > data.size()) return false;
        p = &data[pos + 4];
        pos += 8 + n;
        return true;
    };
    while (pos < data.size()) {
        const unsigned char* p; size_t n;
        HOST_CHECK(rec(p, n) && n == 4, -2, "%s: truncated version record", path.c_str());
        int version = (int)be32(p);
        HOST_CHECK(version == 5, -2, "%s: only version 5 intermediate files are supported (got %d)", path.c_str(), version);
        HOST_CHECK(rec(p, n) && n >= 156, -2, "%s: truncated header record", path.c_str());
        MetRecord r;
        r.hdate = trimmed(p, 24); r.xfcst = be_float(p + 24); r.map_source = trimmed(p + 28, 32); r.field = trimmed(p + 60, 9);
        r.units = trimmed(p + 69, 25); r.desc = trimmed(p + 94, 46);
        r.xlvl = be_float(p + 140); r.nx = (int)be32(p + 144); r.ny = (int)be32(p + 148); r.iproj = (int)be32(p + 152);
        if (r.field == "HGT") r.field = "GHT";   // read_met_module.F:272
        HOST_CHECK(rec(p, n) && n >= 8, -2, "%s: truncated projection record", path.c_str());
        r.startloc = std::string((const char*)p, 8);
        const int nv = (int)((n - 8) / 4);
        float v[8] = {0};
        for (int k = 0; k < nv && k < 8; ++k) v[k] = be_float(p + 8 + 4 * k);
        switch (r.iproj) {
        case 0: r.startlat = v[0]; r.startlon = v[1]; r.deltalat = v[2]; r.deltalon = v[3]; r.earth_radius = v[4]; break;
        case 1: r.startlat = v[0]; r.startlon = v[1]; r.dx = v[2]; r.dy = v[3]; r.truelat1 = v[4]; r.earth_radius = v[5]; break;
        case 3: r.startlat = v[0]; r.startlon = v[1]; r.dx = v[2]; r.dy = v[3]; r.xlonc = v[4]; r.truelat1 = v[5]; r.truelat2 = v[6]; r.earth_radius = v[7]; break;
        case 4: r.startlat = v[0]; r.startlon = v[1]; r.nlats = v[2]; r.deltalon = v[3]; r.earth_radius = v[4]; break;
        case 5: r.startlat = v[0]; r.startlon = v[1]; r.dx = v[2]; r.dy = v[3]; r.xlonc = v[4]; r.truelat1 = v[5]; r.earth_radius = v[6]; break;
        default: HOST_CHECK(false, -2, "%s: unsupported projection code %d", path.c_str(), r.iproj);
        }
        HOST_CHECK(rec(p, n) && n == 4, -2, "%s: truncated is_wind_grid_rel record", path.c_str());
        r.is_wind_grid_rel = be32(p) != 0;
        HOST_CHECK(rec(p, n) && n == (size_t)r.nx * r.ny * 4, -2, "%s: slab record of %zu bytes for %d x %d", path.c_str(), n, r.nx, r.ny);
        r.raw.resize((size_t)r.nx * r.ny);
        memcpy(r.raw.data(), p, n);   // kept in file byte order
        out.push_back(std::move(r));
    }
    return 0;
}

// iproj codes of the intermediate format -> PROJ_* (read_met_module.F:286-386)
static int proj_code_of(int iproj)
{
    switch (iproj) { case 0: return WPSGPU_PROJ_LATLON; case 1: return WPSGPU_PROJ_MERC; case 3: return WPSGPU_PROJ_LC;
                     case 4: return WPSGPU_PROJ_GAUSS; case 5: return WPSGPU_PROJ_PS; default: return -1; }
}

// "Do a simple check to see whether this is a global dataset" (process_domain_module.F:1121-1122)
static bool is_global(const MetRecord& r) { return (r.iproj == 0 && fabsf((float)r.nx * r.deltalon - 360.0f) < 0.0001f) || r.iproj == 4; }

// halo_slab on the host (:1119-1140) -- only needed for the few "<NAME>.mask" fields, which reach the device as ready arrays
static std::vector<float> host_halo_slab(const MetRecord& r)
{
    const int b = is_global(r) ? BDR_WIDTH : 0, snx = r.nx + 2 * b;
    std::vector<float> s((size_t)snx * r.ny);
    for (int j = 0; j < r.ny; ++j)
        for (int i = 0; i < snx; ++i) {
            int k = i - b;
            if (k < 0) k += r.nx; else if (k >= r.nx) k -= r.nx;
            const unsigned char* p = reinterpret_cast<const unsigned char*>(&r.raw[(size_t)j * r.nx + k]);
            s[(size_t)j * snx + i] = be_float(p);
        }
    return s;
}

// ---- storage (storage_module.F; fg_input of datatype_module.F:12-64) ------------------------------------------
struct Field {
    std::string name;
    int level = 0, stagger = WPSGPU_M, nx = 0, ny = 0;
    std::vector<float> r_arr;
    std::vector<int32_t> valid_mask, modified_mask;
    bool fresh = true;
};
typedef std::pair<std::string, int> Key;

}  // namespace wpshost

using namespace wpshost;

struct wpshost_domain {
    wpsgpu_handle_t gpu = nullptr;
    int nx = 0, ny = 0, process_bdy_width = 0;
    bool has_landmask = false;
    std::vector<TableEntry> table;
    std::map<Key, std::unique_ptr<Field>> storage;
    std::vector<Key> order;   // insertion order of the keys (storage_module.F lists)

    void dims(int stagger, int& gx, int& gy) const
    {
        gx = nx + (stagger == WPSGPU_U ? 1 : 0);
        gy = ny + (stagger == WPSGPU_V ? 1 : 0);
    }
    static int nwords(int gx) { return (gx + 31) / 32; }
    Field* find(const std::string& name, int level)
    {
        auto it = storage.find(Key(name, level));
        return it == storage.end() ? nullptr : it->second.get();
    }
    Field* put(const std::string& name, int level, int stagger)
    {
        std::unique_ptr<Field> f(new Field());
        f->name = name; f->level = level; f->stagger = stagger;
        dims(stagger, f->nx, f->ny);
        f->r_arr.assign((size_t)f->nx * f->ny, 0.0f);
        f->valid_mask.assign((size_t)f->ny * nwords(f->nx), 0);
        f->modified_mask.assign((size_t)f->ny * nwords(f->nx), 0);
        Field* p = f.get();
        storage[Key(name, level)] = std::move(f);
        order.push_back(Key(name, level));
        return p;
    }
    // storage_get_levels: ascending vertical_level for time-dependent data (secondary_cmp, datatype_module.F:124)
    std::vector<int> levels_of(const std::string& name) const
    {
        std::vector<int> l;
        for (const auto& kv : storage) if (kv.first.first == name) l.push_back(kv.first.second);
        std::sort(l.begin(), l.end());
        return l;
    }
    // names in storage_get_td_headers order: a new name is linked in at the head, so the newest comes first
    std::vector<std::string> names_newest_first() const
    {
        std::vector<std::string> n;
        for (const Key& k : order) if (std::find(n.begin(), n.end(), k.first) == n.end()) n.push_back(k.first);
        std::reverse(n.begin(), n.end());
        return n;
    }

    // create_level (:3191-3332)
    int create_level(const std::string& name, int stagger, float rlevel, bool constant, float rfillconst, const std::string& srcname, float rsrclevel)
    {
        const int level = (int)lroundf(rlevel);
        if (find(name, level)) return 0;   // "already exists; leaving it alone"
        if (constant) {
            Field* f = put(name, level, stagger);
            GPU_TRY(wpsgpu_create_level(gpu, f->nx, f->ny, rfillconst, nullptr, nullptr, f->r_arr.data(), f->valid_mask.data()));
            f->fresh = false;
            return 0;
        }
        Field* src = find(srcname, (int)lroundf(rsrclevel));
        if (!src) return 0;   // "Couldn't find %s at level %f to fill level %f of %s." (INFORM only)
        int gx, gy;
        dims(stagger, gx, gy);
        HOST_CHECK(src->nx == gx && src->ny == gy, -5,
                   "Dimensions for %s do not match those of %s. This is probably because the staggerings of the fields do not match.",
                   srcname.c_str(), name.c_str());
        Field* f = put(name, level, stagger);
        GPU_TRY(wpsgpu_create_level(gpu, f->nx, f->ny, 0.0f, src->r_arr.data(), src->valid_mask.data(), f->r_arr.data(), f->valid_mask.data()));
        f->fresh = false;
        return 0;
    }

    // fill_field (:3037-3177): walk the entry's fill_lev list in table order
    int fill_field(const std::string& name, const TableEntry& e, int stagger, const std::vector<int>* all_level_list)
    {
        bool filled_all = false;
        for (const auto& kv : e.fill_lev) {
            const std::string &lev_s = kv.first, &spec = kv.second;
            float c = NAN_WPS;
            if (lev_s == "all") {
                if (filled_all) continue;
                filled_all = true;
                const bool is_const = constant_fill_lev(spec, c);
                if (is_const || spec == "vertical_index") {
                    std::vector<int> levels = all_level_list ? *all_level_list : levels_of(e.level_template);
                    for (int l : levels) { int rc = create_level(name, stagger, (float)l, true, is_const ? c : (float)l, "", 0.0f); if (rc) return rc; }
                } else {
                    std::vector<int> levels = all_level_list ? *all_level_list : levels_of(spec);
                    if (!all_level_list && levels.empty()) filled_all = false;
                    for (int l : levels) { int rc = create_level(name, stagger, (float)l, false, 0.0f, spec, (float)l); if (rc) return rc; }
                }
            } else {
                const float rlevel = parse_real(lev_s);
                if (constant_fill_lev(spec, c)) { int rc = create_level(name, stagger, rlevel, true, c, "", 0.0f); if (rc) return rc; }
                else {
                    std::string src; int sl;
                    fill_src_level(spec, src, sl);
                    int rc = create_level(name, stagger, rlevel, false, 0.0f, src, (float)sl);
                    if (rc) return rc;
                }
            }
        }
        return 0;
    }
};

extern "C" {

const char* wpshost_last_error(void) { return g_err; }

int wpshost_domain_create(int device, int nx, int ny, const float* xlat, const float* xlon, const float* xlat_u,
                          const float* xlon_u, const float* xlat_v, const float* xlon_v, const float* landmask,
                          int process_bdy_width, wpshost_domain_t* out)
{
    HOST_CHECK(out && xlat && xlon && xlat_u && xlon_u && xlat_v && xlon_v && nx > 0 && ny > 0, -1, "wpshost_domain_create: bad argument");
    std::unique_ptr<wpshost_domain> d(new wpshost_domain());
    d->nx = nx; d->ny = ny; d->process_bdy_width = process_bdy_width; d->has_landmask = landmask != nullptr;
    GPU_TRY(wpsgpu_create(device, &d->gpu));
    GPU_TRY(wpsgpu_metgrid_set_grid(d->gpu, 0, xlat, xlon, 1, nx, 1, ny));
    GPU_TRY(wpsgpu_metgrid_set_grid(d->gpu, 1, xlat_u, xlon_u, 1, nx + 1, 1, ny));
    GPU_TRY(wpsgpu_metgrid_set_grid(d->gpu, 2, xlat_v, xlon_v, 1, nx, 1, ny + 1));
    if (landmask) GPU_TRY(wpsgpu_metgrid_set_landmask(d->gpu, 0, landmask));
    GPU_TRY(wpsgpu_metgrid_set_domain(d->gpu, 1, nx, 1, ny));
    d->table = read_interp_table("");
    *out = d.release();
    return 0;
}

void wpshost_domain_destroy(wpshost_domain_t d)
{
    if (!d) return;
    if (d->gpu) wpsgpu_destroy(d->gpu);
    delete d;
}

int wpshost_set_table(wpshost_domain_t d, const char* metgrid_tbl_text)
{
    HOST_CHECK(d && metgrid_tbl_text, -1, "wpshost_set_table: NULL argument");
    d->table = read_interp_table(metgrid_tbl_text);
    return (int)d->table.size();
}

int wpshost_process_file(wpshost_domain_t d, const char* path, const char* input_name)
{
    HOST_CHECK(d && path && input_name, -1, "wpshost_process_file: NULL argument");
    std::vector<MetRecord> records;
    int rc = read_intermediate(path, records);
    if (rc) return rc;
    const std::vector<TableEntry>& table = d->table;
    auto renamed = [&](const std::string& field) {
        int idx = find_entry(table, field, input_name);
        return table[idx].output_name != " " ? table[idx].output_name.substr(0, 9) : field;
    };
    // get_interp_masks (:2061-2197): stash the "<NAME>.mask" slabs of this source first.  Only names some entry lists
    // as its interp_mask are stored (:2128); interp_land_mask / interp_water_mask find their field only if it is one of those.
    std::map<std::string, std::vector<float>> masks;
    for (const MetRecord& r : records) {
        std::string name = renamed(r.field);
        bool is_mask = false;
        for (size_t i = 0; i + 1 < table.size(); ++i) if (table[i].interp_mask != " " && table[i].interp_mask == name) is_mask = true;
        if (is_mask && !masks.count(name)) masks[name] = host_halo_slab(r);
    }
    std::vector<Field*> queued;
    for (const MetRecord& r : records) {
        int idx = find_entry(table, r.field, input_name);
        std::string name = r.field;
        if (table[idx].output_name != " ") { name = table[idx].output_name.substr(0, 9); idx = find_entry(table, name, input_name); }
        const TableEntry& e = table[idx];
        // push_source_projection (:1144-1151)
        wpsgpu_proj proj;
        const int code = proj_code_of(r.iproj);
        float starti = 1.0f, startj = 1.0f;
        if (r.startloc.substr(0, 6) == "CENTER") { starti = (float)(r.nx + 1) / 2.0f; startj = (float)(r.ny + 1) / 2.0f; }
        const float re_m = r.earth_radius * 1000.0f;
        GPU_TRY(wpsgpu_push_source_projection(code, r.xlonc, r.truelat1, r.truelat2, r.dx * 1000.0f, r.dy * 1000.0f,
                                              code == WPSGPU_PROJ_GAUSS ? r.nlats : r.deltalat, r.deltalon, starti, startj,
                                              r.startlat, r.startlon, nullptr, &re_m, &proj));
        wpsgpu_field_opts fo;
        memset(&fo, 0, sizeof(fo));
        int n = 0, has_gcell = 0;
        float thr = 1.0f;
        GPU_TRY(wpsgpu_parse_interp_option(e.interp_method.c_str(), fo.ops, fo.opts, WPSGPU_MAX_OPS, &n, &has_gcell, &thr));
        fo.missing_value = e.missing_value; fo.fill_missing = e.fill_missing; fo.masked = e.masked;
        fo.mask_val[0] = e.interp_mask_val; fo.mask_val[1] = e.interp_land_mask_val; fo.mask_val[2] = e.interp_water_mask_val;
        fo.mask_rel[0] = e.interp_mask_relational; fo.mask_rel[1] = e.interp_land_mask_relational; fo.mask_rel[2] = e.interp_water_mask_relational;
        const int stag = (e.output_stagger == WPSGPU_U || e.output_stagger == WPSGPU_V) ? e.output_stagger : WPSGPU_M;
        const int level = (int)lroundf(r.xlvl);
        Field* fld = d->find(name, level);   // storage_query_field (:1225-1233): data and valid mask carry over
        if (fld && std::find(queued.begin(), queued.end(), fld) != queued.end()) {
            // a second record of the same (field, level) in this file: the reference has finished the first one --
            // interpolation and bitarray_merge (:1225-1365) -- before it reads the second, so flush what is queued
            GPU_TRY(wpsgpu_metgrid_flush(d->gpu));
            for (Field* f : queued)
                for (size_t w = 0; w < f->valid_mask.size(); ++w) f->valid_mask[w] |= f->modified_mask[w];
            queued.clear();
        }
        const bool fresh = fld == nullptr;
        if (!fld) fld = d->put(name, level, stag);
        std::fill(fld->modified_mask.begin(), fld->modified_mask.end(), 0);   // bitarray_create(field%modified_mask)
        const int bdr = is_global(r) ? BDR_WIDTH : 0;
        wpsgpu_slab s;
        memset(&s, 0, sizeof(s));
        s.slab = reinterpret_cast<const float*>(r.raw.data());   // file bytes; halo_slab (:1119-1140) is built on the device
        s.raw_nx = r.nx; s.raw_big_endian = 1;
        s.minx = 1 - bdr; s.maxx = r.nx + bdr; s.miny = 1; s.maxy = r.ny; s.bdr = bdr;
        const std::string* mnames[3] = { &e.interp_mask, &e.interp_land_mask, &e.interp_water_mask };
        for (int m = 0; m < 3; ++m) {
            auto it = *mnames[m] == " " ? masks.end() : masks.find(*mnames[m]);
            s.mask[m] = it == masks.end() ? nullptr : it->second.data();
        }
        s.grid_id = stag == WPSGPU_U ? 1 : (stag == WPSGPU_V ? 2 : 0);
        s.field_stagger = stag;
        s.use_landmask = (stag == WPSGPU_M && d->has_landmask) ? 1 : 0;
        s.process_bdy_width = d->process_bdy_width;
        s.r_arr = fld->r_arr.data();
        s.valid_mask = fresh ? nullptr : fld->valid_mask.data();
        s.new_pts = fld->modified_mask.data();
        GPU_TRY(wpsgpu_metgrid_enqueue_slab(d->gpu, &proj, &fo, &s));
        fld->fresh = false;
        queued.push_back(fld);
    }
    GPU_TRY(wpsgpu_metgrid_flush(d->gpu));   // `records` (the raw slabs) and `masks` live until here
    for (Field* f : queued)                  // bitarray_merge(valid_mask, modified_mask) (:1365)
        for (size_t w = 0; w < f->valid_mask.size(); ++w) f->valid_mask[w] |= f->modified_mask[w];
    return (int)records.size();
}

int wpshost_create_derived_fields(wpshost_domain_t d)
{
    HOST_CHECK(d != nullptr, -1, "NULL domain");
    for (size_t i = 0; i + 1 < d->table.size(); ++i)
        if (d->table[i].is_derived_field) {
            int rc = d->fill_field(d->table[i].fieldname, d->table[i], d->table[i].output_stagger, nullptr);
            if (rc) return rc;
        }
    return 0;
}

int wpshost_fill_missing_levels(wpshost_domain_t d)
{
    HOST_CHECK(d != nullptr, -1, "NULL domain");
    std::vector<int> uni;   // union of the levels of all fields, sea level (201300) excluded (:2796-2812)
    for (const auto& kv : d->storage) if (kv.first.second != 201300 && std::find(uni.begin(), uni.end(), kv.first.second) == uni.end()) uni.push_back(kv.first.second);
    if (uni.empty()) return 0;
    std::sort(uni.begin(), uni.end());
    for (const std::string& name : d->names_newest_first()) {
        const TableEntry* e = nullptr;
        for (size_t i = 0; i + 1 < d->table.size(); ++i) if (d->table[i].fieldname == name) { e = &d->table[i]; break; }
        if (!e) continue;
        int stag = WPSGPU_M;
        for (const auto& kv : d->storage) if (kv.first.first == name) { stag = kv.second->stagger; break; }
        int rc = d->fill_field(name, *e, stag, &uni);
        if (rc) return rc;
    }
    for (const std::string& name : d->names_newest_first()) {   // vertical interpolation to missing values (:2862-2921)
        std::vector<int> levels = d->levels_of(name);
        if (levels.size() <= 1) continue;
        std::vector<float*> r;
        std::vector<int32_t*> v;
        Field* f0 = d->find(name, levels[0]);
        for (int l : levels) { Field* f = d->find(name, l); r.push_back(f->r_arr.data()); v.push_back(f->valid_mask.data()); }
        GPU_TRY(wpsgpu_fill_missing_levels(d->gpu, (int)levels.size(), levels.data(), r.data(), v.data(), f0->nx, f0->ny));
    }
    return 0;
}

int wpshost_num_fields(wpshost_domain_t d) { return d ? (int)d->order.size() : 0; }

int wpshost_field_info(wpshost_domain_t d, int index, char name[10], int32_t* level, int32_t* stagger, int32_t* nx, int32_t* ny)
{
    HOST_CHECK(d && index >= 0 && index < (int)d->order.size(), -1, "wpshost_field_info: index out of range");
    Field* f = d->find(d->order[index].first, d->order[index].second);
    if (name) { memset(name, 0, 10); strncpy(name, f->name.c_str(), 9); }
    if (level) *level = f->level;
    if (stagger) *stagger = f->stagger;
    if (nx) *nx = f->nx;
    if (ny) *ny = f->ny;
    return 0;
}

int wpshost_get_field(wpshost_domain_t d, const char* name, int32_t level, float* r_arr, int32_t* valid_mask)
{
    HOST_CHECK(d && name, -1, "wpshost_get_field: NULL argument");
    Field* f = d->find(name, level);
    if (!f) { set_error("field %s at level %d is not in storage", name, level); return 1; }
    if (r_arr) memcpy(r_arr, f->r_arr.data(), f->r_arr.size() * sizeof(float));
    if (valid_mask) memcpy(valid_mask, f->valid_mask.data(), f->valid_mask.size() * sizeof(int32_t));
    return 0;
}

}  // extern "C"
