/* wpshost.h -- C++ host side of the metgrid path above the libwpsgpu C ABI, exposed through a small C interface.
 *
 * WPS is compiled code (Fortran 90); no Fortran compiler exists in this image, so the host side of the drop-in --
 * the part of metgrid.exe that stays on the CPU around the interpolation path -- is restated in C++:
 *   - METGRID.TBL parser                     metgrid/src/interp_option_module.F:31-505, 692-812
 *   - intermediate-format reader (v5)        metgrid/src/read_met_module.F:62-398
 *   - the record loop                        metgrid/src/process_domain_module.F:999-1449 (get_interp_masks :2061-2197,
 *                                            storage semantics of storage_module.F: valid_mask / modified_mask)
 *   - fill_missing_levels / fill_field / create_level / create_derived_fields      :2748-3332
 * Every array operation goes through libwpsgpu (include/wpsgpu.h); records travel in file byte order and are
 * decoded on the device (wpsgpu_slab.raw_nx).  GRIB decoding, netCDF output, namelists and MPI stay out of scope.
 */
#ifndef WPSHOST_H
#define WPSHOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct wpshost_domain *wpshost_domain_t;

const char *wpshost_last_error(void);

/* One model domain (C grid): mass / U / V staggered lat-lon arrays as geogrid wrote them (mass nx x ny, U (nx+1) x ny,
 * V nx x (ny+1), x fastest), LANDMASK on the mass grid (may be NULL), memory and domain bounds starting at 1
 * (process_domain_module.F:405-461).  Creates its own libwpsgpu handle on `device`. */
int wpshost_domain_create(int device, int nx, int ny, const float *xlat, const float *xlon, const float *xlat_u,
                          const float *xlon_u, const float *xlat_v, const float *xlon_v, const float *landmask,
                          int process_bdy_width, wpshost_domain_t *out);
void wpshost_domain_destroy(wpshost_domain_t d);

/* read_interp_table on the text of a METGRID.TBL (interp_option_module.F:31); returns the number of entries incl. the
 * implicit default entry, or a negative error. */
int wpshost_set_table(wpshost_domain_t d, const char *metgrid_tbl_text);

/* process_intermediate_fields for one intermediate file (process_domain_module.F:999-1449): every record becomes one
 * queued slab, ONE flush.  input_name is the fg_name prefix the table's from_input is matched against.
 * Returns the number of records processed, or a negative error. */
int wpshost_process_file(wpshost_domain_t d, const char *path, const char *input_name);

/* create_derived_fields (:2939-3029) and fill_missing_levels (:2748-2933) on the fields stored so far. */
int wpshost_create_derived_fields(wpshost_domain_t d);
int wpshost_fill_missing_levels(wpshost_domain_t d);

/* Storage inspection: fields are identified by (name, nint(level)). */
int wpshost_num_fields(wpshost_domain_t d);
int wpshost_field_info(wpshost_domain_t d, int index, char name[10], int32_t *level, int32_t *stagger, int32_t *nx, int32_t *ny);
/* Copies r_arr (nx*ny floats) and valid_mask (ny*ceil(nx/32) words); either pointer may be NULL. 0 = found. */
int wpshost_get_field(wpshost_domain_t d, const char *name, int32_t level, float *r_arr, int32_t *valid_mask);

#ifdef __cplusplus
}
#endif
#endif
