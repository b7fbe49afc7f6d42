/* orc_internal.h -- CPU ORACLE (test infrastructure): shared internals. Not product code. */
#ifndef ORC_INTERNAL_H
#define ORC_INTERNAL_H
#include "wps_oracle.h"
#include <math.h>
float o_sin(float x); float o_cos(float x); float o_tan(float x); float o_asin(float x);
float o_atan(float x); float o_atan2(float y, float x); float o_log(float x); float o_log10(float x);
float o_exp(float x); float o_acosf_m1(void); float o_pow(float x, float y); float o_sqrt(float x);
#define ORC_PI_F 3.141592653589793f
#define ORC_DEG_PER_RAD (180.0f / ORC_PI_F)
#define ORC_RAD_PER_DEG (ORC_PI_F / 180.0f)
/* Fortran NINT on default REAL: round half away from zero */
static inline int o_nint(float x) { return (int)lroundf(x); }
static inline int o_floor(float x) { return (int)floorf(x); }
static inline int o_ceiling(float x) { return (int)ceilf(x); }
#endif
