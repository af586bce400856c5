// stub: GSL is not available in the build container; only the type name is needed to compile the headers
#ifndef ORACLE_GSL_STUB
#define ORACLE_GSL_STUB
struct gsl_interp2d { int unused; };
struct gsl_interp_accel { int unused; };
#endif
