/* Prototype-free stand-in for netcdf.h: lets the reference translation units that
 * include netcdf_interface.h parse.  No NetCDF function is ever called by the oracle. */
#ifndef MHH_ORACLE_NETCDF_SHIM_H
#define MHH_ORACLE_NETCDF_SHIM_H
typedef int nc_type;
#define NC_UNLIMITED 0L
/* the library's default fill values (netcdf.h), referenced by src/thermo_moist.cxx:54-55 */
#define NC_FILL_FLOAT  (9.9692099683868690e+36f)
#define NC_FILL_DOUBLE (9.9692099683868690e+36)
#endif
