/* Prototype-free stand-in for netcdf.h: lets the reference translation units that
 * include netcdf_interface.h parse.  No NetCDF function is ever called by the oracle. */
#ifndef MHH_ORACLE_NETCDF_SHIM_H
#define MHH_ORACLE_NETCDF_SHIM_H
typedef int nc_type;
#define NC_UNLIMITED 0L
#endif
