/* Minimal NetCDF-C stand-in used ONLY to compile the unmodified CGFD3D reference
 * (and the drop-in integration binary) in a container without libnetcdf.
 * Files are written in a simple dense container ("CGNC1", see netcdf_shim.c and
 * oracle/ncshim.py), not in real NetCDF format. Test infrastructure, not product:
 * a deployment links the real libnetcdf instead. */
#ifndef CGFD_ORACLE_NETCDF_SHIM_H
#define CGFD_ORACLE_NETCDF_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int nc_type;
#define NC_NOERR 0
#define NC_GLOBAL (-1)
#define NC_CHAR 2
#define NC_INT 4
#define NC_FLOAT 5
#define NC_DOUBLE 6
#define NC_NOWRITE 0
#define NC_WRITE 1
#define NC_CLOBBER 0
#define NC_64BIT_OFFSET 0x0200
#define NC_UNLIMITED 0L
int nc_create(const char *path, int cmode, int *ncidp);
int nc_open(const char *path, int mode, int *ncidp);
int nc_close(int ncid);
int nc_enddef(int ncid);
int nc_def_dim(int ncid, const char *name, size_t len, int *idp);
int nc_def_var(int ncid, const char *name, nc_type xtype, int ndims, const int *dimidsp, int *varidp);
int nc_put_att_int(int ncid, int varid, const char *name, nc_type xtype, size_t len, const int *op);
int nc_get_att_int(int ncid, int varid, const char *name, int *ip);
int nc_put_vara_float(int ncid, int varid, const size_t *startp, const size_t *countp, const float *op);
int nc_get_vara_float(int ncid, int varid, const size_t *startp, const size_t *countp, float *ip);
int nc_put_var_float(int ncid, int varid, const float *op);
int nc_get_var_float(int ncid, int varid, float *ip);
int nc_get_var(int ncid, int varid, void *ip);
int nc_put_var1_float(int ncid, int varid, const size_t *indexp, const float *op);
int nc_inq_varid(int ncid, const char *name, int *varidp);
int nc_inq_dimid(int ncid, const char *name, int *idp);
int nc_inq_dimlen(int ncid, int dimid, size_t *lenp);
const char *nc_strerror(int ncerr);
#ifdef __cplusplus
}
#endif
#endif
