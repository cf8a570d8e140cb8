/* oracle HDF5 stub (test infrastructure): the hot-path translation units
 * reference no HDF5 symbol; mcrat.h merely includes the header. */
#ifndef ORACLE_HDF5_STUB_H
#define ORACLE_HDF5_STUB_H
typedef long hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#endif
