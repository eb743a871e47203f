/* Force-included (gcc -include) wherever the reference's headers are compiled: the host glue
 * of this package and the reference sources themselves (oracle/Makefile).
 *
 * HEAD's niperrorhandler.h only defines NIP_NO_ERROR (src/niperrorhandler.h:32)
 * although nip.c / nipparsers.c / util/ still use the NIP_ERROR_* names
 * (e.g. src/nip.c:129, src/nip.c:1853).  The values below are the ones the
 * project's own 2010 snapshot used; only distinctness and "!= 0" matter to the
 * hot path (src/nip.c:2185-2198, util/niptrain.c:153,189).
 */
#ifndef NIP_ERRCODES_SHIM_H
#define NIP_ERRCODES_SHIM_H
#define NIP_ERROR_NULLPOINTER      1
#define NIP_ERROR_DIVBYZERO        2
#define NIP_ERROR_INVALID_ARGUMENT 3
#define NIP_ERROR_OUTOFMEMORY      4
#define NIP_ERROR_IO               5
#define NIP_ERROR_GENERAL          6
#define NIP_ERROR_FILENOTFOUND     7
#define NIP_ERROR_BAD_LUCK         8
#endif
