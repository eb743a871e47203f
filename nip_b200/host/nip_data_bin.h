/* nip_data_bin.h — packed data files behind the unchanged `time_series` struct.
 *
 * read_timeseries() (src/nip.c:512-667) goes through the text tokeniser twice and allocates per
 * token (src/nipparsers.c:122-350); at the size of config C5 (50 M rows) that dwarfs the GPU
 * time.  These two functions store / load the same information as packed int32 state indices
 * (SURVEY section 8 f.2).  The loaded series are ordinary `time_series` objects: everything in
 * nip.h works on them, free them with free_timeseries().
 *
 * File layout (little endian): "NIPB", int32 version = 1, int32 n_series, int32 n_columns,
 * n_columns x { int32 len, len bytes of the variable symbol }, int32 length[n_series],
 * int32 data[sum(length)][n_columns]  (state index, -1 = missing). */
#ifndef NIP_DATA_BIN_H
#define NIP_DATA_BIN_H

#include "nip.h"

/* 0 on success, a NIP error code otherwise; every series must observe the same variables */
int nip_gpu_write_timeseries_bin(time_series* set, int n, const char* filename);
/* like read_timeseries(): returns the number of series (0 on failure) and the array in *results;
 * columns whose symbol the model does not know are skipped; the set is registered for
 * transparent batching (nip_gpu_register_set) */
int nip_gpu_read_timeseries_bin(nip_model model, const char* filename, time_series** results);

#endif
