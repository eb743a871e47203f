/* TEST INFRASTRUCTURE — never linked into the product.
 *
 * util/nipsample.c with one difference: the seed of rand() is an argument (the tool itself
 * seeds from the clock and the pid, src/nip.c:2480-2499, so two runs never agree).  Everything
 * else goes through the reference's public API exactly as the tool does
 * (util/nipsample.c:62-127): parse_model, random_seed, generate_data, write_timeseries.
 * Built twice by oracle/Makefile: against the reference's own code (_cpu) and with the GPU
 * backend linked in its place (_gpu), where generate_data's make_consistent calls
 * (src/nip.c:2449-2461) run on the device.  The same seed must give the same series.
 *
 *   sample_driver <model.net> <n_series> <length> <seed> <out.txt>
 */
#include <stdio.h>
#include <stdlib.h>

#include "nip.h"

int main(int argc, char* argv[]) {
  nip_model model;
  time_series* set;
  long seed;
  int n, t, i, rc;
  if (argc < 6) {
    fprintf(stderr, "usage: %s model.net n_series length seed out.txt\n", argv[0]);
    return 2;
  }
  model = parse_model(argv[1]);
  if (!model) return 1;
  n = atoi(argv[2]);
  t = atoi(argv[3]);
  seed = atol(argv[4]);
  random_seed(&seed);
  set = (time_series*)calloc((size_t)n, sizeof(time_series));
  for (i = 0; i < n; i++) {
    set[i] = generate_data(model, t);
    if (!set[i]) return 1;
  }
  rc = write_timeseries(set, n, argv[5]);
  for (i = 0; i < n; i++) free_timeseries(set[i]);
  free(set);
  free_model(model);
  return rc == NIP_NO_ERROR ? 0 : 1;
}
