/* nip_model_export.h — host glue: snapshot a parsed NIP model into the flat
 * nipgpu_model_desc consumed by the device library (include/nipgpu.h).
 *
 * Compiled against the reference's own headers (src/nip.h); it only READS the
 * structures that parse_model() built (src/nip.c:122-294) and performs no
 * potential arithmetic.
 */
#ifndef NIP_MODEL_EXPORT_H
#define NIP_MODEL_EXPORT_H

#include "nip.h"
#include "nipgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Allocates a description (and all its arrays) for `model`.  Returns NULL on
 * allocation failure or if the model is inconsistent.  Free with
 * nipgpu_desc_free(). */
nipgpu_model_desc* nipgpu_desc_from_model(nip_model model);
void nipgpu_desc_free(nipgpu_model_desc* d);

/* The model's sepsets in the order the description numbers them (walk of every
 * clique's sepset list, first occurrence wins).  Caller frees the array. */
nip_sepset* nipgpu_model_sepsets(nip_model model, int* n);

/* position of `v` in model->variables[], or -1 */
int nipgpu_var_index(nip_model model, nip_variable v);

/* Copies parameters in the description's layout back into the host model:
 * clique->original_p (and ->p) and variable->prior, as m_step leaves them
 * (src/nip.c:2032-2067). */
void nipgpu_desc_store_parameters(nip_model model, const nipgpu_model_desc* d,
                                  const double* clique_tables, const double* var_prior);

#ifdef __cplusplus
}
#endif
#endif
