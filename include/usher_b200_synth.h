/* usher_b200_synth.h — seeded synthetic MAT / sample generator for the benchmark configs of
 * BASELINE.json (recipe: SURVEY.md §8(d)).  Bench and test tooling; not part of the drop-in boundary. */
#ifndef USHER_B200_SYNTH_H
#define USHER_B200_SYNTH_H
#include <stdint.h>
#include "usher_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

#define UB200_SYNTH_UNIFORM 0 /* parent = rng()%i */
#define UB200_SYNTH_SC2 1     /* 60%: attach to one of the 64 most recent nodes (SARS-CoV-2-like backbones) */

#define UB200_FAMILY_SNV40 0  /* config 2/4: <=36 of the origin node's sites + private SNVs up to 40 calls */
#define UB200_FAMILY_LEAF 1   /* config 3: origin genotype minus <=2 sites plus <=5 private SNVs */
#define UB200_FAMILY_AMBIG 2  /* config 5: SNV40 + 10% IUPAC-widened calls + 1-4 N-runs (mean length 200) */

typedef struct ub200_synth ub200_synth;

int ub200_synth_mat_create(uint32_t n_nodes, double mu, uint32_t genome_len, int shape, uint64_t seed,
                           ub200_synth** out);
void ub200_synth_free(ub200_synth* g);
/* Borrowed pointers into the generator's arrays (valid until ub200_synth_free). */
int ub200_synth_flat(ub200_synth* g, ub200_flat_mat* out);
const uint8_t* ub200_synth_reference(ub200_synth* g); /* [genome_len+1] one-hot, index = position */
/* Generate a sample batch; returned arrays are owned by the generator and overwritten by the next call. */
int ub200_synth_samples(ub200_synth* g, uint32_t n_samples, int family, uint64_t seed, const uint64_t** sample_ptr,
                        const ub200_mutation** calls, const uint32_t** origin);

#ifdef __cplusplus
}
#endif
#endif
