/*
 * fpb200_match - C ABI of the B200-native RANSAC minutiae matcher.
 *
 * Replaces, for batches of template pairs, the arithmetic of the reference's
 *   src/matching/match.py:219-275   match_minutiae_pair
 *   src/matching/match.py:129-217   ransac_align_and_match_parallel
 *   src/matching/match.py:75-127    ransac_worker            (one hypothesis, seeded default_rng(42 + i))
 *   src/matching/match.py:33-72     match_with_transform     (KDTree nearest neighbour + gates + score)
 *   src/matching/match.py:10-22     compute_descriptor_weight
 * as called pair by pair from src/matching/FRR.py:105-118 and src/matching/FAR.py:5-24.
 *
 * One matcher handle = one CUDA device + one stream.  Templates (the float64 [n,7] arrays of
 * src/matching/match_features.py:52-62: x, y, type(0/1), orientation, quality, coherence,
 * angular_stability) are uploaded once; pairs are lists of template indices.  The hypotheses of a
 * pair are consumed in SEED ORDER (the reference consumes them in thread-completion order, which is
 * timing-dependent; seed order is what it computes when its futures complete in submission order).
 *
 * Plain C types, status codes (FPB_OK / FPB_E_* of fpb200.h), no exceptions, no CPU path.
 */
#ifndef FPB200_MATCH_H
#define FPB200_MATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fpb_matcher fpb_matcher;

/* keyword arguments of match_minutiae_pair (match.py:219-231); the defaults are the reference's */
typedef struct fpb_match_params {
    double dist_thresh;        /* 10.0 */
    double orient_thresh_deg;  /* 12.0 */
    int    use_type;           /* 1    */
    int    ransac_iter;        /* 300  */
    int    min_inliers;        /* 8    */
    double stop_inlier_ratio;  /* 0.25 */
    int    cross_check;        /* 1    */
} fpb_match_params;

/* the dict match_minutiae_pair returns (match.py:268-274) */
typedef struct fpb_match_result {
    double  final_score;
    double  inlier_ratio;
    double  theta;             /* 0.0 when no model was accepted */
    double  tx, ty;
    int32_t n_matches;         /* len(result["matches"]) */
    int32_t best_iter;         /* hypothesis index that was refined, -1 if none */
} fpb_match_result;

/* max_minutiae <= 256 per template, max_iter <= 4096 hypotheses per pair */
int  fpb_match_create(fpb_matcher** out, int device, int max_templates, int max_minutiae, int max_iter);
void fpb_match_destroy(fpb_matcher* m);
const char* fpb_match_last_error(const fpb_matcher* m);     /* m may be NULL: error of fpb_match_create */

/* templates: n arrays of counts[i] rows x 7 float64 columns, concatenated (row-major); type column must be 0 or 1.
 * Computes on the device, per template: descriptor weights, their sum, the position spread of the early reject
 * (match.py:85-88) and the weighted picks of every hypothesis seed (match.py:93,100). */
int  fpb_match_set_templates(fpb_matcher* m, const double* mins, const int32_t* counts, int n);

/* pairs: n_pairs x (index_a, index_b).  results[n_pairs]; matches (optional, may be NULL):
 * [n_pairs, max_minutiae, 2] int32 (ia, ib), match_scores [n_pairs, max_minutiae] float64.  Synchronous. */
int  fpb_match_pairs(fpb_matcher* m, const int32_t* pairs, int n_pairs, const fpb_match_params* p,
                     fpb_match_result* results, int32_t* matches, double* match_scores);

/* device-resident variant for benchmarking: pairs already uploaded by fpb_match_upload_pairs; asynchronous on the
 * matcher's stream; results stay on the device until fpb_match_download */
int  fpb_match_upload_pairs(fpb_matcher* m, const int32_t* pairs, int n_pairs);
int  fpb_match_run_device(fpb_matcher* m, const fpb_match_params* p);
int  fpb_match_download(fpb_matcher* m, fpb_match_result* results, int32_t* matches, double* match_scores);
int  fpb_match_sync(fpb_matcher* m);
void* fpb_match_stream(fpb_matcher* m);                     /* cudaStream_t, for CUDA-event timing */
long long fpb_match_launch_count(const fpb_matcher* m);

/* host utility (no GPU needed): the two uniform doubles each hypothesis draws, i.e.
 * numpy.random.default_rng(seed0 + i).random() twice (SeedSequence + PCG64 restated); out[n_iter][2] */
int  fpb_match_seed_uniforms(uint64_t seed0, int n_iter, double* out);

#ifdef __cplusplus
}
#endif
#endif /* FPB200_MATCH_H */
