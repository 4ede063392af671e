/* TEST INFRASTRUCTURE ONLY — plain-C restatement of the reference pair-HMM haplotype-likelihood path.
 * See phmm_oracle.c for the reference file:line each function follows. Never linked into the product. */
#ifndef PHMM_ORACLE_H
#define PHMM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_LOWEST (-1.7976931348623157e308)  /* std::numeric_limits<double>::lowest() */

/* Per-base model arrays for one truth (haplotype) sequence, already offset to the start the caller wants.
 * snv_mask == NULL selects the reference's no-SNV overloads; gap_extend == NULL uses gap_extend_scalar;
 * gap_open == NULL uses gap_open_scalar. */
typedef struct {
    const char*   snv_mask;
    const int8_t* snv_prior;
    const int8_t* gap_open;
    const int8_t* gap_extend;
    int gap_open_scalar;
    int gap_extend_scalar;
    int nuc_prior;
} oracle_model;

/* Raw kernel: reference simd::PairHMM::align score-only overloads. truth_len must be target_len + 2*band - 1. */
int oracle_align(int band, const char* truth, const char* target, const int8_t* quals,
                 int truth_len, int target_len, const oracle_model* m);

/* Raw kernel with traceback: reference align(..., first_pos, align1, align2). Buffers >= 2*(target_len+band)+1. */
int oracle_align_tb(int band, const char* truth, const char* target, const int8_t* quals,
                    int truth_len, int target_len, const oracle_model* m,
                    int* first_pos, char* align1, char* align2);

/* reference simd::PairHMM::calculate_flank_score (SNV overload when m->snv_mask != NULL). */
int oracle_flank_score(int truth_len, int lhs_flank, int rhs_flank, const char* target, const int8_t* quals,
                       const oracle_model* m, int first_pos, const char* align1, const char* align2,
                       int* target_mask_size);

/* reference hmm::detail::try_naive_evaluate. Model arrays are for the WHOLE truth (not offset).
 * Returns 1 and sets *phred when the shortcut applies (result is -ln10/10 * *phred), else 0. */
int oracle_try_naive_evaluate(const char* truth, int truth_len, const char* target, const uint8_t* quals, int target_len,
                              int target_offset, const oracle_model* m, int use_flanks, int lhs_flank, int rhs_flank,
                              int* phred);

/* reference hmm::evaluate(truth, target, quals, target_offset, hmm, params) for the MutationModel
 * (use_flanks=1: flank-aware helper, pair_hmm.hpp:723-766) or flank-less models (use_flanks=0, :699-717).
 * If dp_only != 0 the naive shortcut is skipped (what a "shortcut disabled" benchmark computes).
 * Optional outputs (may be NULL): *used_dp (0 shortcut, 1 score-only DP, 2 traceback+flank DP), *raw_score. */
double oracle_evaluate(int band, const char* truth, int truth_len, const char* target, const uint8_t* quals, int target_len,
                       int target_offset, const oracle_model* m, int use_flanks, int lhs_flank, int rhs_flank,
                       int dp_only, int* used_dp, int* raw_score);

/* reference max_score + HaplotypeLikelihoodModel::evaluate (haplotype_likelihood_model.cpp:187-304).
 * positions[0..n_positions) are candidate mapping positions, original_pos = begin_distance(haplotype, read).
 * mapq_cap_trigger < 0 means "no trigger". Returns 0 ok, 1 ShortHaplotypeError (*required_extension set). */
int oracle_model_evaluate(int band, const char* hap, int hap_len, const char* read, const uint8_t* quals, int read_len,
                          const oracle_model* m, int use_flanks, int lhs_flank, int rhs_flank,
                          const int64_t* positions, int n_positions, int64_t original_pos,
                          int use_mapping_quality, int mapping_quality, int mapq_cap, int mapq_cap_trigger,
                          int dp_only, double* out, int* required_extension);

/* reference compute_kmer_hashes<6> + populate_kmer_hash_table<6> + map_query_to_target (utils/kmer_mapper.hpp:43-159).
 * Writes up to max_positions mapping positions, returns how many. */
int oracle_kmer_map(const char* query, int query_len, const char* target, int target_len,
                    int max_positions, int64_t* out_positions);

/* reference compute_optimal_alignment + HaplotypeLikelihoodModel::align (haplotype_likelihood_model.cpp:335-431) over
 * hmm::align (pair_hmm.hpp:321-340, 784-823, 858-874) and make_cigar (:152-188). cigar: text such as "37=1X12=2I98=".
 * Returns 0 ok, 1 ShortHaplotypeError, 2 HMMOverflow. */
int oracle_model_align(int band, const char* hap, int hap_len, const char* read, const uint8_t* quals, int read_len,
                       const oracle_model* m, int use_flanks, int lhs_flank, int rhs_flank,
                       const int64_t* positions, int n_positions, int64_t original_pos,
                       int use_mapping_quality, int mapping_quality, int mapq_cap, int mapq_cap_trigger,
                       int64_t* mapping_position, double* likelihood, char* cigar, int cigar_cap, int* required_extension);

/* reference HaplotypeLikelihoodArray::populate(ReadMap) loop (haplotype_likelihood_array.cpp:51-103), candidate positions as CSR. */
int oracle_populate(int band, int H, const int64_t* hap_off, const char* seq,
                    const char* mask_f, const int8_t* prior_f, const char* mask_r, const int8_t* prior_r,
                    const int8_t* gap_open, const int8_t* gap_extend, const int64_t* hap_begin,
                    int R, const int64_t* read_off, const char* bases, const uint8_t* quals,
                    const uint8_t* mapq, const uint8_t* reverse, const int64_t* read_begin,
                    const int64_t* pos_off, const int32_t* pos,
                    int use_flanks, int lhs_flank, int rhs_flank,
                    int use_mapping_quality, int mapq_cap, int mapq_cap_trigger, int nuc_prior, int dp_only, int map_positions,
                    double* out, int32_t* status);

/* N1: ConstantMixtureGenotypeLikelihoodModel::evaluate over a [H][R] matrix for G genotypes of one ploidy. */
void oracle_genotype_likelihoods(const double* lnl, int H, int R, const int32_t* genotypes, int G, int ploidy, double* out);

#ifdef __cplusplus
}
#endif
#endif
