// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// extern "C" driver around the UNMODIFIED reference error models, compiled from where they lie (oracle/Makefile):
//   src/core/models/error/error_model_factory.cpp               built-in parameter tables (:220-523), label parsing, custom-file loader
//   src/core/models/error/{basic_,custom_,}repeat_based_indel_error_model.cpp   gap_open[] / gap_extend[] from tandem repeats
//   src/core/models/error/repeat_based_snv_error_model.cpp      SNV masks + priors per strand
//   lib/tandem/tandem.cpp + lib/tandem/libdivsufsort/*.c        the exact-tandem-repeat finder both models call
// i.e. everything HaplotypeLikelihoodModel::reset (haplotype_likelihood_model.cpp:60-78) runs per haplotype. The product's
// own implementation (octopus_b200/csrc/phmm_error_model.cpp) is checked against this, array for array.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>
#include <unistd.h>

#include "tandem/tandem.hpp"
#include "core/types/haplotype.hpp"
#include "error_model_factory.hpp"     // the reference's own (found next to the .cpp files: -I$(REF)/src/core/models/error)

namespace {

int run_reset(const octopus::ErrorModel& model, const char* seq, int n, const unsigned char* is_substitution,
              char* mask_f, std::int8_t* prior_f, char* mask_r, std::int8_t* prior_r, std::int8_t* gap_open, std::int8_t* gap_extend)
{
    octopus::Haplotype haplotype {std::string(seq, seq + n), 0};
    if (is_substitution) haplotype.set_substitutions(std::vector<bool>(is_substitution, is_substitution + n));
    int has_snv = 0;
    if (model.snv) {
        octopus::SnvErrorModel::MutationVector fm, rm;
        octopus::SnvErrorModel::PenaltyVector fp, rp;
        model.snv->evaluate(haplotype, fm, fp, rm, rp);
        std::memcpy(mask_f, fm.data(), n); std::memcpy(prior_f, fp.data(), n);
        std::memcpy(mask_r, rm.data(), n); std::memcpy(prior_r, rp.data(), n);
        has_snv = 1;
    } else {                      // haplotype_likelihood_model.cpp:69-73
        std::memset(prior_f, 100, n); std::memcpy(mask_f, seq, n);
        std::memset(prior_r, 100, n); std::memcpy(mask_r, seq, n);
    }
    octopus::IndelErrorModel::PenaltyVector go, ge;
    model.indel->set_penalties(haplotype, go, ge);
    std::memcpy(gap_open, go.data(), n); std::memcpy(gap_extend, ge.data(), n);
    return has_snv;
}

} // namespace

extern "C" {

// tandem::extract_exact_tandem_repeats(seq, min_period, max_period): out receives (pos, length, period) triples in the
// library's output order; returns the number of repeats (may exceed cap: then only cap were written)
int ref_tandem_repeats(const char* seq, int n, int min_period, int max_period, std::uint32_t* out, int cap)
{
    const std::string s(seq, seq + n);
    const auto repeats = tandem::extract_exact_tandem_repeats(s, (std::uint32_t)min_period, (std::uint32_t)max_period);
    int i = 0;
    for (const auto& r : repeats) {
        if (i < cap) { out[3 * i] = r.pos; out[3 * i + 1] = r.length; out[3 * i + 2] = r.period; }
        ++i;
    }
    return i;
}

// HaplotypeLikelihoodModel::reset's arrays for one haplotype under make_error_model(label) ("PCR-free.HiSeq-2500", ...).
// Returns 1 if the configuration has an SNV model, 0 if not (PacBio: masks = sequence, priors = 100), -1 on an unknown label.
int ref_errmodel_reset(const char* label, const char* seq, int n, const unsigned char* is_substitution,
                       char* mask_f, std::int8_t* prior_f, char* mask_r, std::int8_t* prior_r, std::int8_t* gap_open, std::int8_t* gap_extend)
{
    try {
        const auto model = octopus::make_error_model(std::string {label});
        return run_reset(model, seq, n, is_substitution, mask_f, prior_f, mask_r, prior_r, gap_open, gap_extend);
    } catch (const std::exception&) { return -1; }
}

// The custom-model path: make_error_model(file) over the given model text (error_model_factory.cpp:572-589).
int ref_errmodel_reset_custom(const char* model_text, const char* seq, int n, const unsigned char* is_substitution,
                              char* mask_f, std::int8_t* prior_f, char* mask_r, std::int8_t* prior_r, std::int8_t* gap_open, std::int8_t* gap_extend)
{
    char name[] = "/tmp/ref_errmodel_XXXXXX";
    const int fd = mkstemp(name);
    if (fd < 0) return -2;
    { std::ofstream f {name}; f << model_text; }
    close(fd);
    int rc;
    try {
        const auto model = octopus::make_error_model(boost::filesystem::path {name});
        rc = run_reset(model, seq, n, is_substitution, mask_f, prior_f, mask_r, prior_r, gap_open, gap_extend);
    } catch (const std::exception&) { rc = -1; }
    unlink(name);
    return rc;
}

} // extern "C"
