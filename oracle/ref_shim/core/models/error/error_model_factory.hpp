// TEST INFRASTRUCTURE ONLY — stand-in for core/models/error/error_model_factory.hpp. The path takes the error models'
// OUTPUT (per-haplotype mask / prior / gap arrays) as its input, so the "models" here just hand back arrays
// set by the test driver; the abstract interfaces they implement (snv_error_model.hpp, indel_error_model.hpp) are the
// reference's own, compiled from where they lie.
#ifndef REF_SHIM_ERROR_MODEL_FACTORY_HPP
#define REF_SHIM_ERROR_MODEL_FACTORY_HPP
#include <memory>
#include <string>
#include "core/models/error/snv_error_model.hpp"
#include "core/models/error/indel_error_model.hpp"
namespace octopus {
struct FixedPenalties   // filled by the driver before HaplotypeLikelihoodModel::reset
{
    SnvErrorModel::MutationVector forward_mask, reverse_mask;
    SnvErrorModel::PenaltyVector forward_priors, reverse_priors, gap_open, gap_extend;
};
// the arrays of the haplotype being reset: looked up by the address of the Haplotype object the model is handed (the driver
// registers the elements of the block it passes to populate), else the single default set
const FixedPenalties& fixed_penalties(const Haplotype& haplotype) noexcept;   // defined in the driver

class FixedSnvErrorModel : public SnvErrorModel
{
    std::unique_ptr<SnvErrorModel> do_clone() const override { return std::make_unique<FixedSnvErrorModel>(*this); }
    void do_evaluate(const Haplotype& haplotype, MutationVector& forward_snv_mask, PenaltyVector& forward_snv_priors,
                     MutationVector& reverse_snv_mask, PenaltyVector& reverse_snv_priors) const override
    {
        const auto& p = fixed_penalties(haplotype);
        forward_snv_mask = p.forward_mask; forward_snv_priors = p.forward_priors;
        reverse_snv_mask = p.reverse_mask; reverse_snv_priors = p.reverse_priors;
    }
};
class FixedIndelErrorModel : public IndelErrorModel
{
    std::unique_ptr<IndelErrorModel> do_clone() const override { return std::make_unique<FixedIndelErrorModel>(*this); }
    void do_set_penalties(const Haplotype& haplotype, PenaltyVector& gap_open_penalties, PenaltyType& gap_extend_penalty) const override
    {
        gap_open_penalties = fixed_penalties(haplotype).gap_open; gap_extend_penalty = fixed_penalties(haplotype).gap_extend.front();
    }
    void do_set_penalties(const Haplotype& haplotype, PenaltyVector& gap_open_penalties, PenaltyVector& gap_extend_penalties) const override
    {
        gap_open_penalties = fixed_penalties(haplotype).gap_open; gap_extend_penalties = fixed_penalties(haplotype).gap_extend;
    }
};
struct ErrorModel
{
    std::unique_ptr<IndelErrorModel> indel;
    std::unique_ptr<SnvErrorModel> snv;
};
inline std::unique_ptr<SnvErrorModel> make_snv_error_model() { return std::make_unique<FixedSnvErrorModel>(); }
inline std::unique_ptr<IndelErrorModel> make_indel_error_model() { return std::make_unique<FixedIndelErrorModel>(); }
inline ErrorModel make_error_model(const std::string&) { return ErrorModel {make_indel_error_model(), make_snv_error_model()}; }
} // namespace octopus
#endif
