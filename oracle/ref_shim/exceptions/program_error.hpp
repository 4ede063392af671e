// TEST INFRASTRUCTURE ONLY — stand-in for exceptions/program_error.hpp (base of hmm::HMMOverflow, pair_hmm.hpp:47-65).
#ifndef REF_SHIM_PROGRAM_ERROR_HPP
#define REF_SHIM_PROGRAM_ERROR_HPP
#include "error.hpp"
namespace octopus {
class ProgramError : public Error
{
    std::string do_type() const override { return "program"; }
    std::string do_help() const override { return "internal error"; }
public:
    virtual ~ProgramError() = default;
};
} // namespace octopus
#endif
