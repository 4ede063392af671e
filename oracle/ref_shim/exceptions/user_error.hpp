// TEST INFRASTRUCTURE ONLY — stand-in for exceptions/user_error.hpp (included by simd_pair_hmm_wrapper.hpp:11).
#ifndef REF_SHIM_USER_ERROR_HPP
#define REF_SHIM_USER_ERROR_HPP
#include "error.hpp"
namespace octopus {
class UserError : public Error
{
    std::string do_type() const override { return "user"; }
public:
    virtual ~UserError() = default;
};
} // namespace octopus
#endif
