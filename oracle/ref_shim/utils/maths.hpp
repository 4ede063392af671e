// TEST INFRASTRUCTURE ONLY — the one symbol of the reference's utils/maths.hpp that pair_hmm.hpp uses (:94): ln(10)/10,
// with the reference's own literal (utils/maths.hpp:41). The real header needs Boost.Math.
#ifndef REF_SHIM_MATHS_HPP
#define REF_SHIM_MATHS_HPP
namespace octopus { namespace maths { namespace constants {
template <typename T = double>
constexpr T ln10Div10 = T {0.230258509299404568401799145468436420760110148862877297603};
} } } // namespace octopus::maths::constants
#endif
