// TEST INFRASTRUCTURE ONLY — the two symbols of the reference's utils/maths.hpp that the path uses: ln(10)/10 (pair_hmm.hpp:94;
// the reference's own literal, utils/maths.hpp:41) and the two-argument log_sum_exp. The real header needs Boost.Math.
#ifndef REF_SHIM_MATHS_HPP
#define REF_SHIM_MATHS_HPP
#include <algorithm>
#include <cmath>
namespace octopus { namespace maths {
namespace constants {
template <typename T = double>
constexpr T ln10Div10 = T {0.230258509299404568401799145468436420760110148862877297603};
} // namespace constants
// the two-argument overload haplotype_likelihood_model.cpp uses for the mapping-quality mixing (utils/maths.hpp:292-298)
inline double log_sum_exp(const double a, const double b)
{
    const auto r = std::minmax(a, b);
    return r.second + std::log1p(std::exp(r.first - r.second));
}
} } // namespace octopus::maths
#endif
