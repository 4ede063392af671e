// TEST INFRASTRUCTURE ONLY — stand-in for <boost/optional.hpp> (Boost is not in this image) for the UNMODIFIED reference file
// src/core/models/haplotype_likelihood_model.{hpp,cpp}, which uses optional<T>, none, operator bool, * and -> only.
#ifndef REF_SHIM_BOOST_OPTIONAL_HPP
#define REF_SHIM_BOOST_OPTIONAL_HPP
#include <optional>
namespace boost {
template <typename T> using optional = std::optional<T>;
using none_t = std::nullopt_t;
inline constexpr none_t none = std::nullopt;
} // namespace boost
#endif
