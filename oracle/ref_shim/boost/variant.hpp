// TEST INFRASTRUCTURE ONLY — stand-in for <boost/variant.hpp> (Boost is not in this image) so that the UNMODIFIED reference
// header src/core/models/pairhmm/simd_pair_hmm_wrapper.hpp compiles from where it lies. The wrapper only default-constructs
// a variant, assigns alternatives to it and visits it with generic lambdas (:32, :77-203, :222-229): std::variant does all three.
#ifndef REF_SHIM_BOOST_VARIANT_HPP
#define REF_SHIM_BOOST_VARIANT_HPP
#include <utility>
#include <variant>
namespace boost {
template <typename... Ts> using variant = std::variant<Ts...>;
template <typename Visitor, typename Variant>
decltype(auto) apply_visitor(Visitor&& visitor, Variant&& v) { return std::visit(std::forward<Visitor>(visitor), std::forward<Variant>(v)); }
} // namespace boost
#endif
