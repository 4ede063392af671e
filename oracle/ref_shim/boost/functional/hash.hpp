// TEST INFRASTRUCTURE ONLY — stand-in for <boost/functional/hash.hpp>: hash_combine, used by error_model_factory.cpp to key its
// built-in parameter tables by (library, sequencer). Any well-defined hash will do (the map is only looked up).
#ifndef REF_SHIM_BOOST_FUNCTIONAL_HASH_HPP
#define REF_SHIM_BOOST_FUNCTIONAL_HASH_HPP
#include <cstddef>
#include <functional>
namespace boost {
template <typename T>
inline void hash_combine(std::size_t& seed, const T& value)
{
    seed ^= std::hash<std::size_t> {}(static_cast<std::size_t>(value)) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
}
} // namespace boost
#endif
