// TEST INFRASTRUCTURE ONLY — stand-in for boost::numeric_cast (custom_repeat_based_indel_error_model.cpp: int → int8 penalty).
#ifndef REF_SHIM_BOOST_NUMERIC_CAST_HPP
#define REF_SHIM_BOOST_NUMERIC_CAST_HPP
#include <limits>
#include <stdexcept>
namespace boost { namespace numeric {
struct bad_numeric_cast : std::runtime_error { bad_numeric_cast() : std::runtime_error {"bad numeric conversion"} {} };
} // namespace numeric
template <typename Target, typename Source>
inline Target numeric_cast(const Source value)
{
    if (value < static_cast<Source>(std::numeric_limits<Target>::lowest()) || value > static_cast<Source>(std::numeric_limits<Target>::max())) throw numeric::bad_numeric_cast {};
    return static_cast<Target>(value);
}
} // namespace boost
#endif
