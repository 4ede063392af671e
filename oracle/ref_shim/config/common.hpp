// TEST INFRASTRUCTURE ONLY — stand-in for config/common.hpp: the sample-keyed read containers the array's populate() takes
// (config/common.hpp:27-37: MappableMap<SampleName, ...> over flat multisets; here ordered maps of vectors — populate only
// iterates them and indexes samples in iteration order).
#ifndef REF_SHIM_CONFIG_COMMON_HPP
#define REF_SHIM_CONFIG_COMMON_HPP
#include <map>
#include <string>
#include <vector>
#include "basics/aligned_read.hpp"
#include "basics/aligned_template.hpp"
namespace octopus {
using SampleName = std::string;
template <typename K, typename V> using MappableMap = std::map<K, std::vector<V>>;
using ReadMap     = MappableMap<SampleName, AlignedRead>;
using TemplateMap = MappableMap<SampleName, AlignedTemplate>;
} // namespace octopus
#endif
