// TEST INFRASTRUCTURE ONLY — stand-in for config/common.hpp: haplotype_likelihood_model.hpp includes it but uses none of it.
#ifndef REF_SHIM_CONFIG_COMMON_HPP
#define REF_SHIM_CONFIG_COMMON_HPP
#include <string>
namespace octopus { using SampleName = std::string; }
#endif
