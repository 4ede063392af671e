// Stand-in for the reference's CMake-generated config header (src/config/system.hpp.in), which the
// AVX2/AVX-512 instruction-set policies include (avx2_pair_hmm_impl.hpp:20). Nothing from it is used
// by the pair-HMM kernel.
#pragma once
