// TEST INFRASTRUCTURE ONLY — see error.hpp (stand-in for the reference header of this name).
#include "error.hpp"
