// Test-infrastructure shim (NOT product code); see shim_all.hpp
#include "shim_all.hpp"
