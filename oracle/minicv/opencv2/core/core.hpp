// oracle/minicv shim header (test infrastructure): forwards to minicv.hpp
#include "minicv.hpp"
