// The one translation unit that instantiates the C++ shim (include/lbfgsb200_compat.hpp).
// oracle/Makefile links it with the UNMODIFIED sequential-implementation/main.cpp + benchmark.cpp so
// the reference's own driver runs on the B200 drop-in (tests/test_gpu_compat.py).
#define LBFGSB200_COMPAT_IMPLEMENTATION
#include "../include/lbfgsb200_compat.hpp"
