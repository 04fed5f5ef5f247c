// stand-in for the config.h that g2o's CMake generates from Thirdparty/g2o/config.h.in (no OpenMP, shared library): test infrastructure
#ifndef G2O_CONFIG_H
#define G2O_CONFIG_H
#define G2O_SHARED_LIBS 1
#endif
