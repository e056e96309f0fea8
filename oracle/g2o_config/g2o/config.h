/* Hand-written stand-in for the header g2o's CMake would generate from Dependencies/g2o/config.h.in with the options
 * the reference passes (reference Dependencies/CMakeLists.txt:14-22): static libraries, no OpenGL / OpenMP / CHOLMOD /
 * CSPARSE, G2O_NO_IMPLICIT_OWNERSHIP_OF_OBJECTS=ON, double precision. TEST INFRASTRUCTURE (oracle/_ref build only). */
#ifndef G2O_CONFIG_H
#define G2O_CONFIG_H

#define G2O_NO_IMPLICIT_OWNERSHIP_OF_OBJECTS
#define G2O_DELETE_IMPLICITLY_OWNED_OBJECTS 0

#define G2O_NUMBER_FORMAT_STR "%lg"
#ifdef __cplusplus
using number_t = double;
#else
typedef double number_t;
#endif

#define G2O_CXX_COMPILER "g++"

#ifdef __cplusplus
#include <g2o/core/eigen_types.h>
#endif

#endif
