/* oracle/config/qdldl_types.h -- TEST INFRASTRUCTURE (oracle), not product code.
 * Hand-written stand-in for the cmake-generated qdldl_types.h (reference:
 * osqp_sources/lin_sys/direct/qdldl/qdldl_sources/configure/qdldl_types.h.in:1-26,
 * defaults from qdldl_sources/CMakeLists.txt:56-73: double / int / unsigned char).
 */
#ifndef QDLDL_TYPES_H
#define QDLDL_TYPES_H
#include <limits.h>
typedef int           QDLDL_int;
typedef double        QDLDL_float;
typedef unsigned char QDLDL_bool;
#define QDLDL_INT_MAX INT_MAX
#endif
