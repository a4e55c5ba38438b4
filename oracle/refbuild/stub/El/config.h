/* Hand-written El/config.h for the oracle build of the reference's own sources
 * (values follow /root/reference/cmake/configure_files/config.h.in; see
 * SURVEY.md section 8c).  TEST INFRASTRUCTURE ONLY. */
#ifndef EL_CONFIG_H
#define EL_CONFIG_H
#define EL_GIT_SHA1 "oracle-build"
#define EL_VERSION_MAJOR "0"
#define EL_VERSION_MINOR "88-dev"
#define EL_CMAKE_BUILD_TYPE "Release"
#define EL_RELEASE
#define EL_CMAKE_C_COMPILER "gcc"
#define EL_CMAKE_CXX_COMPILER "g++"
#define EL_CXX_FLAGS "-O2 -std=c++14"
#define EL_FORT_LOGICAL int
#define EL_FORT_TRUE 1
#define EL_FORT_FALSE 0
#define EL_MPI_C_COMPILER "none (single-process shim)"
#define EL_MPI_C_INCLUDE_PATH ""
#define EL_MPI_C_COMPILE_FLAGS ""
#define EL_MPI_C_LIBRARIES ""
#define EL_MPI_C_LINK_FLAGS ""
#define EL_MPI_CXX_COMPILER "none (single-process shim)"
#define EL_MPI_CXX_INCLUDE_PATH ""
#define EL_MPI_CXX_COMPILE_FLAGS ""
#define EL_MPI_CXX_LIBRARIES ""
#define EL_MPI_CXX_LINK_FLAGS ""
#define EL_MATH_LIBS "scipy_openblas (LP64)"
#define EL_HAVE_BLAS_SUFFIX
#define EL_HAVE_LAPACK_SUFFIX
#define EL_BLAS_SUFFIX _
#define EL_LAPACK_SUFFIX _
#define EL_RESTRICT __restrict__
#define EL_HAVE_PRETTY_FUNCTION
#define EL_AVOID_COMPLEX_MPI
#define EL_HAVE_CXX11RANDOM
#define EL_HAVE_STEADYCLOCK
#define EL_HAVE_NOEXCEPT
#define EL_HAVE_MPI_REDUCE_SCATTER_BLOCK
#define EL_HAVE_MPI_LONG_LONG
#define EL_HAVE_MPI_COMM_SET_ERRHANDLER
#define EL_HAVE_MPI_INIT_THREAD
#define EL_HAVE_MPI_QUERY_THREAD
#define EL_USE_BYTE_ALLGATHERS
#define EL_EXPORT
#define EL_LOCAL
#endif
