// Drop-in umbrella header for the B200-native Gemm / Cholesky / HPDSolve path:
// code that includes <El.hpp> and uses El::Grid, El::DistMatrix, El::Gemm, El::Cholesky,
// El::HPDSolve, El::Trsm, El::Herk and the Blocksize() API builds against this tree.
#pragma once
#include "elb200/core.hpp"
#include "elb200/level3.hpp"
#include "elb200/factor.hpp"
#include "elb200/io.hpp"
#include "elb200/lu.hpp"
