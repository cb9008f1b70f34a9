// tests/cuda_main_prelude.hpp -- what the main() region of a parallel-implementation/L-BFGS*.cu file needs in
// scope when it is compiled WITHOUT the rest of its file (oracle/Makefile, target `cudamain`): the headers and the
// using-directive of the file's top (par/L-BFGS-Wolfe.cu:1-15), and LBFGS_CUDA with the reference's two signatures,
// which include/lbfgsb200_compat.hpp defines (tests/compat_shim_tu.cpp).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "functions.h" // parallel-implementation/functions.h, found through -I

using namespace std;

// par/L-BFGS.cu:105-112
std::vector<double> LBFGS_CUDA(const std::function<double(std::vector<double>)> f,
                               const std::function<std::vector<double>(std::vector<double>)> grad, const std::vector<double> x0,
                               const std::string line_search_method, const int max_iterations, const int m, const double tolerance);
// par/L-BFGS-Wolfe.cu:105-111 (and the other solvers with an inlined search)
std::vector<double> LBFGS_CUDA(const std::function<double(std::vector<double>)> f,
                               const std::function<std::vector<double>(std::vector<double>)> grad, const std::vector<double> x0,
                               const int max_iterations, const int m, const double tolerance);
