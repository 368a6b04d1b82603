// python_ltsssolver.hpp -- flat LTSS API of the bindings, reference signatures
// (src/cpp/include/python_ltsssolver.hpp:10-23).
#ifndef PS_B200_PYTHON_LTSSSOLVER_HPP
#define PS_B200_PYTHON_LTSSSOLVER_HPP

#include <utility>
#include <vector>

#include <partitionsolvers/LTSS.hpp>

std::vector<int> find_optimal_partition__LTSS(int n, std::vector<float> a, std::vector<float> b, int parametric_dist);

float find_optimal_score__LTSS(int n, std::vector<float> a, std::vector<float> b, int parametric_dist);

std::pair<std::vector<int>, float> optimize_one__LTSS(int n, std::vector<float> a, std::vector<float> b,
                                                      int parametric_dist);

#endif
