// utils/eigen_utils.hpp — Eigen::aligned_vector and friends (L/include/utils/eigen_utils.hpp) live in the Eigen shim of this tree
#include "../Eigen/Dense"
