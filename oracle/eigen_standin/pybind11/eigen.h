// pybind11/eigen.h STAND-IN: tracer/bvh/bvh_helper.h:12 includes it, but no Eigen type crosses the Python boundary of bvh_cpp
// (bvh_build takes and returns py::array_t), so the type casters of the real header are not needed -- only Eigen's own types.
#pragma once
#include <Eigen/Core>
