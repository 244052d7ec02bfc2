// g2o/solvers/csparse/linear_solver_csparse.h -> the B200 backend facade (graph_slam_b200/host/g2o_lite.h): g2o/g2o_graph.cpp of the reference compiles unchanged against it
#pragma once
#include "../../../../graph_slam_b200/host/g2o_lite.h"
