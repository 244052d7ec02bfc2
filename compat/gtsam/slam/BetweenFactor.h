// gtsam/slam/BetweenFactor.h -> the B200 backend facade (graph_slam_b200/host/gtsam_lite.h): the reference sources compile unchanged against it
#pragma once
#include "../../../graph_slam_b200/host/gtsam_lite.h"
