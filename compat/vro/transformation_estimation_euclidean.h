// vro stand-in: rgbdslam's closed-form rigid alignment header; only included, nothing of it is used on the path.
#pragma once
#include "camera_node.h"
