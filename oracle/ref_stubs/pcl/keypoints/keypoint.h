// Stand-in for <pcl/keypoints/keypoint.h> (not installed): everything the reference needs is in kplref_env.h.
#pragma once
#include "kplref_env.h"
