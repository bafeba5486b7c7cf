// Stand-in for <pcl/features/integral_image_normal.h> (not installed): everything the reference needs is in kplref_env.h.
#pragma once
#include "kplref_env.h"
