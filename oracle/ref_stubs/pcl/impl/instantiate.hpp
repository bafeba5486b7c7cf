// Stand-in for <pcl/impl/instantiate.hpp> (not installed): everything the reference needs is in kplref_env.h.
#pragma once
#include "kplref_env.h"
