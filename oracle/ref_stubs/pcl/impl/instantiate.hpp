// Stand-in for <pcl/impl/instantiate.hpp> (PCL is not installed): the reference translation unit
// src/KeypointLearning.cpp only needs it for a commented-out PCL_INSTANTIATE line.
#pragma once
