// Stand-in for <pcl/point_types.h>: nothing of it is used by findAnnulusPair / findBinPair.
#pragma once
