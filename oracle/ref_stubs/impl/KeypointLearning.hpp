// Stand-in for the reference's include/impl/KeypointLearning.hpp when compiling ONLY its helper translation
// unit src/KeypointLearning.cpp (findAnnulusPair / findBinPair) for oracle/_ref.  The real header needs PCL
// and OpenCV; the two helpers need <cmath>, assert and abs().  On the authors' platform (MSVC 2015,
// README.md:70) `abs(float)` is the floating-point overload from <cmath>; g++ would pick ::abs(int) for the
// unqualified call, so std::abs is made visible here ("abs() overload").
#pragma once
#include <cassert>
#include <cmath>
#include <cstdlib>
using std::abs;
using std::floor;
