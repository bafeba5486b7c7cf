// C entry points over the reference's OWN code, compiled from /root/reference into
// oracle/_ref/libkpl_ref.so by oracle/Makefile (target `ref`):
//   * src/KeypointLearning.cpp            -- findAnnulusPair / findBinPair (a separate translation unit);
//   * include/KeypointLearning.h + include/impl/KeypointLearning.hpp -- the detector templates, instantiated
//     here as in src/main_test_detector.cpp:93-95,123 against the stand-in environment ref_stubs/kplref_env.h
//     (which documents exactly what is real and what stands in for PCL / FLANN / Eigen / OpenCV).
// Test infrastructure: tests/test_oracle.py pins the oracle's restatement against these entry points.
#define PCL_NO_PRECOMPILE
#include <algorithm>
#include "KeypointLearning.h"

typedef pcl::keypoints::KeypointLearningDetector<pcl::PointXYZ, pcl::PointXYZI> Detector;
#define KPLREF_API extern "C" __attribute__((visibility("default")))

KPLREF_API void kplref_set_normalize_mode(int reciprocal) { Eigen::kplref_normalize_reciprocal() = reciprocal; }

KPLREF_API void kplref_find_annulus_pair(int n_annulus, float distance, float support, int* index, int* pair, float* weight)
{
    findAnnulusPair(n_annulus, distance, support, *index, *pair, *weight);
}
KPLREF_API void kplref_find_bin_pair(int n_bins, float cosine, int* index, int* pair, float* weight)
{
    findBinPair(n_bins, cosine, *index, *pair, *weight);
}
// whole arrays at once, so a sweep over millions of inputs does not pay a ctypes call each
KPLREF_API void kplref_annulus_sweep(int n_annulus, float support, const float* distance, long n, int* index, int* pair, float* weight)
{
    for (long i = 0; i < n; ++i) findAnnulusPair(n_annulus, distance[i], support, index[i], pair[i], weight[i]);
}
KPLREF_API void kplref_bin_sweep(int n_bins, const float* cosine, long n, int* index, int* pair, float* weight)
{
    for (long i = 0; i < n; ++i) findBinPair(n_bins, cosine[i], index[i], pair[i], weight[i]);
}

// ---- the detector ------------------------------------------------------------------------------------
struct RefJob {
    pcl::PointCloud<pcl::PointXYZ>::Ptr cloud;
    pcl::PointCloud<pcl::Normal>::Ptr normals;
};
static RefJob make_job(const float* xyz, const float* normals4, int64_t n)
{
    RefJob J;
    J.cloud.reset(new pcl::PointCloud<pcl::PointXYZ>());
    J.normals.reset(new pcl::PointCloud<pcl::Normal>());
    J.cloud->points.resize((size_t)n);
    J.normals->points.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        pcl::PointXYZ& p = J.cloud->points[(size_t)i];
        p.x = xyz[3 * i]; p.y = xyz[3 * i + 1]; p.z = xyz[3 * i + 2]; p.pad = 1.0f;
        pcl::Normal& q = J.normals->points[(size_t)i];
        q.normal_x = normals4[4 * i]; q.normal_y = normals4[4 * i + 1]; q.normal_z = normals4[4 * i + 2]; q.pad = 0.0f;
        q.curvature = normals4[4 * i + 3]; q.pad2[0] = q.pad2[1] = q.pad2[2] = 0.0f;
    }
    J.cloud->width = (uint32_t)n; J.cloud->height = 1;
    J.normals->width = (uint32_t)n; J.normals->height = 1;
    return J;
}

// Neighbour lists served by the stand-in search: CSR (offsets[n+1], indices) for the feature radius and, when
// NMS runs, for the NMS radius.  The FIRST entry of every list must be the query itself (sorted-tree slot 0).
KPLREF_API void kplref_set_neighbours(double r_feat, const int64_t* off_feat, const int32_t* idx_feat,
                                      double r_nms, const int64_t* off_nms, const int32_t* idx_nms)
{
    pcl::KplRefSearchData& S = pcl::kplref_search_data();
    S.feat.radius = r_feat; S.feat.offsets = off_feat; S.feat.indices = idx_feat;
    S.nms.radius = off_nms ? r_nms : -1.0; S.nms.offsets = off_nms; S.nms.indices = idx_nms;
}
KPLREF_API void kplref_set_forest(int ntrees, const int32_t* roots, const int32_t* var, const float* thr,
                                  const int32_t* left, const int32_t* right, const float* value)
{
    cv::ml::KplRefForest& F = cv::ml::kplref_forest();
    F.ntrees = ntrees; F.roots = roots; F.var = var; F.thr = thr; F.left = left; F.right = right; F.value = value;
}

// computePointsForTrainingFeatures(indices) (hpp:299-318 -> computePointFeatures hpp:321-376): m x (A*B) rows
KPLREF_API int kplref_features(const float* xyz, const float* normals4, int64_t n, double r_feat, int A, int B,
                               const int32_t* qidx, int64_t m, float* out)
{
    RefJob J = make_job(xyz, normals4, n);
    Detector det;
    det.setNAnnulus(A); det.setNBins(B); det.setRadiusSearch(r_feat);
    det.setInputCloud(J.cloud); det.setNormals(J.normals);
    pcl::PointIndicesPtr ind(new pcl::PointIndices);
    ind->indices.assign(qidx, qidx + m);
    cv::Mat f = det.computePointsForTrainingFeatures(ind);
    if (f.rows != (int)m || f.cols != A * B) return -1;
    std::copy(f.d.begin(), f.d.end(), out);
    return 0;
}

// TestDetector's detector calls (main_test_detector.cpp:123-132,182-187): compute() -> initCompute ->
// detectKeypoints -> runForest + threshold + NMS.  Returns the number of keypoints; kp_idx = keypoints_indices_,
// kp_score = their intensities.  With non_maxima == 0 the output is the response cloud (hpp:189-196).
KPLREF_API int64_t kplref_detect(const float* xyz, const float* normals4, int64_t n, double r_feat, double r_nms, float threshold,
                                 int A, int B, int non_maxima, int draws_remove, float draws_threshold,
                                 int32_t* kp_idx, float* kp_score)
{
    RefJob J = make_job(xyz, normals4, n);
    Detector det;
    det.setNAnnulus(A); det.setNBins(B);
    det.setNonMaxima(non_maxima != 0); det.setNonMaxRadius(r_nms);
    det.setNonMaximaDrawsRemove(draws_remove != 0); det.setNonMaximaDrawsThreshold(draws_threshold);
    det.setPredictionThreshold(threshold);            // a float promoted to double, as in main_test_detector.cpp:118,129
    det.setRadiusSearch(r_feat);
    if (!det.loadForest("stand-in")) return -2;
    det.setInputCloud(J.cloud); det.setNormals(J.normals);
    pcl::PointCloud<pcl::PointXYZI> keypoints;
    det.compute(keypoints);
    pcl::PointIndicesConstPtr ki = det.getKeypointsIndices();
    if (!ki || ki->indices.size() != keypoints.size()) return -3;
    for (size_t k = 0; k < keypoints.size(); ++k) { kp_idx[k] = ki->indices[k]; kp_score[k] = keypoints.points[k].intensity; }
    return (int64_t)keypoints.size();
}
