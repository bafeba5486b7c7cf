// pcl_shim.h -- the handful of PCL 1.8 types and calls the TestDetector path touches, without PCL.
//
// The reference's public class derives from pcl::Keypoint and its driver uses pcl::PointCloud,
// pcl::io::loadPCDFile / savePCDFileASCII, pcl::NormalEstimation and pcl::UniformSampling
// (reference: src/main_test_detector.cpp:37-45,93-95,142-169,212-216).  PCL, FLANN, Eigen and Boost are
// not available in this build, so this header provides layout- and name-compatible stand-ins with the
// same memory layout (PointXYZ 16 B, Normal 32 B, PointXYZI 32 B), which is what lets the detector hand
// `cloud->points.data()` to the C ABI with zero copies.  A build that HAS PCL drops this header and
// includes the real ones: KeypointLearning.h only needs the names below.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "kpl.h"

namespace pcl {

struct alignas(16) PointXYZ {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    PointXYZ() = default;
    PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) Normal {
    union { struct { float normal_x, normal_y, normal_z; }; float normal[3]; };
    float pad0 = 0;
    float curvature = 0;
    float pad1[3] = {0, 0, 0};
    Normal() : normal_x(0), normal_y(0), normal_z(0) {}
};
struct alignas(16) PointXYZI {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    float intensity = 0;
    float pad1[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(Normal) == 32 && sizeof(PointXYZI) == 32, "PCL point layouts");

inline bool isFinite(const PointXYZ& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
inline bool isFinite(const PointXYZI& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
inline bool isFinite(const Normal& n) { return std::isfinite(n.normal_x) && std::isfinite(n.normal_y) && std::isfinite(n.normal_z); }

template <typename PointT>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    float sensor_origin_[4] = {0, 0, 0, 0};                 // PCD VIEWPOINT translation
    float sensor_orientation_[4] = {1, 0, 0, 0};            // PCD VIEWPOINT quaternion (w x y z)
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    bool isOrganized() const { return height > 1; }
    void reserve(size_t n) { points.reserve(n); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
};

struct PointIndices {
    typedef std::shared_ptr<PointIndices> Ptr;
    typedef std::shared_ptr<const PointIndices> ConstPtr;
    std::vector<int> indices;
};
typedef PointIndices::Ptr PointIndicesPtr;
typedef PointIndices::ConstPtr PointIndicesConstPtr;

namespace search {
// Only a tag here: every neighbour search of this build runs on the uniform grid inside libkpl_b200.
template <typename PointT>
class KdTree {
public:
    typedef std::shared_ptr<KdTree<PointT>> Ptr;
    explicit KdTree(bool sorted = true) : sorted_(sorted) {}
    bool sorted_;
};
}  // namespace search

namespace io {
// PCD v0.7 reader: FIELDS containing x y z (any extra fields are skipped), DATA ascii | binary |
// binary_compressed.  Returns 0 on success, -1 on failure (as pcl::io::loadPCDFile).
int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud);
// ASCII writer, FIELDS x y z intensity, precision 8 (pcl::io::savePCDFileASCII<PointXYZI>, main_test_detector.cpp:215).
int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZI>& cloud);
int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZ>& cloud);
int savePCDFileBinary(const std::string& path, const PointCloud<PointXYZ>& cloud);
}  // namespace io

// pcl::NormalEstimation as TestDetector uses it (main_test_detector.cpp:162-169): k-NN or radius PCA
// normals oriented towards the viewpoint, computed on the GPU through kpl_normals().
template <typename PointInT, typename NormalT>
class NormalEstimation {
public:
    typedef typename PointCloud<PointInT>::ConstPtr PointCloudConstPtr;
    void setInputCloud(const PointCloudConstPtr& c)
    {
        input_ = c;
        if (c) { vp_[0] = c->sensor_origin_[0]; vp_[1] = c->sensor_origin_[1]; vp_[2] = c->sensor_origin_[2]; }   // use_sensor_origin_
    }
    void setKSearch(int k) { k_ = k; }
    void setRadiusSearch(double r) { radius_ = r; }
    void setSearchMethod(const typename search::KdTree<PointInT>::Ptr&) {}
    void setViewPoint(float x, float y, float z) { vp_[0] = x; vp_[1] = y; vp_[2] = z; }
    // Returns false (and leaves `out` empty) on failure; the message is printed like PCL_ERROR would.
    bool compute(PointCloud<NormalT>& out, std::string* err = nullptr);
private:
    PointCloudConstPtr input_;
    int k_ = 0;
    double radius_ = 0.0;
    float vp_[3] = {0, 0, 0};
};

// pcl::UniformSampling (main_test_detector.cpp:145-157): one point per leaf-sized voxel, the one
// closest to the voxel centre, computed on the device (kpl_uniform_sample).  PCL emits the survivors in
// hash-map order (platform dependent); here they keep their original relative order.
template <typename PointT>
class UniformSampling {
public:
    typedef std::shared_ptr<UniformSampling<PointT>> Ptr;
    void setRadiusSearch(double leaf) { leaf_ = leaf; }
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
    void filter(PointCloud<PointT>& out);
private:
    typename PointCloud<PointT>::ConstPtr input_;
    double leaf_ = 0.0;
};

}  // namespace pcl

// Stand-in for cv::Mat in computePointsForTrainingFeatures(): CV_32F rows x cols, row-major.
struct KplFeatureMatrix {
    int rows = 0, cols = 0;
    std::vector<float> data;
    bool empty() const { return rows == 0; }
    float at(int r, int c) const { return data[(size_t)r * cols + c]; }
};
