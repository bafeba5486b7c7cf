// kplref_env.h -- the environment the reference's OWN detector templates are compiled against for
// oracle/_ref (test infrastructure).  PCL 1.8, FLANN, Eigen, Boost and OpenCV are not installed, so this header
// provides the few names include/KeypointLearning.h, include/impl/KeypointLearning.hpp and
// src/KeypointLearning.cpp of the reference use.  Nothing of the reference is copied: its files are compiled
// from where they lie under /root/reference (oracle/Makefile, target `ref`).
//
// What is real and what is a stand-in:
//   * REAL (the reference's own code): the detector class, its setters, initCompute, detectKeypoints (threshold,
//     local-maximum NMS, draws-remove skip list), runForest (score = 1 - sum/ntrees), computePointFeatures
//     (slot-0 skip, non-finite normals, cosine, the four histogram updates, per-annulus normalisation, feature
//     layout), computePointsForTrainingFeatures, findAnnulusPair, findBinPair.
//   * STAND-INS for third-party behaviour, written from the pinned upstream versions:
//       pcl::search::KdTree::radiusSearch  -> neighbour lists handed in by the test (FLANN L2_Simple d2 in
//                                             FP32, strict d2 < (float)(r*r), the query itself in slot 0 as a
//                                             sorted tree returns it), in the order the test chose;
//       cv::ml::RTrees::predict(PREDICT_SUM) -> flat-array traversal, left iff x[var] <= c, double sum -> float
//                                             (pinned against the real cv2.ml.RTrees in tests/test_oracle.py);
//       Eigen::Vector3f::dot / MatrixXf row norm / normalize -> a0 + (a1 + a2); sequential sum of squares,
//                                             sqrt, division by the norm;
//       pcl::Keypoint base class           -> initCompute / compute / searchForNeighbors as in PCL 1.8.
//   * g++ two-phase lookup: the reference uses three base-class members without `this->` (name_ in loadForest,
//     input_ / surface_ in computePointsForTrainingFeatures), which MSVC accepts.  They sit on an error
//     message and on a clean-up line; namespace-scope dummies of the same names below let them compile.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

// MSVC's <cmath> puts the floating-point overloads in the global namespace (README.md:70 of the reference)
using std::abs;
using std::floor;
using std::sqrt;

#define PCL_EXPORTS
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
#define pcl_isfinite(x) std::isfinite(x)

namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
}

namespace Eigen {
struct Vector3f {
    float v[3];
    float dot(const Vector3f& o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
    Vector3f operator-(const Vector3f& o) const { return Vector3f{{v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}}; }
    float norm() const { return std::sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])); }
};
// Eigen's `row /= norm` differs between versions: 3.3+ divides every coefficient, 3.2.x (DenseBase::operator/=) multiplies
// by Scalar(1)/norm.  The reference pins PCL 1.8.0 but no Eigen version; the wrapper selects the variant under test.
inline int& kplref_normalize_reciprocal() { static int flag = 0; return flag; }
class MatrixXf {
public:
    struct Row {
        MatrixXf* m; int r;
        float norm() const { float s = 0.0f; for (int c = 0; c < m->cols_; ++c) s += (*m)(r, c) * (*m)(r, c); return std::sqrt(s); }
        void normalize()
        {
            const float n = norm();
            if (kplref_normalize_reciprocal()) { const float inv = 1.0f / n; for (int c = 0; c < m->cols_; ++c) (*m)(r, c) *= inv; }
            else for (int c = 0; c < m->cols_; ++c) (*m)(r, c) /= n;
        }
    };
    static MatrixXf Zero(int rows, int cols) { MatrixXf z; z.rows_ = rows; z.cols_ = cols; z.d.assign((size_t)rows * cols, 0.0f); return z; }
    float& operator()(int r, int c) { return d[(size_t)c * rows_ + r]; }          // column-major like Eigen
    Row row(int r) { return Row{this, r}; }
    int rows_ = 0, cols_ = 0;
    std::vector<float> d;
};
}  // namespace Eigen

namespace pcl {
struct PointXYZ { float x, y, z, pad; Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f{{x, y, z}}; } };
struct PointXYZI { float x, y, z, pad, intensity, pad2[3]; Eigen::Vector3f getVector3fMap() const { return Eigen::Vector3f{{x, y, z}}; } };
struct Normal {
    float normal_x, normal_y, normal_z, pad, curvature, pad2[3];
    Eigen::Vector3f getNormalVector3fMap() const { return Eigen::Vector3f{{normal_x, normal_y, normal_z}}; }
};
template <typename T> inline bool isFinite(const T& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }
template <> inline bool isFinite<Normal>(const Normal& n) { return std::isfinite(n.normal_x) && std::isfinite(n.normal_y) && std::isfinite(n.normal_z); }

template <typename T>
class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
    std::vector<T> points;
    uint32_t width = 0, height = 1;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    void reserve(size_t n) { points.reserve(n); }
    void push_back(const T& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
    bool isOrganized() const { return height > 1; }
};
struct PointIndices { std::vector<int> indices; };
typedef std::shared_ptr<PointIndices> PointIndicesPtr;
typedef std::shared_ptr<const PointIndices> PointIndicesConstPtr;

// The neighbour lists the stand-in search serves: CSR per radius, filled by the test through ref_wrap.cpp.
struct KplRefNeighbours { double radius = -1; const int64_t* offsets = nullptr; const int32_t* indices = nullptr; };
struct KplRefSearchData { KplRefNeighbours feat, nms; };
inline KplRefSearchData& kplref_search_data() { static KplRefSearchData d; return d; }

namespace search {
template <typename PointT>
class KdTree {
public:
    typedef std::shared_ptr<KdTree<PointT>> Ptr;
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { cloud_ = c; }
    // pcl::search::KdTree::radiusSearch(index, radius, ...) -> KdTreeFLANN -> FLANN: the lists come from the test,
    // the squared distances are FLANN's L2_Simple in FP32.
    int radiusSearch(int index, double radius, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int = 0) const
    {
        const KplRefSearchData& S = kplref_search_data();
        const KplRefNeighbours* L = (radius == S.feat.radius) ? &S.feat : ((radius == S.nms.radius) ? &S.nms : nullptr);
        k_indices.clear(); k_sqr_distances.clear();
        if (!L) { std::fprintf(stderr, "kplref: no neighbour lists for radius %g\n", radius); std::abort(); }
        const PointT& q = cloud_->points[(size_t)index];
        for (int64_t t = L->offsets[index]; t < L->offsets[index + 1]; ++t) {
            const int j = L->indices[t];
            const PointT& p = cloud_->points[(size_t)j];
            float result = 0.0f;                                   // flann::L2_Simple<float>
            const float a[3] = {q.x, q.y, q.z}, b[3] = {p.x, p.y, p.z};
            for (int i = 0; i < 3; ++i) { const float diff = a[i] - b[i]; result += diff * diff; }
            k_indices.push_back(j); k_sqr_distances.push_back(result);
        }
        return (int)k_indices.size();
    }
private:
    typename PointCloud<PointT>::ConstPtr cloud_;
};
}  // namespace search

// pcl::Keypoint<PointInT, PointOutT> of PCL 1.8 (keypoints/keypoint.h, impl/keypoint.hpp), radius search only.
template <typename PointInT, typename PointOutT>
class Keypoint {
public:
    typedef PointCloud<PointInT> PointCloudIn;
    typedef PointCloud<PointOutT> PointCloudOut;
    typedef search::KdTree<PointInT> KdTree;
    typedef typename PointCloudIn::ConstPtr PointCloudInConstPtr;
    Keypoint() : search_radius_(0), search_parameter_(0), k_(0) {}
    virtual ~Keypoint() {}
    virtual void setInputCloud(const PointCloudInConstPtr& cloud) { input_ = cloud; }
    void setRadiusSearch(double r) { search_radius_ = r; }
    PointIndicesConstPtr getKeypointsIndices() { return keypoints_indices_; }
    void compute(PointCloudOut& output)
    {
        if (!initCompute()) { PCL_ERROR("[pcl::%s::compute] initCompute failed!\n", name_.c_str()); return; }
        detectKeypoints(output);
        if (input_ == surface_) surface_.reset();
    }
    int searchForNeighbors(int index, double parameter, std::vector<int>& indices, std::vector<float>& distances) const
    {
        return tree_->radiusSearch(index, parameter, indices, distances, 0);
    }
protected:
    virtual bool initCompute()
    {
        if (!input_) return false;
        if (!tree_) tree_.reset(new KdTree());
        if (!surface_) surface_ = input_;
        tree_->setInputCloud(surface_);
        if (search_radius_ == 0.0) { PCL_ERROR("[pcl::%s::initCompute] Neither radius nor K defined!\n", name_.c_str()); return false; }
        search_parameter_ = search_radius_;
        keypoints_indices_.reset(new PointIndices);
        keypoints_indices_->indices.reserve(input_->size());
        return true;
    }
    virtual void detectKeypoints(PointCloudOut& output) = 0;
    PointCloudInConstPtr input_, surface_;
    typename KdTree::Ptr tree_;
    double search_radius_, search_parameter_;
    int k_;
    std::string name_;
    PointIndicesPtr keypoints_indices_;
};

// never executed (the tests always pass normals); declared so that initCompute's fallback branch compiles
template <typename PointInT, typename NormalT>
class NormalEstimation {
public:
    void setInputCloud(const typename PointCloud<PointInT>::ConstPtr&) {}
    void setRadiusSearch(double) {}
    void compute(PointCloud<NormalT>&) { std::fprintf(stderr, "kplref: NormalEstimation stand-in called\n"); std::abort(); }
};
template <typename PointInT, typename NormalT>
class IntegralImageNormalEstimation {
public:
    enum NormalEstimationMethod { COVARIANCE_MATRIX, AVERAGE_3D_GRADIENT, AVERAGE_DEPTH_CHANGE, SIMPLE_3D_GRADIENT };
    void setNormalEstimationMethod(NormalEstimationMethod) {}
    void setInputCloud(const typename PointCloud<PointInT>::ConstPtr&) {}
    void setNormalSmoothingSize(float) {}
    void compute(PointCloud<NormalT>&) { std::abort(); }
};
namespace keypoints {
// see the note on two-phase lookup at the top of this file
static std::string name_ = "Keypoint_Learnining_Detector";
static std::shared_ptr<const int> input_, surface_;
}
}  // namespace pcl

#define CV_32F 5
namespace cv {
class Mat {
public:
    Mat() {}
    Mat(int rows, int cols, int) : rows(rows), cols(cols), d((size_t)rows * cols, 0.0f) {}
    template <typename T> T& at(int r, int c) { return d[(size_t)r * cols + c]; }
    void push_back(const Mat& m) { if (rows == 0) cols = m.cols; d.insert(d.end(), m.d.begin(), m.d.end()); rows += m.rows; }
    int rows = 0, cols = 0;
    std::vector<float> d;
};
template <typename T>
class Ptr {
public:
    Ptr() {}
    Ptr(T* p) : p_(p) {}
    T* operator->() const { return p_.get(); }
    void release() { p_.reset(); }
    bool operator==(long v) const { return v == 0 && !p_; }       // forest_ == NULL
    bool operator!=(long v) const { return !(*this == v); }       // forest_ != 0
private:
    std::shared_ptr<T> p_;
};
namespace ml {
// the flat forest the stand-in RTrees::load hands out (set by the test through ref_wrap.cpp)
struct KplRefForest { int ntrees = 0; const int32_t *roots = nullptr, *var = nullptr, *left = nullptr, *right = nullptr; const float *thr = nullptr, *value = nullptr; };
inline KplRefForest& kplref_forest() { static KplRefForest f; return f; }
class RTrees {
public:
    struct Flags { enum { PREDICT_SUM = 256 }; };
    static Ptr<RTrees> load(const std::string&) { return kplref_forest().ntrees > 0 ? Ptr<RTrees>(new RTrees()) : Ptr<RTrees>(); }
    std::vector<int> getRoots() const { const KplRefForest& F = kplref_forest(); return std::vector<int>(F.roots, F.roots + F.ntrees); }
    // DTreesImpl::predictTrees with PREDICT_SUM (OpenCV 3.2 modules/ml/src/tree.cpp)
    float predict(Mat& sample, Mat&, int) const
    {
        const KplRefForest& F = kplref_forest();
        double sum = 0.0;
        for (int t = 0; t < F.ntrees; ++t) {
            int n = F.roots[t];
            while (F.var[n] >= 0) n = (sample.d[(size_t)F.var[n]] <= F.thr[n]) ? F.left[n] : F.right[n];
            sum += (double)F.value[n];
        }
        return (float)sum;
    }
};
}  // namespace ml
}  // namespace cv
