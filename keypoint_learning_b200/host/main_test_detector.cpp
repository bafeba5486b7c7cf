// TestDetector -- headless drop-in of the reference's src/main_test_detector.cpp on the B200 build.
// Same options and defaults (reference :53-91): --help/-h, --flipNormals, --subSampling, --leaf,
// --pathCloud, --pathRF, --pathKP, --radiusFeatures (20), --radiusNMS (4), --threshold/-t (0.85);
// annuli/bins, compile-time 5/10 in the reference (:105-106), are the run-time options --annuli/--bins.
// Same progress lines on stdout, same exit codes (0 on --help or a parse error :112-113, -1 when the
// forest fails to load :136-139).  The PCLVisualizer window (:190-210) has no headless equivalent and
// is dropped; --stats prints the device-time breakdown instead.  --gpus N (B200 build only) cuts the cloud into N x
// slabs and runs one rank per GPU of this node through kpl_shard_* (NCCL halo exchange inside libkpl_b200.so); the
// keypoints are those of the single-GPU run, bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <string>

#include "KeypointLearning.h"
#include "KeypointLearningSharded.h"

typedef pcl::PointXYZ PointInT;
typedef pcl::Normal PointNormalT;
typedef pcl::PointXYZI KeypointT;

namespace {

struct Options {
    std::map<std::string, std::string> values;
    std::map<std::string, int> flags;
    bool count(const std::string& k) const { return flags.count(k) || values.count(k); }
    float as_float(const std::string& k) const { return std::strtof(values.at(k).c_str(), nullptr); }
    const std::string& as_string(const std::string& k) const { return values.at(k); }
};

void print_help()
{
    std::cout << "Allowed options:\n"
                 "  -h [ --help ]                         produce help message\n"
                 "  --flipNormals                         If present flip normals, some dataset needs normal re-orientation.\n"
                 "  --subSampling                         If present, subsample cloud with leaf.\n"
                 "  --leaf arg                            Leaf size for subsampling.\n"
                 "  --pathCloud arg (=../../../data/point_cloud_test/cheff001.pcd)\n"
                 "                                        Path to dataset.\n"
                 "  --pathRF arg (=../../../data/forest/SHOT-LaserScanner.yaml.gz)\n"
                 "                                        Path to Random Forest.\n"
                 "  --pathKP arg                          Path for keypoints point cloud.\n"
                 "  --radiusFeatures arg (=20)            Radius for features computation.\n"
                 "  --radiusNMS arg (=4)                  Radius for non maxima suppresion.\n"
                 "  -t [ --threshold ] arg (=0.850000024) Threshold for random forest prediction.\n"
                 "  --annuli arg (=5)                     Number of annuli of the feature histogram.\n"
                 "  --bins arg (=10)                      Number of cosine bins of the feature histogram.\n"
                 "  --stats                               Print device timings and counters of the detection.\n"
                 "  --gpus arg (=1)                       GPUs of this node to spread the cloud over (x slabs + halo over NCCL).\n";
}

// boost::program_options subset: --name value, --name=value, -t value, -h, boolean switches
bool parseCommandLine(int argc, char** argv, Options& vm)
{
    vm.values = {{"pathCloud", "../../../data/point_cloud_test/cheff001.pcd"}, {"pathRF", "../../../data/forest/SHOT-LaserScanner.yaml.gz"},
                 {"radiusFeatures", "20"}, {"radiusNMS", "4"}, {"threshold", "0.85"}, {"annuli", "5"}, {"bins", "10"}, {"gpus", "1"}};
    const char* valued[] = {"leaf", "pathCloud", "pathRF", "pathKP", "radiusFeatures", "radiusNMS", "threshold", "annuli", "bins", "gpus"};
    const char* switches[] = {"help", "flipNormals", "subSampling", "stats"};
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], name, val;
        bool has_val = false;
        if (a == "-h") name = "help";
        else if (a == "-t") name = "threshold";
        else if (a.rfind("--", 0) == 0) {
            size_t eq = a.find('=');
            name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
            if (eq != std::string::npos) { val = a.substr(eq + 1); has_val = true; }
        } else {
            std::cerr << "too many positional options have been specified on the command line" << std::endl;
            print_help();
            return false;
        }
        bool is_switch = false, is_valued = false;
        for (const char* s : switches) if (name == s) is_switch = true;
        for (const char* s : valued) if (name == s) is_valued = true;
        if (is_switch) { vm.flags[name] = 1; continue; }
        if (!is_valued) {
            std::cerr << "unrecognised option '" << a << "'" << std::endl;
            print_help();
            return false;
        }
        if (!has_val) {
            if (i + 1 >= argc) {
                std::cerr << "the required argument for option '--" << name << "' is missing" << std::endl;
                print_help();
                return false;
            }
            val = argv[++i];
        }
        vm.values[name] = val;
    }
    if (vm.flags.count("help")) { print_help(); return false; }
    if (vm.flags.count("subSampling") && !vm.values.count("leaf")) {
        std::cout << "Subsampling needs leaf." << std::endl;
        return false;
    }
    return true;
}

}  // namespace

int main(int argc, char** argv)
{
    Options vm;
    if (!parseCommandLine(argc, argv, vm)) return 0;

    const float radius_nms = vm.as_float("radiusNMS");
    const float radius_features = vm.as_float("radiusFeatures");
    const float threshold = vm.as_float("threshold");
    const int annuli = std::atoi(vm.as_string("annuli").c_str());
    const int bins = std::atoi(vm.as_string("bins").c_str());
    const std::string path_rf = vm.as_string("pathRF");
    const std::string path_cloud = vm.as_string("pathCloud");
    const int gpus = std::atoi(vm.as_string("gpus").c_str());

    if (gpus > 1) {
        // the same sequence as below over several GPUs; the k-NN(10) normals of :161-170 are estimated on the devices,
        // slab by slab, with the --flipNormals re-orientation applied there
        pcl::keypoints::ShardedKeypointLearningDetector<PointInT, KeypointT> detector(gpus);
        detector.setNAnnulus(annuli);
        detector.setNBins(bins);
        detector.setNonMaxima(true);
        detector.setNonMaxRadius(radius_nms);
        detector.setNonMaximaDrawsRemove(false);
        detector.setPredictionThreshold(threshold);
        detector.setRadiusSearch(radius_features);
        if (detector.loadForest(path_rf)) std::cout << "Detector created." << std::endl;
        else return -1;
        pcl::PointCloud<PointInT>::Ptr cloud(new pcl::PointCloud<PointInT>());
        pcl::io::loadPCDFile(path_cloud, *cloud);
        if (vm.count("subSampling")) {
            pcl::UniformSampling<PointInT> source_uniform_sampling;
            source_uniform_sampling.setRadiusSearch(vm.as_float("leaf"));
            source_uniform_sampling.setInputCloud(cloud);
            source_uniform_sampling.filter(*cloud);
        }
        std::cout << "Point cloud loaded" << std::endl;
        detector.setNormalEstimation(10, vm.count("flipNormals"));
        if (vm.count("flipNormals")) std::cout << "Flipping " << std::endl;
        detector.setInputCloud(cloud);
        pcl::PointCloud<KeypointT>::Ptr keypoint(new pcl::PointCloud<KeypointT>());
        if (!detector.compute(*keypoint)) return -1;
        std::cout << "Normals Computed" << std::endl;
        std::cout << "Keypoint computed" << std::endl;
        if (vm.count("stats")) {
            const kpl_slab_plan& L = detector.plan();
            std::printf("points %lld keypoints %zu slabs %d halo_cells %d normal_support_cells %d\n", (long long)L.n_points, keypoint->size(), L.world,
                        L.halo, L.normal_support_cells);
            for (int g = 0; g < gpus; ++g) {
                kpl_timings t;
                kpl_stats st;
                kpl_get_timings(detector.rank(g).context(), &t);
                kpl_get_stats(detector.rank(g).context(), &st);
                const kpl_shard_info& I = detector.info()[(size_t)g];
                std::printf("rank %d: columns [%d,%d) owned %lld halo %lld+%lld scored %lld pairs %lld device ms %.3f (features %.3f) exchange ms %.3f\n", g,
                            L.cuts[g], L.cuts[g + 1], (long long)I.n_owned, (long long)I.n_left, (long long)I.n_right, (long long)st.n_scored,
                            (long long)st.feature_pairs, t.total_ms, t.features_ms, I.exchange_ms);
            }
        }
        std::cout << "DONE" << std::endl;
        if (vm.count("pathKP")) pcl::io::savePCDFileASCII(vm.as_string("pathKP"), *keypoint);
        return 0;
    }

    // create detector (reference :123-130)
    pcl::keypoints::KeypointLearningDetector<PointInT, KeypointT>::Ptr detector(new pcl::keypoints::KeypointLearningDetector<PointInT, KeypointT>());
    detector->setNAnnulus(annuli);
    detector->setNBins(bins);
    detector->setNonMaxima(true);
    detector->setNonMaxRadius(radius_nms);
    detector->setNonMaximaDrawsRemove(false);
    detector->setPredictionThreshold(threshold);
    detector->setRadiusSearch(radius_features);

    if (detector->loadForest(path_rf)) std::cout << "Detector created." << std::endl;
    else return -1;

    // load and subsample point cloud (:141-157)
    pcl::PointCloud<PointInT>::Ptr cloud(new pcl::PointCloud<PointInT>());
    pcl::io::loadPCDFile(path_cloud, *cloud);
    if (vm.count("subSampling")) {
        pcl::UniformSampling<PointInT> source_uniform_sampling;
        source_uniform_sampling.setRadiusSearch(vm.as_float("leaf"));
        source_uniform_sampling.setInputCloud(cloud);
        source_uniform_sampling.filter(*cloud);
    }
    std::cout << "Point cloud loaded" << std::endl;

    // compute normals (:161-170)
    pcl::NormalEstimation<PointInT, PointNormalT> ne;
    pcl::PointCloud<PointNormalT>::Ptr normals(new pcl::PointCloud<PointNormalT>);
    ne.setInputCloud(cloud);
    pcl::search::KdTree<PointInT>::Ptr kdtree(new pcl::search::KdTree<PointInT>());
    ne.setKSearch(10);
    ne.setSearchMethod(kdtree);
    ne.compute(*normals);
    std::cout << "Normals Computed" << std::endl;

    if (vm.count("flipNormals")) {                        // :172-179
        std::cout << "Flipping " << std::endl;
        for (auto& q : normals->points) { q.normal[0] *= -1; q.normal[1] *= -1; q.normal[2] *= -1; }
    }

    detector->setInputCloud(cloud);                        // :182-187
    detector->setNormals(normals);
    pcl::PointCloud<KeypointT>::Ptr keypoint(new pcl::PointCloud<KeypointT>());
    detector->compute(*keypoint);
    std::cout << "Keypoint computed" << std::endl;

    if (vm.count("stats")) {
        kpl_timings t;
        kpl_stats s;
        kpl_get_timings(detector->context(), &t);
        kpl_get_stats(detector->context(), &s);
        std::printf("points %lld keypoints %lld above_threshold %lld feature_pairs %lld unscored %lld launches %d\n", (long long)s.n_points,
                    (long long)s.n_keypoints, (long long)s.n_above_threshold, (long long)s.feature_pairs, (long long)s.n_unscored, s.kernel_launches);
        std::printf("device ms: grid %.3f normals %.3f features %.3f forest %.3f nms %.3f total %.3f\n", t.grid_ms, t.normals_ms, t.features_ms,
                    t.forest_ms, t.nms_ms, t.total_ms);
    }
    std::cout << "DONE" << std::endl;

    if (vm.count("pathKP")) pcl::io::savePCDFileASCII(vm.as_string("pathKP"), *keypoint);   // :212-216
    return 0;
}
