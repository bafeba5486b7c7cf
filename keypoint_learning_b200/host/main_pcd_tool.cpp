// pcd_tool -- small CLI over the PCD reader / writers of pcl_shim (used by the CPU tests and handy
// for converting clouds):   pcd_tool dump in.pcd        -> "n width height dense vx vy vz" + one "x y z" line per point (%.9g)
//                           pcd_tool ascii|binary in.pcd out.pcd
//                           pcd_tool subsample leaf in.pcd out.pcd   (pcl::UniformSampling stand-in)
//                           pcd_tool detect forest in.pcd r_feat r_nms threshold
//                               the detector WITHOUT setNormals(): initCompute estimates them itself (hpp:125-148: radius-mode
//                               NormalEstimation for an unorganized cloud, IntegralImageNormalEstimation for an organized one);
//                               prints "n_keypoints" and one "index score" line per keypoint (%.9g)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "pcl_shim.h"
#include "KeypointLearning.h"

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: pcd_tool dump|ascii|binary|subsample ...\n"); return 2; }
    const std::string cmd = argv[1];
    pcl::PointCloud<pcl::PointXYZ>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZ>());
    if (cmd == "dump") {
        if (pcl::io::loadPCDFile(argv[2], *cloud)) return 1;
        std::printf("%zu %u %u %d %.9g %.9g %.9g\n", cloud->size(), cloud->width, cloud->height, (int)cloud->is_dense, cloud->sensor_origin_[0],
                    cloud->sensor_origin_[1], cloud->sensor_origin_[2]);
        for (const auto& p : cloud->points) std::printf("%.9g %.9g %.9g\n", p.x, p.y, p.z);
        return 0;
    }
    if ((cmd == "ascii" || cmd == "binary") && argc == 4) {
        if (pcl::io::loadPCDFile(argv[2], *cloud)) return 1;
        return cmd == "ascii" ? pcl::io::savePCDFileASCII(argv[3], *cloud) : pcl::io::savePCDFileBinary(argv[3], *cloud);
    }
    if (cmd == "subsample" && argc == 5) {
        if (pcl::io::loadPCDFile(argv[3], *cloud)) return 1;
        pcl::UniformSampling<pcl::PointXYZ> us;
        us.setRadiusSearch(std::atof(argv[2]));
        us.setInputCloud(cloud);
        us.filter(*cloud);
        return pcl::io::savePCDFileASCII(argv[4], *cloud);
    }
    if (cmd == "detect" && argc == 7) {
        if (pcl::io::loadPCDFile(argv[3], *cloud)) return 1;
        pcl::keypoints::KeypointLearningDetector<pcl::PointXYZ, pcl::PointXYZI> det;
        det.setNAnnulus(5); det.setNBins(10); det.setNonMaxima(true); det.setNonMaximaDrawsRemove(false);
        det.setRadiusSearch(std::atof(argv[4])); det.setNonMaxRadius(std::atof(argv[5])); det.setPredictionThreshold((float)std::atof(argv[6]));
        if (!det.loadForest(argv[2])) return 1;
        det.setInputCloud(cloud);
        pcl::PointCloud<pcl::PointXYZI> kp;
        det.compute(kp);
        std::printf("%zu\n", kp.size());
        for (size_t k = 0; k < kp.size(); ++k) std::printf("%d %.9g\n", det.getKeypointsIndices()->indices[k], kp.points[k].intensity);
        return 0;
    }
    std::fprintf(stderr, "bad arguments\n");
    return 2;
}
