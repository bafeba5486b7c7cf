// pcd_tool -- small CLI over the PCD reader / writers of pcl_shim (used by the CPU tests and handy
// for converting clouds):   pcd_tool dump in.pcd        -> "n width height dense vx vy vz" + one "x y z" line per point (%.9g)
//                           pcd_tool ascii|binary in.pcd out.pcd
//                           pcd_tool subsample leaf in.pcd out.pcd   (pcl::UniformSampling stand-in)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "pcl_shim.h"

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
    std::fprintf(stderr, "bad arguments\n");
    return 2;
}
