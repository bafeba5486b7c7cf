// pcl_shim.cpp -- PCD v0.7 IO; NormalEstimation and UniformSampling forward to the C ABI (see pcl_shim.h).
#include "pcl_shim.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <unordered_map>

namespace pcl {
namespace io {

namespace {

struct Field { std::string name; int size = 4; char type = 'F'; int count = 1; int offset = 0; };

// LZF decompression (the codec of PCD "binary_compressed"); returns the number of bytes produced.
size_t lzf_decompress(const unsigned char* in, size_t in_len, unsigned char* out, size_t out_len)
{
    const unsigned char* ip = in;
    const unsigned char* const in_end = in + in_len;
    unsigned char* op = out;
    unsigned char* const out_end = out + out_len;
    while (ip < in_end) {
        unsigned ctrl = *ip++;
        if (ctrl < 32) {                 // literal run of ctrl+1 bytes
            ctrl++;
            if (op + ctrl > out_end || ip + ctrl > in_end) return 0;
            std::memcpy(op, ip, ctrl);
            op += ctrl; ip += ctrl;
        } else {                         // back reference
            unsigned len = ctrl >> 5;
            if (ip >= in_end) return 0;
            const unsigned char* ref = op - ((ctrl & 0x1f) << 8) - 1;
            if (len == 7) { len += *ip++; if (ip >= in_end) return 0; }
            ref -= *ip++;
            if (op + len + 2 > out_end || ref < out) return 0;
            len += 2;
            while (len--) *op++ = *ref++;
        }
    }
    return (size_t)(op - out);
}

double read_scalar(const unsigned char* p, const Field& f)
{
    switch (f.type) {
    case 'F':
        if (f.size == 4) { float v; std::memcpy(&v, p, 4); return v; }
        if (f.size == 8) { double v; std::memcpy(&v, p, 8); return v; }
        break;
    case 'I':
        if (f.size == 1) { int8_t v; std::memcpy(&v, p, 1); return v; }
        if (f.size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
        if (f.size == 4) { int32_t v; std::memcpy(&v, p, 4); return v; }
        break;
    case 'U':
        if (f.size == 1) { uint8_t v; std::memcpy(&v, p, 1); return v; }
        if (f.size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; }
        if (f.size == 4) { uint32_t v; std::memcpy(&v, p, 4); return v; }
        break;
    }
    return 0.0;
}

}  // namespace

int loadPCDFile(const std::string& path, PointCloud<PointXYZ>& cloud)
{
    std::ifstream f(path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "[pcl::io::loadPCDFile] cannot open %s\n", path.c_str()); return -1; }
    std::string raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<Field> fields;
    size_t npoints = 0, width = 0, height = 1, pos = 0;
    std::string data_kind;
    bool have_points = false;
    while (pos < raw.size()) {
        size_t nl = raw.find('\n', pos);
        if (nl == std::string::npos) nl = raw.size();
        std::string line = raw.substr(pos, nl - pos);
        pos = nl + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ls(line);
        std::string key;
        if (!(ls >> key) || key[0] == '#') continue;
        if (key == "FIELDS" || key == "COLUMNS") {
            std::string nm;
            while (ls >> nm) { Field fd; fd.name = nm; fields.push_back(fd); }
        } else if (key == "SIZE") { for (auto& fd : fields) ls >> fd.size; }
        else if (key == "TYPE") { for (auto& fd : fields) ls >> fd.type; }
        else if (key == "COUNT") { for (auto& fd : fields) ls >> fd.count; }
        else if (key == "WIDTH") ls >> width;
        else if (key == "HEIGHT") ls >> height;
        else if (key == "VIEWPOINT") {
            for (int i = 0; i < 3; ++i) ls >> cloud.sensor_origin_[i];
            for (int i = 0; i < 4; ++i) ls >> cloud.sensor_orientation_[i];
        } else if (key == "POINTS") { ls >> npoints; have_points = true; }
        else if (key == "DATA") { ls >> data_kind; break; }
    }
    if (!have_points) npoints = width * height;
    int ix = -1, iy = -1, iz = -1, point_step = 0;
    for (size_t i = 0; i < fields.size(); ++i) {
        fields[i].offset = point_step;
        point_step += fields[i].size * fields[i].count;
        if (fields[i].name == "x") ix = (int)i;
        if (fields[i].name == "y") iy = (int)i;
        if (fields[i].name == "z") iz = (int)i;
    }
    if (ix < 0 || iy < 0 || iz < 0 || data_kind.empty()) { std::fprintf(stderr, "[pcl::io::loadPCDFile] %s: no x y z fields / DATA line\n", path.c_str()); return -1; }
    cloud.points.assign(npoints, PointXYZ());
    bool dense = true;
    if (data_kind == "ascii") {
        const char* p = raw.c_str() + pos;
        const char* const end = raw.c_str() + raw.size();
        int ncols = 0;
        for (auto& fd : fields) ncols += fd.count;
        std::vector<int> col_of(3);
        int c = 0;
        for (size_t i = 0; i < fields.size(); ++i) {
            if ((int)i == ix) col_of[0] = c;
            if ((int)i == iy) col_of[1] = c;
            if ((int)i == iz) col_of[2] = c;
            c += fields[i].count;
        }
        for (size_t n = 0; n < npoints; ++n) {
            float v[3] = {0, 0, 0};
            for (int col = 0; col < ncols; ++col) {
                while (p < end && std::isspace((unsigned char)*p)) ++p;
                if (p >= end) { std::fprintf(stderr, "[pcl::io::loadPCDFile] %s: truncated ascii data\n", path.c_str()); cloud.points.resize(n); npoints = n; goto done; }
                char* q = nullptr;
                float val = std::strtof(p, &q);
                if (q == p) { while (p < end && !std::isspace((unsigned char)*p)) ++p; val = NAN; } else p = q;
                for (int a = 0; a < 3; ++a) if (col_of[a] == col) v[a] = val;
            }
            cloud.points[n] = PointXYZ(v[0], v[1], v[2]);
            if (!isFinite(cloud.points[n])) dense = false;
        }
    } else if (data_kind == "binary" || data_kind == "binary_compressed") {
        std::vector<unsigned char> buf;
        const unsigned char* base = reinterpret_cast<const unsigned char*>(raw.data()) + pos;
        size_t avail = raw.size() - pos;
        bool soa = false;
        if (data_kind == "binary_compressed") {
            if (avail < 8) return -1;
            uint32_t csz, usz;
            std::memcpy(&csz, base, 4); std::memcpy(&usz, base + 4, 4);
            if (avail < 8 + (size_t)csz || usz < npoints * (size_t)point_step) return -1;
            buf.resize(usz);
            if (lzf_decompress(base + 8, csz, buf.data(), usz) != usz) { std::fprintf(stderr, "[pcl::io::loadPCDFile] %s: LZF stream corrupt\n", path.c_str()); return -1; }
            base = buf.data(); avail = usz; soa = true;      // compressed payload is stored field by field
        }
        if (avail < npoints * (size_t)point_step) { std::fprintf(stderr, "[pcl::io::loadPCDFile] %s: truncated binary data\n", path.c_str()); return -1; }
        size_t soa_off[3] = {0, 0, 0};
        if (soa) {
            size_t off = 0;
            for (size_t i = 0; i < fields.size(); ++i) {
                if ((int)i == ix) soa_off[0] = off;
                if ((int)i == iy) soa_off[1] = off;
                if ((int)i == iz) soa_off[2] = off;
                off += (size_t)fields[i].size * fields[i].count * npoints;
            }
        }
        const int fi[3] = {ix, iy, iz};
        for (size_t n = 0; n < npoints; ++n) {
            float v[3];
            for (int a = 0; a < 3; ++a) {
                const Field& fd = fields[fi[a]];
                const unsigned char* p = soa ? base + soa_off[a] + n * (size_t)fd.size * fd.count : base + n * (size_t)point_step + fd.offset;
                v[a] = (float)read_scalar(p, fd);
            }
            cloud.points[n] = PointXYZ(v[0], v[1], v[2]);
            if (!isFinite(cloud.points[n])) dense = false;
        }
    } else {
        std::fprintf(stderr, "[pcl::io::loadPCDFile] %s: unsupported DATA %s\n", path.c_str(), data_kind.c_str());
        return -1;
    }
done:
    cloud.width = (uint32_t)(height > 1 ? width : npoints);
    cloud.height = (uint32_t)(height > 1 ? height : 1);
    cloud.is_dense = dense;
    return 0;
}

namespace {
template <typename CloudT, typename RowFn>
int save_ascii(const std::string& path, const CloudT& cloud, const char* fields, const char* sizes, const char* types, const char* counts, RowFn row)
{
    std::ofstream fs(path);
    if (!fs) { std::fprintf(stderr, "[pcl::io::savePCDFile] cannot write %s\n", path.c_str()); return -1; }
    fs.precision(8);
    fs.imbue(std::locale::classic());
    const size_t n = cloud.points.size();
    fs << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS " << fields << "\nSIZE " << sizes << "\nTYPE " << types << "\nCOUNT " << counts
       << "\nWIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA ascii\n";
    for (size_t i = 0; i < n; ++i) { row(fs, cloud.points[i]); fs << '\n'; }
    return fs.good() ? 0 : -1;
}
}  // namespace

int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZI>& cloud)
{
    return save_ascii(path, cloud, "x y z intensity", "4 4 4 4", "F F F F", "1 1 1 1",
                      [](std::ofstream& fs, const PointXYZI& p) { fs << p.x << ' ' << p.y << ' ' << p.z << ' ' << p.intensity; });
}
int savePCDFileASCII(const std::string& path, const PointCloud<PointXYZ>& cloud)
{
    return save_ascii(path, cloud, "x y z", "4 4 4", "F F F", "1 1 1",
                      [](std::ofstream& fs, const PointXYZ& p) { fs << p.x << ' ' << p.y << ' ' << p.z; });
}
int savePCDFileBinary(const std::string& path, const PointCloud<PointXYZ>& cloud)
{
    std::ofstream fs(path, std::ios::binary);
    if (!fs) return -1;
    const size_t n = cloud.points.size();
    fs << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH " << n
       << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
    for (size_t i = 0; i < n; ++i) fs.write(reinterpret_cast<const char*>(&cloud.points[i]), 12);
    return fs.good() ? 0 : -1;
}

}  // namespace io

// ---- NormalEstimation -> kpl_normals -----------------------------------------------------------
template <typename PointInT, typename NormalT>
bool NormalEstimation<PointInT, NormalT>::compute(PointCloud<NormalT>& out, std::string* err)
{
    out.points.clear(); out.width = out.height = 0;
    auto fail = [&](const std::string& m) { if (err) *err = m; std::fprintf(stderr, "[pcl::NormalEstimation::compute] %s\n", m.c_str()); return false; };
    if (!input_) return fail("no input cloud");
    if ((k_ > 0) == (radius_ > 0)) return fail("set exactly one of setKSearch / setRadiusSearch");
    kpl_ctx* ctx = nullptr;
    if (kpl_create(0, &ctx) != KPL_OK) return fail("no usable sm_100 device; there is no CPU path");
    kpl_params p;
    kpl_params_default(&p);
    if (k_ > 0) { p.normals_mode = KPL_NORMALS_KNN; p.k_normals = k_; }
    else { p.normals_mode = KPL_NORMALS_RADIUS; p.radius_features = (float)radius_; }
    p.viewpoint[0] = vp_[0]; p.viewpoint[1] = vp_[1]; p.viewpoint[2] = vp_[2];
    bool ok = kpl_set_params(ctx, &p) == KPL_OK;
    const size_t n = input_->size();
    std::vector<float> buf(n * 4);
    // non-dense clouds: PCL's search ignores NaN points and gives them NaN normals; only the finite points go to the device
    int nrc = ok ? kpl_normals(ctx, reinterpret_cast<const float*>(input_->points.data()), (int32_t)sizeof(PointInT), (int64_t)n, buf.data()) : KPL_E_INVALID;
    if (nrc != KPL_E_NONFINITE) ok = ok && nrc == KPL_OK;
    else {
        std::vector<size_t> finite;
        for (size_t i = 0; i < n; ++i) if (isFinite(input_->points[i])) finite.push_back(i);
        std::vector<PointInT> pts(finite.size());
        std::vector<float> sub(finite.size() * 4);
        for (size_t k = 0; k < finite.size(); ++k) pts[k] = input_->points[finite[k]];
        ok = ok && kpl_normals(ctx, reinterpret_cast<const float*>(pts.data()), (int32_t)sizeof(PointInT), (int64_t)pts.size(), sub.data()) == KPL_OK;
        std::fill(buf.begin(), buf.end(), std::nanf(""));
        for (size_t k = 0; ok && k < finite.size(); ++k) std::copy(sub.begin() + 4 * k, sub.begin() + 4 * k + 4, buf.begin() + 4 * finite[k]);
    }
    std::string msg = ok ? "" : kpl_last_error(ctx);
    kpl_destroy(ctx);
    if (!ok) return fail(msg);
    out.points.resize(n);
    bool dense = true;
    for (size_t i = 0; i < n; ++i) {
        NormalT& q = out.points[i];
        q.normal_x = buf[4 * i]; q.normal_y = buf[4 * i + 1]; q.normal_z = buf[4 * i + 2]; q.curvature = buf[4 * i + 3];
        if (!isFinite(q)) dense = false;
    }
    out.width = (uint32_t)n; out.height = 1; out.is_dense = dense;
    return true;
}
template class NormalEstimation<PointXYZ, Normal>;

// ---- UniformSampling -> kpl_uniform_sample ---------------------------------------------------------
template <typename PointT>
void UniformSampling<PointT>::filter(PointCloud<PointT>& out)
{
    typename PointCloud<PointT>::ConstPtr in = input_;
    std::vector<PointT> kept;
    if (in && leaf_ > 0 && !in->points.empty()) {
        kpl_ctx* ctx = nullptr;
        if (kpl_create(0, &ctx) != KPL_OK) {
            std::fprintf(stderr, "[pcl::UniformSampling::filter] no usable sm_100 device; there is no CPU path\n");
        } else {
            const int64_t n = (int64_t)in->points.size();
            std::vector<int32_t> idx((size_t)n);
            int64_t m = 0;
            if (kpl_uniform_sample(ctx, reinterpret_cast<const float*>(in->points.data()), (int32_t)sizeof(PointT), n, (float)leaf_, idx.data(), &m) != KPL_OK)
                std::fprintf(stderr, "[pcl::UniformSampling::filter] %s\n", kpl_last_error(ctx));
            else {
                kept.reserve((size_t)m);
                for (int64_t k = 0; k < m; ++k) kept.push_back(in->points[(size_t)idx[(size_t)k]]);
            }
            kpl_destroy(ctx);
        }
    }
    out.points.swap(kept);          // `out` may alias the input cloud (main_test_detector.cpp:156)
    out.width = (uint32_t)out.points.size(); out.height = 1; out.is_dense = true;
}
template class UniformSampling<PointXYZ>;

}  // namespace pcl
