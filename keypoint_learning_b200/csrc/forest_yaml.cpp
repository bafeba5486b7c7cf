// forest_yaml.cpp -- opencv_ml_rtrees YAML / YAML.gz reader (host side of loadForest,
// impl/KeypointLearning.hpp:159-176, which calls cv::ml::RTrees::load).  OpenCV is not a dependency:
// the file is tokenised directly.  Schema (OpenCV 3.x "format: 3"): nodes of each tree are listed in
// pre-order; a node with a `splits` entry is internal, the next node is its left child and the first
// node after the left subtree is its right child; `{ var:v, quality:.., le:c }` means
// "go left iff x[v] <= c".  Only ordered (`le`) splits can occur for this feature set; categorical
// splits are rejected.
#include <zlib.h>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "kpl_internal.h"

namespace kpl {

static bool read_all(const char* path, std::string& out, std::string& err)
{
    gzFile f = gzopen(path, "rb");   // transparently reads plain files too
    if (!f) { err = std::string("cannot open forest file ") + path; return false; }
    char buf[1 << 16];
    int got;
    while ((got = gzread(f, buf, sizeof buf)) > 0) out.append(buf, (size_t)got);
    bool ok = got == 0;
    gzclose(f);
    if (!ok) err = std::string("read error in forest file ") + path;
    return ok;
}

int parse_forest_yaml(const char* path, HostForestArrays& F, std::string& err)
{
    std::string txt;
    if (!read_all(path, txt, err)) return KPL_E_FOREST;
    if (txt.find("opencv_ml_rtrees") == std::string::npos) { err = "not an opencv_ml_rtrees file"; return KPL_E_FOREST; }
    F = HostForestArrays();
    struct Open { int32_t node; int children; };
    std::vector<Open> stack;
    int32_t cur = -1;
    bool cur_has_var = false, cur_has_thr = false, in_trees = false;
    long declared_ntrees = -1, is_classifier = 1;
    std::vector<long> var_idx;
    std::vector<uint8_t> has_thr;
    const char* s = txt.c_str();
    const size_t L = txt.size();
    size_t i = 0;
    auto is_id = [](char c) { return std::isalnum((unsigned char)c) || c == '_'; };
    while (i < L) {
        if (!(std::isalpha((unsigned char)s[i]) || s[i] == '_') || (i > 0 && is_id(s[i - 1]))) { ++i; continue; }
        size_t b = i;
        while (i < L && is_id(s[i])) ++i;
        size_t e = i;
        size_t j = i;
        while (j < L && (s[j] == ' ' || s[j] == '\t')) ++j;
        if (j >= L || s[j] != ':') continue;
        ++j;
        while (j < L && (s[j] == ' ' || s[j] == '\t')) ++j;
        std::string key(s + b, e - b);
        bool has_num = j < L && (std::isdigit((unsigned char)s[j]) || s[j] == '-' || s[j] == '+' || s[j] == '.');
        double num = 0.0;
        if (has_num) { char* endp = nullptr; num = std::strtod(s + j, &endp); if (endp == s + j) has_num = false; }
        i = j;
        if (key == "trees") { in_trees = true; continue; }
        if (!in_trees) {
            if (key == "ntrees" && has_num) declared_ntrees = (long)num;
            else if (key == "var_count" && has_num) F.var_count = (int32_t)num;
            else if (key == "is_classifier" && has_num) is_classifier = (long)num;
            else if (key == "var_idx" && j < L && s[j] == '[') {
                size_t k = j + 1;
                while (k < L && s[k] != ']') {
                    if (std::isdigit((unsigned char)s[k])) { char* endp; var_idx.push_back(std::strtol(s + k, &endp, 10)); k = (size_t)(endp - s); }
                    else ++k;
                }
                i = k;
            }
            continue;
        }
        if (key == "nodes") {
            if (!stack.empty()) { err = "forest: truncated tree"; return KPL_E_FOREST; }
            F.roots.push_back((int32_t)F.var.size());
            cur = -1;
        } else if (key == "depth") {
            if (F.roots.empty()) { err = "forest: node outside a tree"; return KPL_E_FOREST; }
            cur = (int32_t)F.var.size();
            has_thr.push_back(0);
            F.var.push_back(-1); F.thr.push_back(0.f); F.left.push_back(-1); F.right.push_back(-1); F.value.push_back(0.f);
            cur_has_var = cur_has_thr = false;
            if (!stack.empty()) {
                Open& p = stack.back();
                if (p.children == 0) F.left[p.node] = cur; else F.right[p.node] = cur;
                if (++p.children == 2) stack.pop_back();
            } else if (cur != F.roots.back()) { err = "forest: node after a closed tree"; return KPL_E_FOREST; }
        } else if (key == "value" && cur >= 0 && has_num) {
            F.value[cur] = (float)num;
        } else if (key == "var" && cur >= 0 && has_num) {
            if (!cur_has_var) { F.var[cur] = (int32_t)num; cur_has_var = true; stack.push_back({cur, 0}); }
        } else if (key == "le" && cur >= 0) {
            // `.Inf` / `.Nan` thresholds are not numbers strtod accepts: refuse them instead of keeping 0
            if (!has_num || !std::isfinite(num)) { err = "forest: non-numeric or non-finite 'le' threshold"; return KPL_E_FOREST; }
            if (!cur_has_thr) { F.thr[cur] = (float)num; cur_has_thr = true; has_thr[(size_t)cur] = 1; }
        } else if (key == "gt" || key == "in" || key == "not_in") {
            err = "forest: unsupported split type '" + key + "' (only ordered 'le' splits)"; return KPL_E_FOREST;
        }
    }
    if (!stack.empty()) { err = "forest: truncated tree"; return KPL_E_FOREST; }
    if (F.roots.empty()) { err = "forest: no trees (getRoots().size() == 0)"; return KPL_E_FOREST; }
    if (declared_ntrees >= 0 && declared_ntrees != (long)F.roots.size()) { err = "forest: ntrees does not match the tree list"; return KPL_E_FOREST; }
    if (!is_classifier) { err = "forest: not a classifier"; return KPL_E_FOREST; }
    for (size_t k = 0; k < var_idx.size(); ++k)
        if (var_idx[k] != (long)k) { err = "forest: non-identity var_idx is not supported"; return KPL_E_FOREST; }
    for (size_t k = 0; k < F.var.size(); ++k)
        if (F.var[k] >= 0 && (F.left[k] < 0 || F.right[k] < 0)) { err = "forest: internal node without two children"; return KPL_E_FOREST; }
    for (size_t k = 0; k < F.var.size(); ++k)
        if (F.var[k] >= 0 && !has_thr[k]) { err = "forest: split without an 'le' threshold"; return KPL_E_FOREST; }
    return KPL_OK;
}

}  // namespace kpl
