// C entry points over the reference's OWN findAnnulusPair / findBinPair (src/KeypointLearning.cpp:41-92),
// compiled from /root/reference into oracle/_ref/libkpl_ref_helpers.so by oracle/Makefile.  Test
// infrastructure: tests/test_oracle.py pins the oracle's restatement of the two helpers against it.
void findAnnulusPair(int n_annulus, float distance, float support, int& annulus_index, int& annulus_index_pair, float& annulus_weight);
void findBinPair(int n_bins, float cosine, int& bin_index, int& bin_index_pair, float& bin_weight);

extern "C" __attribute__((visibility("default")))
void kplref_find_annulus_pair(int n_annulus, float distance, float support, int* index, int* pair, float* weight)
{
    findAnnulusPair(n_annulus, distance, support, *index, *pair, *weight);
}
extern "C" __attribute__((visibility("default")))
void kplref_find_bin_pair(int n_bins, float cosine, int* index, int* pair, float* weight)
{
    findBinPair(n_bins, cosine, *index, *pair, *weight);
}
// whole arrays at once, so a sweep over millions of inputs does not pay a ctypes call each
extern "C" __attribute__((visibility("default")))
void kplref_annulus_sweep(int n_annulus, float support, const float* distance, long n, int* index, int* pair, float* weight)
{
    for (long i = 0; i < n; ++i) findAnnulusPair(n_annulus, distance[i], support, index[i], pair[i], weight[i]);
}
extern "C" __attribute__((visibility("default")))
void kplref_bin_sweep(int n_bins, const float* cosine, long n, int* index, int* pair, float* weight)
{
    for (long i = 0; i < n; ++i) findBinPair(n_bins, cosine[i], index[i], pair[i], weight[i]);
}
