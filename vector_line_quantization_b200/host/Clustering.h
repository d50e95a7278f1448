// k-means driver (reference Clustering.{h,cpp}): same parameters, same order of operations (sub-sampling by
// rand_perm(seed), initialisation from rand_perm(seed+1), niter x {assign, mean update with the empty-cluster split}),
// but the loop body runs on the device: assignment = tcgen05 / fp32 kernels, mean update = vlq_km_update.
#pragma once
#include <vector>

#include "GpuIndexFlat.h"

namespace faiss {

struct ClusteringParameters {
  int niter;
  int nredo;
  bool verbose;
  bool spherical;
  bool update_index;
  int min_points_per_centroid;
  int max_points_per_centroid;
  int seed;
  ClusteringParameters();  // reference defaults Clustering.cpp:27-35
};

struct Clustering : ClusteringParameters {
  typedef Index::idx_t idx_t;
  size_t d;
  size_t k;
  std::vector<float> centroids;  // (k * d) on the host after train()
  std::vector<float> obj;        // objective (sum of squared distances) of every iteration

  Clustering(int d, int k);
  Clustering(int d, int k, const ClusteringParameters& cp);

  /// x: n*d, host or device.  `index` must be a (empty) GpuIndexFlat: it serves as the assigner and holds the final
  /// centroids on return, exactly like the reference (Clustering.cpp:154-192).
  void train(idx_t n, const float* x, gpu::GpuIndexFlat& index);
};

/// glibc random_r generator on an 8-byte state, as the reference RandomGenerator (utils.cpp:135-160)
struct RandomGenerator {
  explicit RandomGenerator(long seed = 1234);
  int rand_int();
  float rand_float();

 private:
  char state_[8];
  char data_[64];  // struct random_data
};
void rand_perm(int* perm, size_t n, long seed);  // utils.cpp:307-317

}  // namespace faiss
