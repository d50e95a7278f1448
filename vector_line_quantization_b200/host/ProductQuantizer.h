// Product quantizer codebook container + training (reference ProductQuantizer.{h,cpp}): M independent k-means of
// ksub = 2^nbits centroids over the dsub = d/M slices (ProductQuantizer.cpp:236-308), run on the device.
#pragma once
#include <vector>

#include "Clustering.h"

namespace faiss {

struct ProductQuantizer {
  size_t d, M, nbits;
  size_t dsub, ksub, code_size;
  bool verbose;
  ClusteringParameters cp;
  std::vector<float> centroids;  // (M, ksub, dsub)

  ProductQuantizer(size_t d, size_t M, size_t nbits);
  void train(int n, const float* x, gpu::GpuResources* res);  // x: n*d host or device
  const float* get_centroids(size_t m, size_t i) const { return &centroids[(m * ksub + i) * dsub]; }
};

}  // namespace faiss
