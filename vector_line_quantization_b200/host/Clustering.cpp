#include "Clustering.h"

#include <cmath>

#include <stdlib.h>

#include <algorithm>
#include <cstring>

#include "ProductQuantizer.h"

namespace faiss {

// ------------------------------------------------------------------------------------------------ RNG
RandomGenerator::RandomGenerator(long seed) {
  static_assert(sizeof(struct random_data) <= sizeof(data_), "random_data does not fit");
  std::memset(data_, 0, sizeof(data_));
  initstate_r((unsigned)seed, state_, sizeof(state_), reinterpret_cast<struct random_data*>(data_));
}
int RandomGenerator::rand_int() {
  int32_t a;
  random_r(reinterpret_cast<struct random_data*>(data_), &a);
  return a;
}
float RandomGenerator::rand_float() { return rand_int() / float(1L << 31); }

void rand_perm(int* perm, size_t n, long seed) {
  for (size_t i = 0; i < n; i++) perm[i] = (int)i;
  RandomGenerator rng(seed);
  for (size_t i = 0; i + 1 < n; i++) {
    int i2 = (int)(i + rng.rand_int() % (n - i));
    std::swap(perm[i], perm[i2]);
  }
}

// ------------------------------------------------------------------------------------------------ parameters
ClusteringParameters::ClusteringParameters()
    : niter(25), nredo(1), verbose(false), spherical(false), update_index(false), min_points_per_centroid(39),
      max_points_per_centroid(256), seed(1234) {}

Clustering::Clustering(int d_, int k_) : d(d_), k(k_) {}
Clustering::Clustering(int d_, int k_, const ClusteringParameters& cp) : ClusteringParameters(cp), d(d_), k(k_) {}

// empty-cluster split of km_update_centroids (reference utils.cpp:1419-1446): sequential RNG logic on the counts; the
// touched centroid rows are patched on the host copy and written back
static int split_empty(std::vector<float>& cent, std::vector<long>& hassign, size_t d, size_t k, size_t n) {
  int nsplit = 0;
  RandomGenerator rng(1234);
  const float EPS = 1 / 1024.f;
  for (size_t ci = 0; ci < k; ci++) {
    if (hassign[ci] != 0) continue;
    size_t cj;
    for (cj = 0; true; cj = (cj + 1) % k) {
      float p = (hassign[cj] - 1.0) / (float)(n - k);
      float r = rng.rand_float();
      if (r < p) break;
    }
    std::memcpy(&cent[ci * d], &cent[cj * d], sizeof(float) * d);
    for (size_t j = 0; j < d; j++) {
      if (j % 2 == 0) {
        cent[ci * d + j] *= 1 + EPS;
        cent[cj * d + j] *= 1 - EPS;
      } else {
        cent[ci * d + j] *= 1 - EPS;
        cent[cj * d + j] *= 1 + EPS;
      }
    }
    hassign[ci] = hassign[cj] / 2;
    hassign[cj] -= hassign[ci];
    nsplit++;
  }
  return nsplit;
}

void Clustering::train(idx_t nx, const float* x_in, gpu::GpuIndexFlat& index) {
  VLQ_THROW_IF_NOT_MSG((size_t)nx >= k, "number of training points should be at least as large as number of clusters");
  VLQ_THROW_IF_NOT(index.d == (int)d);
  VLQ_THROW_IF_NOT_MSG(nredo == 1 && !spherical, "nredo > 1 / spherical k-means are not on the VLQ path");
  gpu::GpuResources* res = index.resources();
  gpu::DeviceScope scope(index.device());
  vlq_stream_t st = res->getDefaultStream();

  // the training set lives on the device for the whole loop (one H2D copy instead of one per iteration)
  gpu::DeviceBuffer xin;
  const float* dx_all = static_cast<const float*>(gpu::toDevice(x_in, (size_t)nx * d * sizeof(float), xin, st));
  gpu::DeviceBuffer xsub;
  const float* dx = dx_all;
  if ((size_t)nx > k * (size_t)max_points_per_centroid) {  // Clustering.cpp:81-93
    std::vector<int> perm(nx);
    rand_perm(perm.data(), nx, seed);
    nx = (idx_t)(k * max_points_per_centroid);
    std::vector<int64_t> rows(perm.begin(), perm.begin() + nx);
    gpu::DeviceBuffer drows((size_t)nx * sizeof(int64_t));
    VLQ_CALL(vlq_memcpy_h2d(drows.get(), rows.data(), drows.bytes(), st));
    xsub.resize((size_t)nx * d * sizeof(float));
    VLQ_CALL(vlq_gather_rows(dx_all, (int)d, drows.as<int64_t>(), nx, xsub.as<float>(), st));
    res->syncDefaultStream();
    dx = xsub.as<float>();
  }
  centroids.resize(d * k);
  {  // initial centroids = rows perm[0..k) with rand_perm(seed + 1) (Clustering.cpp:132-141)
    std::vector<int> perm(nx);
    rand_perm(perm.data(), nx, seed + 1);
    std::vector<int64_t> rows(perm.begin(), perm.begin() + k);
    gpu::DeviceBuffer drows(k * sizeof(int64_t)), dc(k * d * sizeof(float));
    VLQ_CALL(vlq_memcpy_h2d(drows.get(), rows.data(), drows.bytes(), st));
    VLQ_CALL(vlq_gather_rows(dx, (int)d, drows.as<int64_t>(), (int64_t)k, dc.as<float>(), st));
    VLQ_CALL(vlq_memcpy_d2h(centroids.data(), dc.get(), dc.bytes(), st));
    res->syncDefaultStream();
  }
  VLQ_THROW_IF_NOT_MSG(index.ntotal == 0, "the assigner index must be empty (Clustering.cpp:147-150)");
  index.add((idx_t)k, centroids.data());

  gpu::DeviceBuffer dassign((size_t)nx * sizeof(int)), ddis((size_t)nx * sizeof(float));
  gpu::DeviceBuffer dcent(k * d * sizeof(float)), dcount(k * sizeof(int));
  gpu::DeviceBuffer ws(vlq_km_update_workspace_bytes(nx, (int)k));
  std::vector<int> counts(k);
  std::vector<long> hassign(k);
  obj.clear();
  std::vector<float> dis(nx);
  for (int it = 0; it < niter; it++) {  // Clustering.cpp:161-193
    index.assignDevice(dx, nx, dassign.as<int>(), ddis.as<float>(), true);
    VLQ_CALL(vlq_memcpy_d2h(dis.data(), ddis.get(), ddis.bytes(), st));  // the objective of every iteration, like the
    VLQ_CALL(vlq_km_update(dx, nx, (int)d, dassign.as<int>(), (int)k, dcent.as<float>(), dcount.as<int>(), ws.get(),  // reference
                           ws.bytes(), st));
    VLQ_CALL(vlq_memcpy_d2h(counts.data(), dcount.get(), dcount.bytes(), st));
    VLQ_CALL(vlq_memcpy_d2h(centroids.data(), dcent.get(), dcent.bytes(), st));
    res->syncDefaultStream();
    {
      double err = 0;
      for (float v : dis) err += v;
      obj.push_back((float)err);
      if (verbose) printf("  Iteration %d  objective=%g\n", it, err);
      // rows with NaN / Inf components get no centroid (label -1) and would silently drop out of the means: the
      // reference refuses such input up front (Clustering.cpp:70-73 "input contains NaN's or Inf's")
      long assigned = 0;
      for (size_t c = 0; c < k; c++) assigned += counts[c];
      VLQ_THROW_IF_NOT_MSG(assigned == (long)nx && std::isfinite(err), "input contains NaN's or Inf's");
    }
    for (size_t c = 0; c < k; c++) hassign[c] = counts[c];
    split_empty(centroids, hassign, d, k, (size_t)nx);
    index.reset();
    index.add((idx_t)k, centroids.data());
  }
}

// ------------------------------------------------------------------------------------------------ PQ training
ProductQuantizer::ProductQuantizer(size_t d_, size_t M_, size_t nbits_) : d(d_), M(M_), nbits(nbits_), verbose(false) {
  VLQ_THROW_IF_NOT_MSG(M > 0 && d % M == 0, "d must be a multiple of M");
  dsub = d / M;
  ksub = (size_t)1 << nbits;
  code_size = (nbits * M + 7) / 8;
  centroids.resize(d * ksub);
  cp.niter = 25;
}

void ProductQuantizer::train(int n, const float* x, gpu::GpuResources* res) {
  gpu::DeviceScope scope(res->getDevice());
  vlq_stream_t st = res->getDefaultStream();
  // slice on the host side of the boundary (ProductQuantizer.cpp:262-275), k-means per slice on the device
  std::vector<float> host;
  const float* hx = x;
  if (vlq_pointer_is_device(x) == 1) {
    host.resize((size_t)n * d);
    VLQ_CALL(vlq_memcpy_d2h(host.data(), x, host.size() * sizeof(float), st));
    res->syncDefaultStream();
    hx = host.data();
  }
  std::vector<float> xslice((size_t)n * dsub);
  for (size_t m = 0; m < M; m++) {
    for (int i = 0; i < n; i++) std::memcpy(&xslice[(size_t)i * dsub], hx + (size_t)i * d + m * dsub, dsub * sizeof(float));
    Clustering clus((int)dsub, (int)ksub, cp);
    gpu::GpuIndexFlatConfig cfg;
    cfg.device = res->getDevice();
    gpu::GpuIndexFlatL2 assigner(res, (int)dsub, cfg);
    clus.train(n, xslice.data(), assigner);
    std::memcpy(&centroids[m * ksub * dsub], clus.centroids.data(), ksub * dsub * sizeof(float));
  }
}

}  // namespace faiss
