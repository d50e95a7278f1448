#include "IndexProxy.h"

#include <algorithm>
#include <exception>
#include <thread>

#include "DeviceBuffer.h"

namespace faiss {

// run fn(i) for i in [0, n) on one thread each; rethrow the first exception (reference: WorkerThread + futures)
template <typename F>
static void parallelFor(size_t n, bool threaded, F fn) {
  if (!threaded || n <= 1) {
    for (size_t i = 0; i < n; i++) fn(i);
    return;
  }
  std::vector<std::thread> th;
  std::vector<std::exception_ptr> err(n);
  for (size_t i = 0; i < n; i++)
    th.emplace_back([&, i] {
      try {
        fn(i);
      } catch (...) {
        err[i] = std::current_exception();
      }
    });
  for (auto& t : th) t.join();
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}

namespace gpu {

void IndexProxy::addIndex(faiss::Index* index) {
  if (indices_.empty()) {
    d = index->d;
    metric_type = index->metric_type;
    is_trained = index->is_trained;
    ntotal = index->ntotal;
  } else {
    VLQ_THROW_IF_NOT_MSG(d == index->d && metric_type == index->metric_type && ntotal == index->ntotal,
                         "IndexProxy sub-indexes must be replicas (same d, metric, ntotal)");
  }
  indices_.push_back(index);
}
void IndexProxy::removeIndex(faiss::Index* index) {
  indices_.erase(std::remove(indices_.begin(), indices_.end(), index), indices_.end());
}
void IndexProxy::runOnIndex(void (*f)(faiss::Index*, void*), void* arg) {
  parallelFor(indices_.size(), true, [&](size_t i) { f(indices_[i], arg); });
}
void IndexProxy::reset() {
  parallelFor(indices_.size(), true, [&](size_t i) { indices_[i]->reset(); });
  ntotal = 0;
}
void IndexProxy::train(Index::idx_t n, const float* x) {
  parallelFor(indices_.size(), true, [&](size_t i) { indices_[i]->train(n, x); });
  is_trained = true;
}
void IndexProxy::add(Index::idx_t n, const float* x) {
  parallelFor(indices_.size(), true, [&](size_t i) { indices_[i]->add(n, x); });
  ntotal += n;
}
void IndexProxy::search(Index::idx_t n, const float* x, Index::idx_t k, float* distances, Index::idx_t* labels) const {
  VLQ_THROW_IF_NOT_MSG(!indices_.empty(), "no sub-index");
  if (n == 0) return;
  const Index::idx_t per = (n + (Index::idx_t)indices_.size() - 1) / (Index::idx_t)indices_.size();  // IndexProxy.cpp:136-150
  parallelFor(indices_.size(), true, [&](size_t i) {
    const Index::idx_t b = (Index::idx_t)i * per;
    if (b >= n) return;
    const Index::idx_t m = std::min(per, n - b);
    indices_[i]->search(m, x + (size_t)b * d, k, distances + (size_t)b * k, labels + (size_t)b * k);
  });
}

}  // namespace gpu

IndexShards::IndexShards(idx_t d_, bool threaded_, bool successive_ids_)
    : Index(d_, METRIC_L2), threaded(threaded_), successive_ids(successive_ids_) {}

void IndexShards::add_shard(Index* index) {
  VLQ_THROW_IF_NOT(index->d == d);
  shard_indexes.push_back(index);
  ntotal += index->ntotal;
  is_trained = index->is_trained;
}
void IndexShards::train(idx_t n, const float* x) {
  parallelFor(shard_indexes.size(), threaded, [&](size_t i) { shard_indexes[i]->train(n, x); });
  is_trained = true;
}
void IndexShards::reset() {
  for (auto* s : shard_indexes) s->reset();
  ntotal = 0;
}
void IndexShards::add(idx_t n, const float* x) {
  const idx_t ns = (idx_t)shard_indexes.size();
  VLQ_THROW_IF_NOT(ns > 0);
  parallelFor((size_t)ns, threaded, [&](size_t i) {
    const idx_t i0 = (idx_t)i * n / ns, i1 = ((idx_t)i + 1) * n / ns;
    if (successive_ids) {
      shard_indexes[i]->add(i1 - i0, x + (size_t)i0 * d);
    } else {
      std::vector<long> ids(i1 - i0);
      for (idx_t j = i0; j < i1; j++) ids[j - i0] = ntotal + j;
      shard_indexes[i]->add_with_ids(i1 - i0, x + (size_t)i0 * d, ids.data());
    }
  });
  ntotal += n;
}
void IndexShards::search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const {
  const size_t ns = shard_indexes.size();
  VLQ_THROW_IF_NOT(ns > 0);
  if (n == 0) return;
  std::vector<float> allD(ns * n * k);
  std::vector<idx_t> allI(ns * n * k);
  parallelFor(ns, threaded, [&](size_t i) {
    shard_indexes[i]->search(n, x, k, allD.data() + i * n * k, allI.data() + i * n * k);
  });
  if (successive_ids) {  // translate shard-local ids (MetaIndexes.cpp:536-546)
    idx_t shift = 0;
    for (size_t i = 0; i < ns; i++) {
      idx_t* p = allI.data() + i * n * k;
      if (shift)
        for (size_t j = 0; j < (size_t)n * k; j++)
          if (p[j] >= 0) p[j] += shift;
      shift += shard_indexes[i]->ntotal;
    }
  }
  // merge on the current device (vlq_merge_topk == mergekernel / merge_tables semantics)
  gpu::DeviceBuffer dD(allD.size() * sizeof(float)), dI(allI.size() * sizeof(int64_t));
  gpu::DeviceBuffer oD((size_t)n * k * sizeof(float)), oI((size_t)n * k * sizeof(int64_t));
  VLQ_CALL(vlq_memcpy_h2d(dD.get(), allD.data(), dD.bytes(), nullptr));
  VLQ_CALL(vlq_memcpy_h2d(dI.get(), allI.data(), dI.bytes(), nullptr));
  VLQ_CALL(vlq_merge_topk(dD.as<float>(), dI.as<int64_t>(), (int)ns, n, (int)k, oD.as<float>(), oI.as<int64_t>(), nullptr));
  VLQ_CALL(vlq_memcpy_d2h(distances, oD.get(), oD.bytes(), nullptr));
  VLQ_CALL(vlq_memcpy_d2h(labels, oI.get(), oI.bytes(), nullptr));
  VLQ_CALL(vlq_stream_synchronize(nullptr));
}

}  // namespace faiss
