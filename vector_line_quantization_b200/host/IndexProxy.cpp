#include "IndexProxy.h"

#include <algorithm>
#include <exception>
#include <thread>

#include "DeviceBuffer.h"
#include "GpuIndexFlat.h"
#include "GpuIndexIVFPQ.h"

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
// device a shard lives on, -1 for an index that is not one of the GPU classes
static int deviceOfShard(const Index* ix) {
  if (auto* p = dynamic_cast<const gpu::GpuIndexIVF*>(ix)) return p->getDevice();
  if (auto* p = dynamic_cast<const gpu::GpuIndexFlat*>(ix)) return p->resources()->getDevice();
  return -1;
}

void IndexShards::search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const {
  const size_t ns = shard_indexes.size();
  VLQ_THROW_IF_NOT(ns > 0);
  if (n == 0) return;
  std::vector<int> dev(ns);
  bool all_gpu = true;
  for (size_t i = 0; i < ns; i++) {
    dev[i] = deviceOfShard(shard_indexes[i]);
    all_gpu = all_gpu && dev[i] >= 0;
  }
  if (all_gpu) {
    searchPeers_(dev, n, x, k, distances, labels);
    return;
  }
  // generic shards (any faiss::Index): results staged through host memory, merged on the current device
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
  gpu::DeviceBuffer dD(allD.size() * sizeof(float)), dI(allI.size() * sizeof(int64_t));
  gpu::DeviceBuffer oD((size_t)n * k * sizeof(float)), oI((size_t)n * k * sizeof(int64_t));
  VLQ_CALL(vlq_memcpy_h2d(dD.get(), allD.data(), dD.bytes(), nullptr));
  VLQ_CALL(vlq_memcpy_h2d(dI.get(), allI.data(), dI.bytes(), nullptr));
  VLQ_CALL(vlq_merge_topk(dD.as<float>(), dI.as<int64_t>(), (int)ns, n, (int)k, oD.as<float>(), oI.as<int64_t>(), nullptr));
  VLQ_CALL(vlq_memcpy_d2h(distances, oD.get(), oD.bytes(), nullptr));
  VLQ_CALL(vlq_memcpy_d2h(labels, oI.get(), oI.bytes(), nullptr));
  VLQ_CALL(vlq_stream_synchronize(nullptr));
}

// GPU shards: nothing but the queries and the final (n, k) result crosses PCIe.  Every shard searches with DEVICE result
// buffers on its own GPU (one thread per shard), shifts its labels there, and the merge kernel on the first shard's GPU
// reads the shards' results straight from their memories over NVLink (vlq_merge_topk_peers: gather + merge in one
// kernel).  Replaces the host staging of the reference drivers (gpu/test/sift1b16_query.cpp:389-430: MPI_Gather to
// host, cudaMemcpy back, mergekernel, gpu/GpuIndexIVFPQ.cu:1467-1591).
void IndexShards::searchPeers_(const std::vector<int>& dev, idx_t n, const float* x, idx_t k, float* distances,
                               idx_t* labels) const {
  const size_t ns = shard_indexes.size();
  const size_t nk = (size_t)n * k;
  const size_t i_off = (nk * sizeof(float) + 15) / 16 * 16;  // [D f32 (n, k) | I int64 (n, k)] per shard
  const size_t buf_bytes = i_off + nk * sizeof(int64_t);
  int x_on = -1;  // GPU the queries live on (-1: host memory)
  VLQ_CALL(vlq_pointer_device(x, &x_on));
  const bool out_dev = vlq_pointer_is_device(distances) == 1;
  std::vector<idx_t> shift(ns, 0);
  if (successive_ids)
    for (size_t i = 1; i < ns; i++) shift[i] = shift[i - 1] + shard_indexes[i - 1]->ntotal;
  std::vector<gpu::DeviceBuffer> res(ns), xq(ns);
  parallelFor(ns, threaded, [&](size_t i) {
    gpu::DeviceScope scope(dev[i]);
    res[i].resize(buf_bytes);
    const float* xi = x;
    if (x_on != dev[i]) {  // one copy of the queries per GPU: from the host, or from the GPU they live on (NVLink)
      xq[i].resize((size_t)n * d * sizeof(float));
      if (x_on < 0) VLQ_CALL(vlq_memcpy_h2d(xq[i].get(), x, xq[i].bytes(), nullptr));
      else VLQ_CALL(vlq_memcpy_d2d(xq[i].get(), x, xq[i].bytes(), nullptr));
      VLQ_CALL(vlq_stream_synchronize(nullptr));
      xi = xq[i].as<float>();
    }
    float* Di = res[i].as<float>();
    idx_t* Ii = reinterpret_cast<idx_t*>(res[i].as<unsigned char>() + i_off);
    shard_indexes[i]->search(n, xi, k, Di, Ii);  // device pointers in and out; returns with the results complete
    if (shift[i]) {
      VLQ_CALL(vlq_shift_ids(reinterpret_cast<int64_t*>(Ii), (int64_t)nk, shift[i], nullptr));
      VLQ_CALL(vlq_stream_synchronize(nullptr));
    }
  });
  gpu::DeviceScope scope(dev[0]);
  for (size_t i = 1; i < ns; i++) VLQ_CALL(vlq_enable_peer_access(dev[i]));
  std::vector<const void*> ptrs(ns);
  for (size_t i = 0; i < ns; i++) ptrs[i] = res[i].get();
  gpu::DeviceBuffer table(ns * sizeof(void*));
  VLQ_CALL(vlq_memcpy_h2d(table.get(), ptrs.data(), table.bytes(), nullptr));
  gpu::DeviceBuffer oD, oI;
  float* outD = distances;
  idx_t* outI = labels;
  if (!out_dev) {
    oD.resize(nk * sizeof(float));
    oI.resize(nk * sizeof(int64_t));
    outD = oD.as<float>();
    outI = oI.as<idx_t>();
  }
  VLQ_CALL(vlq_merge_topk_peers(table.as<const void*>(), 0, i_off, (int)ns, n, (int)k, outD,
                                reinterpret_cast<int64_t*>(outI), nullptr));
  if (!out_dev) {
    VLQ_CALL(vlq_memcpy_d2h(distances, oD.get(), oD.bytes(), nullptr));
    VLQ_CALL(vlq_memcpy_d2h(labels, oI.get(), oI.bytes(), nullptr));
  }
  VLQ_CALL(vlq_stream_synchronize(nullptr));
}

}  // namespace faiss
