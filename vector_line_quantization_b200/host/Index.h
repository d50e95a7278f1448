// faiss::Index -- the operator interface the VLQ hot path sits behind (drop-in boundary).
// Written from scratch to mirror the reference's public contract (reference Index.h:60-188): same member names,
// argument meaning, padding conventions (-1 labels / FLT_MAX distances) and error behaviour (FaissException for user
// errors).  idx_t = long; matrices are row-major compact float32; pointers may be host or device.
#pragma once
#include <cstdio>
#include <exception>
#include <string>

namespace faiss {

enum MetricType { METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1 };

/// user-facing errors (reference FaissAssert.h:54-91 throws; invariants abort)
class FaissException : public std::exception {
 public:
  explicit FaissException(const std::string& m) : msg(m) {}
  FaissException(const std::string& m, const char* func, const char* file, int line) {
    char buf[1024];
    snprintf(buf, sizeof(buf), "Error in %s at %s:%d: %s", func, file, line, m.c_str());
    msg = buf;
  }
  const char* what() const noexcept override { return msg.c_str(); }
  std::string msg;
};

#define VLQ_THROW_MSG(MSG) throw ::faiss::FaissException(MSG, __PRETTY_FUNCTION__, __FILE__, __LINE__)
#define VLQ_THROW_IF_NOT_MSG(COND, MSG) \
  do {                                   \
    if (!(COND)) VLQ_THROW_MSG(MSG);     \
  } while (0)
#define VLQ_THROW_IF_NOT(COND) VLQ_THROW_IF_NOT_MSG(COND, "'" #COND "' failed")

struct Index {
  typedef long idx_t;

  int d;
  idx_t ntotal;
  bool verbose;
  bool is_trained;
  MetricType metric_type;

  explicit Index(idx_t d_ = 0, MetricType metric = METRIC_L2)
      : d((int)d_), ntotal(0), verbose(false), is_trained(true), metric_type(metric) {}
  virtual ~Index() {}

  virtual void train(idx_t /*n*/, const float* /*x*/) {}
  virtual void add(idx_t n, const float* x) = 0;
  virtual void add_with_ids(idx_t /*n*/, const float* /*x*/, const long* /*xids*/) {
    VLQ_THROW_MSG("add_with_ids not implemented for this type of index");
  }
  virtual void search(idx_t n, const float* x, idx_t k, float* distances, idx_t* labels) const = 0;
  virtual void reset() = 0;

  /// stored vector `key` (reference Index.h:157; indexes that cannot decode throw, Index.cpp:49-51)
  virtual void reconstruct(idx_t /*key*/, float* /*recons*/) const { VLQ_THROW_MSG("reconstruct not implemented for this type of index"); }
  /// stored vectors [i0, i0 + ni) (reference Index.cpp:54-59: one reconstruct per row unless overridden)
  virtual void reconstruct_n(idx_t i0, idx_t ni, float* recons) const {
    for (idx_t i = 0; i < ni; i++) reconstruct(i0 + i, recons + i * d);
  }

  /// labels of the k nearest neighbours (= search without the distances; reference Index.cpp:23-29)
  void assign(idx_t n, const float* x, idx_t* labels, idx_t k = 1);
};

}  // namespace faiss
