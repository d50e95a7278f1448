// RAII device / pinned-host buffers and error mapping on top of the C-ABI helpers (no CUDA headers in the host layer).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>

#include "Index.h"
#include "vlq_b200.h"

namespace faiss {
namespace gpu {

/// C-ABI return code -> FaissException (the C-ABI never throws; the host layer does, like the reference's FAISS_THROW_*)
inline void vlqCheck(int rc, const char* what) {
  if (rc != VLQ_OK) throw FaissException(std::string(what) + ": " + vlq_error_string(rc));
}
#define VLQ_CALL(expr) ::faiss::gpu::vlqCheck((expr), #expr)

class DeviceBuffer {
 public:
  DeviceBuffer() : p_(nullptr), bytes_(0) {}
  explicit DeviceBuffer(size_t bytes) : p_(nullptr), bytes_(0) { resize(bytes); }
  ~DeviceBuffer() { release(); }
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  DeviceBuffer(DeviceBuffer&& o) noexcept : p_(o.p_), bytes_(o.bytes_) { o.p_ = nullptr; o.bytes_ = 0; }
  DeviceBuffer& operator=(DeviceBuffer&& o) noexcept {
    if (this != &o) {
      release();
      p_ = o.p_; bytes_ = o.bytes_; o.p_ = nullptr; o.bytes_ = 0;
    }
    return *this;
  }
  /// discard contents and hold exactly `bytes`
  void resize(size_t bytes) {
    if (bytes == bytes_) return;
    release();
    if (bytes) VLQ_CALL(vlq_malloc(&p_, bytes));
    bytes_ = bytes;
  }
  /// grow-only scratch (contents discarded on growth)
  void reserve(size_t bytes) {
    if (bytes > bytes_) resize(bytes);
  }
  void release() {
    if (p_) vlq_free(p_);
    p_ = nullptr;
    bytes_ = 0;
  }
  void swap(DeviceBuffer& o) { std::swap(p_, o.p_); std::swap(bytes_, o.bytes_); }
  template <typename T> T* as() const { return static_cast<T*>(p_); }
  void* get() const { return p_; }
  size_t bytes() const { return bytes_; }

 private:
  void* p_;
  size_t bytes_;
};

}  // namespace gpu
}  // namespace faiss
