// Per-device resources (replaces gpu/GpuResources.h + gpu/StandardGpuResources.cpp:18-172): one stream, a grow-only
// device scratch arena and a pinned staging buffer.  The reference's cuBLAS handle / temp-memory stack are not needed:
// the GEMMs are our own tcgen05 kernels and every C-ABI call receives an explicit workspace.
#pragma once
#include "DeviceBuffer.h"

namespace faiss {
namespace gpu {

class GpuResources {
 public:
  virtual ~GpuResources() {}
  virtual int getDevice() const = 0;
  virtual vlq_stream_t getDefaultStream() = 0;
  virtual void syncDefaultStream() = 0;
  /// second stream for host->device staging that overlaps compute (reference: getAsyncCopyStream)
  virtual vlq_stream_t getAsyncCopyStream() = 0;
  /// stream of the device->host result copies: kept apart from the uploads so that the compute stream, which waits for
  /// the upload of the next query tile, never waits behind the download of the previous one
  virtual vlq_stream_t getAsyncDownloadStream() { return getAsyncCopyStream(); }
  /// second COMPUTE stream: the scan of query tile i runs here beside the coarse stage of tile i + 1 on the default
  /// stream (default: the default stream itself = no overlap)
  virtual vlq_stream_t getScanStream() { return getDefaultStream(); }
};

class StandardGpuResources : public GpuResources {
 public:
  explicit StandardGpuResources(int device = 0);
  ~StandardGpuResources() override;
  int getDevice() const override { return device_; }
  vlq_stream_t getDefaultStream() override { return stream_; }
  void syncDefaultStream() override;
  vlq_stream_t getAsyncCopyStream() override { return copyStream_; }
  vlq_stream_t getAsyncDownloadStream() override { return downStream_; }
  vlq_stream_t getScanStream() override { return scanStream_; }
  /// kept for source compatibility with the reference (gpu/StandardGpuResources.h): sizes are managed on demand
  void noTempMemory() {}
  void setTempMemory(size_t) {}
  void setPinnedMemory(size_t) {}

 private:
  int device_;
  vlq_stream_t stream_;
  vlq_stream_t copyStream_;
  vlq_stream_t downStream_;
  vlq_stream_t scanStream_;
};

/// binds the calling thread to the resource's device for the lifetime of the scope (reference DeviceScope)
class DeviceScope {
 public:
  explicit DeviceScope(int device);
  ~DeviceScope();

 private:
  int prev_;
};

}  // namespace gpu
}  // namespace faiss
