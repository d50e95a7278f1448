// Memory / stream helpers and error strings of the C-ABI (include/vlq_b200.h).  These are the only entry points that
// allocate or synchronise; they exist so the C++ host layer (csrc/host) and ctypes callers need no CUDA toolkit.
#include <cstdio>

#include "common.cuh"

extern "C" {

const char* vlq_version(void) { return "vlq_b200 0.1 (sm_100a)"; }

const char* vlq_error_string(int code) {
  switch (code) {
    case VLQ_OK: return "ok";
    case VLQ_EINVAL: return "vlq: invalid argument";
    case VLQ_EWORKSPACE: return "vlq: workspace too small";
    case VLQ_EUNSUPPORTED: return "vlq: unsupported configuration";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "vlq: unknown error";
}

int vlq_device_count(int* count) { return (int)cudaGetDeviceCount(count); }
int vlq_set_device(int device) { return (int)cudaSetDevice(device); }
int vlq_get_device(int* device) { return (int)cudaGetDevice(device); }
int vlq_mem_info(size_t* free_bytes, size_t* total_bytes) { return (int)cudaMemGetInfo(free_bytes, total_bytes); }
int vlq_malloc(void** ptr, size_t bytes) {
  if (!ptr) return VLQ_EINVAL;
  *ptr = nullptr;
  if (bytes == 0) return VLQ_OK;
  return (int)cudaMalloc(ptr, bytes);
}
int vlq_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : VLQ_OK; }
int vlq_malloc_host(void** ptr, size_t bytes) {
  if (!ptr) return VLQ_EINVAL;
  *ptr = nullptr;
  if (bytes == 0) return VLQ_OK;
  return (int)cudaMallocHost(ptr, bytes);
}
int vlq_free_host(void* ptr) { return ptr ? (int)cudaFreeHost(ptr) : VLQ_OK; }
int vlq_memcpy_h2d(void* dst, const void* src, size_t bytes, vlq_stream_t stream) {
  if (bytes == 0) return VLQ_OK;
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, vlq::as_stream(stream));
}
int vlq_memcpy_d2h(void* dst, const void* src, size_t bytes, vlq_stream_t stream) {
  if (bytes == 0) return VLQ_OK;
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, vlq::as_stream(stream));
}
int vlq_memcpy_d2d(void* dst, const void* src, size_t bytes, vlq_stream_t stream) {
  if (bytes == 0) return VLQ_OK;
  return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, vlq::as_stream(stream));
}
int vlq_memset(void* dst, int value, size_t bytes, vlq_stream_t stream) {
  if (bytes == 0) return VLQ_OK;
  return (int)cudaMemsetAsync(dst, value, bytes, vlq::as_stream(stream));
}
int vlq_pointer_is_device(const void* ptr) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? 1 : 0;
}
int vlq_pointer_device(const void* ptr, int* device) {
  if (!device) return VLQ_EINVAL;
  *device = -1;
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return VLQ_OK;
  }
  if (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) *device = attr.device;
  return VLQ_OK;
}
int vlq_enable_peer_access(int peer_device) {
  int cur = 0;
  cudaError_t e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return (int)e;
  if (cur == peer_device) return VLQ_OK;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, cur, peer_device);
  if (e != cudaSuccess) return (int)e;
  if (!can) return VLQ_EUNSUPPORTED;
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return VLQ_OK;
  }
  return (int)e;
}
int vlq_stream_create(vlq_stream_t* stream) {
  cudaStream_t s;
  cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  if (e != cudaSuccess) return (int)e;
  *stream = (vlq_stream_t)s;
  return VLQ_OK;
}
int vlq_stream_destroy(vlq_stream_t stream) { return (int)cudaStreamDestroy(vlq::as_stream(stream)); }
int vlq_stream_synchronize(vlq_stream_t stream) { return (int)cudaStreamSynchronize(vlq::as_stream(stream)); }
int vlq_stream_wait(vlq_stream_t waiter, vlq_stream_t producer) {
  cudaEvent_t ev;
  cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (e != cudaSuccess) return (int)e;
  e = cudaEventRecord(ev, vlq::as_stream(producer));
  if (e == cudaSuccess) e = cudaStreamWaitEvent(vlq::as_stream(waiter), ev, 0);
  cudaEventDestroy(ev);  // released once the recorded work has completed
  return (int)e;
}

// events for pipelines whose wait must be placed later than the record (two compute streams working on alternating buffers)
int vlq_event_create(vlq_event_t* ev) {
  if (!ev) return VLQ_EINVAL;
  cudaEvent_t e;
  cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  *ev = rc == cudaSuccess ? static_cast<vlq_event_t>(e) : nullptr;
  return (int)rc;
}
int vlq_event_destroy(vlq_event_t ev) { return ev ? (int)cudaEventDestroy(static_cast<cudaEvent_t>(ev)) : VLQ_OK; }
int vlq_event_record(vlq_event_t ev, vlq_stream_t stream) {
  return ev ? (int)cudaEventRecord(static_cast<cudaEvent_t>(ev), vlq::as_stream(stream)) : VLQ_EINVAL;
}
int vlq_stream_wait_event(vlq_stream_t waiter, vlq_event_t ev) {
  return ev ? (int)cudaStreamWaitEvent(vlq::as_stream(waiter), static_cast<cudaEvent_t>(ev), 0) : VLQ_EINVAL;
}

}  // extern "C"
