#include "filehelper.h"

#include <cstdio>
#include <cstring>
#include <memory>

#include "Index.h"

namespace faiss {

namespace {
struct FileCloser {
  void operator()(FILE* f) const {
    if (f) fclose(f);
  }
};
using File = std::unique_ptr<FILE, FileCloser>;

File open_or_throw(const std::string& path, const char* mode) {
  File f(fopen(path.c_str(), mode));
  if (!f) throw FaissException("cannot open " + path);
  return f;
}

size_t file_size(FILE* f) {
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  return sz < 0 ? 0 : (size_t)sz;
}

template <typename T>
std::vector<T> vecs_read(const std::string& path, size_t* n_out, size_t* d_out, size_t start, size_t num) {
  File f = open_or_throw(path, "rb");
  const size_t bytes = file_size(f.get());
  int32_t d32 = 0;
  if (bytes < sizeof(int32_t) || fread(&d32, sizeof(int32_t), 1, f.get()) != 1 || d32 <= 0)
    throw FaissException("bad TexMex header in " + path);
  const size_t d = (size_t)d32, rec = sizeof(int32_t) + d * sizeof(T);
  if (bytes % rec != 0) throw FaissException("TexMex file size is not a multiple of the record size: " + path);
  const size_t total = bytes / rec;
  if (start > total) throw FaissException("start beyond the end of " + path);
  if (num == 0 || start + num > total) num = total - start;
  std::vector<T> out(num * d);
  std::vector<unsigned char> buf(rec);
  fseek(f.get(), (long)(start * rec), SEEK_SET);
  for (size_t i = 0; i < num; i++) {
    if (fread(buf.data(), 1, rec, f.get()) != rec) throw FaissException("short read in " + path);
    int32_t di;
    std::memcpy(&di, buf.data(), sizeof(int32_t));
    if ((size_t)di != d) throw FaissException("inconsistent dimension in " + path);
    std::memcpy(out.data() + i * d, buf.data() + sizeof(int32_t), d * sizeof(T));
  }
  if (n_out) *n_out = num;
  if (d_out) *d_out = d;
  return out;
}

template <typename T>
void vecs_write(const std::string& path, const T* x, size_t n, size_t d) {
  File f = open_or_throw(path, "wb");
  const int32_t d32 = (int32_t)d;
  for (size_t i = 0; i < n; i++) {
    if (fwrite(&d32, sizeof(int32_t), 1, f.get()) != 1 || fwrite(x + i * d, sizeof(T), d, f.get()) != d)
      throw FaissException("write error on " + path);
  }
}
}  // namespace

void vecs_header(const std::string& path, size_t elem_size, size_t* n, size_t* d) {
  File f = open_or_throw(path, "rb");
  const size_t bytes = file_size(f.get());
  int32_t d32 = 0;
  if (fread(&d32, sizeof(int32_t), 1, f.get()) != 1 || d32 <= 0) throw FaissException("bad TexMex header in " + path);
  const size_t rec = sizeof(int32_t) + (size_t)d32 * elem_size;
  if (n) *n = bytes / rec;
  if (d) *d = (size_t)d32;
}

std::vector<float> fvecs_read(const std::string& p, size_t* n, size_t* d, size_t s, size_t m) { return vecs_read<float>(p, n, d, s, m); }
std::vector<int32_t> ivecs_read(const std::string& p, size_t* n, size_t* d, size_t s, size_t m) { return vecs_read<int32_t>(p, n, d, s, m); }
std::vector<uint8_t> bvecs_read(const std::string& p, size_t* n, size_t* d, size_t s, size_t m) { return vecs_read<uint8_t>(p, n, d, s, m); }
void fvecs_write(const std::string& p, const float* x, size_t n, size_t d) { vecs_write(p, x, n, d); }
void ivecs_write(const std::string& p, const int32_t* x, size_t n, size_t d) { vecs_write(p, x, n, d); }
void bvecs_write(const std::string& p, const uint8_t* x, size_t n, size_t d) { vecs_write(p, x, n, d); }

void umem_header(const std::string& path, size_t* num, size_t* dim) {
  File f = open_or_throw(path, "rb");
  unsigned long long a = 0, b = 0;
  if (fscanf(f.get(), "%llu %llu", &a, &b) != 2) throw FaissException("bad .umem header in " + path);
  if (num) *num = (size_t)a;
  if (dim) *dim = (size_t)b;
}

void umem_write(const std::string& path, size_t num, size_t dim, const void* ptr, size_t elem_size, size_t len,
                size_t offset) {
  File f = open_or_throw(path, offset == 0 ? "wb" : "rb+");
  if (offset == 0) {
    char hdr[kUmemPayloadOffset + 1];
    std::memset(hdr, 0, sizeof(hdr));
    const int w = snprintf(hdr, sizeof(hdr), "%zu\n%zu\n", num, dim);
    if (w < 0 || (size_t)w > kUmemPayloadOffset) throw FaissException("header does not fit in 20 bytes: " + path);
    if (fwrite(hdr, 1, kUmemPayloadOffset, f.get()) != kUmemPayloadOffset) throw FaissException("write error on " + path);
  }
  fseek(f.get(), (long)(kUmemPayloadOffset + elem_size * offset), SEEK_SET);
  if (len && fwrite(ptr, elem_size, len, f.get()) != len) throw FaissException("write error on " + path);
}

void umem_read(const std::string& path, void* ptr, size_t elem_size, size_t len, size_t offset) {
  File f = open_or_throw(path, "rb");
  fseek(f.get(), (long)(kUmemPayloadOffset + elem_size * offset), SEEK_SET);
  if (len && fread(ptr, elem_size, len, f.get()) != len) throw FaissException("short read in " + path);
}

}  // namespace faiss
