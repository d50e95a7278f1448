// Dataset / matrix file formats of the reference drivers (reference filehelper.{h,cpp}):
//   * TexMex .fvecs / .ivecs / .bvecs ("Jegou" format): every vector is int32 dim followed by dim values of
//     float / int32 / uint8                                              (reference readJegou, filehelper.cpp:106-250)
//   * .umem / .imem matrices: ASCII header "<num>\n<dim>\n", binary payload starting at byte 20
//                                                                         (reference write/read/header, :252-345)
// Written from scratch; only the formats are shared.  Errors throw faiss::FaissException.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace faiss {

/// header of a TexMex file with elements of `elem_size` bytes: number of vectors (from the file size) and dimension
void vecs_header(const std::string& path, size_t elem_size, size_t* n, size_t* d);
/// vectors [start, start + num) of a TexMex file, dimension prefix stripped; num == 0 means "to the end"
std::vector<float> fvecs_read(const std::string& path, size_t* n, size_t* d, size_t start = 0, size_t num = 0);
std::vector<int32_t> ivecs_read(const std::string& path, size_t* n, size_t* d, size_t start = 0, size_t num = 0);
std::vector<uint8_t> bvecs_read(const std::string& path, size_t* n, size_t* d, size_t start = 0, size_t num = 0);
void fvecs_write(const std::string& path, const float* x, size_t n, size_t d);
void ivecs_write(const std::string& path, const int32_t* x, size_t n, size_t d);
void bvecs_write(const std::string& path, const uint8_t* x, size_t n, size_t d);

constexpr size_t kUmemPayloadOffset = 20;  // reference filehelper.cpp:267,320
/// "<num>\n<dim>\n" header
void umem_header(const std::string& path, size_t* num, size_t* dim);
/// write `len` elements at element offset `offset` of the payload; offset == 0 (re)creates the file with its header
void umem_write(const std::string& path, size_t num, size_t dim, const void* ptr, size_t elem_size, size_t len,
                size_t offset = 0);
/// read `len` elements from element offset `offset` of the payload
void umem_read(const std::string& path, void* ptr, size_t elem_size, size_t len, size_t offset = 0);

}  // namespace faiss
