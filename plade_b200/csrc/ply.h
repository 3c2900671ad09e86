#pragma once
#include <cstddef>
#include <string>
#include <vector>
namespace plade {
std::string file_extension(const std::string &file_name);
// interleaved x y z nx ny nz per vertex; false on any failure (message on std::cerr)
bool load_ply_xyzn(const std::string &file_name, std::vector<float> &out);
// same, decoding straight into a buffer the caller provides once the vertex count is known (batch mode reads
// the binary records directly into page-locked memory: one copy from the page cache, no pageable staging)
typedef float *(*PlyAlloc)(size_t n_floats, void *user);
bool load_ply_xyzn_into(const std::string &file_name, PlyAlloc alloc, void *user, size_t &n_points);
}  // namespace plade
