#pragma once
#include <string>
#include <vector>
namespace plade {
std::string file_extension(const std::string &file_name);
// interleaved x y z nx ny nz per vertex; false on any failure (message on std::cerr)
bool load_ply_xyzn(const std::string &file_name, std::vector<float> &out);
}  // namespace plade
