// PLY ingest for the file-name overload (PLADE/plade.cpp:665-707).  The reference goes through rply
// callbacks -> double arrays -> "point"/"normal" vec3 properties -> PointNormal (PLADE/ply_reader.cpp:46-148,
// PLADE/util.cpp:1505-1546).  Here: one header parse, then either a straight read of the binary
// little-endian `float x y z nx ny nz` records (the layout of all sample_data files) or a generic
// per-property decode (ascii / other scalar types / extra properties), into interleaved float[6].
// Restriction (documented): `vertex` must be the first element of the file, as in every file the reference ships.
#include "ply.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

namespace plade {

namespace {
struct Prop { std::string type, name; int size; };
int type_size(const std::string &t) {
  if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
  if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
  if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
  if (t == "double" || t == "float64") return 8;
  return 0;
}
double decode(const unsigned char *p, const std::string &t, bool swap_bytes = false) {
  unsigned char tmp[8];
  if (swap_bytes) {                       // binary_big_endian files (rply reads both byte orders)
    const int n = type_size(t);
    for (int i = 0; i < n; ++i) tmp[i] = p[n - 1 - i];
    p = tmp;
  }
  if (t == "float" || t == "float32") { float v; memcpy(&v, p, 4); return v; }
  if (t == "double" || t == "float64") { double v; memcpy(&v, p, 8); return v; }
  if (t == "char" || t == "int8") return (signed char) p[0];
  if (t == "uchar" || t == "uint8") return p[0];
  if (t == "short" || t == "int16") { short v; memcpy(&v, p, 2); return v; }
  if (t == "ushort" || t == "uint16") { unsigned short v; memcpy(&v, p, 2); return v; }
  if (t == "int" || t == "int32") { int v; memcpy(&v, p, 4); return v; }
  if (t == "uint" || t == "uint32") { unsigned v; memcpy(&v, p, 4); return v; }
  return 0;
}
}  // namespace

std::string file_extension(const std::string &file_name) {      // extension(), PLADE/util.cpp:525-531
  std::string::size_type dot = file_name.find_last_of('.');
  std::string::size_type slash = file_name.find_last_of("/\\");
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return std::string("");
  return std::string(file_name.begin() + dot + 1, file_name.end());
}

bool load_ply_xyzn(const std::string &file_name, std::vector<float> &out) {
  out.clear();
  struct Grow {
    static float *alloc(size_t n_floats, void *user) {
      std::vector<float> *v = static_cast<std::vector<float> *>(user);
      v->resize(n_floats);
      return v->data();
    }
  };
  size_t n = 0;
  if (!load_ply_xyzn_into(file_name, &Grow::alloc, &out, n)) { out.clear(); return false; }
  return true;
}

bool load_ply_xyzn_into(const std::string &file_name, PlyAlloc alloc, void *user, size_t &n_points) {
  n_points = 0;
  FILE *f = fopen(file_name.c_str(), "rb");
  if (!f) { std::cerr << "failed to open ply file: " << file_name << std::endl; return false; }
  char line[1024];
  bool binary_le = false, binary_be = false, ascii = false, in_vertex = false, vertex_first = true, seen_elem = false;
  size_t n_vertex = 0;
  std::vector<Prop> props;
  bool header_ok = false;
  if (!fgets(line, sizeof(line), f) || strncmp(line, "ply", 3) != 0) { fclose(f); std::cerr << "failed to read ply header" << std::endl; return false; }
  while (fgets(line, sizeof(line), f)) {
    std::istringstream ss(line);
    std::string tok;
    ss >> tok;
    if (tok == "format") { std::string fmt; ss >> fmt; binary_le = fmt == "binary_little_endian"; binary_be = fmt == "binary_big_endian"; ascii = fmt == "ascii"; }
    else if (tok == "element") {
      std::string name; size_t cnt;
      ss >> name >> cnt;
      in_vertex = name == "vertex";
      if (in_vertex) { n_vertex = cnt; vertex_first = !seen_elem; }
      seen_elem = true;
    } else if (tok == "property" && in_vertex) {
      Prop p;
      ss >> p.type;
      if (p.type == "list") { fclose(f); std::cerr << "failed to read ply header" << std::endl; return false; }
      ss >> p.name;
      p.size = type_size(p.type);
      // an unknown scalar type (rply knows the eight above) would silently shift every later field offset
      if (p.size == 0) { fclose(f); std::cerr << "failed to read ply header" << std::endl; return false; }
      props.push_back(p);
    } else if (tok == "end_header") { header_ok = true; break; }
  }
  if (!header_ok || (!binary_le && !binary_be && !ascii) || !vertex_first) {
    fclose(f);
    std::cerr << "failed to read ply header" << std::endl;
    return false;
  }
  int ix[6] = {-1, -1, -1, -1, -1, -1};
  const char *want[6] = {"x", "y", "z", "nx", "ny", "nz"};
  for (size_t i = 0; i < props.size(); ++i)
    for (int k = 0; k < 6; ++k) if (props[i].name == want[k]) ix[k] = (int) i;
  if (ix[0] < 0 || ix[1] < 0 || ix[2] < 0 || ix[3] < 0 || ix[4] < 0 || ix[5] < 0) {
    fclose(f);
    std::cerr << "the number of points does not equal to the number of normals in the file" << std::endl;
    return false;
  }
  // the header's vertex count is untrusted: a binary file must actually hold n_vertex records (and an ascii one at
  // least two characters per value) before anything of that size is allocated
  {
    size_t rec = 0;
    for (const Prop &q : props) rec += (size_t) q.size;
    const long here = ftell(f);
    long end = here;
    if (here >= 0 && fseek(f, 0, SEEK_END) == 0) { end = ftell(f); fseek(f, here, SEEK_SET); }
    const unsigned long long remaining = end > here ? (unsigned long long) (end - here) : 0ull;
    const unsigned long long need = (binary_le || binary_be) ? (unsigned long long) n_vertex * rec : (unsigned long long) n_vertex * props.size() * 2ull;
    if (need > remaining + 1) { fclose(f); std::cerr << "failed to read ply file: " << file_name << std::endl; return false; }
  }
  float *out = alloc(n_vertex * 6, user);
  if (!out && n_vertex) { fclose(f); std::cerr << "failed to read ply file: " << file_name << std::endl; return false; }
  bool ok = true;
  if (binary_le || binary_be) {
    bool fast = binary_le && props.size() == 6;
    for (int k = 0; k < 6 && fast; ++k) fast = ix[k] == k && (props[k].type == "float" || props[k].type == "float32");
    if (fast) {
      ok = fread(out, sizeof(float) * 6, n_vertex, f) == n_vertex;
    } else {
      size_t rec = 0;
      std::vector<size_t> off(props.size());
      for (size_t i = 0; i < props.size(); ++i) { off[i] = rec; rec += props[i].size; }
      std::vector<unsigned char> buf(rec * 4096);
      size_t done = 0;
      while (done < n_vertex && ok) {
        size_t chunk = std::min<size_t>(4096, n_vertex - done);
        ok = fread(buf.data(), rec, chunk, f) == chunk;
        for (size_t r = 0; r < chunk && ok; ++r)
          for (int k = 0; k < 6; ++k) out[(done + r) * 6 + k] = (float) decode(buf.data() + r * rec + off[ix[k]], props[ix[k]].type, binary_be);
        done += chunk;
      }
    }
  } else {
    std::vector<double> vals(props.size());
    for (size_t r = 0; r < n_vertex && ok; ++r) {
      for (size_t i = 0; i < props.size(); ++i) if (fscanf(f, "%lf", &vals[i]) != 1) { ok = false; break; }
      for (int k = 0; k < 6 && ok; ++k) out[r * 6 + k] = (float) vals[ix[k]];
    }
  }
  fclose(f);
  if (!ok) { std::cerr << "failed to read ply file: " << file_name << std::endl; return false; }
  n_points = n_vertex;
  return n_vertex > 0;
}

}  // namespace plade
