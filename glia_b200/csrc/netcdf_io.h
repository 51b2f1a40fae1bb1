// netcdf_io.h -- host-side reader / writer of the NetCDF classic formats (CDF-1, CDF-2 "64-bit
// offset", and the CDF-5 header widths) for the one layout GLIA uses: a fixed-size variable
// "data" over dimensions (x, y, z), x slowest (dataIn / dataOut, src/utils/IO.cpp:511-614, which
// go through PnetCDF; the file format is the public NetCDF classic specification).  Plain C++,
// no CUDA: the engine moves the block to / from the device.
#pragma once
#include <fcntl.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace glia {
namespace nc {

struct Error { std::string msg; };

enum { NC_BYTE = 1, NC_CHAR = 2, NC_SHORT = 3, NC_INT = 4, NC_FLOAT = 5, NC_DOUBLE = 6,
       NC_UBYTE = 7, NC_USHORT = 8, NC_UINT = 9, NC_INT64 = 10, NC_UINT64 = 11 };
enum { TAG_DIMENSION = 0x0A, TAG_VARIABLE = 0x0B, TAG_ATTRIBUTE = 0x0C };

inline int type_size(int t) {
  switch (t) {
    case NC_BYTE: case NC_CHAR: case NC_UBYTE: return 1;
    case NC_SHORT: case NC_USHORT: return 2;
    case NC_INT: case NC_UINT: case NC_FLOAT: return 4;
    case NC_DOUBLE: case NC_INT64: case NC_UINT64: return 8;
    default: throw Error{"netcdf: unknown nc_type " + std::to_string(t)};
  }
}

struct Reader {
  FILE* f = nullptr;
  int version = 0;
  explicit Reader(const char* path) {
    f = std::fopen(path, "rb");
    if (!f) throw Error{std::string("netcdf: cannot open ") + path};
  }
  ~Reader() { if (f) std::fclose(f); }
  void bytes(void* p, size_t n) {
    if (std::fread(p, 1, n, f) != n) throw Error{"netcdf: truncated file"};
  }
  uint32_t u32() { unsigned char b[4]; bytes(b, 4); return (uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3]; }
  uint64_t u64() { const uint64_t hi = u32(); return hi << 32 | u32(); }
  uint64_t nonneg() { return version == 5 ? u64() : u32(); }   // NON_NEG: 8 bytes in CDF-5
  uint64_t offset() { return version == 1 ? u32() : u64(); }   // begin: 4 bytes only in CDF-1
  std::string name() {
    const uint64_t n = nonneg();
    std::string s((size_t)n, '\0');
    if (n) bytes(&s[0], (size_t)n);
    const size_t pad = (4 - n % 4) % 4;
    char junk[4];
    if (pad) bytes(junk, pad);
    return s;
  }
  void skip(uint64_t n) {
    if (std::fseek(f, (long)n, SEEK_CUR) != 0) throw Error{"netcdf: seek failed"};
  }
  void skip_attrs() {
    const uint32_t tag = u32();
    const uint64_t n = nonneg();
    if (tag == 0 && n == 0) return;
    if (tag != TAG_ATTRIBUTE) throw Error{"netcdf: malformed attribute list"};
    for (uint64_t i = 0; i < n; ++i) {
      name();
      const int t = (int)u32();
      const uint64_t ne = nonneg();
      uint64_t b = ne * (uint64_t)type_size(t);
      b += (4 - b % 4) % 4;
      skip(b);
    }
  }
};

struct VarInfo {
  int type = 0;
  uint64_t begin = 0;
  std::vector<uint64_t> shape;
};

// locate a fixed-size variable and return its type, shape and file offset
inline VarInfo find_var(Reader& r, const char* var) {
  char magic[4];
  r.bytes(magic, 4);
  if (std::memcmp(magic, "CDF", 3) != 0 || !(magic[3] == 1 || magic[3] == 2 || magic[3] == 5))
    throw Error{"netcdf: not a classic-format file (CDF-1/2/5)"};
  r.version = magic[3];
  r.nonneg();  // numrecs
  std::vector<uint64_t> dims;
  {
    const uint32_t tag = r.u32();
    const uint64_t n = r.nonneg();
    if (!(tag == 0 && n == 0)) {
      if (tag != TAG_DIMENSION) throw Error{"netcdf: malformed dimension list"};
      for (uint64_t i = 0; i < n; ++i) { r.name(); dims.push_back(r.nonneg()); }
    }
  }
  r.skip_attrs();
  const uint32_t tag = r.u32();
  const uint64_t nv = r.nonneg();
  if (tag == 0 && nv == 0) throw Error{"netcdf: no variables"};
  if (tag != TAG_VARIABLE) throw Error{"netcdf: malformed variable list"};
  for (uint64_t i = 0; i < nv; ++i) {
    const std::string nm = r.name();
    const uint64_t rank = r.nonneg();
    VarInfo v;
    for (uint64_t d = 0; d < rank; ++d) {
      const uint64_t id = r.nonneg();
      if (id >= dims.size()) throw Error{"netcdf: dimension id out of range"};
      v.shape.push_back(dims[(size_t)id]);
    }
    r.skip_attrs();
    v.type = (int)r.u32();
    r.nonneg();  // vsize
    v.begin = r.offset();
    if (nm == var) return v;
  }
  throw Error{std::string("netcdf: variable '") + var + "' not found"};
}

template <typename S>
inline S load_be(const unsigned char* p) {
  unsigned char b[sizeof(S)];
  for (size_t i = 0; i < sizeof(S); ++i) b[i] = p[sizeof(S) - 1 - i];
  S v;
  std::memcpy(&v, b, sizeof(S));
  return v;
}

// read rows [x0, x0 + nx) of variable "data" (shape n[0..2]) into out[nx * n1 * n2], converting to T
template <typename T>
inline void read_block(const char* path, const int n[3], int x0, int nx, T* out) {
  Reader r(path);
  const VarInfo v = find_var(r, "data");
  if (v.shape.size() != 3 || v.shape[0] != (uint64_t)n[0] || v.shape[1] != (uint64_t)n[1] || v.shape[2] != (uint64_t)n[2])
    throw Error{std::string("netcdf: 'data' in ") + path + " does not have the grid's shape"};
  const int ts = type_size(v.type);
  const size_t plane = (size_t)n[1] * n[2], count = (size_t)nx * plane;
  if (std::fseek(r.f, (long)(v.begin + (uint64_t)x0 * plane * ts), SEEK_SET) != 0) throw Error{"netcdf: seek failed"};
  std::vector<unsigned char> raw(count * ts);
  r.bytes(raw.data(), raw.size());
  for (size_t i = 0; i < count; ++i) {
    const unsigned char* p = raw.data() + i * ts;
    switch (v.type) {
      case NC_BYTE: out[i] = (T)(signed char)p[0]; break;
      case NC_CHAR: case NC_UBYTE: out[i] = (T)p[0]; break;
      case NC_SHORT: out[i] = (T)load_be<int16_t>(p); break;
      case NC_USHORT: out[i] = (T)load_be<uint16_t>(p); break;
      case NC_INT: out[i] = (T)load_be<int32_t>(p); break;
      case NC_UINT: out[i] = (T)load_be<uint32_t>(p); break;
      case NC_FLOAT: out[i] = (T)load_be<float>(p); break;
      case NC_DOUBLE: out[i] = (T)load_be<double>(p); break;
      case NC_INT64: out[i] = (T)load_be<int64_t>(p); break;
      default: out[i] = (T)load_be<uint64_t>(p); break;
    }
  }
}

struct Buf {
  std::vector<unsigned char> b;
  void u32(uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((unsigned char)(v >> s)); }
  void u64(uint64_t v) { u32((uint32_t)(v >> 32)); u32((uint32_t)v); }
  void name(const char* s) {
    const size_t n = std::strlen(s);
    u32((uint32_t)n);
    for (size_t i = 0; i < n; ++i) b.push_back((unsigned char)s[i]);
    while (b.size() % 4) b.push_back(0);
  }
};

// write rows [x0, x0 + nx) of a CDF-2 file laid out exactly like dataOut's (dims x y z, variable
// "data" NC_FLOAT / NC_DOUBLE, global attribute "CDF-5 mode" = 0).  Every writer of a
// slab-decomposed field calls this with its own rows; whoever has x0 == 0 also writes the header.
template <typename T>
inline void write_block(const char* path, const int n[3], int x0, int nx, const T* in) {
  Buf h;
  h.b = {'C', 'D', 'F', 2};
  h.u32(0);  // numrecs
  h.u32(TAG_DIMENSION); h.u32(3);
  h.name("x"); h.u32((uint32_t)n[0]);
  h.name("y"); h.u32((uint32_t)n[1]);
  h.name("z"); h.u32((uint32_t)n[2]);
  h.u32(TAG_ATTRIBUTE); h.u32(1);
  h.name("CDF-5 mode"); h.u32(NC_INT); h.u32(1); h.u32(0);
  h.u32(TAG_VARIABLE); h.u32(1);
  h.name("data"); h.u32(3); h.u32(0); h.u32(1); h.u32(2);
  h.u32(0); h.u32(0);  // no variable attributes
  h.u32(sizeof(T) == 4 ? NC_FLOAT : NC_DOUBLE);
  const uint64_t vbytes = (uint64_t)n[0] * n[1] * n[2] * sizeof(T);
  h.u32(vbytes > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)vbytes);
  const uint64_t begin = h.b.size() + 8;
  h.u64(begin);
  // POSIX open without truncation: the ranks of a slab-decomposed field write their rows in any order
  const int fd = ::open(path, O_CREAT | O_WRONLY, 0644);
  if (fd < 0) throw Error{std::string("netcdf: cannot create ") + path};
  bool ok = true;
  if (x0 == 0) {
    ok = ::ftruncate(fd, (off_t)(begin + vbytes)) == 0;
    ok = ok && ::pwrite(fd, h.b.data(), h.b.size(), 0) == (ssize_t)h.b.size();
  }
  const size_t plane = (size_t)n[1] * n[2], count = (size_t)nx * plane;
  std::vector<unsigned char> raw(count * sizeof(T));
  for (size_t i = 0; i < count; ++i) {
    unsigned char b[sizeof(T)];
    std::memcpy(b, &in[i], sizeof(T));
    for (size_t j = 0; j < sizeof(T); ++j) raw[i * sizeof(T) + j] = b[sizeof(T) - 1 - j];
  }
  size_t done = 0;
  const off_t at = (off_t)(begin + (uint64_t)x0 * plane * sizeof(T));
  while (ok && done < raw.size()) {
    const ssize_t w = ::pwrite(fd, raw.data() + done, raw.size() - done, at + (off_t)done);
    if (w <= 0) ok = false; else done += (size_t)w;
  }
  ok = (::close(fd) == 0) && ok;
  if (!ok) throw Error{std::string("netcdf: write failed: ") + path};
}

}  // namespace nc
}  // namespace glia
