// HDF5 files of the reference: tnc/src/io/hdf5.rs (load_tensor :28-34, load_data :37-43, store_data :46-52,
// read_tensor :54-88, read_data :90-103, write_data :105-113) and the TensorData::File payload
// (tensornetwork/tensordata.rs:43-49).
//
// The reference goes through the crate hdf5-metno 0.12.4 -> libhdf5 (C), neither of which is part of the reference
// tree or of this image.  This file is a self-contained restatement of the PUBLISHED file format ("HDF5 File Format
// Specification Version 3.0") for the subset those calls produce and consume:
//
//   read : superblock 0-3 (user block / base address), object headers 1 and 2 (continuations), groups as symbol
//          tables (B-tree v1 + local heap + SNOD, any depth) or compact link messages, dataspace 1/2, datatypes
//          fixed-point / IEEE float / compound of two floats (= Complex, any member names, versions 1-3), data layout
//          1-4 contiguous / compact / chunked-by-B-tree-v1 with deflate, shuffle and fletcher32, attributes 1-3.
//          Dense link / attribute storage (fractal heaps), layout-4 chunk indexes, external links, virtual and
//          external storage, committed (shared) datatypes -> TNCB_ERR_UNSUPPORTED naming the feature.
//   write: what libhdf5 writes with default property lists ("earliest" format bounds), i.e. superblock 0, symbol-table
//          groups (leaf K 4, internal K 16), version-1 object headers, contiguous little-endian datasets of the
//          compound {re: f64 @0, im: f64 @8} (hdf5-types' H5Type for num_complex::Complex<f64>), u64 attributes.
//
// Pinned by: a file written by libhdf5 itself that ships in this image (scipy's MATLAB 7.3 fixture, read in
// tests/test_hdf5.py), an independent Python parser of the writer's output in the same test file, and the reference's
// own three tests (hdf5.rs:196-257) replayed through writer + reader.  There is no byte-level fixture in the reference.
// Host code only; every malformed input is reported as TNCB_ERR_IO, never followed out of the mapping.
#include "internal.h"
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace tncb {
namespace h5 {

constexpr uint64_t UNDEF = ~0ull;
static const uint8_t kSignature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};

struct IoError : std::runtime_error {
  int code;
  IoError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
[[noreturn]] static void bad(const std::string& m) { throw IoError(TNCB_ERR_IO, "HDF5: " + m); }
[[noreturn]] static void unsupported(const std::string& m) { throw IoError(TNCB_ERR_UNSUPPORTED, "HDF5: " + m + " is outside the supported subset"); }

// ---- bounded little-endian cursor -------------------------------------------------------
struct Cur {
  const uint8_t* p;
  const uint8_t* end;
  const uint8_t* take(size_t n) {
    if ((size_t)(end - p) < n) bad("truncated structure");
    const uint8_t* q = p; p += n; return q;
  }
  void skip(size_t n) { take(n); }
  size_t left() const { return (size_t)(end - p); }
  uint64_t uN(int n) {
    if (n < 1 || n > 8) bad("bad integer width");
    const uint8_t* q = take((size_t)n);
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)q[i] << (8 * i);
    return v;
  }
  uint8_t u8() { return (uint8_t)uN(1); }
  uint16_t u16() { return (uint16_t)uN(2); }
  uint32_t u32() { return (uint32_t)uN(4); }
  uint64_t u64() { return uN(8); }
};

struct Datatype {
  int cls = -1, version = 0;
  uint32_t size = 0;
  bool little = true, is_signed = false;
  bool ieee = false;                       // class 1 with the IEEE binary32 / binary64 field placement
  std::vector<std::string> member_name;    // class 6
  std::vector<uint32_t> member_offset;
  std::vector<Datatype> member_type;
};

struct Attribute {
  std::string name;
  Datatype type;
  std::vector<uint64_t> dims;
  const uint8_t* data = nullptr;
  size_t bytes = 0;
};

struct Filter { int id; uint32_t flags; std::vector<uint32_t> cd; };

struct Dataset {
  std::string name;
  Datatype type;
  bool have_type = false, have_space = false, have_layout = false;
  int space_kind = 1;                     // 0 scalar, 1 simple, 2 null
  std::vector<uint64_t> dims;
  int layout = -1;                        // 0 compact, 1 contiguous, 2 chunked (B-tree v1)
  uint64_t addr = UNDEF, bytes = 0;       // contiguous
  const uint8_t* compact = nullptr;       // compact
  std::vector<uint32_t> chunk;            // chunk dims in elements, then the element size
  uint64_t chunk_btree = UNDEF;
  std::vector<Filter> filters;
  std::vector<Attribute> attrs;
  uint64_t elems() const {
    if (space_kind == 2) return 0;
    uint64_t e = 1;
    for (uint64_t d : dims) e *= d;
    return e;
  }
};

struct Msg { int type; int flags; const uint8_t* data; size_t size; };

struct File {
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t len = 0;
  uint64_t base = 0;
  int O = 8, L = 8;
  uint64_t root_header = UNDEF;
  mutable uint64_t nodes_visited = 0;      // B-tree nodes walked so far: a cyclic (corrupted) tree must end in an error, not in a loop
  void visit_node() const { if (++nodes_visited > (1u << 22)) bad("B-tree with more than 2^22 nodes (cyclic?)"); }
  std::vector<Dataset> tensors;            // members of the opened group, ascending by name (H5_INDEX_NAME, H5_ITER_INC)

  ~File() {
    if (map) munmap(const_cast<uint8_t*>(map), len);
    if (fd >= 0) close(fd);
  }
  const uint8_t* at(uint64_t addr, uint64_t n) const {   // addr is relative to the base address
    if (addr == UNDEF) bad("undefined address followed");
    uint64_t a = base + addr;
    if (a < addr || a > len || n > len - a) bad("address " + std::to_string(addr) + " (+" + std::to_string(n) + ") lies outside the file");
    return map + a;
  }
  Cur cur(uint64_t addr, uint64_t n) const { const uint8_t* p = at(addr, n); return Cur{p, p + n}; }
  Cur cur_to_end(uint64_t addr) const { const uint8_t* p = at(addr, 0); return Cur{p, map + len}; }
};

// ---- superblock ----------------------------------------------------------------------------
static void read_superblock(File& f) {
  uint64_t off = 0;
  bool found = false;
  while (off + 8 <= f.len) {                // 0, 512, 1024, 2048, ... (user block sizes)
    if (std::memcmp(f.map + off, kSignature, 8) == 0) { found = true; break; }
    off = off ? off * 2 : 512;
  }
  if (!found) bad("not an HDF5 file (no signature)");
  Cur c{f.map + off + 8, f.map + f.len};
  int version = c.u8();
  if (version == 0 || version == 1) {
    c.skip(3);                               // free-space version, root-group version, reserved
    c.skip(1);                               // shared-header version
    f.O = c.u8(); f.L = c.u8();
    c.skip(1);
    c.skip(4);                               // group leaf K, group internal K
    c.skip(4);                               // consistency flags
    if (version == 1) c.skip(4);             // indexed-storage K, reserved
    if (f.O < 2 || f.O > 8 || f.L < 2 || f.L > 8) bad("bad size of offsets / lengths");
    uint64_t base = c.uN(f.O);
    c.skip((size_t)f.O);                     // free-space info
    c.skip((size_t)f.O);                     // end of file
    c.skip((size_t)f.O);                     // driver info
    // root group symbol table entry
    c.skip((size_t)f.O);                     // link name offset
    f.root_header = c.uN(f.O);
    f.base = base;
  } else if (version == 2 || version == 3) {
    f.O = c.u8(); f.L = c.u8();
    c.skip(1);
    if (f.O < 2 || f.O > 8 || f.L < 2 || f.L > 8) bad("bad size of offsets / lengths");
    f.base = c.uN(f.O);
    c.skip((size_t)f.O);                     // superblock extension
    c.skip((size_t)f.O);                     // end of file
    f.root_header = c.uN(f.O);
  } else {
    bad("unknown superblock version " + std::to_string(version));
  }
  if (f.O < 8 && f.root_header == (UNDEF >> (64 - 8 * f.O))) f.root_header = UNDEF;
}

static uint64_t read_addr(const File& f, Cur& c) {
  uint64_t v = c.uN(f.O);
  if (f.O < 8 && v == (UNDEF >> (64 - 8 * f.O))) return UNDEF;
  return v;
}

// ---- object headers -------------------------------------------------------------------------
static std::vector<Msg> read_object_header(const File& f, uint64_t addr) {
  std::vector<Msg> msgs;
  const uint8_t* h = f.at(addr, 16);
  struct Chunk { uint64_t addr, size; };
  std::vector<Chunk> chunks;
  if (std::memcmp(h, "OHDR", 4) == 0) {                      // version 2
    Cur c = f.cur_to_end(addr);
    c.skip(4);
    if (c.u8() != 2) bad("object header: unknown version");
    int flags = c.u8();
    if (flags & 0x20) c.skip(16);
    if (flags & 0x10) c.skip(4);
    uint64_t size0 = c.uN(1 << (flags & 3));
    uint64_t first = (uint64_t)(c.p - f.map) - f.base;
    chunks.push_back({first, size0});
    const size_t mh = 4 + ((flags & 0x04) ? 2 : 0);
    for (size_t ci = 0; ci < chunks.size(); ci++) {
      if (chunks.size() > 4096) bad("object header: too many continuation blocks");
      Cur m = f.cur(chunks[ci].addr, chunks[ci].size);
      while (m.left() >= mh) {
        int type = m.u8();
        size_t sz = m.u16();
        int mflags = m.u8();
        if (flags & 0x04) m.skip(2);
        const uint8_t* d = m.take(sz);
        if (type == 0x10) {
          Cur k{d, d + sz};
          uint64_t ca = read_addr(f, k), cl = k.uN(f.L);
          if (cl < 8) bad("object header: continuation block too small");
          if (std::memcmp(f.at(ca, 4), "OCHK", 4) != 0) bad("object header: continuation block without its signature");
          chunks.push_back({ca + 4, cl - 8});
        } else if (type != 0) {
          msgs.push_back({type, mflags, d, sz});
        }
      }
    }
    return msgs;
  }
  Cur c = f.cur(addr, 16);
  if (c.u8() != 1) bad("object header: unknown version at address " + std::to_string(addr));
  c.skip(1);
  size_t nmsgs = c.u16();
  c.skip(4);                                                   // reference count
  uint64_t size0 = c.u32();
  chunks.push_back({addr + 16, size0});
  size_t seen = 0;
  for (size_t ci = 0; ci < chunks.size() && seen < nmsgs; ci++) {
    if (chunks.size() > 4096) bad("object header: too many continuation blocks");
    Cur m = f.cur(chunks[ci].addr, chunks[ci].size);
    while (m.left() >= 8 && seen < nmsgs) {
      int type = m.u16();
      size_t sz = m.u16();
      int mflags = m.u8();
      m.skip(3);
      const uint8_t* d = m.take(sz);
      seen++;
      if (type == 0x10) {
        Cur k{d, d + sz};
        uint64_t ca = read_addr(f, k), cl = k.uN(f.L);
        chunks.push_back({ca, cl});
      } else if (type != 0) {
        msgs.push_back({type, mflags, d, sz});
      }
    }
  }
  return msgs;
}

// ---- datatypes ---------------------------------------------------------------------------------
static size_t pad8(size_t n) { return (n + 7) & ~(size_t)7; }

static Datatype read_datatype(Cur& c, int depth = 0) {
  if (depth > 4) bad("datatype nesting too deep");
  Datatype t;
  uint8_t cv = c.u8();
  t.cls = cv & 0x0f; t.version = cv >> 4;
  uint32_t bits = (uint32_t)c.uN(3);
  t.size = c.u32();
  if (t.version < 1 || t.version > 5) bad("datatype: unknown version");
  switch (t.cls) {
    case 0: {                                 // fixed point
      t.little = !(bits & 1);
      t.is_signed = (bits & 8) != 0;
      unsigned bit_off = c.u16(), prec = c.u16();
      if (bit_off != 0 || prec != 8 * t.size) unsupported("integer type with padding bits");
      if (t.size != 1 && t.size != 2 && t.size != 4 && t.size != 8) unsupported("integer type of " + std::to_string(t.size) + " bytes");
      break;
    }
    case 1: {                                 // floating point
      if (bits & 0x40) unsupported("VAX byte order");
      t.little = !(bits & 1);
      unsigned sign_loc = (bits >> 8) & 0xff;
      unsigned bit_off = c.u16(), prec = c.u16();
      unsigned e_loc = c.u8(), e_size = c.u8(), m_loc = c.u8(), m_size = c.u8();
      uint32_t bias = c.u32();
      bool f64 = t.size == 8 && prec == 64 && sign_loc == 63 && e_loc == 52 && e_size == 11 && m_loc == 0 && m_size == 52 && bias == 1023;
      bool f32 = t.size == 4 && prec == 32 && sign_loc == 31 && e_loc == 23 && e_size == 8 && m_loc == 0 && m_size == 23 && bias == 127;
      t.ieee = bit_off == 0 && (f64 || f32) && ((bits >> 4) & 3) == 2;
      if (!t.ieee) unsupported("non-IEEE floating-point type");
      break;
    }
    case 6: {                                 // compound
      unsigned n = bits & 0xffff;
      if (n > 64) unsupported("compound type with more than 64 members");
      for (unsigned i = 0; i < n; i++) {
        const uint8_t* s = c.p;
        size_t nl = 0;
        while (true) { if (nl >= c.left()) bad("compound member name not terminated"); if (s[nl] == 0) break; nl++; }
        t.member_name.emplace_back((const char*)s, nl);
        c.skip(t.version < 3 ? pad8(nl + 1) : nl + 1);
        uint32_t off;
        if (t.version < 3) off = c.u32();
        else { int w = t.size < 0x100 ? 1 : t.size < 0x10000 ? 2 : t.size < 0x1000000 ? 3 : 4; off = (uint32_t)c.uN(w); }
        if (t.version == 1) {
          int dimensionality = c.u8();
          c.skip(3 + 4 + 4 + 16);
          if (dimensionality != 0) unsupported("array members of a compound type");
        }
        t.member_offset.push_back(off);
        t.member_type.push_back(read_datatype(c, depth + 1));
      }
      break;
    }
    case 2: unsupported("time datatype");
    case 3: {                                 // string (attributes of other producers; never converted)
      break;
    }
    case 4: c.skip(4); break;                 // bit field
    case 5: {                                 // opaque: tag padded to 8
      c.skip(pad8(bits & 0xff));
      break;
    }
    default:
      unsupported("datatype class " + std::to_string(t.cls));
  }
  return t;
}

// dataspace message: kind (0 scalar / 1 simple / 2 null) and dims
static void read_dataspace(const File& f, Cur& c, int* kind, std::vector<uint64_t>& dims) {
  int version = c.u8();
  int rank = c.u8();
  int flags = c.u8();
  if (version == 1) { c.skip(5); *kind = rank == 0 ? 0 : 1; }
  else if (version == 2) { *kind = c.u8(); if (*kind > 2) bad("dataspace: unknown type"); }
  else bad("dataspace: unknown version");
  if (rank > 32) bad("dataspace: rank above 32");
  dims.clear();
  for (int i = 0; i < rank; i++) dims.push_back(c.uN(f.L));
  if (flags & 1) c.skip((size_t)rank * f.L);
  uint64_t e = 1;
  for (uint64_t d : dims) { if (d && e > (UNDEF >> 1) / d) bad("dataspace: element count overflows"); e *= d; }
}

static Attribute read_attribute(const File& f, const Msg& m) {
  if (m.flags & 0x02) unsupported("shared attribute message");
  Cur c{m.data, m.data + m.size};
  Attribute a;
  int version = c.u8();
  int flags = c.u8();
  if (version < 1 || version > 3) bad("attribute: unknown version");
  size_t name_sz = c.u16(), type_sz = c.u16(), space_sz = c.u16();
  if (version == 3) c.skip(1);
  if (version >= 2 && (flags & 3)) unsupported("attribute with a shared datatype / dataspace");
  auto pad = [&](size_t n) { return version == 1 ? pad8(n) : n; };
  const uint8_t* nm = c.take(pad(name_sz));
  a.name.assign((const char*)nm, strnlen((const char*)nm, name_sz));
  { const uint8_t* p = c.take(pad(type_sz)); Cur t{p, p + type_sz}; a.type = read_datatype(t); }
  { const uint8_t* p = c.take(pad(space_sz)); Cur s{p, p + space_sz}; int kind; read_dataspace(f, s, &kind, a.dims); if (kind == 2) a.dims.assign(1, 0); }
  uint64_t n = 1;
  for (uint64_t d : a.dims) n *= d;
  if (a.type.size && n > c.left() / a.type.size) bad("attribute '" + a.name + "': data shorter than its dataspace");
  a.bytes = (size_t)n * a.type.size;
  a.data = c.take(a.bytes);
  return a;
}

static void read_filters(const Msg& m, std::vector<Filter>& out) {
  Cur c{m.data, m.data + m.size};
  int version = c.u8();
  int n = c.u8();
  if (version == 1) c.skip(6);
  else if (version != 2) bad("filter pipeline: unknown version");
  for (int i = 0; i < n; i++) {
    Filter fl;
    fl.id = c.u16();
    size_t name_len = (version == 1 || fl.id >= 256) ? c.u16() : 0;
    fl.flags = c.u16();
    int ncd = c.u16();
    c.skip(version == 1 ? pad8(name_len) : name_len);
    for (int k = 0; k < ncd; k++) fl.cd.push_back(c.u32());
    if (version == 1 && (ncd & 1)) c.skip(4);
    out.push_back(std::move(fl));
  }
}

static void read_layout(const File& f, const Msg& m, Dataset& d) {
  Cur c{m.data, m.data + m.size};
  int version = c.u8();
  if (version == 1 || version == 2) {
    int dimensionality = c.u8();
    int cls = c.u8();
    c.skip(5);
    if (dimensionality < 1 || dimensionality > 33) bad("data layout: bad dimensionality");
    uint64_t addr = UNDEF;
    if (cls != 0) addr = read_addr(f, c);
    std::vector<uint32_t> sz;
    for (int i = 0; i < dimensionality; i++) sz.push_back(c.u32());
    if (cls == 2) { sz.push_back(c.u32()); d.layout = 2; d.chunk_btree = addr; d.chunk = sz; }
    else if (cls == 1) { d.layout = 1; d.addr = addr; d.bytes = UNDEF; }       // size follows from dataspace x datatype
    else if (cls == 0) { uint32_t n = c.u32(); d.layout = 0; d.bytes = n; d.compact = c.take(n); }
    else bad("data layout: unknown class");
  } else if (version == 3 || version == 4) {
    int cls = c.u8();
    if (cls == 0) { size_t n = c.u16(); d.layout = 0; d.bytes = n; d.compact = c.take(n); }
    else if (cls == 1) { d.layout = 1; d.addr = read_addr(f, c); d.bytes = c.uN(f.L); }
    else if (cls == 2) {
      if (version == 4) unsupported("version-4 chunked layout (chunk indexes of the 'latest' file format)");
      int dimensionality = c.u8();
      if (dimensionality < 2 || dimensionality > 33) bad("data layout: bad dimensionality");
      d.layout = 2; d.chunk_btree = read_addr(f, c);
      for (int i = 0; i < dimensionality; i++) d.chunk.push_back(c.u32());
    } else if (cls == 3) unsupported("virtual dataset");
    else bad("data layout: unknown class");
  } else {
    bad("data layout: unknown version");
  }
  d.have_layout = true;
}

static Dataset read_dataset(const File& f, const std::string& name, uint64_t header) {
  Dataset d;
  d.name = name;
  for (const Msg& m : read_object_header(f, header)) {
    switch (m.type) {
      case 0x01: { Cur c{m.data, m.data + m.size}; read_dataspace(f, c, &d.space_kind, d.dims); d.have_space = true; break; }
      case 0x03: {
        if (m.flags & 0x02) unsupported("committed (shared) datatype of dataset '" + name + "'");
        Cur c{m.data, m.data + m.size}; d.type = read_datatype(c); d.have_type = true; break;
      }
      case 0x08: read_layout(f, m, d); break;
      case 0x0b: read_filters(m, d.filters); break;
      case 0x0c:                              // an attribute of a type outside the subset (strings of other producers,
        try { d.attrs.push_back(read_attribute(f, m)); }      // references, enums ...) is skipped, not fatal: only the
        catch (const IoError& e) { if (e.code != TNCB_ERR_UNSUPPORTED) throw; }   // integer arrays `bids` / `tids` are read
        break;
      case 0x15: {                            // attribute info: dense storage when the fractal heap exists
        Cur c{m.data, m.data + m.size};
        c.skip(1); int flags = c.u8();
        if (flags & 1) c.skip(2);
        if (read_addr(f, c) != UNDEF) unsupported("dense attribute storage (more than 8 attributes on '" + name + "')");
        break;
      }
      case 0x07: unsupported("external data files");
      default: break;
    }
  }
  if (!d.have_space || !d.have_type || !d.have_layout) bad("'" + name + "' is not a dataset (dataspace / datatype / layout message missing)");
  return d;
}

// ---- groups -----------------------------------------------------------------------------------------
struct Link { std::string name; uint64_t header; };

static void walk_group_btree(const File& f, uint64_t node, const uint8_t* heap, size_t heap_size, std::vector<Link>& out, int depth,
                             int expect_level = -1) {
  if (depth > 16) bad("group B-tree deeper than 16 levels");
  f.visit_node();
  Cur c = f.cur_to_end(node);
  if (std::memcmp(c.take(4), "TREE", 4) != 0) bad("group B-tree node without its signature");
  if (c.u8() != 0) bad("group B-tree node of the wrong type");
  int level = c.u8();
  if (expect_level >= 0 && level != expect_level) bad("group B-tree: a child is not one level below its parent");
  size_t used = c.u16();
  c.skip(2 * (size_t)f.O);
  if (used > 65535 / 2) bad("group B-tree: too many entries");
  for (size_t i = 0; i < used; i++) {
    c.skip((size_t)f.L);                      // key i
    uint64_t child = read_addr(f, c);
    if (level > 0) { walk_group_btree(f, child, heap, heap_size, out, depth + 1, level - 1); continue; }
    Cur s = f.cur_to_end(child);
    if (std::memcmp(s.take(4), "SNOD", 4) != 0) bad("symbol table node without its signature");
    s.skip(2);
    size_t nsym = s.u16();
    for (size_t k = 0; k < nsym; k++) {
      uint64_t name_off = s.uN(f.O);
      uint64_t header = read_addr(f, s);
      unsigned cache = s.u32();
      s.skip(4 + 16);
      if (name_off >= heap_size) bad("symbol name outside the local heap");
      size_t nl = strnlen((const char*)heap + name_off, heap_size - name_off);
      if (nl == heap_size - name_off) bad("symbol name not terminated");
      if (cache == 2) unsupported("symbolic links");
      out.push_back({std::string((const char*)heap + name_off, nl), header});
      if (out.size() > (1u << 24)) bad("group with more than 2^24 members");
    }
  }
}

static std::vector<Link> list_group(const File& f, uint64_t header) {
  std::vector<Link> links;
  for (const Msg& m : read_object_header(f, header)) {
    Cur c{m.data, m.data + m.size};
    if (m.type == 0x11) {                     // symbol table: B-tree + local heap
      uint64_t btree = read_addr(f, c), heap = read_addr(f, c);
      Cur h = f.cur_to_end(heap);
      if (std::memcmp(h.take(4), "HEAP", 4) != 0) bad("local heap without its signature");
      h.skip(4);
      uint64_t seg_size = h.uN(f.L);
      h.skip((size_t)f.L);
      uint64_t seg = read_addr(f, h);
      const uint8_t* data = f.at(seg, seg_size);
      walk_group_btree(f, btree, data, (size_t)seg_size, links, 0);
    } else if (m.type == 0x06) {              // link message (compact new-style group)
      if (c.u8() != 1) bad("link message: unknown version");
      int flags = c.u8();
      int type = (flags & 0x08) ? c.u8() : 0;
      if (flags & 0x04) c.skip(8);
      if (flags & 0x10) c.skip(1);
      uint64_t nl = c.uN(1 << (flags & 3));
      const uint8_t* nm = c.take((size_t)nl);
      if (type != 0) unsupported("soft / external links");
      links.push_back({std::string((const char*)nm, (size_t)nl), read_addr(f, c)});
    } else if (m.type == 0x02) {              // link info: dense storage when the fractal heap exists
      c.skip(1); int flags = c.u8();
      if (flags & 1) c.skip(8);
      if (read_addr(f, c) != UNDEF) unsupported("dense link storage (groups of the 'latest' file format with many members)");
    }
  }
  std::sort(links.begin(), links.end(), [](const Link& a, const Link& b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
  return links;
}

static uint64_t find_group(const File& f, const std::string& path) {
  uint64_t cur = f.root_header;
  size_t i = 0;
  while (i < path.size()) {
    while (i < path.size() && path[i] == '/') i++;
    size_t j = i;
    while (j < path.size() && path[j] != '/') j++;
    if (j == i) break;
    std::string part = path.substr(i, j - i);
    i = j;
    if (part == ".") continue;
    bool hit = false;
    for (const Link& l : list_group(f, cur)) if (l.name == part) { cur = l.header; hit = true; break; }
    if (!hit) bad("group '" + path + "' does not exist");
  }
  return cur;
}

// ---- element conversion ---------------------------------------------------------------------------------
static inline double load_float(const uint8_t* p, const Datatype& t) {
  if (t.size == 8) {
    uint64_t v; std::memcpy(&v, p, 8);
    if (!t.little) v = __builtin_bswap64(v);
    double d; std::memcpy(&d, &v, 8); return d;
  }
  uint32_t v; std::memcpy(&v, p, 4);
  if (!t.little) v = __builtin_bswap32(v);
  float s; std::memcpy(&s, &v, 4); return (double)s;
}

struct ElemKind { int kind; const Datatype* re; const Datatype* im; uint32_t off_re, off_im; };   // 0 complex compound, 1 real float

static ElemKind classify(const Dataset& d) {
  const Datatype& t = d.type;
  if (t.cls == 6 && t.member_type.size() == 2 && t.member_type[0].cls == 1 && t.member_type[1].cls == 1 &&
      t.member_type[0].size == t.member_type[1].size && t.member_offset[0] != t.member_offset[1] &&
      (uint64_t)t.member_offset[0] + t.member_type[0].size <= t.size && (uint64_t)t.member_offset[1] + t.member_type[1].size <= t.size) {
    int r = t.member_offset[0] < t.member_offset[1] ? 0 : 1;          // real part first in memory (num_complex layout)
    return {0, &t.member_type[r], &t.member_type[1 - r], t.member_offset[r], t.member_offset[1 - r]};
  }
  if (t.cls == 1) return {1, &t, nullptr, 0, 0};
  throw IoError(TNCB_ERR_INVALID, "HDF5: dataset '" + d.name + "' holds neither complex ({re, im} compound of two floats) nor real floating-point data");
}

static void convert(const ElemKind& k, uint32_t esize, const uint8_t* src, uint64_t n, double* out) {
  if (k.kind == 0 && esize == 16 && k.off_re == 0 && k.off_im == 8 && k.re->little && k.im->little && k.re->size == 8) {
    std::memcpy(out, src, (size_t)n * 16);                             // the native layout: one copy
    return;
  }
  for (uint64_t i = 0; i < n; i++) {
    const uint8_t* e = src + i * esize;
    out[2 * i] = load_float(e + k.off_re, *k.re);
    out[2 * i + 1] = k.kind == 0 ? load_float(e + k.off_im, *k.im) : 0.0;
  }
}

// ---- chunked storage ------------------------------------------------------------------------------------------
static void unshuffle(std::vector<uint8_t>& buf, size_t esize) {
  if (esize <= 1) return;
  size_t n = buf.size() / esize;
  std::vector<uint8_t> out(buf.size());
  for (size_t b = 0; b < esize; b++)
    for (size_t i = 0; i < n; i++) out[i * esize + b] = buf[b * n + i];
  std::copy(buf.begin() + n * esize, buf.end(), out.begin() + n * esize);
  buf.swap(out);
}

static void read_chunks(const File& f, const Dataset& d, const ElemKind& k, uint64_t node, double* out, int depth, int expect_level = -1) {
  if (depth > 16) bad("chunk B-tree deeper than 16 levels");
  f.visit_node();
  const size_t rank = d.dims.size();
  const uint32_t esize = d.type.size;
  Cur c = f.cur_to_end(node);
  if (std::memcmp(c.take(4), "TREE", 4) != 0) bad("chunk B-tree node without its signature");
  if (c.u8() != 1) bad("chunk B-tree node of the wrong type");
  int level = c.u8();
  if (expect_level >= 0 && level != expect_level) bad("chunk B-tree: a child is not one level below its parent");
  size_t used = c.u16();
  c.skip(2 * (size_t)f.O);
  uint64_t chunk_elems = 1;
  for (size_t i = 0; i < rank; i++) {
    if (chunk_elems > 0xffffffffull / esize / d.chunk[i]) bad("dataset '" + d.name + "': chunk above 4 GiB");   // the format's limit
    chunk_elems *= d.chunk[i];
  }
  for (size_t i = 0; i < used; i++) {
    uint32_t nbytes = c.u32();
    uint32_t mask = c.u32();
    std::vector<uint64_t> off(rank + 1);
    for (size_t q = 0; q <= rank; q++) off[q] = c.u64();
    uint64_t child = read_addr(f, c);
    if (level > 0) { read_chunks(f, d, k, child, out, depth + 1, level - 1); continue; }
    const uint8_t* raw = f.at(child, nbytes);
    std::vector<uint8_t> buf;
    const uint8_t* data = raw;
    if (!d.filters.empty()) {
      buf.assign(raw, raw + nbytes);
      for (size_t q = d.filters.size(); q-- > 0;) {
        if (mask & (1u << q)) continue;
        const Filter& fl = d.filters[q];
        if (fl.id == 1) {                      // deflate (at most ~1032 : 1)
          if (chunk_elems * esize > (uint64_t)buf.size() * 1040 + 4096) bad("dataset '" + d.name + "': a deflated chunk does not inflate");
          std::vector<uint8_t> o(chunk_elems * esize + 4);
          uLongf ol = (uLongf)o.size();
          if (uncompress(o.data(), &ol, buf.data(), (uLong)buf.size()) != Z_OK) bad("dataset '" + d.name + "': a deflated chunk does not inflate");
          o.resize(ol); buf.swap(o);
        } else if (fl.id == 2) {
          unshuffle(buf, fl.cd.empty() ? esize : fl.cd[0]);
        } else if (fl.id == 3) {               // fletcher32: checksum appended
          if (buf.size() < 4) bad("fletcher32 chunk too short");
          buf.resize(buf.size() - 4);
        } else {
          unsupported("filter " + std::to_string(fl.id) + " on dataset '" + d.name + "'");
        }
      }
      data = buf.data();
      if (buf.size() < chunk_elems * esize) bad("dataset '" + d.name + "': chunk shorter than its dimensions");
    } else if (nbytes < chunk_elems * esize) {
      bad("dataset '" + d.name + "': chunk shorter than its dimensions");
    }
    for (size_t q = 0; q < rank; q++) if (off[q] >= d.dims[q]) bad("chunk offset outside the dataset");
    // copy the part of the chunk that lies inside the dataset, one innermost run at a time
    if (rank == 0) { convert(k, esize, data, 1, out); continue; }
    std::vector<uint64_t> idx(rank, 0), ext(rank);
    for (size_t q = 0; q < rank; q++) ext[q] = std::min<uint64_t>(d.chunk[q], d.dims[q] - off[q]);
    const uint64_t run = ext[rank - 1];
    while (true) {
      uint64_t src = 0, dst = 0;
      for (size_t q = 0; q < rank; q++) { src = src * d.chunk[q] + idx[q]; dst = dst * d.dims[q] + off[q] + idx[q]; }
      convert(k, esize, data + src * esize, run, out + 2 * dst);
      int q = (int)rank - 2;
      while (q >= 0 && ++idx[q] == ext[q]) idx[q--] = 0;
      if (q < 0) break;
    }
  }
}

static void read_elements(const File& f, const Dataset& d, double* out) {
  const ElemKind k = classify(d);
  const uint64_t n = d.elems();
  const uint32_t esize = d.type.size;
  if (n == 0) return;
  if (esize == 0 || n > (UNDEF >> 1) / esize) bad("dataset '" + d.name + "': size overflows");
  if (d.layout == 0) {
    if (d.bytes < n * esize) bad("dataset '" + d.name + "': compact data shorter than its dataspace");
    convert(k, esize, d.compact, n, out);
  } else if (d.layout == 1) {
    if (d.addr == UNDEF) { std::memset(out, 0, (size_t)n * 16); return; }   // never written: fill value 0
    if (d.bytes != UNDEF && d.bytes < n * esize) bad("dataset '" + d.name + "': contiguous block shorter than its dataspace");
    convert(k, esize, f.at(d.addr, n * esize), n, out);
  } else {
    if (d.chunk.size() != d.dims.size() + 1) bad("dataset '" + d.name + "': chunk rank differs from the dataspace");
    for (size_t q = 0; q < d.dims.size(); q++) if (d.chunk[q] == 0) bad("zero chunk dimension");
    std::memset(out, 0, (size_t)n * 16);
    if (d.chunk_btree != UNDEF) read_chunks(f, d, k, d.chunk_btree, out, 0);
  }
}

static std::unique_ptr<File> open_file(const char* path, const char* group) {
  std::unique_ptr<File> f(new File);
  f->fd = ::open(path, O_RDONLY | O_CLOEXEC);
  if (f->fd < 0) bad(std::string("cannot open '") + path + "': " + std::strerror(errno));
  struct stat st;
  if (fstat(f->fd, &st) != 0 || st.st_size < 16) bad(std::string("'") + path + "' is too short to be an HDF5 file");
  f->len = (size_t)st.st_size;
  void* m = mmap(nullptr, f->len, PROT_READ, MAP_PRIVATE, f->fd, 0);
  if (m == MAP_FAILED) { f->map = nullptr; bad(std::string("cannot map '") + path + "'"); }
  f->map = (const uint8_t*)m;
  read_superblock(*f);
  uint64_t g = find_group(*f, group ? group : "/tensors");
  for (const Link& l : list_group(*f, g)) f->tensors.push_back(read_dataset(*f, l.name, l.header));
  return f;
}

// ======================================= writer =======================================
struct Out {
  std::vector<uint8_t> b;
  void u8(unsigned v) { b.push_back((uint8_t)v); }
  void uN(uint64_t v, int n) { for (int i = 0; i < n; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void u16(unsigned v) { uN(v, 2); }
  void u32(uint64_t v) { uN(v, 4); }
  void u64(uint64_t v) { uN(v, 8); }
  void raw(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void pad_to8() { while (b.size() & 7) b.push_back(0); }
  size_t size() const { return b.size(); }
};

struct WTensor {
  std::string name;
  std::vector<uint64_t> dims;
  const double* data = nullptr;             // null: dataset declared but never written (the "-1" output tensor)
  bool has_bids = false;
  std::vector<uint64_t> bids;
};

static void put_message(Out& o, int type, int flags, const Out& body) {
  size_t sz = pad8(body.size());
  if (sz > 65528) throw IoError(TNCB_ERR_UNSUPPORTED, "HDF5 writer: header message above 64 KiB (attribute too long)");
  o.u16(type); o.u16((unsigned)sz); o.u8(flags); o.zeros(3);
  o.raw(body.b.data(), body.size());
  o.zeros(sz - body.size());
}

static void put_f64_type(Out& o) {           // IEEE binary64, little endian
  o.u8(0x11); o.u8(0x20); o.u8(0x3f); o.u8(0x00); o.u32(8);
  o.u16(0); o.u16(64); o.u8(52); o.u8(11); o.u8(0); o.u8(52); o.u32(1023);
}

static void put_complex_type(Out& o) {       // compound {re @0, im @8}, version 1 (libhdf5's default encoding)
  o.u8(0x16); o.u8(2); o.u8(0); o.u8(0); o.u32(16);
  const char* names[2] = {"re", "im"};
  for (int i = 0; i < 2; i++) {
    o.raw(names[i], 3); o.zeros(5);
    o.u32(8 * i);
    o.u8(0); o.zeros(3); o.u32(0); o.u32(0); o.zeros(16);
    put_f64_type(o);
  }
}

static Out dataset_header(const WTensor& t, uint64_t data_addr) {
  Out msgs;
  int n = 0;
  { Out m; m.u8(1); m.u8((unsigned)t.dims.size()); m.u8(0); m.zeros(5); for (uint64_t d : t.dims) m.u64(d); put_message(msgs, 0x01, 0, m); n++; }
  { Out m; put_complex_type(m); put_message(msgs, 0x03, 1, m); n++; }
  { Out m; m.u8(2); m.u8(2); m.u8(2); m.u8(1); m.u32(0); put_message(msgs, 0x05, 1, m); n++; }     // fill value: default, allocate late
  { uint64_t e = 1; for (uint64_t d : t.dims) e *= d;
    Out m; m.u8(3); m.u8(1); m.u64(data_addr); m.u64(e * 16); put_message(msgs, 0x08, 0, m); n++; }
  if (t.has_bids) {
    Out m;
    m.u8(1); m.u8(0); m.u16(5); m.u16(12); m.u16(16);
    m.raw("bids", 5); m.zeros(3);
    m.u8(0x10); m.u8(0); m.u8(0); m.u8(0); m.u32(8); m.u16(0); m.u16(64); m.zeros(4);          // u64 little endian (padded to 16)
    m.u8(1); m.u8(1); m.u8(0); m.zeros(5); m.u64(t.bids.size());                                  // 1-d dataspace
    for (uint64_t v : t.bids) m.u64(v);
    put_message(msgs, 0x0c, 0, m); n++;
  }
  Out h;
  h.u8(1); h.u8(0); h.u16((unsigned)n); h.u32(1); h.u32(msgs.size()); h.zeros(4);
  h.raw(msgs.b.data(), msgs.size());
  return h;
}

static Out group_header(uint64_t btree, uint64_t heap) {
  Out m; m.u64(btree); m.u64(heap);
  Out msgs; put_message(msgs, 0x11, 0, m);
  Out h; h.u8(1); h.u8(0); h.u16(1); h.u32(1); h.u32(msgs.size()); h.zeros(4);
  h.raw(msgs.b.data(), msgs.size());
  return h;
}

constexpr int kLeafK = 4, kInternalK = 16;
constexpr size_t kSnodBytes = 8 + 2 * kLeafK * 40, kTreeBytes = 24 + (2 * kInternalK + 1) * 8 + 2 * kInternalK * 8;

struct Symbol { uint64_t name_off, header; unsigned cache; uint64_t btree, heap; };

// Symbol-table group: local heap + SNODs + B-tree over them.  `symbols` are in ascending name order.  Everything is
// appended to `o` (whose start is file address `origin`); returns the addresses of the B-tree root and the heap.
static void put_group_index(Out& o, uint64_t origin, const std::vector<std::string>& names, std::vector<Symbol> symbols,
                            uint64_t* btree_root, uint64_t* heap_addr) {
  // local heap: offset 0 holds the empty string
  Out seg; seg.zeros(8);
  for (size_t i = 0; i < names.size(); i++) { symbols[i].name_off = seg.size(); seg.raw(names[i].c_str(), names[i].size() + 1); seg.pad_to8(); }
  *heap_addr = origin + o.size();
  o.raw("HEAP", 4); o.zeros(4); o.u64(seg.size()); o.u64(1 /* H5HL_FREE_NULL: no free block */); o.u64(origin + o.size() + 8);
  o.raw(seg.b.data(), seg.size());
  // leaves
  struct Node { uint64_t addr; uint64_t last_name; };
  std::vector<Node> level;
  for (size_t i = 0; i < symbols.size(); i += 2 * kLeafK) {
    size_t n = std::min<size_t>(2 * kLeafK, symbols.size() - i);
    level.push_back({origin + o.size(), symbols[i + n - 1].name_off});
    o.raw("SNOD", 4); o.u8(1); o.u8(0); o.u16((unsigned)n);
    for (size_t k = 0; k < n; k++) {
      const Symbol& s = symbols[i + k];
      o.u64(s.name_off); o.u64(s.header); o.u32(s.cache); o.u32(0);
      if (s.cache == 1) { o.u64(s.btree); o.u64(s.heap); } else o.zeros(16);
    }
    o.zeros((2 * kLeafK - n) * 40);
  }
  // B-tree levels, bottom up
  int lvl = 0;
  while (true) {
    std::vector<Node> up;
    size_t n_nodes = std::max<size_t>(1, (level.size() + 2 * kInternalK - 1) / (2 * kInternalK));
    uint64_t first = origin + o.size();
    for (size_t q = 0; q < n_nodes; q++) {
      size_t i = q * 2 * kInternalK;
      size_t n = level.empty() ? 0 : std::min<size_t>(2 * kInternalK, level.size() - i);
      up.push_back({origin + o.size(), n ? level[i + n - 1].last_name : 0});
      o.raw("TREE", 4); o.u8(0); o.u8((unsigned)lvl); o.u16((unsigned)n);
      o.u64(q == 0 ? UNDEF : first + (q - 1) * kTreeBytes);
      o.u64(q + 1 == n_nodes ? UNDEF : first + (q + 1) * kTreeBytes);
      o.u64(i == 0 ? 0 : level[i - 1].last_name);                   // key 0: everything to the left
      for (size_t k = 0; k < n; k++) { o.u64(level[i + k].addr); o.u64(level[i + k].last_name); }
      o.zeros((2 * kInternalK - n) * 16);
    }
    level.swap(up);
    lvl++;
    if (level.size() == 1) break;
  }
  *btree_root = level[0].addr;
}

static void write_file(const char* path, std::vector<WTensor> tensors) {
  std::sort(tensors.begin(), tensors.end(), [](const WTensor& a, const WTensor& b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
  for (size_t i = 0; i < tensors.size(); i++) {
    if (tensors[i].name.empty() || tensors[i].name.find('/') != std::string::npos) throw IoError(TNCB_ERR_INVALID, "HDF5 writer: bad dataset name");
    if (i && tensors[i].name == tensors[i - 1].name) throw IoError(TNCB_ERR_INVALID, "HDF5 writer: duplicate dataset name '" + tensors[i].name + "'");
    if (tensors[i].dims.size() > 32) throw IoError(TNCB_ERR_INVALID, "HDF5 writer: rank above 32");
  }
  // 1. dataset headers need the data addresses, the group index needs the header addresses: lay the metadata out twice
  //    (sizes do not depend on the addresses).
  const uint64_t superblock = 96;
  std::vector<uint64_t> header_addr(tensors.size()), data_addr(tensors.size(), UNDEF);
  Out meta;
  uint64_t data_start = 0;
  for (int pass = 0; pass < 2; pass++) {
    meta = Out();
    // root group: header, then heap + SNOD + B-tree with the single member "tensors"
    const uint64_t root_header = superblock;
    Out rh = group_header(0, 0);              // patched below
    const uint64_t root_index = root_header + rh.size();
    Out tensors_index;                        // index of /tensors, placed after the root's
    Out root_idx;
    // sizes are deterministic: build the /tensors index first at a provisional origin to learn the size of the root index
    std::vector<std::string> names;
    std::vector<Symbol> syms;
    for (size_t i = 0; i < tensors.size(); i++) { names.push_back(tensors[i].name); syms.push_back({0, header_addr[i], 0, 0, 0}); }
    // root index size: heap 32 + 16, one SNOD, one TREE node
    const uint64_t root_index_size = 32 + 16 + kSnodBytes + kTreeBytes;
    const uint64_t tensors_header = root_index + root_index_size;
    Out th = group_header(0, 0);
    const uint64_t tensors_index_origin = tensors_header + th.size();
    uint64_t t_btree, t_heap;
    put_group_index(tensors_index, tensors_index_origin, names, syms, &t_btree, &t_heap);
    th = group_header(t_btree, t_heap);
    uint64_t r_btree, r_heap;
    put_group_index(root_idx, root_index, {"tensors"}, {Symbol{0, tensors_header, 1, t_btree, t_heap}}, &r_btree, &r_heap);
    if (root_idx.size() != root_index_size) throw IoError(TNCB_ERR_INVALID, "HDF5 writer: internal layout error");
    rh = group_header(r_btree, r_heap);
    // superblock
    meta.raw(kSignature, 8);
    meta.u8(0); meta.u8(0); meta.u8(0); meta.u8(0); meta.u8(0); meta.u8(8); meta.u8(8); meta.u8(0);
    meta.u16(kLeafK); meta.u16(kInternalK); meta.u32(0);
    meta.u64(0); meta.u64(UNDEF);
    const size_t eof_pos = meta.size();
    meta.u64(0); meta.u64(UNDEF);
    meta.u64(0); meta.u64(root_header); meta.u32(1); meta.u32(0); meta.u64(r_btree); meta.u64(r_heap);
    if (meta.size() != superblock) throw IoError(TNCB_ERR_INVALID, "HDF5 writer: internal layout error");
    meta.raw(rh.b.data(), rh.size());
    meta.raw(root_idx.b.data(), root_idx.size());
    meta.raw(th.b.data(), th.size());
    meta.raw(tensors_index.b.data(), tensors_index.size());
    for (size_t i = 0; i < tensors.size(); i++) {
      header_addr[i] = meta.size();
      Out dh = dataset_header(tensors[i], data_addr[i]);
      meta.raw(dh.b.data(), dh.size());
    }
    meta.pad_to8();
    data_start = meta.size();
    uint64_t at = data_start;
    for (size_t i = 0; i < tensors.size(); i++) {
      uint64_t e = 1; for (uint64_t d : tensors[i].dims) e *= d;
      if (tensors[i].data && e) { data_addr[i] = at; at += e * 16; } else data_addr[i] = UNDEF;
    }
    for (int q = 0; q < 8; q++) meta.b[eof_pos + q] = (uint8_t)(at >> (8 * q));
  }
  FILE* fp = std::fopen(path, "wb");
  if (!fp) bad(std::string("cannot create '") + path + "': " + std::strerror(errno));
  bool ok = std::fwrite(meta.b.data(), 1, meta.size(), fp) == meta.size();
  for (size_t i = 0; ok && i < tensors.size(); i++) {
    if (data_addr[i] == UNDEF) continue;
    uint64_t e = 1; for (uint64_t d : tensors[i].dims) e *= d;
    ok = std::fwrite(tensors[i].data, 16, (size_t)e, fp) == (size_t)e;
  }
  ok = (std::fclose(fp) == 0) && ok;
  if (!ok) bad(std::string("short write to '") + path + "'");
}

// load_data (hdf5.rs:37-43,90-103) + matrix_adjoint_inplace (gates.rs:82-99) for a TensorData::File leaf: the first member
// of /tensors, checked against the bond dimensions the network gives the leaf.
int load_file_leaf(const char* path, bool adjoint, int rank, const uint64_t* dims, double* out_re_im) {
  try {
    std::unique_ptr<File> f = open_file(path, nullptr);
    if (f->tensors.empty()) return fail(TNCB_ERR_IO, std::string("HDF5: '") + path + "' has no member in /tensors");
    const Dataset& d = f->tensors[0];
    std::vector<uint64_t> shape = d.dims;
    const size_t r = shape.size();
    if (adjoint && r > 0) {
      if (r & (r - 1)) return fail(TNCB_ERR_SHAPE, "adjoint of a File tensor needs a power-of-two rank (gates.rs:84)");
      std::rotate(shape.begin(), shape.begin() + r / 2, shape.end());
    }
    if ((int)r != rank) return fail(TNCB_ERR_SHAPE, std::string("File leaf '") + path + "': rank differs from the leaf's bond dimensions");
    for (size_t i = 0; i < r; i++)
      if (shape[i] != dims[i]) return fail(TNCB_ERR_SHAPE, std::string("File leaf '") + path + "': shape differs from the leaf's bond dimensions");
    const uint64_t n = d.elems();
    if (!adjoint || r == 0) {
      read_elements(*f, d, out_re_im);
      if (adjoint) for (uint64_t i = 0; i < n; i++) out_re_im[2 * i + 1] = -out_re_im[2 * i + 1];
      return TNCB_OK;
    }
    std::vector<double> tmp((size_t)n * 2);
    read_elements(*f, d, tmp.data());
    uint64_t rows = 1, cols = 1;
    for (size_t i = 0; i < r / 2; i++) rows *= d.dims[i];
    for (size_t i = r / 2; i < r; i++) cols *= d.dims[i];
    for (uint64_t i = 0; i < rows; i++)
      for (uint64_t j = 0; j < cols; j++) {
        out_re_im[2 * (j * rows + i)] = tmp[2 * (i * cols + j)];
        out_re_im[2 * (j * rows + i) + 1] = -tmp[2 * (i * cols + j) + 1];
      }
    return TNCB_OK;
  } catch (const IoError& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(TNCB_ERR_OOM, "HDF5: out of host memory");
  }
}

} // namespace h5
} // namespace tncb

struct tncb_h5file { std::unique_ptr<tncb::h5::File> f; };

#define TNCB_H5_GUARD(...)                                                       \
  try { __VA_ARGS__ } catch (const tncb::h5::IoError& e) { return tncb::fail(e.code, e.what()); } \
  catch (const std::bad_alloc&) { return tncb::fail(TNCB_ERR_OOM, "HDF5: out of host memory"); } \
  catch (const std::exception& e) { return tncb::fail(TNCB_ERR_IO, std::string("HDF5: ") + e.what()); }

extern "C" {

int tncb_hdf5_open(const char* path, const char* group, tncb_h5file** out) {
  if (!path || !out) return tncb::fail(TNCB_ERR_INVALID, "null argument");
  *out = nullptr;
  TNCB_H5_GUARD(
    std::unique_ptr<tncb_h5file> h(new tncb_h5file);
    h->f = tncb::h5::open_file(path, group);
    *out = h.release();
    return TNCB_OK;
  )
}

void tncb_hdf5_close(tncb_h5file* file) { delete file; }

size_t tncb_hdf5_count(const tncb_h5file* file) { return file ? file->f->tensors.size() : 0; }

const char* tncb_hdf5_name(const tncb_h5file* file, size_t i) {
  if (!file || i >= file->f->tensors.size()) return nullptr;
  return file->f->tensors[i].name.c_str();
}

int tncb_hdf5_shape(const tncb_h5file* file, size_t i, int* rank, uint64_t* dims, uint64_t* elems) {
  if (!file || i >= file->f->tensors.size()) return tncb::fail(TNCB_ERR_INVALID, "tensor index out of range");
  const tncb::h5::Dataset& d = file->f->tensors[i];
  if (rank) *rank = (int)d.dims.size();
  if (dims) for (size_t q = 0; q < d.dims.size(); q++) dims[q] = d.dims[q];
  if (elems) *elems = d.elems();
  return TNCB_OK;
}

int tncb_hdf5_attr(const tncb_h5file* file, size_t i, const char* name, size_t cap, int64_t* out, size_t* n) {
  if (!file || i >= file->f->tensors.size() || !name || !n) return tncb::fail(TNCB_ERR_INVALID, "bad argument");
  for (const tncb::h5::Attribute& a : file->f->tensors[i].attrs) {
    if (a.name != name) continue;
    const tncb::h5::Datatype& t = a.type;
    if (t.cls != 0) return tncb::fail(TNCB_ERR_INVALID, std::string("HDF5: attribute '") + name + "' is not an integer array");
    size_t cnt = t.size ? a.bytes / t.size : 0;
    *n = cnt;
    if (!out) return TNCB_OK;
    if (cnt > cap) return tncb::fail(TNCB_ERR_INVALID, "attribute buffer too small");
    for (size_t q = 0; q < cnt; q++) {
      const uint8_t* p = a.data + q * t.size;
      uint64_t v = 0;
      for (uint32_t b = 0; b < t.size; b++) v |= (uint64_t)p[t.little ? b : t.size - 1 - b] << (8 * b);
      if (t.is_signed && t.size < 8 && (v >> (8 * t.size - 1))) v |= ~0ull << (8 * t.size);
      if (!t.is_signed && t.size == 8 && (v >> 63)) return tncb::fail(TNCB_ERR_INVALID, "attribute value above 2^63");
      out[q] = (int64_t)v;
    }
    return TNCB_OK;
  }
  return tncb::fail(TNCB_ERR_INVALID, std::string("HDF5: dataset '") + file->f->tensors[i].name + "' has no attribute '" + name + "'");
}

int tncb_hdf5_read(const tncb_h5file* file, size_t i, double* out_re_im) {
  if (!file || i >= file->f->tensors.size() || !out_re_im) return tncb::fail(TNCB_ERR_INVALID, "bad argument");
  TNCB_H5_GUARD(
    tncb::h5::read_elements(*file->f, file->f->tensors[i], out_re_im);
    return TNCB_OK;
  )
}

int tncb_hdf5_load_leaf(const char* path, int adjoint, int rank, const uint64_t* dims, double* out_re_im) {
  if (!path || !out_re_im || rank < 0 || (rank > 0 && !dims)) return tncb::fail(TNCB_ERR_INVALID, "bad argument");
  return tncb::h5::load_file_leaf(path, adjoint != 0, rank, dims, out_re_im);
}

int tncb_hdf5_store(const char* path, size_t n, const char* const* names, const int* ranks, const uint64_t* const* dims,
                    const double* const* data_re_im, const int64_t* n_bids, const uint64_t* const* bids) {
  if (!path || (n && (!names || !ranks || !dims || !data_re_im))) return tncb::fail(TNCB_ERR_INVALID, "null argument");
  TNCB_H5_GUARD(
    std::vector<tncb::h5::WTensor> ts(n);
    for (size_t i = 0; i < n; i++) {
      if (!names[i] || ranks[i] < 0 || (ranks[i] > 0 && !dims[i])) return tncb::fail(TNCB_ERR_INVALID, "bad tensor description");
      ts[i].name = names[i];
      ts[i].dims.assign(dims[i], dims[i] + ranks[i]);
      ts[i].data = data_re_im[i];
      if (n_bids && n_bids[i] >= 0) {
        if (n_bids[i] > 0 && (!bids || !bids[i])) return tncb::fail(TNCB_ERR_INVALID, "bids missing");
        ts[i].has_bids = true;
        if (n_bids[i] > 0) ts[i].bids.assign(bids[i], bids[i] + n_bids[i]);
      }
    }
    tncb::h5::write_file(path, std::move(ts));
    return TNCB_OK;
  )
}

int tncb_hdf5_store_data(const char* path, int rank, const uint64_t* dims, const double* data_re_im) {
  const char* name = "-1";                                           // write_data: the single dataset is called "-1" (hdf5.rs:111)
  if (!data_re_im) return tncb::fail(TNCB_ERR_INVALID, "null argument");
  return tncb_hdf5_store(path, 1, &name, &rank, &dims, &data_re_im, nullptr, nullptr);
}

} // extern "C"
