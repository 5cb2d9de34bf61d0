// git_store.hpp - read-only access to a git object store for `tosem-scan history` (S8 on
// real revision history).  Host work only: the bytes of the (old, new) blobs it hands out are diffed on the GPU through
// tsm_diff_pairs_detail.  Reads what `git` itself writes - loose objects (zlib), version-2 packfiles with their
// version-2 .idx (OFS_DELTA and REF_DELTA chains), HEAD, refs/ and packed-refs - from the published on-disk formats
// (Documentation/gitformat-pack.txt, gitformat-index is not needed); no alternates, no multi-pack-index, no SHA-256
// repositories.  Nothing is verified cryptographically: objects are looked up by the name the repository gives them.
// The reference package ships no repository; the tests pin this reader against `git` itself.
#pragma once
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace gitstore {

struct Oid {
  uint8_t b[20];
  bool operator<(const Oid& o) const { return memcmp(b, o.b, 20) < 0; }
  bool operator==(const Oid& o) const { return memcmp(b, o.b, 20) == 0; }
  bool operator!=(const Oid& o) const { return !(*this == o); }
  std::string hex() const {
    static const char* d = "0123456789abcdef";
    std::string s(40, '0');
    for (int i = 0; i < 20; ++i) { s[2 * i] = d[b[i] >> 4]; s[2 * i + 1] = d[b[i] & 15]; }
    return s;
  }
  static bool from_hex(const char* h, size_t n, Oid& out) {
    if (n < 40) return false;
    for (int i = 0; i < 20; ++i) {
      int v = 0;
      for (int k = 0; k < 2; ++k) {
        const char c = h[2 * i + k];
        int x;
        if (c >= '0' && c <= '9') x = c - '0'; else if (c >= 'a' && c <= 'f') x = c - 'a' + 10; else if (c >= 'A' && c <= 'F') x = c - 'A' + 10; else return false;
        v = v * 16 + x;
      }
      out.b[i] = (uint8_t)v;
    }
    return true;
  }
};

enum { OBJ_COMMIT = 1, OBJ_TREE = 2, OBJ_BLOB = 3, OBJ_TAG = 4, OBJ_OFS_DELTA = 6, OBJ_REF_DELTA = 7 };

struct Object { int type = 0; std::vector<uint8_t> data; };

struct TreeEntry { uint32_t mode; std::string name; Oid oid; bool is_tree() const { return mode == 040000; } bool is_blob() const { return (mode & 0170000) == 0100000; } };

struct Commit { Oid tree; std::vector<Oid> parents; long long time = 0; std::string subject; };

static inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

// zlib stream at `src` (at most `avail` bytes readable) into exactly `want` bytes.
static inline bool inflate_exact(const uint8_t* src, size_t avail, size_t want, std::vector<uint8_t>& out) {
  out.resize(want);
  z_stream z;
  memset(&z, 0, sizeof z);
  if (inflateInit(&z) != Z_OK) return false;
  z.next_in = const_cast<Bytef*>(src);
  z.avail_in = (uInt)std::min<size_t>(avail, 0x7fffffffu);
  uint8_t dummy;
  z.next_out = want ? out.data() : &dummy;
  z.avail_out = want ? (uInt)want : 1u;
  int rc = Z_OK;
  while (rc == Z_OK) rc = inflate(&z, Z_FINISH);
  const bool ok = (rc == Z_STREAM_END) && z.total_out == want;
  inflateEnd(&z);
  return ok;
}

class Store {
 public:
  ~Store() { for (Pack& p : packs_) { if (p.map) munmap(const_cast<uint8_t*>(p.map), p.size); if (p.idx) munmap(const_cast<uint8_t*>(p.idx), p.idx_size); } }

  // `path`: a work tree (with .git, directory or `gitdir:` file) or a bare repository.
  bool open(const std::string& path, std::string& err) {
    struct stat st;
    std::string g = path + "/.git";
    if (stat(g.c_str(), &st) == 0) {
      if (S_ISREG(st.st_mode)) {                            // "gitdir: <path>" (worktrees, submodules)
        std::ifstream is(g);
        std::string line;
        std::getline(is, line);
        if (line.rfind("gitdir: ", 0) != 0) { err = "unreadable .git file"; return false; }
        g = line.substr(8);
        if (!g.empty() && g[0] != '/') g = path + "/" + g;
      }
    } else g = path;
    if (stat((g + "/objects").c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) { err = "no git object store under " + path; return false; }
    dir_ = g;
    const std::string pd = g + "/objects/pack";
    if (DIR* d = opendir(pd.c_str())) {
      std::vector<std::string> names;
      while (dirent* e = readdir(d)) { const std::string n = e->d_name; if (n.size() > 4 && n.substr(n.size() - 4) == ".idx") names.push_back(n); }
      closedir(d);
      std::sort(names.begin(), names.end());
      for (const std::string& n : names) {
        Pack p;
        if (!map_file(pd + "/" + n, p.idx, p.idx_size) || !map_file(pd + "/" + n.substr(0, n.size() - 4) + ".pack", p.map, p.size)) { err = "cannot map " + n; return false; }
        if (p.idx_size < 8 + 256 * 4 || memcmp(p.idx, "\377tOc", 4) != 0 || be32(p.idx + 4) != 2) { err = "pack index is not version 2: " + n; return false; }
        p.n = be32(p.idx + 8 + 255 * 4);
        if (p.idx_size < 8 + 1024 + (size_t)p.n * 28 + 40 || p.size < 12 || memcmp(p.map, "PACK", 4) != 0 || be32(p.map + 4) != 2) { err = "pack is not version 2: " + n; return false; }
        packs_.push_back(p);
      }
    }
    load_packed_refs();
    return true;
  }

  bool read(const Oid& id, Object& out, int depth = 0) {
    if (depth > 64) return false;
    for (const Pack& p : packs_) {
      uint64_t off;
      if (find_in_pack(p, id, off)) return read_packed(p, off, out, depth);
    }
    return read_loose(id, out);
  }

  // "HEAD", a full ref name, a branch or tag name, or 40 hex digits; tags are peeled to the commit.
  bool resolve(const std::string& rev, Oid& out) {
    Oid id;
    bool ok = false;
    if (rev.size() == 40 && Oid::from_hex(rev.data(), 40, id)) ok = true;
    for (const char* prefix : {"", "refs/", "refs/tags/", "refs/heads/", "refs/remotes/"}) {
      if (ok) break;
      ok = read_ref(std::string(prefix) + rev, id, 0);
    }
    if (!ok) return false;
    for (int i = 0; i < 8; ++i) {                           // peel annotated tags
      Object o;
      if (!read(id, o)) return false;
      if (o.type != OBJ_TAG) break;
      const std::string s((const char*)o.data.data(), o.data.size());
      if (s.rfind("object ", 0) != 0 || !Oid::from_hex(s.data() + 7, s.size() - 7, id)) return false;
    }
    out = id;
    return true;
  }

  std::vector<std::string> tag_names() {
    std::vector<std::string> names;
    list_dir_refs("refs/tags", names);
    for (auto& kv : packed_) if (kv.first.rfind("refs/tags/", 0) == 0) names.push_back(kv.first.substr(10));
    std::sort(names.begin(), names.end());
    names.erase(std::unique(names.begin(), names.end()), names.end());
    return names;
  }

  bool commit(const Oid& id, Commit& c) {
    Object o;
    if (!read(id, o) || o.type != OBJ_COMMIT) return false;
    const std::string s((const char*)o.data.data(), o.data.size());
    size_t pos = 0;
    c = Commit{};
    bool have_tree = false;
    while (pos < s.size()) {
      const size_t e = std::min(s.find('\n', pos), s.size());
      if (e == pos) { pos = e + 1; break; }                  // blank line: the message follows
      const std::string line = s.substr(pos, e - pos);
      if (line.rfind("tree ", 0) == 0) have_tree = Oid::from_hex(line.data() + 5, line.size() - 5, c.tree);
      else if (line.rfind("parent ", 0) == 0) { Oid p; if (Oid::from_hex(line.data() + 7, line.size() - 7, p)) c.parents.push_back(p); }
      else if (line.rfind("committer ", 0) == 0) {
        const size_t gt = line.rfind('>');
        if (gt != std::string::npos) c.time = atoll(line.c_str() + gt + 1);
      }
      pos = e + 1;
    }
    if (pos < s.size()) c.subject = s.substr(pos, std::min(s.find('\n', pos), s.size()) - pos);
    return have_tree;
  }

  bool tree(const Oid& id, std::vector<TreeEntry>& out) {
    Object o;
    out.clear();
    if (!read(id, o) || o.type != OBJ_TREE) return false;
    const uint8_t* p = o.data.data();
    const uint8_t* end = p + o.data.size();
    while (p < end) {
      TreeEntry e;
      e.mode = 0;
      while (p < end && *p != ' ') { if (*p < '0' || *p > '7') return false; e.mode = e.mode * 8 + (uint32_t)(*p - '0'); ++p; }
      if (p >= end) return false;
      ++p;
      const uint8_t* z = (const uint8_t*)memchr(p, 0, (size_t)(end - p));
      if (!z || end - z < 21) return false;
      e.name.assign((const char*)p, (size_t)(z - p));
      memcpy(e.oid.b, z + 1, 20);
      p = z + 21;
      out.push_back(std::move(e));
    }
    return true;
  }

 private:
  struct Pack { const uint8_t* map = nullptr; size_t size = 0; const uint8_t* idx = nullptr; size_t idx_size = 0; uint32_t n = 0; };
  std::string dir_;
  std::vector<Pack> packs_;
  std::map<std::string, Oid> packed_;

  static bool map_file(const std::string& path, const uint8_t*& p, size_t& n) {
    const int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { close(fd); return false; }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return false;
    p = (const uint8_t*)m; n = (size_t)st.st_size;
    return true;
  }

  static bool find_in_pack(const Pack& p, const Oid& id, uint64_t& off) {
    const uint8_t* fan = p.idx + 8;
    uint32_t lo = id.b[0] ? be32(fan + 4 * (id.b[0] - 1)) : 0u, hi = be32(fan + 4 * id.b[0]);
    const uint8_t* names = fan + 1024;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo) / 2;
      const int c = memcmp(names + 20 * (size_t)mid, id.b, 20);
      if (c == 0) {
        const uint8_t* offs = names + 20 * (size_t)p.n + 4 * (size_t)p.n;      // behind the CRC table
        const uint32_t o = be32(offs + 4 * (size_t)mid);
        if (o & 0x80000000u) {
          const uint8_t* big = offs + 4 * (size_t)p.n + 8 * (size_t)(o & 0x7fffffffu);
          if (big + 8 > p.idx + p.idx_size) return false;
          off = ((uint64_t)be32(big) << 32) | be32(big + 4);
        } else off = o;
        return off < p.size;
      }
      if (c < 0) lo = mid + 1; else hi = mid;
    }
    return false;
  }

  bool read_packed(const Pack& p, uint64_t off, Object& out, int depth) {
    if (depth > 64 || off >= p.size) return false;
    const uint8_t* q = p.map + off;
    const uint8_t* end = p.map + p.size;
    uint8_t c = *q++;
    const int type = (c >> 4) & 7;
    uint64_t size = c & 15;
    int shift = 4;
    while (c & 0x80) { if (q >= end) return false; c = *q++; size |= (uint64_t)(c & 0x7f) << shift; shift += 7; }
    if (type == OBJ_COMMIT || type == OBJ_TREE || type == OBJ_BLOB || type == OBJ_TAG) {
      out.type = type;
      return inflate_exact(q, (size_t)(end - q), (size_t)size, out.data);
    }
    Object base;
    if (type == OBJ_OFS_DELTA) {
      if (q >= end) return false;
      c = *q++;
      uint64_t back = c & 0x7f;
      while (c & 0x80) { if (q >= end) return false; c = *q++; back = ((back + 1) << 7) | (c & 0x7f); }
      if (back == 0 || back > off) return false;
      if (!read_packed(p, off - back, base, depth + 1)) return false;
    } else if (type == OBJ_REF_DELTA) {
      if (end - q < 20) return false;
      Oid b;
      memcpy(b.b, q, 20);
      q += 20;
      if (!read(b, base, depth + 1)) return false;
    } else return false;
    std::vector<uint8_t> delta;
    if (!inflate_exact(q, (size_t)(end - q), (size_t)size, delta)) return false;
    out.type = base.type;
    return apply_delta(base.data, delta, out.data);
  }

  static bool apply_delta(const std::vector<uint8_t>& src, const std::vector<uint8_t>& d, std::vector<uint8_t>& out) {
    size_t i = 0;
    auto varint = [&](uint64_t& v) { v = 0; int sh = 0; uint8_t c; do { if (i >= d.size()) return false; c = d[i++]; v |= (uint64_t)(c & 0x7f) << sh; sh += 7; } while (c & 0x80); return true; };
    uint64_t ssz, dsz;
    if (!varint(ssz) || !varint(dsz) || ssz != src.size()) return false;
    out.clear();
    out.reserve((size_t)dsz);
    while (i < d.size()) {
      const uint8_t op = d[i++];
      if (op & 0x80) {
        uint32_t off = 0, len = 0;
        for (int k = 0; k < 4; ++k) if (op & (1 << k)) { if (i >= d.size()) return false; off |= (uint32_t)d[i++] << (8 * k); }
        for (int k = 0; k < 3; ++k) if (op & (0x10 << k)) { if (i >= d.size()) return false; len |= (uint32_t)d[i++] << (8 * k); }
        if (len == 0) len = 0x10000;
        if ((uint64_t)off + len > src.size()) return false;
        out.insert(out.end(), src.begin() + off, src.begin() + off + len);
      } else if (op) {
        if (i + op > d.size()) return false;
        out.insert(out.end(), d.begin() + (long)i, d.begin() + (long)(i + op));
        i += op;
      } else return false;
    }
    return out.size() == dsz;
  }

  bool read_loose(const Oid& id, Object& out) {
    const std::string h = id.hex();
    const uint8_t* m;
    size_t n;
    if (!map_file(dir_ + "/objects/" + h.substr(0, 2) + "/" + h.substr(2), m, n)) return false;
    // header "type size\0" first: inflate a little, then the whole object
    bool ok = false;
    z_stream z;
    memset(&z, 0, sizeof z);
    if (inflateInit(&z) == Z_OK) {
      uint8_t head[64];
      z.next_in = const_cast<Bytef*>(m); z.avail_in = (uInt)n;
      z.next_out = head; z.avail_out = sizeof head;
      const int rc = inflate(&z, Z_SYNC_FLUSH);
      const size_t got = sizeof head - z.avail_out;
      const uint8_t* nul = (const uint8_t*)memchr(head, 0, got);
      if ((rc == Z_OK || rc == Z_STREAM_END) && nul) {
        const std::string hd((const char*)head, (size_t)(nul - head));
        const size_t sp = hd.find(' ');
        if (sp != std::string::npos) {
          const std::string t = hd.substr(0, sp);
          const size_t size = (size_t)atoll(hd.c_str() + sp + 1);
          out.type = t == "commit" ? OBJ_COMMIT : t == "tree" ? OBJ_TREE : t == "blob" ? OBJ_BLOB : t == "tag" ? OBJ_TAG : 0;
          std::vector<uint8_t> all;
          if (out.type && inflate_exact(m, n, hd.size() + 1 + size, all)) { out.data.assign(all.begin() + (long)hd.size() + 1, all.end()); ok = true; }
        }
      }
      inflateEnd(&z);
    }
    munmap(const_cast<uint8_t*>(m), n);
    return ok;
  }

  void load_packed_refs() {
    std::ifstream is(dir_ + "/packed-refs");
    std::string line;
    while (std::getline(is, line)) {
      if (line.empty() || line[0] == '#' || line[0] == '^') continue;
      Oid id;
      if (line.size() > 41 && line[40] == ' ' && Oid::from_hex(line.data(), 40, id)) packed_[line.substr(41)] = id;
    }
  }

  bool read_ref(const std::string& name, Oid& out, int depth) {
    if (depth > 8 || name.empty() || name.find("..") != std::string::npos) return false;
    std::ifstream is(dir_ + "/" + name);
    std::string line;
    struct stat st;
    if (stat((dir_ + "/" + name).c_str(), &st) == 0 && S_ISREG(st.st_mode) && std::getline(is, line)) {
      while (!line.empty() && (line.back() == '\n' || line.back() == '\r' || line.back() == ' ')) line.pop_back();
      if (line.rfind("ref: ", 0) == 0) return read_ref(line.substr(5), out, depth + 1);
      return Oid::from_hex(line.data(), line.size(), out);
    }
    auto it = packed_.find(name);
    if (it == packed_.end()) return false;
    out = it->second;
    return true;
  }

  void list_dir_refs(const std::string& rel, std::vector<std::string>& out, const std::string& prefix = "") {
    DIR* d = opendir((dir_ + "/" + rel).c_str());
    if (!d) return;
    while (dirent* e = readdir(d)) {
      const std::string n = e->d_name;
      if (n == "." || n == "..") continue;
      struct stat st;
      if (stat((dir_ + "/" + rel + "/" + n).c_str(), &st) != 0) continue;
      if (S_ISDIR(st.st_mode)) list_dir_refs(rel + "/" + n, out, prefix + n + "/");
      else out.push_back(prefix + n);
    }
    closedir(d);
  }
};

}  // namespace gitstore
