// rg_data.cu -- host side of the tile data path: what the reference gets from the C extensions
// `lmdb` and `lz4framed` in PatchRNADataset (src/read_data.py:284-371).  No device code here.
//
//   * a read-only, memory-mapped reader of the LMDB file format (one file per slide, `subdir=False, readonly=True,
//     lock=False` in the reference): meta page selection, B+tree descent with the default key order, overflow pages;
//   * an LZ4 *frame* decoder (magic 0x184D2204; linked or independent blocks, optional content size / dictionary id /
//     block and content checksum fields are parsed and skipped).
//
// Both are written from the published format descriptions (lmdb.h / mdb.c structure layout of LMDB 0.9, 64-bit;
// lz4_Frame_format.md and lz4_Block_format.md); the LZ4 decoder is pinned against the system liblz4 in
// tests/test_host_cpu.py, the LMDB reader against files laid out by a test-side writer (py-lmdb is not installed in
// this image: parity with liblmdb-written files is unpinned).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include "rg_host.cuh"

namespace rg {

namespace {

constexpr uint32_t kMdbMagic = 0xBEEFC0DEu;
constexpr size_t kPageHdr = 16;             // pgno(8) pad(2) flags(2) lower(2) upper(2) -- or pages(4) for overflow
constexpr uint16_t P_BRANCH = 0x01, P_LEAF = 0x02, P_OVERFLOW = 0x04, P_META = 0x08, P_LEAF2 = 0x20;
constexpr uint16_t F_BIGDATA = 0x01, F_SUBDATA = 0x02, F_DUPDATA = 0x04;
constexpr size_t kNodeHdr = 8;              // lo(2) hi(2) flags(2) ksize(2)

struct Lmdb {
  int fd = -1;
  const uint8_t* map = nullptr;
  size_t size = 0;
  uint32_t psize = 0;
  uint64_t root = 0, entries = 0, last_pg = 0;
  uint16_t depth = 0;
};

template <typename T>
inline T rd(const uint8_t* p) {
  T v;
  memcpy(&v, p, sizeof(T));
  return v;
}

// MDB_meta after the page header: magic(4) version(4) address(8) mapsize(8) dbs[2] x 48 last_pg(8) txnid(8)
// MDB_db: pad(4) flags(2) depth(2) branch_pages(8) leaf_pages(8) overflow_pages(8) entries(8) root(8)
constexpr size_t kMetaDb0 = 24, kDbSize = 48, kMetaLastPg = kMetaDb0 + 2 * kDbSize, kMetaTxn = kMetaLastPg + 8;

bool parse_meta(const uint8_t* page, size_t avail, uint64_t* txn, Lmdb* out) {
  if (avail < kPageHdr + kMetaTxn + 8) return false;
  if ((rd<uint16_t>(page + 10) & P_META) == 0) return false;
  const uint8_t* m = page + kPageHdr;
  if (rd<uint32_t>(m) != kMdbMagic || rd<uint32_t>(m + 4) != 1u) return false;
  const uint8_t* main_db = m + kMetaDb0 + kDbSize;
  out->psize = rd<uint32_t>(m + kMetaDb0);                 // the free DB's md_pad holds the page size
  out->depth = rd<uint16_t>(main_db + 6);
  out->entries = rd<uint64_t>(main_db + 32);
  out->root = rd<uint64_t>(main_db + 40);
  out->last_pg = rd<uint64_t>(m + kMetaLastPg);
  *txn = rd<uint64_t>(m + kMetaTxn);
  return true;
}

inline int key_cmp(const uint8_t* a, size_t an, const uint8_t* b, size_t bn) {
  const int c = memcmp(a, b, std::min(an, bn));
  if (c != 0) return c;
  return an < bn ? -1 : (an > bn ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------ LZ4
// one LZ4 block; `dst` may already hold `done` bytes of earlier (linked) blocks that matches can reach back into
long long lz4_block(const uint8_t* src, size_t n, uint8_t* dst, size_t done, size_t cap) {
  size_t ip = 0, op = done;
  while (ip < n) {
    const uint8_t token = src[ip++];
    size_t lit = token >> 4;
    if (lit == 15) {
      uint8_t b;
      do {
        if (ip >= n) return -1;
        b = src[ip++];
        lit += b;
      } while (b == 255);
    }
    if (ip + lit > n) return -1;
    if (op + lit > cap) return -2;
    memcpy(dst + op, src + ip, lit);
    ip += lit;
    op += lit;
    if (ip >= n) break;                                   // the last sequence has literals only
    if (ip + 2 > n) return -1;
    const size_t off = src[ip] | (static_cast<size_t>(src[ip + 1]) << 8);
    ip += 2;
    if (off == 0 || off > op) return -1;
    size_t ml = token & 15;
    if (ml == 15) {
      uint8_t b;
      do {
        if (ip >= n) return -1;
        b = src[ip++];
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (op + ml > cap) return -2;
    const uint8_t* from = dst + op - off;
    if (off >= ml) {
      memcpy(dst + op, from, ml);
    } else {
      for (size_t i = 0; i < ml; ++i) dst[op + i] = from[i];   // overlapping copy replicates the pattern
    }
    op += ml;
  }
  return static_cast<long long>(op - done);
}

}  // namespace

}  // namespace rg

using namespace rg;

extern "C" {

void* rg_lmdb_open(const char* path) {
  if (!path) {
    set_error("rg_lmdb_open: null path");
    return nullptr;
  }
  Lmdb db;
  db.fd = open(path, O_RDONLY);
  if (db.fd < 0) {
    set_error("rg_lmdb_open: cannot open %s", path);
    return nullptr;
  }
  struct stat st;
  if (fstat(db.fd, &st) != 0 || st.st_size < 2 * 512) {
    set_error("rg_lmdb_open: %s is too small to be an LMDB file", path);
    close(db.fd);
    return nullptr;
  }
  db.size = static_cast<size_t>(st.st_size);
  void* m = mmap(nullptr, db.size, PROT_READ, MAP_SHARED, db.fd, 0);
  if (m == MAP_FAILED) {
    set_error("rg_lmdb_open: mmap of %s failed", path);
    close(db.fd);
    return nullptr;
  }
  db.map = static_cast<const uint8_t*>(m);
  // meta page 0 gives the page size; meta page 1 sits one page later; the one with the larger transaction id is current
  Lmdb m0, m1;
  uint64_t t0 = 0, t1 = 0;
  const bool ok0 = parse_meta(db.map, db.size, &t0, &m0);
  bool ok1 = false;
  if (ok0 && m0.psize >= 512 && static_cast<size_t>(m0.psize) * 2 <= db.size)
    ok1 = parse_meta(db.map + m0.psize, db.size - m0.psize, &t1, &m1);
  if (!ok0 && !ok1) {
    set_error("rg_lmdb_open: %s has no valid LMDB meta page (magic / version mismatch)", path);
    munmap(m, db.size);
    close(db.fd);
    return nullptr;
  }
  const Lmdb& cur = (ok1 && (!ok0 || t1 > t0)) ? m1 : m0;
  db.psize = ok0 ? m0.psize : m1.psize;
  db.root = cur.root;
  db.entries = cur.entries;
  db.depth = cur.depth;
  db.last_pg = cur.last_pg;
  if (db.psize < 512 || (db.psize & (db.psize - 1)) != 0) {
    set_error("rg_lmdb_open: implausible page size %u in %s", db.psize, path);
    munmap(m, db.size);
    close(db.fd);
    return nullptr;
  }
  return new Lmdb(db);
}

void rg_lmdb_close(void* h) {
  Lmdb* db = static_cast<Lmdb*>(h);
  if (!db) return;
  if (db->map) munmap(const_cast<uint8_t*>(db->map), db->size);
  if (db->fd >= 0) close(db->fd);
  delete db;
}

int rg_lmdb_stat(void* h, unsigned long long* entries, unsigned* psize, unsigned* depth) {
  Lmdb* db = static_cast<Lmdb*>(h);
  RG_CHECK_ARG(db, "rg_lmdb_stat: null handle");
  if (entries) *entries = db->entries;
  if (psize) *psize = db->psize;
  if (depth) *depth = db->depth;
  return 0;
}

// *val points into the mapping (valid until rg_lmdb_close).  Returns 0, RG_ENOTFOUND, or RG_EINVAL on a corrupt tree.
int rg_lmdb_get(void* h, const void* key, size_t klen, const void** val, size_t* vlen) {
  Lmdb* db = static_cast<Lmdb*>(h);
  RG_CHECK_ARG(db && key && val && vlen, "rg_lmdb_get: null argument");
  const uint8_t* k = static_cast<const uint8_t*>(key);
  if (db->root == ~0ull || db->entries == 0) return RG_ENOTFOUND;
  uint64_t pgno = db->root;
  for (int level = 0; level < 64; ++level) {
    if ((pgno + 1) * db->psize > db->size) {
      set_error("rg_lmdb_get: page %llu lies outside the file", static_cast<unsigned long long>(pgno));
      return RG_EINVAL;
    }
    const uint8_t* pg = db->map + pgno * db->psize;
    const uint16_t flags = rd<uint16_t>(pg + 10);
    const uint16_t lower = rd<uint16_t>(pg + 12);
    if (lower < kPageHdr || lower > db->psize || (flags & (P_BRANCH | P_LEAF)) == 0 || (flags & P_LEAF2)) {
      set_error("rg_lmdb_get: unexpected page (flags 0x%x) in the main tree", flags);
      return RG_EINVAL;
    }
    const int nkeys = (lower - static_cast<int>(kPageHdr)) >> 1;
    auto node = [&](int i) -> const uint8_t* {
      const uint16_t off = rd<uint16_t>(pg + kPageHdr + 2 * i);
      return (off + kNodeHdr <= db->psize) ? pg + off : nullptr;
    };
    if (flags & P_BRANCH) {
      // the first separator of a branch page is implicit (empty key): child i covers keys >= key(i)
      int lo = 1, hi = nkeys - 1, pick = 0;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const uint8_t* nd = node(mid);
        if (!nd) return RG_EINVAL;
        const uint16_t ks = rd<uint16_t>(nd + 6);
        if (key_cmp(k, klen, nd + kNodeHdr, ks) >= 0) {
          pick = mid;
          lo = mid + 1;
        } else {
          hi = mid - 1;
        }
      }
      const uint8_t* nd = node(pick);
      if (!nd || nkeys < 1) return RG_EINVAL;
      pgno = static_cast<uint64_t>(rd<uint16_t>(nd)) | (static_cast<uint64_t>(rd<uint16_t>(nd + 2)) << 16) |
             (static_cast<uint64_t>(rd<uint16_t>(nd + 4)) << 32);
      continue;
    }
    int lo = 0, hi = nkeys - 1;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1;
      const uint8_t* nd = node(mid);
      if (!nd) return RG_EINVAL;
      const uint16_t ks = rd<uint16_t>(nd + 6);
      const int c = key_cmp(k, klen, nd + kNodeHdr, ks);
      if (c == 0) {
        const uint16_t nf = rd<uint16_t>(nd + 4);
        const size_t dsz = static_cast<size_t>(rd<uint16_t>(nd)) | (static_cast<size_t>(rd<uint16_t>(nd + 2)) << 16);
        if (nf & (F_SUBDATA | F_DUPDATA)) {
          set_error("rg_lmdb_get: sub-databases / duplicate keys are not supported");
          return RG_EINVAL;
        }
        const uint8_t* data = nd + kNodeHdr + ks;
        if (nf & F_BIGDATA) {
          const uint64_t opg = rd<uint64_t>(data);
          const uint8_t* ov = db->map + opg * db->psize;
          if ((opg + 1) * db->psize > db->size || (rd<uint16_t>(ov + 10) & P_OVERFLOW) == 0 ||
              opg * db->psize + kPageHdr + dsz > db->size) {
            set_error("rg_lmdb_get: bad overflow page %llu", static_cast<unsigned long long>(opg));
            return RG_EINVAL;
          }
          *val = ov + kPageHdr;
        } else {
          if (static_cast<size_t>(data - pg) + dsz > db->psize) return RG_EINVAL;
          *val = data;
        }
        *vlen = dsz;
        return 0;
      }
      if (c < 0) hi = mid - 1;
      else lo = mid + 1;
    }
    return RG_ENOTFOUND;
  }
  set_error("rg_lmdb_get: tree deeper than 64 levels (corrupt file)");
  return RG_EINVAL;
}

// Decompress one LZ4 frame.  Returns the number of bytes written, -1 on a malformed frame, -2 when `cap` is too small
// (call again with a larger buffer; when the frame header carries the content size, *content_size receives it).
long long rg_lz4f_decompress(const void* src_, size_t n, void* dst_, size_t cap, long long* content_size) {
  const uint8_t* src = static_cast<const uint8_t*>(src_);
  uint8_t* dst = static_cast<uint8_t*>(dst_);
  if (content_size) *content_size = -1;
  if (!src || n < 7 || rd<uint32_t>(src) != 0x184D2204u) return -1;
  const uint8_t flg = src[4];
  if ((flg >> 6) != 1) return -1;                                  // version
  const bool block_checksum = flg & 0x10, has_size = flg & 0x08, content_checksum = flg & 0x04, has_dict = flg & 0x01;
  const bool linked = (flg & 0x20) == 0;
  size_t ip = 6;                                                   // magic(4) FLG BD
  if (has_size) {
    if (ip + 8 > n) return -1;
    if (content_size) *content_size = static_cast<long long>(rd<uint64_t>(src + ip));
    ip += 8;
  }
  if (has_dict) ip += 4;
  ip += 1;                                                         // header checksum byte
  size_t op = 0;
  for (;;) {
    if (ip + 4 > n) return -1;
    const uint32_t bs = rd<uint32_t>(src + ip);
    ip += 4;
    if (bs == 0) break;                                            // EndMark
    const size_t len = bs & 0x7FFFFFFFu;
    if (ip + len > n) return -1;
    if (bs & 0x80000000u) {                                        // stored uncompressed
      if (!dst || op + len > cap) return -2;
      memcpy(dst + op, src + ip, len);
      op += len;
    } else {
      if (!dst) return -2;
      // linked blocks may reference everything decoded so far; independent ones only their own output
      const long long got = linked ? lz4_block(src + ip, len, dst, op, cap)
                                   : lz4_block(src + ip, len, dst + op, 0, cap - op);
      if (got < 0) return got;
      op += static_cast<size_t>(got);
    }
    ip += len + (block_checksum ? 4 : 0);
  }
  (void)content_checksum;                                          // trailing xxh32 is not verified
  return static_cast<long long>(op);
}

}  // extern "C"
