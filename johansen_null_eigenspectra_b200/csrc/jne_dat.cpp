// Batched EIGENVALS_V6 writer / reader / resume scan (include/jne_dat.h).  Host-only.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/vfs.h>
#include <unistd.h>

#include "../../include/jne.h"
#include "../../include/jne_dat.h"

namespace {

const char kMagic[12] = {'E', 'I', 'G', 'E', 'N', 'V', 'A', 'L', 'S', '_', 'V', '6'};
const char kEof[8] = {'E', 'O', 'F', '_', 'M', 'A', 'R', 'K'};
constexpr long kHeader = 18, kTrailer = 17;

thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return JNE_ERR_IO; }

// Read-only view of a whole file (mmap): the record walk touches every page once at memory speed instead of one
// stdio seek + refill per record (a 10^7-record file is scanned in well under a second from the page cache).
struct Reader {
  const unsigned char* p = nullptr;
  long len = 0;
  int fd = -1;
  ~Reader() { close_map(); }
  void close_map() {
    if (p && len > 0) munmap(const_cast<unsigned char*>(p), (size_t)len);
    if (fd >= 0) ::close(fd);
    p = nullptr; fd = -1; len = 0;
  }
  bool open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
    len = (long)st.st_size;
    if (len > 0) {
      void* m = mmap(nullptr, (size_t)len, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m == MAP_FAILED) { ::close(fd); fd = -1; len = 0; return false; }
      p = static_cast<const unsigned char*>(m);
      madvise(m, (size_t)len, MADV_SEQUENTIAL);
    }
    return true;
  }
};

struct Header { uint8_t model, dim; uint32_t steps; };

// 0 ok, 1 file too short for a header, -1 magic mismatch
int read_header(const Reader& r, Header* h) {
  if (r.len < kHeader) return 1;
  const unsigned char* b = r.p;
  if (memcmp(b, kMagic, 12) != 0) return -1;
  h->model = b[12]; h->dim = b[13];
  h->steps = (uint32_t)b[14] | ((uint32_t)b[15] << 8) | ((uint32_t)b[16] << 16) | ((uint32_t)b[17] << 24);
  return 0;
}

bool read_trailer(const Reader& r, uint64_t* count, uint32_t* per_run) {
  if (r.len < kHeader + kTrailer) return false;
  const unsigned char* b = r.p + (r.len - kTrailer);
  if (memcmp(b, kEof, 8) != 0) return false;
  uint64_t c = 0;
  for (int i = 0; i < 8; ++i) c |= (uint64_t)b[8 + i] << (8 * i);
  *count = c; *per_run = b[16];
  return true;
}

// Reads one ULEB128 u32 at offset pos: >0 bytes consumed, 0 clean end of file before the first byte, <0 error.
inline int read_uleb(const Reader& r, long pos, uint32_t* v) {
  uint32_t result = 0; int shift = 0, n = 0;
  for (;;) {
    if (pos + n >= r.len) return n == 0 ? 0 : -1;
    const unsigned c = r.p[pos + n];
    ++n;
    if (n > 5) return -2;
    const uint32_t bits = c & 0x7F;
    if (shift == 28 && bits > 0x0F) return -3;
    result |= bits << shift;
    shift += 7;
    if (!(c & 0x80)) { *v = result; return n; }
  }
}

// Walks the records behind the header.  on_record(seed, count, payload offset) is called per COMPLETE record;
// returns the offset just behind the last complete record.  data_end: end of the record area (file length, or
// file length minus the trailer).  limit: stop after this many records (trailer-known fast path) or UINT64_MAX
// (scan path, reader.rs:157-216).
template <class F>
long walk(const Reader& r, long data_end, uint64_t limit, bool scan, F&& on_record, uint64_t* n_out, std::string* err) {
  long good_end = kHeader, pos = kHeader;
  uint64_t n = 0;
  while (n < limit) {
    uint32_t seed;
    const int used = read_uleb(r, pos, &seed);
    if (used <= 0) { if (!scan && err) *err = "Incomplete ULEB128 encoding"; break; }
    if (pos + used >= data_end) break;                 // the count byte must lie inside the record area
    const int cnt = r.p[pos + used];
    const long payload = pos + used + 1;
    if (cnt == 0) {
      if (scan) {   // a zero count followed by the EOF marker ends the data (reader.rs:178-190)
        if (payload + 8 <= r.len && memcmp(r.p + payload, kEof, 8) == 0) break;
      } else { if (err) *err = "Invalid eigenvalue count: cannot be zero"; break; }
    }
    if (payload + 8L * cnt > data_end) break;        // torn record: dropped (reader.rs:199-210)
    if (!on_record(seed, (uint32_t)cnt, payload)) break;
    pos = good_end = payload + 8L * cnt;
    ++n;
  }
  *n_out = n;
  return good_end;
}

}  // namespace

struct jne_dat_writer {
  int fd = -1;
  uint64_t pos = 0;       // end of the valid data == offset of the next record
  bool tmpfs = false;
  std::vector<unsigned char> buf;
  uint64_t written = 0;
  uint32_t per_run = 0;   // 0 = not yet known
  uint8_t model = 0, dim = 0;
  uint32_t steps = 0;
};

namespace {

bool write_all(int fd, const unsigned char* p, size_t n, uint64_t off) {
  while (n) {
    const ssize_t k = pwrite(fd, p, n, (off_t)off);
    if (k < 0) { if (errno == EINTR) continue; return false; }
    p += k; n -= (size_t)k; off += (uint64_t)k;
  }
  return true;
}

inline size_t encode_record(unsigned char* q, uint32_t seed, const double* vals, uint32_t p) {
  const int k = jne_uleb128_encode(seed, q);
  q[k] = (unsigned char)p;
  memcpy(q + k + 1, vals, 8 * (size_t)p);           // f64 little-endian == host representation (x86-64 / aarch64)
  return (size_t)k + 1 + 8 * (size_t)p;
}

}  // namespace

extern "C" {

const char* jne_dat_last_error(void) { return g_err.c_str(); }

int jne_uleb128_encode(uint32_t value, uint8_t out[5]) {
  int n = 0;
  do {
    uint8_t byte = value & 0x7F;
    value >>= 7;
    if (value) byte |= 0x80;
    out[n++] = byte;
  } while (value);
  return n;
}

int jne_uleb128_encoded_size(uint32_t v) { return v < (1u << 7) ? 1 : v < (1u << 14) ? 2 : v < (1u << 21) ? 3 : v < (1u << 28) ? 4 : 5; }

int jne_uleb128_decode(const uint8_t* bytes, size_t len, uint32_t* value) {
  uint32_t result = 0; int shift = 0;
  for (size_t i = 0; i < len; ++i) {
    if (i + 1 > 5) return -2;                                   // EncodingTooLong
    const uint32_t bits = bytes[i] & 0x7F;
    if (shift == 28 && bits > 0x0F) return -3;                  // ValueTooLarge
    result |= bits << shift;
    shift += 7;
    if (!(bytes[i] & 0x80)) { if (value) *value = result; return (int)i + 1; }
  }
  return -1;                                                    // IncompleteEncoding
}

uint64_t jne_dat_expected_file_size(uint64_t num_runs, uint32_t per_run) {
  // total ULEB128 bytes of seeds 1..=num_runs by size class (file_format.rs:29-86)
  auto upto = [](uint64_t n) {   // sum of encoded sizes of 1..n
    const uint64_t lim[5] = {127, 16383, 2097151, 268435455, 4294967295ull};
    uint64_t total = 0, lo = 0;
    for (int k = 0; k < 5 && n > lo; ++k) { const uint64_t hi = n < lim[k] ? n : lim[k]; total += (hi - lo) * (k + 1); lo = lim[k]; }
    return total;
  };
  const uint64_t seeds = num_runs == 0 ? 1 : upto(num_runs);
  return kHeader + seeds + num_runs + 8ull * per_run * num_runs + kTrailer;
}

int jne_dat_info(const char* path, uint8_t* model, uint8_t* dim, uint32_t* steps, uint64_t* n_records,
                 uint32_t* per_run, int* has_trailer) {
  Reader r;
  if (!r.open(path)) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
  const long len = r.len;
  Header h{};
  const int hr = read_header(r, &h);
  if (hr < 0) return fail("File format error: magic header mismatch");
  if (hr > 0) return fail("file too short for a header");
  uint64_t count = 0; uint32_t pr = 0;
  const bool trailer = read_trailer(r, &count, &pr);
  if (!trailer) {
    uint32_t first = 0;
    walk(r, len, UINT64_MAX, true, [&](uint32_t, uint32_t c, long) { if (!first) first = c; return true; }, &count, nullptr);
    pr = first;
  }
  if (model) *model = h.model;
  if (dim) *dim = h.dim;
  if (steps) *steps = h.steps;
  if (n_records) *n_records = count;
  if (per_run) *per_run = pr;
  if (has_trailer) *has_trailer = trailer ? 1 : 0;
  return JNE_OK;
}

int jne_dat_read(const char* path, uint32_t* seeds, double* eigs, uint64_t capacity, uint32_t p, uint64_t* n_read) {
  Reader r;
  if (!r.open(path)) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
  const long len = r.len;
  Header h{};
  const int hr = read_header(r, &h);
  if (hr < 0) return fail("File format error: magic header mismatch");
  if (hr > 0) return fail("file too short for a header");
  uint64_t count = 0; uint32_t pr = 0;
  const bool trailer = read_trailer(r, &count, &pr);
  std::string err;
  bool mismatch = false;
  uint64_t n = 0, walked = 0;
  const uint64_t limit = trailer ? (count < capacity ? count : capacity) : capacity;
  walk(r, trailer ? len - kTrailer : len, limit, !trailer,
       [&](uint32_t seed, uint32_t c, long payload) {
         if (c != p) { mismatch = true; err = "Eigenvalue count mismatch: expected " + std::to_string(p) + ", actual " + std::to_string(c); return false; }
         seeds[n] = seed;
         memcpy(eigs + n * p, r.p + payload, 8 * (size_t)p);      // little-endian host assumed (x86-64 / aarch64)
         ++n;
         return true;
       }, &walked, &err);
  if (mismatch) return fail(err);
  // Fast path (reader.rs:107-160): every one of the trailer's `count` records must be there -- a zero count, a bad
  // ULEB128 or a record cut short is InvalidData / UnexpectedEof in the reference, not a short read.
  if (trailer && n < limit)
    return fail(!err.empty() ? err : "unexpected end of file: the trailer promises " + std::to_string(count) +
                                     " records, " + std::to_string(n) + " are readable");
  if (n_read) *n_read = n;
  return JNE_OK;
}

int jne_dat_open(const char* path, uint8_t model, uint8_t dim, uint32_t steps, uint64_t* existing, jne_dat_writer** out) {
  if (!path || !out) return fail("path/out is NULL");
  *out = nullptr;
  uint64_t have = 0; uint32_t per_run = 0;
  bool fresh = access(path, F_OK) != 0;
  if (!fresh) {
    Reader r;
    if (!r.open(path)) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
    const long len = r.len;
    Header h{};
    const int hr = read_header(r, &h);
    if (hr < 0) {
      fresh = true;                                  // foreign magic: recreate (writer.rs:116-150)
    } else if (hr > 0) {
      fresh = true;                                  // shorter than a header: nothing to keep
    } else {
      if (h.model != model) return fail("Model mismatch: file has model " + std::to_string(h.model) + ", expected " + std::to_string(model));
      if (h.dim != dim) return fail("Dimension mismatch: file has dim " + std::to_string(h.dim) + ", expected " + std::to_string(dim));
      if (h.steps != steps) return fail("Steps mismatch: file has steps " + std::to_string(h.steps) + ", expected " + std::to_string(steps));
      uint64_t count = 0; uint32_t pr = 0;
      const bool trailer = read_trailer(r, &count, &pr);
      uint32_t first = 0;
      const long good_end = walk(r, trailer ? len - kTrailer : len, trailer ? count : UINT64_MAX, !trailer,
                                 [&](uint32_t, uint32_t c, long) { if (!first) first = c; return true; }, &have, nullptr);
      per_run = first;
      r.close_map();
      if (trailer && (have != count || good_end != len - kTrailer)) {
        // A FINISHED file whose records do not add up to its trailer is damaged.  The reference's progress check
        // treats it as "no progress" (progress.rs:51) and its writer then appends behind the damaged bytes
        // (writer.rs:150-158), which leaves an unreadable file.  Here nothing is thrown away silently: the damaged
        // file is set aside as <path>.damaged and the job restarts on a fresh file, as the progress check implies.
        const std::string aside = std::string(path) + ".damaged";
        fprintf(stderr, "WARNING: %s: trailer promises %llu records, %llu readable; kept as %s, starting a fresh file\n", path,
                (unsigned long long)count, (unsigned long long)have, aside.c_str());
        if (rename(path, aside.c_str()) != 0) return fail(std::string("cannot set the damaged file aside: ") + strerror(errno));
        fresh = true; have = 0; per_run = 0;
      } else if (truncate(path, good_end) != 0) {
        // drop the trailer (writer.rs:181-203) and, in an interrupted file, the torn tail so that appended records stay parseable
        return fail(std::string("truncate failed: ") + strerror(errno));
      }
    }
  }
  jne_dat_writer* w = new jne_dat_writer();
  w->model = model; w->dim = dim; w->steps = steps; w->written = have; w->per_run = per_run;
  w->fd = ::open(path, O_RDWR | O_CREAT | (fresh ? O_TRUNC : 0), 0644);
  if (w->fd < 0) { delete w; return fail(std::string("cannot open ") + path + " for writing: " + strerror(errno)); }
  struct stat st;
  if (fstat(w->fd, &st) != 0) { ::close(w->fd); delete w; return fail(std::string("fstat failed: ") + strerror(errno)); }
  w->pos = (uint64_t)st.st_size;                      // batches are appended whole at this offset: no stdio buffer
  struct statfs fs;
  w->tmpfs = fstatfs(w->fd, &fs) == 0 && (unsigned long)fs.f_type == 0x01021994ul;   // TMPFS_MAGIC
  if (fresh) {
    unsigned char h[kHeader];
    memcpy(h, kMagic, 12); h[12] = model; h[13] = dim;
    for (int i = 0; i < 4; ++i) h[14 + i] = (steps >> (8 * i)) & 0xFF;
    if (!write_all(w->fd, h, kHeader, 0)) { ::close(w->fd); delete w; return fail("header write failed"); }
    w->pos = kHeader;
  }
  if (existing) *existing = have;
  *out = w;
  return JNE_OK;
}

int jne_dat_append_batch(jne_dat_writer* w, const uint32_t* seeds, const double* eigs, uint64_t n, uint32_t p) {
  return jne_dat_append_batch_strided(w, seeds, eigs, n, p, p);
}

int jne_dat_append_batch_strided(jne_dat_writer* w, const uint32_t* seeds, const double* eigs, uint64_t n, uint32_t p,
                                 uint64_t stride) {
  if (!w || w->fd < 0) return fail("writer is closed");
  if (stride < p) return fail("stride is smaller than the eigenvalue count");
  if (p > 255) return fail("Too many eigenvalues: " + std::to_string(p) + " exceeds maximum of 255");
  if (n == 0) return JNE_OK;
  if (w->per_run == 0) w->per_run = p;
  if (p != w->per_run)
    return fail("Eigenvalue count mismatch: expected " + std::to_string(w->per_run) + ", actual " + std::to_string(p) +
                " (model " + std::to_string(w->model) + ", dim " + std::to_string(w->dim) + ", steps " + std::to_string(w->steps) + ")");
  const size_t rec_max = 5 + 1 + 8 * (size_t)p;
  const uint64_t chunk = 1u << 16;
  for (uint64_t a = 0; a < n; a += chunk) {
    const uint64_t m = n - a < chunk ? n - a : chunk;
    w->buf.resize(m * rec_max);
    unsigned char* q = w->buf.data();
    for (uint64_t i = 0; i < m; ++i) q += encode_record(q, seeds[a + i], eigs + (a + i) * stride, p);
    const size_t bytes = q - w->buf.data();
    if (!write_all(w->fd, w->buf.data(), bytes, w->pos)) return fail(std::string("write failed: ") + strerror(errno));
    w->pos += bytes;
  }
  w->written += n;
  return JNE_OK;
}

// ---- random-access batch (see include/jne_dat.h) ----
// Record sizes are known from the seeds, so every record of a batch has its place before any is written: the file grows
// by the batch, the new range is mapped MAP_SHARED, and whoever has rows encodes them straight into the file's pages
// (page allocation in parallel; no intermediate buffer, no write() through the inode lock).
}  // extern "C"

struct jne_dat_batch {
  static constexpr uint64_t kBlk = 1024;      // prefix table granularity, in records
  jne_dat_writer* w = nullptr;
  const uint32_t* seeds = nullptr;
  uint64_t n = 0, total = 0;
  uint32_t p = 0;
  std::vector<uint64_t> blk;                  // byte offset of record i * kBlk inside the batch
  bool mapped = false;                         // false: ranges are encoded into a private buffer and pwrite()n in place
  void* base = nullptr;
  size_t map_len = 0;
  unsigned char* dst = nullptr;
  long page = 4096;
  unsigned char first[5 + 1 + 8 * 255];
  size_t first_len = 0;
  std::atomic<uint64_t> filled{0};
  std::atomic<int> err{0};
};

namespace {
void batch_fill_range(jne_dat_batch* b, uint64_t first, uint64_t count, const double* rows, uint64_t stride) {
  const uint32_t p = b->p;
  uint64_t off = b->blk[first / jne_dat_batch::kBlk];
  for (uint64_t j = (first / jne_dat_batch::kBlk) * jne_dat_batch::kBlk; j < first; ++j)
    off += (uint64_t)jne_uleb128_encoded_size(b->seeds[j]) + 1 + 8 * (uint64_t)p;
  uint64_t bytes = count * (1 + 8 * (uint64_t)p);
  for (uint64_t j = first; j < first + count; ++j) bytes += (uint64_t)jne_uleb128_encoded_size(b->seeds[j]);
  unsigned char* q;
  thread_local std::vector<unsigned char> scratch;
  if (b->mapped) {
    q = b->dst + off;
#ifdef MADV_POPULATE_WRITE
    {   // allocate this range's pages in one call (whole pages inside it; the edges fault on first touch).  On tmpfs this
        // is also what turns a full file system into an error code instead of a SIGBUS.
      const uintptr_t lo = ((uintptr_t)q + b->page - 1) & ~(uintptr_t)(b->page - 1), hi = ((uintptr_t)q + bytes) & ~(uintptr_t)(b->page - 1);
      if (hi > lo && madvise((void*)lo, hi - lo, MADV_POPULATE_WRITE) != 0 && errno == EFAULT) { b->err.store(ENOSPC); return; }
    }
#endif
  } else {
    try { if (scratch.size() < bytes) scratch.resize(bytes); } catch (...) { b->err.store(ENOMEM); return; }
    q = scratch.data();
  }
  unsigned char* const q0 = q;
  uint64_t i = first;
  size_t skip = 0;
  if (first == 0 && count > 0) {               // record 0 goes in last (jne_dat_batch_end): the batch stays invalid until then
    b->first_len = encode_record(b->first, b->seeds[0], rows, p);
    q += b->first_len;
    skip = b->first_len;
    ++i;
  }
  for (; i < first + count; ++i) q += encode_record(q, b->seeds[i], rows + (i - first) * stride, p);
  if (!b->mapped && bytes > skip && !write_all(b->w->fd, q0 + skip, bytes - skip, b->w->pos + off + skip)) {
    b->err.store(errno ? errno : EIO);
    return;
  }
  b->filled.fetch_add(count);
}
}  // namespace

extern "C" {

int jne_dat_batch_begin(jne_dat_writer* w, const uint32_t* seeds, uint64_t n, uint32_t p, jne_dat_batch** out) {
  if (!out) return fail("out is NULL");
  *out = nullptr;
  if (!w || w->fd < 0) return fail("writer is closed");
  if (p < 1 || p > 255) return fail("Too many eigenvalues: " + std::to_string(p) + " exceeds maximum of 255");
  if (n == 0 || !seeds) return fail("empty batch");
  if (w->per_run == 0) w->per_run = p;
  if (p != w->per_run)
    return fail("Eigenvalue count mismatch: expected " + std::to_string(w->per_run) + ", actual " + std::to_string(p) +
                " (model " + std::to_string(w->model) + ", dim " + std::to_string(w->dim) + ", steps " + std::to_string(w->steps) + ")");
  jne_dat_batch* b = new (std::nothrow) jne_dat_batch();
  if (!b) return fail("out of memory");
  b->w = w; b->seeds = seeds; b->n = n; b->p = p;
  b->page = sysconf(_SC_PAGESIZE);
  try { b->blk.resize((n + jne_dat_batch::kBlk - 1) / jne_dat_batch::kBlk + 1); } catch (...) { delete b; return fail("out of memory"); }
  uint64_t off = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (i % jne_dat_batch::kBlk == 0) b->blk[i / jne_dat_batch::kBlk] = off;
    off += (uint64_t)jne_uleb128_encoded_size(seeds[i]) + 1 + 8 * (uint64_t)p;
  }
  b->total = off;
  // Two ways to put a range's bytes in place (JNE_DAT_MMAP=1 selects the second):
  //   pwrite (default)  a private buffer per range, then one positioned pwrite() -- on the 8-GPU box write() produces
  //                     new tmpfs pages at 20 GB/s into five files, a mapping at 12 GB/s, and tearing a mapping down
  //                     costs another 0.7 us per page (142 ms per 830 MB batch and file), profiles/r2_io_bench_8gpu_box.txt
  //   mapped pages      MAP_SHARED + MADV_POPULATE_WRITE, encoders write straight into the page cache
  static const bool want_map = [] { const char* e = getenv("JNE_DAT_MMAP"); return e && e[0] == '1'; }();
  b->mapped = want_map;
  // grow the file; on a disk file system reserve the blocks now so that a full disk is an error code here (on tmpfs the
  // pwrite / MADV_POPULATE_WRITE of the range reports it)
  if (!w->tmpfs) {
    const int e = posix_fallocate(w->fd, (off_t)w->pos, (off_t)b->total);
    if (e != 0 && e != EOPNOTSUPP && e != EINVAL) { delete b; return fail(std::string("cannot reserve file space: ") + strerror(e)); }
  }
  if (ftruncate(w->fd, (off_t)(w->pos + b->total)) != 0) { const int e = errno; delete b; return fail(std::string("cannot grow the file: ") + strerror(e)); }
  const unsigned char poison[5] = {0xFF, 0xFF, 0xFF, 0xFF, 0xFF};     // a scan stops here until the batch is committed
  if (b->mapped) {
    const uint64_t map_off = w->pos & ~(uint64_t)(b->page - 1), delta = w->pos - map_off;
    b->map_len = (size_t)(delta + b->total);
    b->base = mmap(nullptr, b->map_len, PROT_READ | PROT_WRITE, MAP_SHARED, w->fd, (off_t)map_off);
    if (b->base == MAP_FAILED) {
      const int e = errno;
      if (ftruncate(w->fd, (off_t)w->pos) != 0) { /* keep the first error */ }
      delete b;
      return fail(std::string("cannot map the file: ") + strerror(e));
    }
    b->dst = static_cast<unsigned char*>(b->base) + delta;
#ifdef MADV_POPULATE_WRITE
    if (madvise(b->base, (size_t)std::min<uint64_t>(b->map_len, (uint64_t)b->page), MADV_POPULATE_WRITE) != 0 && errno == EFAULT) {
      jne_dat_batch_end(b, 0);
      return fail(std::string("write failed: ") + strerror(ENOSPC));
    }
#endif
    memcpy(b->dst, poison, 5);
  } else if (!write_all(w->fd, poison, 5, w->pos)) {
    const int e = errno;
    if (ftruncate(w->fd, (off_t)w->pos) != 0) { /* keep the first error */ }
    delete b;
    return fail(std::string("write failed: ") + strerror(e));
  }
  *out = b;
  return JNE_OK;
}

int jne_dat_batch_fill(jne_dat_batch* b, uint64_t first, uint64_t count, const double* rows, uint64_t stride, int threads) {
  if (!b || !rows) return fail("batch / rows is NULL");
  if (first + count > b->n || stride < b->p) return fail("record range outside the batch");
  if (count == 0) return JNE_OK;
  const uint64_t bytes = count * 8 * (uint64_t)b->p;
  if (threads <= 1 || bytes < ((uint64_t)4 << 20)) { batch_fill_range(b, first, count, rows, stride); }
  else {
    if (threads > 16) threads = 16;
    const uint64_t per = (count + threads - 1) / threads;
    std::vector<std::thread> th;
    int started = 1;                               // range 0 is this thread's
    try {
      for (; started < threads; ++started) {
        const uint64_t a = std::min<uint64_t>((uint64_t)started * per, count), e = std::min<uint64_t>(a + per, count);
        if (a < e) th.emplace_back(batch_fill_range, b, first + a, e - a, rows + a * stride, stride);
      }
    } catch (...) { /* ranges started .. threads-1 did not get a helper: filled below */ }
    batch_fill_range(b, first, std::min<uint64_t>(per, count), rows, stride);
    const uint64_t a = std::min<uint64_t>((uint64_t)started * per, count);
    if (started < threads && a < count) batch_fill_range(b, first + a, count - a, rows + a * stride, stride);
    for (auto& x : th) x.join();
  }
  const int e = b->err.load();
  return e ? fail(std::string("write failed: ") + strerror(e)) : JNE_OK;
}

int jne_dat_batch_end(jne_dat_batch* b, int commit) {
  if (!b) return fail("batch is NULL");
  jne_dat_writer* w = b->w;
  int rc = JNE_OK;
  const bool ok = commit && b->err.load() == 0 && b->filled.load() == b->n && b->first_len > 0;
  if (commit && !ok)
    rc = fail(b->err.load() ? std::string("write failed: ") + strerror(b->err.load())
                            : "batch committed with " + std::to_string(b->filled.load()) + " of " + std::to_string(b->n) + " records filled");
  bool ok2 = ok;
  if (ok) {                                             // the batch becomes valid
    if (b->mapped) memcpy(b->dst, b->first, b->first_len);
    else if (!write_all(w->fd, b->first, b->first_len, w->pos)) { ok2 = false; rc = fail(std::string("write failed: ") + strerror(errno)); }
  }
  if (b->mapped && b->base && b->base != MAP_FAILED) munmap(b->base, b->map_len);
  if (ok2) { w->pos += b->total; w->written += b->n; }
  else if (ftruncate(w->fd, (off_t)w->pos) != 0 && rc == JNE_OK && commit) rc = fail(std::string("truncate failed: ") + strerror(errno));
  delete b;
  return rc;
}

int jne_dat_append_batch_strided_mt(jne_dat_writer* w, const uint32_t* seeds, const double* eigs, uint64_t n, uint32_t p,
                                    uint64_t stride, int threads) {
  if (threads <= 1 || n < 4096) return jne_dat_append_batch_strided(w, seeds, eigs, n, p, stride);
  if (stride < p) return fail("stride is smaller than the eigenvalue count");
  jne_dat_batch* b = nullptr;
  int rc = jne_dat_batch_begin(w, seeds, n, p, &b);
  if (rc != JNE_OK) return rc;
  rc = jne_dat_batch_fill(b, 0, n, eigs, stride, threads);
  const int rc2 = jne_dat_batch_end(b, rc == JNE_OK);
  return rc != JNE_OK ? rc : rc2;
}

int jne_dat_flush(jne_dat_writer* w) {
  if (!w || w->fd < 0) return fail("writer is closed");
  return JNE_OK;      // nothing is buffered in user space: every batch is already in the OS page cache
}

int jne_dat_finish(jne_dat_writer* w) {
  if (!w || w->fd < 0) return fail("writer is closed");
  unsigned char t[kTrailer];
  memcpy(t, kEof, 8);
  for (int i = 0; i < 8; ++i) t[8 + i] = (w->written >> (8 * i)) & 0xFF;
  t[16] = (unsigned char)w->per_run;
  const bool ok = write_all(w->fd, t, kTrailer, w->pos);
  const bool closed = ::close(w->fd) == 0;
  delete w;
  return ok && closed ? JNE_OK : fail(std::string("trailer write failed: ") + strerror(errno));
}

void jne_dat_abandon(jne_dat_writer* w) {
  if (!w) return;
  if (w->fd >= 0) ::close(w->fd);
  delete w;
}

int jne_dat_completed_bitmap(const char* path, uint8_t model, uint8_t dim, uint32_t steps, uint64_t num_runs,
                             uint8_t* bitmap, uint64_t* completed) {
  if (!bitmap) return fail("bitmap is NULL");
  memset(bitmap, 0, (num_runs + 7) / 8);
  if (completed) *completed = 0;
  if (access(path, F_OK) != 0) return JNE_OK;        // progress.rs:17-19
  Reader r;
  if (!r.open(path)) return JNE_OK;                  // unreadable: start over (progress.rs:51)
  const long len = r.len;
  Header h{};
  if (read_header(r, &h) != 0) return JNE_OK;      // damaged: start over
  if (h.model != model) return fail("Model mismatch: file has model " + std::to_string(h.model) + ", expected " + std::to_string(model));
  if (h.dim != dim) return fail("Dimension mismatch: file has dim " + std::to_string(h.dim) + ", expected " + std::to_string(dim));
  if (h.steps != steps) return fail("Steps mismatch: file has steps " + std::to_string(h.steps) + ", expected " + std::to_string(steps));
  uint64_t count = 0; uint32_t pr = 0;
  const bool trailer = read_trailer(r, &count, &pr);
  uint64_t n = 0;
  const long good_end = walk(r, trailer ? len - kTrailer : len, trailer ? count : UINT64_MAX, !trailer,
       [&](uint32_t seed, uint32_t, long) {
         if (seed >= 1 && seed <= num_runs) bitmap[(seed - 1) >> 3] |= (uint8_t)(1u << ((seed - 1) & 7));
         return true;
       }, &n, nullptr);
  if (trailer && (n != count || good_end != len - kTrailer)) {
    // the reference's fast read fails on such a file (reader.rs:107-160) and check_append_progress then reports
    // no progress at all (progress.rs:51): restart
    memset(bitmap, 0, (num_runs + 7) / 8);
    n = 0;
  }
  if (completed) *completed = n;      // ALL records of the file, also seeds beyond num_runs (progress.rs:47)
  return JNE_OK;
}

uint64_t jne_dat_remaining_seeds(const uint8_t* bitmap, uint64_t num_runs, uint32_t* out, uint64_t capacity) {
  uint64_t k = 0;
  for (uint64_t s = 1; s <= num_runs; ++s) {
    if (!(bitmap[(s - 1) >> 3] & (1u << ((s - 1) & 7)))) {
      if (out && k < capacity) out[k] = (uint32_t)s;
      ++k;
    }
  }
  return k;
}

}  // extern "C"
