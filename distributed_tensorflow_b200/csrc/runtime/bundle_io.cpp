// Tensor-bundle data file writer/reader: positional I/O of many tensors into one
// `<prefix>.data-00000-of-00001` file, 64-byte aligned, parallelised over a small thread pool, plus CRC32.
// The JSON index is written by Python (train/saver.py); this is the bulk-bytes path (TF: SaveV2/RestoreV2).
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

uint32_t crc_table[256];
std::atomic<bool> crc_ready{false};

void crc_init() {
  if (crc_ready.load()) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
    crc_table[i] = c;
  }
  crc_ready.store(true);
}

int pwrite_all(int fd, const uint8_t* p, int64_t n, int64_t off) {
  while (n > 0) {
    ssize_t w = pwrite(fd, p, (size_t)n, (off_t)off);
    if (w < 0) {
      if (errno == EINTR) continue;
      return -errno;
    }
    p += w; n -= w; off += w;
  }
  return 0;
}

int pread_all(int fd, uint8_t* p, int64_t n, int64_t off) {
  while (n > 0) {
    ssize_t r = pread(fd, p, (size_t)n, (off_t)off);
    if (r < 0) {
      if (errno == EINTR) continue;
      return -errno;
    }
    if (r == 0) return -EIO;      // truncated file
    p += r; n -= r; off += r;
  }
  return 0;
}

// ---- CRC32C (Castagnoli): the checksum of TensorFlow's checkpoint (tensor bundle) and event-file formats ----------
uint32_t crc32c_table[256];
std::atomic<bool> crc32c_ready{false};

void crc32c_init() {
  if (crc32c_ready.load()) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (0x82F63B78u ^ (c >> 1)) : (c >> 1);
    crc32c_table[i] = c;
  }
  crc32c_ready.store(true);
}

uint32_t crc32c_sw(uint32_t c, const uint8_t* p, int64_t n) {
  crc32c_init();
  for (int64_t i = 0; i < n; ++i) c = crc32c_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c;
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(uint32_t c, const uint8_t* p, int64_t n) {
  uint64_t c64 = c;
  while (n >= 8) {
    uint64_t v;
    memcpy(&v, p, 8);
    c64 = __builtin_ia32_crc32di(c64, v);
    p += 8; n -= 8;
  }
  c = (uint32_t)c64;
  while (n > 0) {
    c = __builtin_ia32_crc32qi(c, *p);
    ++p; --n;
  }
  return c;
}
bool have_sse42() { return __builtin_cpu_supports("sse4.2"); }
#else
uint32_t crc32c_hw(uint32_t c, const uint8_t* p, int64_t n) { return crc32c_sw(c, p, n); }
bool have_sse42() { return false; }
#endif

}  // namespace

extern "C" {

// Plain (unmasked) CRC32C of a buffer; `seed` = a previous result to continue a running checksum (0 to start).
uint32_t dtf_crc32c(const void* data, int64_t n, uint32_t seed) {
  static const bool hw = have_sse42();
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = seed ^ 0xFFFFFFFFu;
  c = hw ? crc32c_hw(c, p, n) : crc32c_sw(c, p, n);
  return c ^ 0xFFFFFFFFu;
}

uint32_t dtf_crc32(const void* data, int64_t n) {
  crc_init();
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = 0xFFFFFFFFu;
  for (int64_t i = 0; i < n; ++i) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}

// Write `count` blobs at the given offsets; file is truncated to total_bytes and fsync'ed.  0 on success.
int dtf_bundle_write(const char* path, int count, const void* const* ptrs, const int64_t* sizes, const int64_t* offsets,
                     int64_t total_bytes, int threads) {
  int fd = open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
  if (fd < 0) return -errno;
  if (ftruncate(fd, (off_t)total_bytes) != 0) {
    int e = -errno;
    close(fd);
    return e;
  }
  std::atomic<int> next{0}, err{0};
  auto work = [&] {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= count) return;
      int rc = pwrite_all(fd, static_cast<const uint8_t*>(ptrs[i]), sizes[i], offsets[i]);
      if (rc) err.store(rc);
    }
  };
  if (threads < 1) threads = 1;
  if (threads > count) threads = count > 0 ? count : 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  int rc = err.load();
  if (!rc && fsync(fd) != 0) rc = -errno;
  close(fd);
  return rc;
}

// Read `count` regions into caller-provided buffers.  0 on success, -EIO on truncation.
int dtf_bundle_read(const char* path, int count, void* const* ptrs, const int64_t* sizes, const int64_t* offsets,
                    int threads) {
  int fd = open(path, O_RDONLY);
  if (fd < 0) return -errno;
  std::atomic<int> next{0}, err{0};
  auto work = [&] {
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= count) return;
      int rc = pread_all(fd, static_cast<uint8_t*>(ptrs[i]), sizes[i], offsets[i]);
      if (rc) err.store(rc);
    }
  };
  if (threads < 1) threads = 1;
  if (threads > count) threads = count > 0 ? count : 1;
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  close(fd);
  return err.load();
}

// ---- TFRecord files (the record framing of TensorFlow's input files and event logs) ----------------------------------------
// record = uint64 length | masked crc32c(length) | payload | masked crc32c(payload)        (little endian)
// Scans a whole file that the caller mapped / read into memory: verifies both checksums of every record (SSE4.2 crc32c, no
// interpreter involved) and writes the payload offsets / lengths.  Returns the number of complete records found (a truncated
// tail -- a writer that is still appending -- ends the scan), or -(index + 1) of the first record whose checksum does not match.
// `max_records` < 0: count only.
static inline uint32_t dtf_mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xA282EAD8u; }

int64_t dtf_tfrecord_scan(const void* data, int64_t nbytes, int64_t* offsets, int64_t* lengths, int64_t max_records, int verify) {
  const uint8_t* p = static_cast<const uint8_t*>(data);
  int64_t at = 0, n = 0;
  while (at + 12 <= nbytes) {
    if (max_records >= 0 && n >= max_records) return n;           // the caller's arrays are full (or it asked for a prefix)
    uint64_t len;
    uint32_t hc;
    memcpy(&len, p + at, 8);
    memcpy(&hc, p + at + 8, 4);
    if (verify && dtf_mask_crc(dtf_crc32c(p + at, 8, 0)) != hc) return -(n + 1);
    if (len > (uint64_t)(nbytes - at - 16) || at + 16 > nbytes) break;                 // truncated tail
    const int64_t body = at + 12;
    uint32_t pc;
    memcpy(&pc, p + body + (int64_t)len, 4);
    if (verify && dtf_mask_crc(dtf_crc32c(p + body, (int64_t)len, 0)) != pc) return -(n + 1);
    if (max_records >= 0) {
      offsets[n] = body;
      lengths[n] = (int64_t)len;
    }
    ++n;
    at = body + (int64_t)len + 4;
  }
  return n;
}

}  // extern "C"
