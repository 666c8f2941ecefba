// K4, host half: frame the device-encoded DEFLATE pieces of every figure as a PNG file and write it.
//
// The device (png.cu) encodes only the scanlines with new content; the file is
//   signature, IHDR, one IDAT = zlib header + [content-row pieces | "repeat the line above" runs]* + final
//   empty stored block + Adler-32, IEND
// The pieces are read in place from the pinned read-back buffer (no copy: CRC-32 pass, then writev).  This
// is what png.assemble_png + a Python writer thread did; it runs here on native threads because the Python
// version held the interpreter lock while the next chunk of the directory was being planned (measured:
// planning 0.21 s -> 1.33 s when the two overlapped).  png.assemble_png stays as the in-memory variant and
// the tests' cross-check.
#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/uio.h>
#include <unistd.h>
#include <zlib.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "csgpu.h"

namespace {
constexpr uint32_t kAdler = 65521;

// n_lines scanlines of filter type 2 (Up) + 4 * width zero bytes as a byte-aligned raw-DEFLATE piece that
// ends on an empty stored block (Z_SYNC_FLUSH), so that pieces concatenate.  A constant per (width, n).
const std::string& zero_run(int width, int n_lines) {
  static std::mutex lock;
  static std::map<std::pair<int, int>, std::string> cache;
  {
    std::lock_guard<std::mutex> hold(lock);
    auto it = cache.find({width, n_lines});
    if (it != cache.end()) return it->second;
  }
  std::string line((size_t)1 + 4 * (size_t)width, '\0');
  line[0] = 2;
  z_stream z;
  memset(&z, 0, sizeof z);
  deflateInit2(&z, 6, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  std::string out;
  std::vector<unsigned char> buf(1 << 16);
  for (int k = 0; k < n_lines; ++k) {
    z.next_in = (Bytef*)line.data();
    z.avail_in = (uInt)line.size();
    const int flush = k + 1 == n_lines ? Z_SYNC_FLUSH : Z_NO_FLUSH;
    do {
      z.next_out = buf.data();
      z.avail_out = (uInt)buf.size();
      deflate(&z, flush);
      out.append((const char*)buf.data(), buf.size() - z.avail_out);
    } while (z.avail_out == 0);
  }
  deflateEnd(&z);
  std::lock_guard<std::mutex> hold(lock);
  if (cache.size() > 8192) cache.clear();
  return cache.emplace(std::make_pair(width, n_lines), std::move(out)).first->second;
}

void put_be32(unsigned char* p, uint32_t v) {
  p[0] = (unsigned char)(v >> 24), p[1] = (unsigned char)(v >> 16), p[2] = (unsigned char)(v >> 8), p[3] = (unsigned char)v;
}

int write_all(int fd, std::vector<iovec>& io) {
  size_t k = 0;
  while (k < io.size()) {
    const int n = (int)std::min<size_t>(io.size() - k, 512);
    ssize_t put = writev(fd, &io[k], n);
    if (put < 0) {
      if (errno == EINTR) continue;
      return errno;
    }
    while (put > 0 && k < io.size()) {  // advance over what was written (short writes keep the rest)
      if ((size_t)put >= io[k].iov_len) {
        put -= (ssize_t)io[k].iov_len;
        ++k;
      } else {
        io[k].iov_base = (char*)io[k].iov_base + put;
        io[k].iov_len -= (size_t)put;
        put = 0;
      }
    }
    while (k < io.size() && io[k].iov_len == 0) ++k;
  }
  return 0;
}

int finish_one(csg_png_file& job, const int32_t* rows_all, const uint8_t* packed, const int64_t* offsets, const uint32_t* adler) {
  const int W = job.width, H = job.height, per_row = job.segs_per_row, n_rows = job.n_rows;
  const int32_t* rows = rows_all + job.row_first;
  const int64_t s0 = job.seg_first;
  const int64_t line_bytes = 1 + 4 * (int64_t)W;
  // Adler-32 of the filtered stream from the pieces' partial sums: with A the running byte sum (+1) before a
  // piece of n bytes, the piece adds n * A + weighted to B (see png.adler32_of_segments)
  uint64_t a = 1, b = 0;
  std::vector<iovec> io;
  io.reserve(8 + 2 * (size_t)n_rows);
  unsigned char head[8 + 25 + 8 + 2];
  io.push_back({head, sizeof head});
  uint64_t idat = 2;
  int64_t run_start = s0;  // first segment of the device bytes not yet listed
  std::vector<const std::string*> runs;
  for (int r = 0; r < n_rows; ++r) {
    for (int c = 0; c < per_row; ++c) {
      const int64_t s = s0 + (int64_t)r * per_row + c;
      const int npx = std::min(1024, W - 1024 * c);
      const uint64_t n = 4 * (uint64_t)npx + (c == 0 ? 1 : 0);
      b = (b + (n % kAdler) * a + adler[2 * s + 1]) % kAdler;
      a = (a + adler[2 * s]) % kAdler;
    }
    const int next = r + 1 < n_rows ? rows[r + 1] : H;
    const int64_t g = (int64_t)next - rows[r] - 1;  // repeated lines after this content row
    if (g > 0) {
      const int64_t end = s0 + (int64_t)(r + 1) * per_row;
      io.push_back({(void*)(packed + offsets[run_start]), (size_t)(offsets[end] - offsets[run_start])});
      idat += io.back().iov_len;
      run_start = end;
      const std::string& z = zero_run(W, (int)g);
      io.push_back({(void*)z.data(), z.size()});
      idat += z.size();
      // the run's only non-zero bytes are the filter bytes (2) opening each of its g lines
      const uint64_t L = (uint64_t)g * (uint64_t)line_bytes;
      const uint64_t sum = (2 * (uint64_t)g) % kAdler;
      // sum over lines k of (L - k * line_bytes) * 2
      const unsigned __int128 w = (unsigned __int128)2 * ((unsigned __int128)g * L - (unsigned __int128)line_bytes * ((unsigned __int128)g * (g - 1) / 2));
      b = (b + (L % kAdler) * a + (uint64_t)(w % kAdler)) % kAdler;
      a = (a + sum) % kAdler;
    }
  }
  const int64_t s1 = s0 + (int64_t)n_rows * per_row;
  if (run_start < s1) {
    io.push_back({(void*)(packed + offsets[run_start]), (size_t)(offsets[s1] - offsets[run_start])});
    idat += io.back().iov_len;
  }
  unsigned char tail[5 + 4 + 4 + 12] = {0x01, 0x00, 0x00, 0xff, 0xff};
  put_be32(tail + 5, (uint32_t)((b << 16) | a));
  idat += 9;
  if (idat > 0x7fffffffu) return EFBIG;
  // ---- head: signature, IHDR, IDAT length + tag + zlib header
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  memcpy(head, sig, 8);
  put_be32(head + 8, 13);
  memcpy(head + 12, "IHDR", 4);
  put_be32(head + 16, (uint32_t)W);
  put_be32(head + 20, (uint32_t)H);
  head[24] = 8, head[25] = 6, head[26] = 0, head[27] = 0, head[28] = 0;
  put_be32(head + 29, (uint32_t)crc32_z(0, head + 12, 17));
  put_be32(head + 33, (uint32_t)idat);
  memcpy(head + 37, "IDAT", 4);
  head[41] = 0x78, head[42] = 0x01;
  uLong crc = crc32_z(0, head + 37, 6);
  for (size_t k = 1; k < io.size(); ++k) crc = crc32_z(crc, (const Bytef*)io[k].iov_base, io[k].iov_len);
  crc = crc32_z(crc, tail, 9);
  put_be32(tail + 9, (uint32_t)crc);
  put_be32(tail + 13, 0);
  memcpy(tail + 17, "IEND", 4);
  put_be32(tail + 21, (uint32_t)crc32_z(0, tail + 17, 4));
  io.push_back({tail, sizeof tail});
  int64_t total = 0;
  for (auto& v : io) total += (int64_t)v.iov_len;
  job.file_bytes = total;
  const int fd = open(job.path, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
  if (fd < 0) return errno;
  int err = write_all(fd, io);
  if (close(fd) != 0 && !err) err = errno;
  return err;
}
}  // namespace

extern "C" int csg_png_write_files(csg_png_file* files, int n_files, const int32_t* rows, const uint8_t* packed,
                                   const int64_t* offsets, const uint32_t* adler, int n_threads) {
  if (n_files < 0 || (n_files > 0 && (!files || !rows || !packed || !offsets || !adler))) return CSG_ERR_ARG;
  std::atomic<int> next{0}, failed{0};
  auto work = [&]() {
    for (int k = next.fetch_add(1); k < n_files; k = next.fetch_add(1)) {
      files[k].status = files[k].path ? finish_one(files[k], rows, packed, offsets, adler) : EINVAL;
      if (files[k].status) failed.fetch_add(1);
    }
  };
  const int n = std::max(1, std::min(n_threads, n_files));
  std::vector<std::thread> pool;
  for (int t = 1; t < n; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return failed.load() ? CSG_ERR_IO : CSG_OK;
}
