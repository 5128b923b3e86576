// wav.cu -- WAV container walk (host metadata only) and PCM payload decode (device).
// "next" row f3 of SURVEY.md §8: mindaudio/data/io.py:347-747 (`read`, `_fmt_chunk`, `_data_chunk`,
// `_skip_unknown_chunk`).  The walk is the reference's state machine restated over a byte buffer, quirks included
// (the `offset` argument skips BYTES, `duration` counts items over all channels, parsing resumes right after the
// items that were read).  The arithmetic of the "unified output format" (io.py:741-746) runs on the device.
#include "common.cuh"

#include <errno.h>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <thread>

namespace mafe {
namespace {

struct ByteFile {           // file-object semantics over a buffer: short reads at EOF, seeks may pass EOF
  const uint8_t* p;
  int64_t n;
  int64_t pos;
  int64_t read(uint8_t* dst, int64_t k) {
    int64_t avail = pos < n ? n - pos : 0;
    if (k > avail) k = avail;
    if (k > 0) { memcpy(dst, p + pos, (size_t)k); pos += k; }
    return k > 0 ? k : 0;
  }
  void skip_read(int64_t k) { int64_t avail = pos < n ? n - pos : 0; pos += k < avail ? k : avail; }   // read(k), result dropped
  void seek_rel(int64_t k) { pos += k; }
};

inline uint32_t rd_u32(const uint8_t* b, bool be) {
  return be ? ((uint32_t)b[0] << 24 | (uint32_t)b[1] << 16 | (uint32_t)b[2] << 8 | b[3])
            : ((uint32_t)b[3] << 24 | (uint32_t)b[2] << 16 | (uint32_t)b[1] << 8 | b[0]);
}
inline uint16_t rd_u16(const uint8_t* b, bool be) { return be ? (uint16_t)(b[0] << 8 | b[1]) : (uint16_t)(b[1] << 8 | b[0]); }

int fail(mafe_wav_info* info, int kind) { info->error_kind = kind; return MAFE_E_INVALID_ARG; }

const char* kSupported = "PCM, IEEE_FLOAT";

// io.py:520-538; false = the size field is cut short (struct.error in the reference)
bool skip_unknown_chunk(ByteFile& f, bool be) {
  uint8_t b[4];
  const int64_t got = f.read(b, 4);
  if (got == 0) return true;                     // `if data:` -- nothing left, nothing to skip
  if (got < 4) return false;
  const uint32_t size = rd_u32(b, be);
  f.seek_rel(size);
  if (size & 1) f.seek_rel(1);
  return true;
}

}  // namespace
}  // namespace mafe

using namespace mafe;

extern "C" int mafe_wav_parse(const void* bytes, int64_t n_bytes, double offset_s, double duration_s, int32_t filelike,
                              mafe_wav_info* info) {
  MAFE_REQUIRE(info != nullptr, "mafe_wav_parse: info is NULL");
  memset(info, 0, sizeof(*info));
  MAFE_REQUIRE(bytes != nullptr || n_bytes == 0, "mafe_wav_parse: NULL buffer");
  MAFE_REQUIRE(n_bytes >= 0, "mafe_wav_parse: negative size");
  ByteFile f{(const uint8_t*)bytes, n_bytes, 0};
  uint8_t b[40];
  // ---- riff chunk (io.py:652-676)
  int64_t got = f.read(b, 4);
  bool be;
  if (got == 4 && memcmp(b, "RIFF", 4) == 0) be = false;
  else if (got == 4 && memcmp(b, "RIFX", 4) == 0) be = true;
  else {
    set_error("File format b'%.*s' not understood. Only 'RIFF' and 'RIFX' supported.", (int)got, (const char*)b);
    return fail(info, MAFE_WAV_ERR_VALUE);
  }
  info->big_endian = be;
  if (f.read(b, 4) != 4) { set_error("unpack requires a buffer of 4 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
  const int64_t file_size = (int64_t)rd_u32(b, be) + 8;
  got = f.read(b, 4);
  if (!(got == 4 && memcmp(b, "WAVE", 4) == 0)) {
    // the reference does `raise (f"Not a WAV file. ...")`, i.e. raises a str: Python answers with a TypeError
    set_error("exceptions must derive from BaseException");
    return fail(info, MAFE_WAV_ERR_TYPE);
  }

  bool have_fmt = false, have_data = false;
  while (f.pos < file_size) {
    uint8_t id[4];
    got = f.read(id, 4);
    if (got == 0) {
      if (have_data) { info->warnings |= MAFE_WAV_WARN_EOF; break; }    // io.py:682-693
      set_error("Unexpected end of file.");
      return fail(info, MAFE_WAV_ERR_VALUE);
    }
    if (got < 4) {                                                        // io.py:696-702
      if (have_fmt && have_data) { info->warnings |= MAFE_WAV_WARN_INCOMPLETE_ID; }
      else { set_error("Incomplete chunk ID: b'%.*s'", (int)got, (const char*)id); return fail(info, MAFE_WAV_ERR_VALUE); }
      memset(id + got, 0, (size_t)(4 - got));
    }
    if (got == 4 && memcmp(id, "fmt ", 4) == 0) {
      // ---- io.py:347-424
      have_fmt = true;
      if (f.read(b, 4) != 4) { set_error("unpack requires a buffer of 4 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
      const uint32_t chunk_size = rd_u32(b, be);
      if (chunk_size < 16) { set_error("Binary structure of wave file is not compliant"); return fail(info, MAFE_WAV_ERR_VALUE); }
      int64_t bytes_read = 16;
      if (f.read(b, 16) != 16) { set_error("unpack requires a buffer of 16 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
      uint32_t format_tag = rd_u16(b, be);
      info->channels = rd_u16(b + 2, be);
      info->sample_rate = (int32_t)rd_u32(b + 4, be);
      info->bytes_per_second = (int32_t)rd_u32(b + 8, be);
      info->block_align = rd_u16(b + 12, be);
      info->bit_depth = rd_u16(b + 14, be);
      if (format_tag == 0xFFFE && chunk_size >= (uint32_t)(bytes_read + 2)) {
        if (f.read(b, 2) != 2) { set_error("unpack requires a buffer of 2 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
        const uint16_t ext = rd_u16(b, be);
        bytes_read += 2;
        if (ext >= 22) {
          uint8_t x[22];
          const int64_t gx = f.read(x, 22);
          bytes_read += 22;
          static const uint8_t tail_le[12] = {0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71};
          static const uint8_t tail_be[12] = {0x00, 0x00, 0x00, 0x10, 0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71};
          if (gx == 22 && memcmp(x + 6 + 4, be ? tail_be : tail_le, 12) == 0) format_tag = rd_u32(x + 6, be);
        } else {
          set_error("Binary structure of wave file is not compliant");
          return fail(info, MAFE_WAV_ERR_VALUE);
        }
      }
      info->format_tag = (int32_t)format_tag;
      if (format_tag != 1 && format_tag != 3) {
        set_error("Unknown wave file format: %#06x. Supported formats: %s", format_tag, kSupported);
        return fail(info, MAFE_WAV_ERR_VALUE);
      }
      if ((int64_t)chunk_size > bytes_read) f.skip_read((int64_t)chunk_size - bytes_read);
      if (chunk_size & 1) f.seek_rel(1);
      if (format_tag == 1 && (int64_t)(uint32_t)info->bytes_per_second != (int64_t)(uint32_t)info->sample_rate * info->block_align) {
        set_error("WAV header is invalid: nAvgBytesPerSec must equal product of nSamplesPerSec and nBlockAlign, but file has "
                  "nSamplesPerSec = %u, nBlockAlign = %d, and nAvgBytesPerSec = %u",
                  (uint32_t)info->sample_rate, info->block_align, (uint32_t)info->bytes_per_second);
        return fail(info, MAFE_WAV_ERR_VALUE);
      }
    } else if (got == 4 && memcmp(id, "data", 4) == 0) {
      // ---- io.py:427-517
      have_data = true;
      if (!have_fmt) { set_error("No fmt chunk before data"); return fail(info, MAFE_WAV_ERR_VALUE); }
      if (f.read(b, 4) != 4) { set_error("unpack requires a buffer of 4 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
      const int64_t size = rd_u32(b, be);
      if (info->channels == 0) { set_error("integer division or modulo by zero"); return fail(info, MAFE_WAV_ERR_ZERODIV); }
      const int bps = info->block_align / info->channels;
      if (bps == 0) { set_error("integer division or modulo by zero"); return fail(info, MAFE_WAV_ERR_ZERODIV); }
      const int64_t n_samples = size / bps;
      int kind, item;
      bool raw = false;
      if (info->format_tag == 1) {
        if (info->bit_depth >= 1 && info->bit_depth <= 8) { kind = MAFE_WAV_U8; item = 1; }
        else if (bps == 3 || bps == 5 || bps == 6 || bps == 7) { kind = bps == 3 ? MAFE_WAV_I24 : MAFE_WAV_I40 + (bps - 5); item = 1; raw = true; }
        else if (info->bit_depth <= 64) {
          if (bps == 1) kind = MAFE_WAV_I8; else if (bps == 2) kind = MAFE_WAV_I16; else if (bps == 4) kind = MAFE_WAV_I32;
          else if (bps == 8) kind = MAFE_WAV_I64;
          else { set_error("data type '%si%d' not understood", be ? ">" : "<", bps); return fail(info, MAFE_WAV_ERR_TYPE); }
          item = bps;
        } else {
          set_error("Unsupported bit depth: the WAV file has %d-bit integer data.", info->bit_depth);
          return fail(info, MAFE_WAV_ERR_VALUE);
        }
      } else {
        if (info->bit_depth == 32 || info->bit_depth == 64) {
          if (bps == 4) kind = MAFE_WAV_F32; else if (bps == 8) kind = MAFE_WAV_F64;
          else { set_error("data type '%sf%d' not supported", be ? ">" : "<", bps); return fail(info, MAFE_WAV_ERR_TYPE); }
          item = bps;
        } else {
          set_error("Unsupported bit depth: the WAV file has %d-bit floating-point data.", info->bit_depth);
          return fail(info, MAFE_WAV_ERR_VALUE);
        }
      }
      info->sample_kind = kind;
      info->bytes_per_sample = bps;
      info->data_chunk_bytes = size;
      int64_t ignore = 0;
      if (offset_s > 0) {
        const double want = offset_s * (double)(uint32_t)info->sample_rate;
        ignore = want < 4.0e18 ? (int64_t)want : (int64_t)4000000000000000000LL;   // int() truncates; beyond any file anyway
        f.skip_read(ignore);                                                  // the reference reads `ignore` BYTES
      }
      const int64_t start = f.pos;
      int64_t items;
      if (!filelike) {
        int64_t count = raw ? size : n_samples;
        if (ignore <= count) count -= ignore;
        if (duration_s != 0.0 && duration_s * (double)(uint32_t)info->sample_rate < (double)count) {
          // a negative duration makes the reference's count negative and np.fromfile(count < 0) reads to the end of the
          // file (io.py:500-503): keep the full count
          if (duration_s > 0) count = (int64_t)(duration_s * (double)(uint32_t)info->sample_rate);
        }
        const int64_t avail = start < f.n ? (f.n - start) / item : 0;
        items = count < avail ? count : avail;                               // np.fromfile stops at EOF
        f.pos = start + items * item;
      } else {
        // no C-level file descriptor (BytesIO): the reference falls back to read(size) and ignores `duration`
        const int64_t avail = start < f.n ? f.n - start : 0;
        const int64_t nb = size < avail ? size : avail;
        if (nb % item) { set_error("buffer size must be a multiple of element size"); return fail(info, MAFE_WAV_ERR_VALUE); }
        items = nb / item;
        f.pos = start + nb;
      }
      if (raw) {
        if (items % bps) {
          set_error("cannot reshape array of size %lld into shape (%d)", (long long)items, bps);
          return fail(info, MAFE_WAV_ERR_VALUE);
        }
        items /= bps;
      }
      info->data_offset = start;
      info->n_items = items;
      if (size & 1) f.seek_rel(1);
      if (info->channels > 1 && items % info->channels) {
        set_error("cannot reshape array of size %lld into shape (%d)", (long long)items, info->channels);
        return fail(info, MAFE_WAV_ERR_VALUE);
      }
    } else if (got == 4 && (memcmp(id, "fact", 4) == 0 || memcmp(id, "LIST", 4) == 0 || memcmp(id, "JUNK", 4) == 0 ||
                            memcmp(id, "Fake", 4) == 0)) {
      if (!skip_unknown_chunk(f, be)) { set_error("unpack requires a buffer of 4 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
    } else {
      info->warnings |= MAFE_WAV_WARN_UNKNOWN_CHUNK;
      if (!skip_unknown_chunk(f, be)) { set_error("unpack requires a buffer of 4 bytes"); return fail(info, MAFE_WAV_ERR_STRUCT); }
    }
  }
  if (!have_data) {
    set_error("cannot access local variable 'audio' where it is not associated with a value");
    return fail(info, MAFE_WAV_ERR_UNBOUND);
  }
  return MAFE_OK;
}

// ---- batch staging: many files -> one contiguous payload buffer (pinned by the caller), parsed and copied by a few
// host threads.  This is the host half of load_batch: Python hands over the file contents, the library finds the
// payloads and packs them so that one cudaMemcpyAsync moves the whole batch.
namespace mafe {
namespace {
inline int64_t payload_bytes(const mafe_wav_info& i) {
  static const int kItem[] = {0, 1, 1, 2, 3, 4, 5, 6, 7, 8, 4, 8};
  return i.n_items * kItem[i.sample_kind];
}
}  // namespace
}  // namespace mafe

extern "C" int mafe_wav_stage(const void* const* blobs, const int64_t* blob_bytes, int32_t n_files, int32_t n_threads,
                              mafe_wav_info* infos, int64_t* payload_offsets, void* stage, int64_t stage_bytes,
                              int32_t* failed_index) {
  MAFE_REQUIRE(n_files >= 0, "mafe_wav_stage: negative file count");
  MAFE_REQUIRE(n_files == 0 || (blobs && blob_bytes && infos), "mafe_wav_stage: NULL argument");
  MAFE_REQUIRE(payload_offsets != nullptr, "mafe_wav_stage: payload_offsets is NULL");
  if (failed_index) *failed_index = -1;
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  int nt = n_threads > 0 ? n_threads : (hw < 16 ? hw : 16);
  if (nt > n_files) nt = n_files > 0 ? n_files : 1;

  // pass 1: walk every container
  std::atomic<int32_t> next{0};
  std::vector<std::string> msgs((size_t)nt);
  std::vector<int32_t> bad_of((size_t)nt, INT32_MAX);
  auto walk = [&](int t) {
    for (;;) {
      const int32_t k = next.fetch_add(1);
      if (k >= n_files) break;
      const int rc = mafe_wav_parse(blobs[k], blob_bytes[k], 0.0, 0.0, 0, &infos[k]);
      if (rc != MAFE_OK && k < bad_of[(size_t)t]) { bad_of[(size_t)t] = k; msgs[(size_t)t] = mafe_last_error(); }
    }
  };
  {
    // a container walk takes a few microseconds: more than one thread per 64 files costs more than it saves
    const int nw = std::min(nt, n_files / 64 + 1);
    std::vector<std::thread> pool;
    for (int t = 1; t < nw; ++t) pool.emplace_back(walk, t);
    walk(0);
    for (auto& th : pool) th.join();
  }
  int32_t bad = INT32_MAX, bad_t = -1;
  for (int t = 0; t < nt; ++t) if (bad_of[(size_t)t] < bad) { bad = bad_of[(size_t)t]; bad_t = t; }
  if (bad_t >= 0) {   // the first failing file, with the exception class of the reference in infos[bad].error_kind
    if (failed_index) *failed_index = bad;
    set_error("%s", msgs[(size_t)bad_t].c_str());
    return MAFE_E_INVALID_ARG;
  }
  payload_offsets[0] = 0;
  for (int32_t k = 0; k < n_files; ++k) payload_offsets[k + 1] = payload_offsets[k] + payload_bytes(infos[k]);
  if (!stage) return MAFE_OK;
  MAFE_REQUIRE(stage_bytes >= payload_offsets[n_files], "mafe_wav_stage: staging buffer of %lld bytes, %lld needed",
               (long long)stage_bytes, (long long)payload_offsets[n_files]);

  // pass 2: pack the payloads
  next.store(0);
  auto pack = [&]() {
    for (;;) {
      const int32_t k = next.fetch_add(1);
      if (k >= n_files) break;
      const int64_t nb = payload_offsets[k + 1] - payload_offsets[k];
      if (nb > 0) memcpy((char*)stage + payload_offsets[k], (const char*)blobs[k] + infos[k].data_offset, (size_t)nb);
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(pack);
    pack();
    for (auto& th : pool) th.join();
  }
  return MAFE_OK;
}

// ---- the same, from paths: the files are mapped (not read) by the host threads, so the only copy of a payload is the
// one from the page cache into the caller's pinned staging buffer.
struct mafe_wav_files {
  struct Map { void* p = nullptr; size_t n = 0; };
  std::vector<Map> maps;
  std::vector<const void*> ptrs;
  std::vector<int64_t> sizes;
  std::vector<mafe_wav_info> infos;
  std::vector<int64_t> offsets;
  std::vector<std::string> paths;
  struct Ident { dev_t dev = 0; ino_t ino = 0; off_t size = 0; struct timespec mtime = {0, 0}; };
  std::vector<Ident> idents;   // identity of every file at walk time: the pack refuses a file replaced or rewritten since
  int n_threads = 0;
  ~mafe_wav_files() {
    for (auto& m : maps) if (m.p && m.n) munmap(m.p, m.n);
  }
};

extern "C" int mafe_wav_files_open(const char* const* paths, int32_t n_files, int32_t n_threads, mafe_wav_files** out,
                                   mafe_wav_info* infos, int64_t* payload_offsets, int32_t* failed_index) {
  MAFE_REQUIRE(out != nullptr, "mafe_wav_files_open: out is NULL");
  *out = nullptr;
  MAFE_REQUIRE(n_files >= 0 && (n_files == 0 || (paths && infos)) && payload_offsets, "mafe_wav_files_open: bad argument");
  if (failed_index) *failed_index = -1;
  mafe_wav_files* h = new (std::nothrow) mafe_wav_files();
  if (!h) { set_error("out of host memory"); return MAFE_E_OOM; }
  h->maps.resize((size_t)n_files);
  h->ptrs.assign((size_t)n_files, nullptr);
  h->sizes.assign((size_t)n_files, 0);
  h->idents.resize((size_t)n_files);
  h->n_threads = n_threads;
  for (int32_t k = 0; k < n_files; ++k) h->paths.emplace_back(paths[k]);
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  int nt = n_threads > 0 ? n_threads : (hw < 16 ? hw : 16);
  if (nt > n_files) nt = n_files > 0 ? n_files : 1;
  std::atomic<int32_t> next{0};
  std::vector<int32_t> bad_of((size_t)nt, INT32_MAX);
  std::vector<std::string> msgs((size_t)nt);
  static const char kEmpty = 0;
  auto map_all = [&](int t) {
    for (;;) {
      const int32_t k = next.fetch_add(1);
      if (k >= n_files) break;
      const int fd = open(paths[k], O_RDONLY | O_CLOEXEC);
      struct stat st;
      if (fd < 0 || fstat(fd, &st) != 0) {
        if (k < bad_of[(size_t)t]) { bad_of[(size_t)t] = k; msgs[(size_t)t] = std::string(paths[k]) + ": " + strerror(errno); }
        if (fd >= 0) close(fd);
        continue;
      }
      if (st.st_size > 0) {
        void* p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (p == MAP_FAILED) {
          if (k < bad_of[(size_t)t]) { bad_of[(size_t)t] = k; msgs[(size_t)t] = std::string(paths[k]) + ": mmap: " + strerror(errno); }
          close(fd);
          continue;
        }
        h->maps[(size_t)k].p = p;
        h->maps[(size_t)k].n = (size_t)st.st_size;
        h->ptrs[(size_t)k] = p;
      } else {
        h->ptrs[(size_t)k] = &kEmpty;
      }
      h->sizes[(size_t)k] = (int64_t)st.st_size;
      h->idents[(size_t)k].dev = st.st_dev; h->idents[(size_t)k].ino = st.st_ino;
      h->idents[(size_t)k].size = st.st_size; h->idents[(size_t)k].mtime = st.st_mtim;
      close(fd);
    }
  };
  {
    const int nw = std::min(nt, n_files / 64 + 1);   // open + fstat + mmap: a few microseconds per file
    std::vector<std::thread> pool;
    for (int t = 1; t < nw; ++t) pool.emplace_back(map_all, t);
    map_all(0);
    for (auto& th : pool) th.join();
  }
  int32_t bad = INT32_MAX, bad_t = -1;
  for (int t = 0; t < nt; ++t) if (bad_of[(size_t)t] < bad) { bad = bad_of[(size_t)t]; bad_t = t; }
  if (bad_t >= 0) {
    if (failed_index) *failed_index = bad;
    if (infos) { memset(&infos[bad], 0, sizeof(mafe_wav_info)); infos[bad].error_kind = MAFE_WAV_ERR_OS; }
    set_error("%s", msgs[(size_t)bad_t].c_str());
    delete h;
    return MAFE_E_INVALID_ARG;
  }
  const int rc = mafe_wav_stage(h->ptrs.data(), h->sizes.data(), n_files, n_threads, infos, payload_offsets, nullptr, 0, failed_index);
  if (rc != MAFE_OK) { delete h; return rc; }
  h->infos.assign(infos, infos + n_files);
  h->offsets.assign(payload_offsets, payload_offsets + n_files + 1);
  *out = h;
  return MAFE_OK;
}

extern "C" int mafe_wav_files_pack(mafe_wav_files* h, void* stage, int64_t stage_bytes) {
  MAFE_REQUIRE(h != nullptr && stage != nullptr, "mafe_wav_files_pack: NULL argument");
  const int32_t n = (int32_t)h->ptrs.size();
  MAFE_REQUIRE(stage_bytes >= h->offsets[(size_t)n], "mafe_wav_files_pack: staging buffer of %lld bytes, %lld needed",
               (long long)stage_bytes, (long long)h->offsets[(size_t)n]);
  // pread straight into the staging buffer: one kernel copy from the page cache, no page faults on a mapping
  int hw = (int)std::thread::hardware_concurrency();
  if (hw < 1) hw = 1;
  int nt = h->n_threads > 0 ? h->n_threads : (hw < 16 ? hw : 16);
  if (nt > n) nt = n > 0 ? n : 1;
  std::atomic<int32_t> next{0}, bad{INT32_MAX};
  auto pack = [&]() {
    for (;;) {
      const int32_t k = next.fetch_add(1);
      if (k >= n) break;
      int64_t nb = h->offsets[(size_t)k + 1] - h->offsets[(size_t)k];
      if (nb <= 0) continue;
      char* dst = (char*)stage + h->offsets[(size_t)k];
      const int fd = open(h->paths[(size_t)k].c_str(), O_RDONLY | O_CLOEXEC);
      int64_t pos = h->infos[(size_t)k].data_offset;
      bool ok = fd >= 0;
      if (ok) {   // the same file as the one that was walked? (a path can be replaced or rewritten between open and pack)
        struct stat st;
        const mafe_wav_files::Ident& id = h->idents[(size_t)k];
        ok = fstat(fd, &st) == 0 && st.st_dev == id.dev && st.st_ino == id.ino && st.st_size == id.size &&
             st.st_mtim.tv_sec == id.mtime.tv_sec && st.st_mtim.tv_nsec == id.mtime.tv_nsec;
      }
      while (ok && nb > 0) {
        const ssize_t got = pread(fd, dst, (size_t)nb, (off_t)pos);
        if (got <= 0) { if (got < 0 && errno == EINTR) continue; ok = false; break; }
        dst += got; pos += got; nb -= got;
      }
      if (fd >= 0) close(fd);
      if (!ok) {   // the file changed under us since it was walked
        int32_t cur = bad.load();
        while (k < cur && !bad.compare_exchange_weak(cur, k)) {}
      }
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(pack);
    pack();
    for (auto& th : pool) th.join();
  }
  if (bad.load() != INT32_MAX) {
    set_error("%s: short read while packing the payload", h->paths[(size_t)bad.load()].c_str());
    return MAFE_E_INVALID_ARG;
  }
  return MAFE_OK;
}

extern "C" int mafe_wav_files_close(mafe_wav_files* h) {
  delete h;
  return MAFE_OK;
}

// ---- payload decode: one thread per item, bytes assembled by endianness, left-justified like the reference's numpy view
namespace mafe {

template <typename OUT>
__global__ void __launch_bounds__(256) wav_decode_kernel(const uint8_t* __restrict__ payload, int64_t n_items, int kind, int bps,
                                                        int big_endian, double scale, OUT* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
    const uint8_t* s = payload + i * bps;
    uint64_t u = 0;
    if (big_endian) for (int k = 0; k < bps; ++k) u = u << 8 | s[k];
    else for (int k = bps - 1; k >= 0; --k) u = u << 8 | s[k];
    double v;
    switch (kind) {
      case MAFE_WAV_U8: v = (double)(uint8_t)u; break;
      case MAFE_WAV_I8: v = (double)(int8_t)u; break;
      case MAFE_WAV_I16: v = (double)(int16_t)u * (1.0 / 32768.0); break;                     // io.py:745-746
      case MAFE_WAV_I24: v = (double)(int32_t)((uint32_t)u << 8) * (1.0 / 2147483648.0); break;   // left-justified int32, io.py:743-744
      case MAFE_WAV_I32: v = (double)(int32_t)u * (1.0 / 2147483648.0); break;
      case MAFE_WAV_I40: case MAFE_WAV_I48: case MAFE_WAV_I56: v = (double)(int64_t)(u << (8 * (8 - bps))); break;
      case MAFE_WAV_I64: v = (double)(int64_t)u; break;
      case MAFE_WAV_F32: v = (double)__uint_as_float((uint32_t)u); break;
      default: v = __longlong_as_double((long long)u); break;
    }
    out[i] = (OUT)(v * scale);
  }
}

// PCM16 -> raw int16 in host byte order (RIFX files for the front-end's MAFE_WAVE_I16 input)
__global__ void __launch_bounds__(256) wav_i16_kernel(const uint8_t* __restrict__ payload, int64_t n_items, int big_endian,
                                                     int16_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
    const uint8_t a = payload[2 * i], b = payload[2 * i + 1];
    out[i] = (int16_t)(big_endian ? (a << 8 | b) : (b << 8 | a));
  }
}

}  // namespace mafe

extern "C" int mafe_wav_decode(mafe_ctx* ctx, const void* payload_dev, int64_t n_items, int32_t sample_kind, int32_t big_endian,
                               int32_t out_dtype, double scale, void* out_dev) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(n_items >= 0, "mafe_wav_decode: negative item count");
  static const int kBytes[] = {0, 1, 1, 2, 3, 4, 5, 6, 7, 8, 4, 8};
  MAFE_REQUIRE(sample_kind >= MAFE_WAV_U8 && sample_kind <= MAFE_WAV_F64, "mafe_wav_decode: unknown sample kind %d", sample_kind);
  MAFE_REQUIRE(out_dtype == MAFE_WAV_OUT_F32 || out_dtype == MAFE_WAV_OUT_F64 || out_dtype == MAFE_WAV_OUT_I16,
               "mafe_wav_decode: unknown output type %d", out_dtype);
  MAFE_REQUIRE(out_dtype != MAFE_WAV_OUT_I16 || sample_kind == MAFE_WAV_I16, "mafe_wav_decode: raw int16 output needs PCM16 input");
  if (n_items == 0) return MAFE_OK;
  MAFE_REQUIRE(payload_dev && out_dev, "mafe_wav_decode: NULL buffer");
  cudaSetDevice(ctx->device);
  const int bps = kBytes[sample_kind];
  const int64_t want = (n_items + 255) / 256;
  const unsigned grid = (unsigned)(want < (int64_t)ctx->sm_count * 32 ? want : (int64_t)ctx->sm_count * 32);
  const uint8_t* p = (const uint8_t*)payload_dev;
  if (out_dtype == MAFE_WAV_OUT_I16) wav_i16_kernel<<<grid, 256, 0, ctx->stream>>>(p, n_items, big_endian, (int16_t*)out_dev);
  else if (out_dtype == MAFE_WAV_OUT_F32)
    wav_decode_kernel<float><<<grid, 256, 0, ctx->stream>>>(p, n_items, sample_kind, bps, big_endian, scale, (float*)out_dev);
  else
    wav_decode_kernel<double><<<grid, 256, 0, ctx->stream>>>(p, n_items, sample_kind, bps, big_endian, scale, (double*)out_dev);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}
