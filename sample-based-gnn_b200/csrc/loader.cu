// loader.cu -- feature / label / mask ingestion (host side; no kernels).
//
// Replaces GNNDatum::readFeature_Label_Mask (core/ntsDataloador.hpp:999-1063 of the reference): three text files read token by
// token with operator>> on one thread -- the slowest part of the reference's start-up (cora: 2708 x 1433 floats as text). Here the
// files are mapped, cut at line boundaries, and parsed by all host threads with strtof / strtol (the conversions operator>> itself
// ends in, so every value is bit-identical); the parsed table is also written next to the text file as a raw binary cache
// (<feature_file>.nb_f32: "NBF1", |V|, F, then |V|*F floats in vertex order) that later runs read with one read().
// Line formats (as the reference reads them): feature "id v0 v1 ... v{F-1}", label "id label", mask "id train|eval|val|test|...".
#include <errno.h>
#include <fcntl.h>
#include <omp.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace {
struct Mapped {
  const char *p = nullptr;
  size_t n = 0;
  int fd = -1;
  bool open(const char *path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) return false;
    n = (size_t)st.st_size;
    if (n == 0) { p = ""; return true; }
    void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) return false;
    p = (const char *)m;
    madvise(m, n, MADV_SEQUENTIAL);
    return true;
  }
  ~Mapped() {
    if (p && n) munmap((void *)p, n);
    if (fd >= 0) close(fd);
  }
};
// [begin, end) of every non-empty line
std::vector<std::pair<size_t, size_t>> lines_of(const Mapped &m) {
  std::vector<std::pair<size_t, size_t>> out;
  size_t a = 0;
  while (a < m.n) {
    const void *nl = memchr(m.p + a, '\n', m.n - a);
    const size_t b = nl ? (size_t)((const char *)nl - m.p) : m.n;
    size_t s = a;
    while (s < b && (m.p[s] == ' ' || m.p[s] == '\t' || m.p[s] == '\r')) s++;
    if (s < b) out.push_back(std::make_pair(s, b));
    a = b + 1;
  }
  return out;
}
const uint32_t CACHE_MAGIC = 0x3146424eu;  // "NBF1"
}  // namespace

extern "C" {

int nb_read_feature_table(const char *path, uint32_t n_vertices, uint32_t feature_size, uint32_t id_begin, uint32_t id_end, float *out,
                          int use_binary_cache, int *from_cache_out) {
  NB_REQUIRE(path && out && feature_size > 0 && id_begin <= id_end && id_end <= n_vertices, NB_ERR_ARG, "nb_read_feature_table: bad argument");
  if (from_cache_out) *from_cache_out = 0;
  const std::string cache = std::string(path) + ".nb_f32";
  const size_t rows = (size_t)id_end - id_begin, F = feature_size;
  if (use_binary_cache) {   // a cache written by an earlier run, at least as new as the text file
    struct stat st_t, st_c;
    if (stat(path, &st_t) == 0 && stat(cache.c_str(), &st_c) == 0 && st_c.st_mtime >= st_t.st_mtime &&
        (size_t)st_c.st_size == 12 + (size_t)n_vertices * F * 4) {
      int fd = ::open(cache.c_str(), O_RDONLY);
      uint32_t hdr[3] = {0, 0, 0};
      if (fd >= 0 && pread(fd, hdr, 12, 0) == 12 && hdr[0] == CACHE_MAGIC && hdr[1] == n_vertices && hdr[2] == feature_size) {
        size_t done = 0, want = rows * F * 4;
        const off_t base = 12 + (off_t)id_begin * F * 4;
        while (done < want) {
          ssize_t r = pread(fd, (char *)out + done, want - done, base + (off_t)done);
          if (r <= 0) break;
          done += (size_t)r;
        }
        close(fd);
        if (done == want) { if (from_cache_out) *from_cache_out = 1; return NB_OK; }
      } else if (fd >= 0) close(fd);
    }
  }
  Mapped m;
  NB_REQUIRE(m.open(path), NB_ERR_ARG, "nb_read_feature_table: cannot open %s (%s)", path, strerror(errno));
  std::vector<std::pair<size_t, size_t>> lines = lines_of(m);
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : bad)
  for (long i = 0; i < (long)lines.size(); i++) {
    // a line is copied into a NUL-terminated buffer: strtof must not run past its end into the next line
    std::string buf(m.p + lines[i].first, lines[i].second - lines[i].first);
    char *q = &buf[0], *e = nullptr;
    const unsigned long id = strtoul(q, &e, 10);
    if (e == q) { bad++; continue; }
    if (id < id_begin || id >= id_end) continue;
    float *row = out + (size_t)(id - id_begin) * F;
    q = e;
    for (size_t k = 0; k < F; k++) {
      row[k] = strtof(q, &e);
      if (e == q) { bad++; break; }
      q = e;
    }
  }
  NB_REQUIRE(bad == 0, NB_ERR_ARG, "nb_read_feature_table: %d malformed line(s) in %s", bad, path);
  if (use_binary_cache && id_begin == 0 && id_end == n_vertices) {   // best effort: a read-only directory just means no cache
    const std::string tmp = cache + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (f) {
      const uint32_t hdr[3] = {CACHE_MAGIC, n_vertices, feature_size};
      const bool ok = fwrite(hdr, 4, 3, f) == 3 && fwrite(out, 4, rows * F, f) == rows * F;
      fclose(f);
      if (ok) rename(tmp.c_str(), cache.c_str()); else unlink(tmp.c_str());
    }
  }
  return NB_OK;
}

/* labels: "id label" per line -> label_out[id - id_begin]; masks: "id word" -> 0 train, 1 eval / val, 2 test, 3 anything else */
int nb_read_label_mask(const char *label_path, const char *mask_path, uint32_t id_begin, uint32_t id_end, int64_t *label_out, int32_t *mask_out) {
  NB_REQUIRE(id_begin <= id_end, NB_ERR_ARG, "nb_read_label_mask: bad range");
  if (label_path && label_out) {
    Mapped m;
    NB_REQUIRE(m.open(label_path), NB_ERR_ARG, "nb_read_label_mask: cannot open %s (%s)", label_path, strerror(errno));
    std::vector<std::pair<size_t, size_t>> lines = lines_of(m);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)lines.size(); i++) {
      std::string buf(m.p + lines[i].first, lines[i].second - lines[i].first);
      char *e = nullptr;
      const unsigned long id = strtoul(buf.c_str(), &e, 10);
      if (e == buf.c_str() || id < id_begin || id >= id_end) continue;
      label_out[id - id_begin] = strtol(e, nullptr, 10);
    }
  }
  if (mask_path && mask_out) {
    Mapped m;
    NB_REQUIRE(m.open(mask_path), NB_ERR_ARG, "nb_read_label_mask: cannot open %s (%s)", mask_path, strerror(errno));
    std::vector<std::pair<size_t, size_t>> lines = lines_of(m);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)lines.size(); i++) {
      std::string buf(m.p + lines[i].first, lines[i].second - lines[i].first);
      char *e = nullptr;
      const unsigned long id = strtoul(buf.c_str(), &e, 10);
      if (e == buf.c_str() || id < id_begin || id >= id_end) continue;
      while (*e == ' ' || *e == '\t') e++;
      size_t len = 0;
      while (e[len] && e[len] != ' ' && e[len] != '\t' && e[len] != '\r') len++;
      const std::string w(e, len);
      mask_out[id - id_begin] = w == "train" ? 0 : (w == "eval" || w == "val") ? 1 : w == "test" ? 2 : 3;
    }
  }
  return NB_OK;
}

}  // extern "C"
