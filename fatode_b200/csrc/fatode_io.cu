// Batched structure-of-arrays container and the reference's text formats (host code only).
//
// The reference has no batch concept and only writes one system's result as text, one E24.16 number per line
// (EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_ros_adj_dr.F90:134-164: cbm4_ros_sol.txt, cbm4_output_sens.txt).  A batched engine needs
// a format for y0 / y(T) / Lambda / Mu / status arrays of many cells, on disk and on the wire:
//
//   "FATODEB1" container (little endian)
//     bytes 0..7    magic "FATODEB1"
//     u32 version (1), u32 model, u64 ncells, u32 nvar, u32 np, u32 nadj, u32 narrays, f64 tin, f64 tout      (56 bytes)
//     narrays directory entries of 48 bytes: char name[16]; u32 dtype (0 = f64, 1 = i32); u32 rank; u64 dims[2] (per cell);
//                                            u64 offset (from the start of the file, 64-byte aligned)
//     the arrays, each [ncells][dims...] exactly as the C ABI passes them (include/fatode_b200.h), so a section can be
//     handed to fatode_*_batch() — or mapped / streamed — without any re-layout
//   names: "y", "lambda", "mu", "istatus", "rstatus", "ierr" (any subset)
//
// The text routines write and read the reference's own files for one system, so existing post-processing keeps working.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/fatode_b200.h"

namespace {
struct DirEnt { char name[16]; uint32_t dtype, rank; uint64_t dims[2]; uint64_t offset; };
static_assert(sizeof(DirEnt) == 48, "directory entry layout");
struct Header { char magic[8]; uint32_t version, model; uint64_t ncells; uint32_t nvar, np, nadj, narrays; double tin, tout; };
static_assert(sizeof(Header) == 56, "header layout");

uint64_t ent_bytes(const DirEnt& e, uint64_t ncells) {
  uint64_t n = ncells;
  for (uint32_t r = 0; r < e.rank; ++r) n *= e.dims[r];
  return n * (e.dtype == 0 ? 8u : 4u);
}

// Fortran E24.16: "  0.dddddddddddddddddE+ee" (the mantissa is 0.d1..d16), right-justified in 24 columns
void e24_16(double v, char* out /* >= 32 */) {
  char t[40], body[40];
  if (v != v) { std::snprintf(out, 32, "%24s", "NaN"); return; }
  if (v - v != 0.0) { std::snprintf(out, 32, "%24s", v > 0 ? "Infinity" : "-Infinity"); return; }
  std::snprintf(t, sizeof t, "%.15E", v < 0 ? -v : v);           // d.dddddddddddddddE+XX (16 significant digits)
  char digits[20];
  digits[0] = t[0];
  std::memcpy(digits + 1, t + 2, 15);
  digits[16] = 0;
  int ex = std::atoi(std::strchr(t, 'E') + 1);
  if (v != 0.0) ex += 1;
  const bool neg = v < 0 || (v == 0.0 && 1.0 / v < 0);
  if (ex > 99 || ex < -99) std::snprintf(body, sizeof body, "%s0.%s%c%03d", neg ? "-" : "", digits, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
  else std::snprintf(body, sizeof body, "%s0.%sE%c%02d", neg ? "-" : "", digits, ex < 0 ? '-' : '+', ex < 0 ? -ex : ex);
  std::snprintf(out, 32, "%24s", body);
}

bool read_number(FILE* f, double* v) {
  char line[256];
  while (std::fgets(line, sizeof line, f)) {
    char* p = line;
    while (*p == ' ' || *p == '\t') ++p;
    if (*p == 0 || *p == '\n' || *p == '\r') continue;
    if ((*p >= 'A' && *p <= 'Z' && *p != 'E' && *p != 'I' && *p != 'N') || (*p >= 'a' && *p <= 'z')) continue;   // "Lambda: column  1"
    for (char* q = p; *q; ++q) if (*q == 'D' || *q == 'd') *q = 'E';
    char* end = nullptr;
    const double x = std::strtod(p, &end);
    if (end == p) continue;
    *v = x;
    return true;
  }
  return false;
}
}  // namespace

extern "C" int fatode_file_write(const char* path, int model, int64_t ncells, int nvar, int np, int nadj, double tin, double tout,
                                 const double* y, const double* lambda, const double* mu, const int* istatus,
                                 const double* rstatus, const int* ierr) {
  if (!path || ncells < 0 || nvar < 1) return FATODE_EINVAL;
  std::vector<DirEnt> dir;
  std::vector<const void*> src;
  auto add = [&](const char* name, const void* p, uint32_t dtype, uint32_t rank, uint64_t d0, uint64_t d1) {
    if (!p) return;
    DirEnt e;
    std::memset(&e, 0, sizeof e);
    std::strncpy(e.name, name, sizeof e.name - 1);
    e.dtype = dtype; e.rank = rank; e.dims[0] = d0; e.dims[1] = d1;
    dir.push_back(e);
    src.push_back(p);
  };
  add("y", y, 0, 1, (uint64_t)nvar, 1);
  if (nadj > 0) add("lambda", lambda, 0, 2, (uint64_t)nadj, (uint64_t)nvar);
  if (nadj > 0 && np > 0) add("mu", mu, 0, 2, (uint64_t)nadj, (uint64_t)np);
  add("istatus", istatus, 1, 1, 20, 1);
  add("rstatus", rstatus, 0, 1, 20, 1);
  add("ierr", ierr, 1, 0, 1, 1);
  Header h;
  std::memset(&h, 0, sizeof h);
  std::memcpy(h.magic, "FATODEB1", 8);
  h.version = 1; h.model = (uint32_t)model; h.ncells = (uint64_t)ncells; h.nvar = (uint32_t)nvar; h.np = (uint32_t)np;
  h.nadj = (uint32_t)nadj; h.narrays = (uint32_t)dir.size(); h.tin = tin; h.tout = tout;
  uint64_t off = sizeof(Header) + dir.size() * sizeof(DirEnt);
  for (DirEnt& e : dir) { off = (off + 63) / 64 * 64; e.offset = off; off += ent_bytes(e, h.ncells); }
  FILE* f = std::fopen(path, "wb");
  if (!f) return FATODE_EINVAL;
  bool ok = std::fwrite(&h, sizeof h, 1, f) == 1 && (dir.empty() || std::fwrite(dir.data(), sizeof(DirEnt), dir.size(), f) == dir.size());
  uint64_t pos = sizeof(Header) + dir.size() * sizeof(DirEnt);
  static const char zeros[64] = {0};
  for (size_t i = 0; ok && i < dir.size(); ++i) {
    if (dir[i].offset > pos) ok = std::fwrite(zeros, 1, dir[i].offset - pos, f) == dir[i].offset - pos;
    const uint64_t nb = ent_bytes(dir[i], h.ncells);
    if (ok && nb) ok = std::fwrite(src[i], 1, nb, f) == nb;
    pos = dir[i].offset + nb;
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? 0 : FATODE_EINVAL;
}

// header of a container: any output pointer may be NULL.  has[6] = presence of y, lambda, mu, istatus, rstatus, ierr
extern "C" int fatode_file_info(const char* path, int* model, int64_t* ncells, int* nvar, int* np, int* nadj, double* tin,
                                double* tout, int* has) {
  FILE* f = path ? std::fopen(path, "rb") : nullptr;
  if (!f) return FATODE_EINVAL;
  Header h;
  bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "FATODEB1", 8) == 0 && h.version == 1 && h.narrays <= 64;
  std::vector<DirEnt> dir(ok ? h.narrays : 0);
  if (ok && h.narrays) ok = std::fread(dir.data(), sizeof(DirEnt), dir.size(), f) == dir.size();
  std::fclose(f);
  if (!ok) return FATODE_EINVAL;
  if (model) *model = (int)h.model;
  if (ncells) *ncells = (int64_t)h.ncells;
  if (nvar) *nvar = (int)h.nvar;
  if (np) *np = (int)h.np;
  if (nadj) *nadj = (int)h.nadj;
  if (tin) *tin = h.tin;
  if (tout) *tout = h.tout;
  if (has) {
    static const char* names[6] = {"y", "lambda", "mu", "istatus", "rstatus", "ierr"};
    for (int k = 0; k < 6; ++k) {
      has[k] = 0;
      for (const DirEnt& e : dir) if (std::strncmp(e.name, names[k], 16) == 0) has[k] = 1;
    }
  }
  return 0;
}

// one array of a container, cells [first_cell, first_cell + ncells) of it, into dst (the caller sizes dst from fatode_file_info)
extern "C" int fatode_file_read(const char* path, const char* name, int64_t first_cell, int64_t ncells, void* dst) {
  FILE* f = (path && name && dst) ? std::fopen(path, "rb") : nullptr;
  if (!f) return FATODE_EINVAL;
  Header h;
  bool ok = std::fread(&h, sizeof h, 1, f) == 1 && std::memcmp(h.magic, "FATODEB1", 8) == 0 && h.version == 1 && h.narrays <= 64;
  std::vector<DirEnt> dir(ok ? h.narrays : 0);
  if (ok && h.narrays) ok = std::fread(dir.data(), sizeof(DirEnt), dir.size(), f) == dir.size();
  int rc = FATODE_EINVAL;
  if (ok && first_cell >= 0 && ncells >= 0 && (uint64_t)(first_cell + ncells) <= h.ncells)
    for (const DirEnt& e : dir)
      if (std::strncmp(e.name, name, 16) == 0) {
        const uint64_t per = ent_bytes(e, 1);
        if (std::fseek(f, (long)(e.offset + per * (uint64_t)first_cell), SEEK_SET) == 0 &&
            (ncells == 0 || std::fread(dst, per, (size_t)ncells, f) == (size_t)ncells))
          rc = 0;
        break;
      }
  std::fclose(f);
  return rc;
}

// ---- the reference's text files for ONE system (cbm4_ros_adj_dr.F90:134-164)
extern "C" int fatode_text_write_solution(const char* path, int nvar, const double* y) {
  FILE* f = (path && y) ? std::fopen(path, "w") : nullptr;
  if (!f) return FATODE_EINVAL;
  char b[32];
  for (int i = 0; i < nvar; ++i) { e24_16(y[i], b); std::fprintf(f, "%s\n", b); }
  return std::fclose(f) == 0 ? 0 : FATODE_EINVAL;
}
// lambda[nadj][nvar] (= Fortran Lambda(NVAR,NADJ)), mu[nadj][np] (may be NULL with np = 0)
extern "C" int fatode_text_write_sens(const char* path, int nvar, int np, int nadj, const double* lambda, const double* mu) {
  FILE* f = (path && lambda) ? std::fopen(path, "w") : nullptr;
  if (!f) return FATODE_EINVAL;
  char b[32];
  for (int i = 0; i < nadj; ++i) {
    std::fprintf(f, "Lambda: column%3d\n", i + 1);
    for (int j = 0; j < nvar; ++j) { e24_16(lambda[(size_t)i * nvar + j], b); std::fprintf(f, "%s\n", b); }
  }
  if (mu)
    for (int i = 0; i < nadj; ++i) {
      std::fprintf(f, "Mu: column%3d\n", i + 1);
      for (int j = 0; j < np; ++j) { e24_16(mu[(size_t)i * np + j], b); std::fprintf(f, "%s\n", b); }
    }
  return std::fclose(f) == 0 ? 0 : FATODE_EINVAL;
}
extern "C" int fatode_text_read_solution(const char* path, int nvar, double* y) {
  FILE* f = (path && y) ? std::fopen(path, "r") : nullptr;
  if (!f) return FATODE_EINVAL;
  int n = 0;
  while (n < nvar && read_number(f, y + n)) ++n;
  std::fclose(f);
  return n == nvar ? 0 : FATODE_EINVAL;
}
extern "C" int fatode_text_read_sens(const char* path, int nvar, int np, int nadj, double* lambda, double* mu) {
  FILE* f = (path && lambda) ? std::fopen(path, "r") : nullptr;
  if (!f) return FATODE_EINVAL;
  size_t n = 0;
  const size_t nl = (size_t)nadj * nvar, nm = mu ? (size_t)nadj * np : 0;
  while (n < nl && read_number(f, lambda + n)) ++n;
  size_t m = 0;
  while (m < nm && read_number(f, mu + m)) ++m;
  std::fclose(f);
  return (n == nl && m == nm) ? 0 : FATODE_EINVAL;
}
