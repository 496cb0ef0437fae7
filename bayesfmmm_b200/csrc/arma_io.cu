// arma_io.cu -- writers/readers of the reference's stored-sample files (include/bfmmm_io.h).
// Host-only code (built by nvcc with the rest of the library).  Format notes (appendix B of
// SURVEY.md, checked byte-for-byte against inst/test-data/Functional_trace in tests/test_io.py):
//   text:   "ARMA_MAT_TXT_FN008\n<rows> <cols>\n" / "ARMA_CUB_TXT_FN008\n<rows> <cols> <slices>\n", then one
//           matrix row per line, every value written as ' ' followed by "%24.16e", slices back to back
//   binary: "ARMA_FLD_BIN\n<n_rows>\n<n_cols>\n", then per element "ARMA_CUB_BIN_FN008\n<r> <c> <s>\n" + raw
//           little-endian column-major doubles
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bfmmm_io.h"
#include "common.cuh"

namespace {
int iofail(const std::string& m) { return bf::set_error(m.c_str()); }

void put_value(std::string& line, double v) {
  char buf[64];
  if (std::isnan(v)) std::snprintf(buf, sizeof buf, " %24s", "nan");
  else if (std::isinf(v)) std::snprintf(buf, sizeof buf, " %24s", v > 0 ? "inf" : "-inf");
  else std::snprintf(buf, sizeof buf, " %24.16e", v);
  line += buf;
}

int write_txt(const char* path, const char* head, const double* d, int64_t r, int64_t c, int64_t s, bool cube) {
  FILE* f = std::fopen(path, "wb");
  if (!f) return iofail(std::string("cannot open for writing: ") + path);
  if (cube) std::fprintf(f, "%s\n%lld %lld %lld\n", head, (long long)r, (long long)c, (long long)s);
  else std::fprintf(f, "%s\n%lld %lld\n", head, (long long)r, (long long)c);
  std::string line;
  for (int64_t k = 0; k < s; k++)
    for (int64_t i = 0; i < r; i++) {
      line.clear();
      for (int64_t j = 0; j < c; j++) put_value(line, d[(size_t)k * r * c + (size_t)j * r + i]);
      line += '\n';
      std::fwrite(line.data(), 1, line.size(), f);
    }
  std::fclose(f);
  return 0;
}

bool read_line(FILE* f, std::string& out) {
  out.clear();
  int ch;
  while ((ch = std::fgetc(f)) != EOF) {
    if (ch == '\n') return true;
    out += (char)ch;
  }
  return !out.empty();
}

struct Info { int kind = 0; long long r = 0, c = 0, s = 1, fr = 1, fc = 1; };

int parse_elem_header(FILE* f, Info& in, bool in_field) {
  std::string h, d;
  if (!read_line(f, h) || !read_line(f, d)) return 1;
  if (h == "ARMA_MAT_TXT_FN008" || h == "ARMA_MAT_BIN_FN008") {
    if (std::sscanf(d.c_str(), "%lld %lld", &in.r, &in.c) != 2) return 1;
    in.s = 1;
    if (!in_field) in.kind = h == "ARMA_MAT_TXT_FN008" ? BFMMM_FILE_MAT_TXT : BFMMM_FILE_MAT_BIN;
    else in.kind = BFMMM_FILE_FIELD_MAT_BIN;
    return 0;
  }
  if (h == "ARMA_CUB_TXT_FN008" || h == "ARMA_CUB_BIN_FN008") {
    if (std::sscanf(d.c_str(), "%lld %lld %lld", &in.r, &in.c, &in.s) != 3) return 1;
    if (!in_field) in.kind = h == "ARMA_CUB_TXT_FN008" ? BFMMM_FILE_CUBE_TXT : BFMMM_FILE_CUBE_BIN;
    else in.kind = BFMMM_FILE_FIELD_CUBE_BIN;
    return 0;
  }
  return 1;
}

int read_info(FILE* f, Info& in) {
  long pos = std::ftell(f);
  std::string h;
  if (!read_line(f, h)) return 1;
  if (h == "ARMA_FLD_BIN") {
    std::string a, b;
    if (!read_line(f, a) || !read_line(f, b)) return 1;
    in.fr = std::atoll(a.c_str()); in.fc = std::atoll(b.c_str());
    long p2 = std::ftell(f);
    if (parse_elem_header(f, in, true)) return 1;
    std::fseek(f, p2, SEEK_SET);
    return 0;
  }
  std::fseek(f, pos, SEEK_SET);
  int rc = parse_elem_header(f, in, false);
  std::fseek(f, pos, SEEK_SET);          // leave the stream at the element header
  return rc;
}
}  // namespace

extern "C" {

int bfmmm_save_mat_txt(const char* path, const double* data, int64_t r, int64_t c) {
  return write_txt(path, "ARMA_MAT_TXT_FN008", data, r, c, 1, false);
}
int bfmmm_save_cube_txt(const char* path, const double* data, int64_t r, int64_t c, int64_t s) {
  return write_txt(path, "ARMA_CUB_TXT_FN008", data, r, c, s, true);
}
int bfmmm_save_field_cube_bin(const char* path, const double* data, int64_t fr, int64_t fc, int64_t r, int64_t c,
                              int64_t s) {
  FILE* f = std::fopen(path, "wb");
  if (!f) return iofail(std::string("cannot open for writing: ") + path);
  std::fprintf(f, "ARMA_FLD_BIN\n%lld\n%lld\n", (long long)fr, (long long)fc);
  const size_t per = (size_t)r * c * s;
  for (int64_t e = 0; e < fr * fc; e++) {
    std::fprintf(f, "ARMA_CUB_BIN_FN008\n%lld %lld %lld\n", (long long)r, (long long)c, (long long)s);
    std::fwrite(data + (size_t)e * per, sizeof(double), per, f);
  }
  std::fclose(f);
  return 0;
}

int bfmmm_file_info(const char* path, int32_t* kind, int64_t* dims) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return iofail(std::string("cannot open: ") + path);
  Info in;
  int rc = read_info(f, in);
  std::fclose(f);
  if (rc) return iofail(std::string("not an Armadillo mat/cube/field file: ") + path);
  *kind = in.kind;
  dims[0] = in.r; dims[1] = in.c; dims[2] = in.s; dims[3] = in.fr; dims[4] = in.fc;
  return 0;
}

int bfmmm_load(const char* path, double* out, int64_t capacity) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return iofail(std::string("cannot open: ") + path);
  Info in;
  if (read_info(f, in)) { std::fclose(f); return iofail(std::string("not an Armadillo mat/cube/field file: ") + path); }
  const size_t per = (size_t)in.r * in.c * in.s;
  const size_t total = per * (size_t)(in.fr * in.fc);
  if ((int64_t)total > capacity) { std::fclose(f); return iofail("bfmmm_load: output buffer too small"); }
  if (in.kind == BFMMM_FILE_MAT_TXT || in.kind == BFMMM_FILE_CUBE_TXT) {
    Info dummy; parse_elem_header(f, dummy, false);
    for (long long k = 0; k < in.s; k++)
      for (long long i = 0; i < in.r; i++)
        for (long long j = 0; j < in.c; j++) {
          char tok[64];
          if (std::fscanf(f, "%63s", tok) != 1) { std::fclose(f); return iofail(std::string("truncated file: ") + path); }
          out[(size_t)k * in.r * in.c + (size_t)j * in.r + i] = std::strtod(tok, nullptr);
        }
  } else if (in.kind == BFMMM_FILE_MAT_BIN || in.kind == BFMMM_FILE_CUBE_BIN) {
    Info dummy; parse_elem_header(f, dummy, false);
    if (std::fread(out, sizeof(double), per, f) != per) { std::fclose(f); return iofail(std::string("truncated file: ") + path); }
  } else {
    for (long long e = 0; e < in.fr * in.fc; e++) {
      Info el;
      if (parse_elem_header(f, el, true) || (size_t)el.r * el.c * el.s != per) { std::fclose(f); return iofail("field elements of unequal size"); }
      if (std::fread(out + (size_t)e * per, sizeof(double), per, f) != per) { std::fclose(f); return iofail(std::string("truncated file: ") + path); }
    }
  }
  std::fclose(f);
  return 0;
}

}  // extern "C"
