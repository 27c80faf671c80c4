// libsvm_host.cu -- HOST-side libsvm tokenizer behind the C ABI (no CUDA in this file).
//
// Replaces decode_libsvm (DeepFM.py:65-81): tf.string_split(line, ' ') -> label = first token
// (string_to_number float32), every other token split on ':' into exactly (id, value), ids -> int32,
// values -> float32.  tf.string_split skips empty tokens, so runs of spaces are tolerated.
// The reference never checks the pair count; batching then needs every line to have the same number
// of pairs and the model reshapes to [-1, field_size] (DeepFM.py:120-122) -- here a line whose pair
// count differs from F is reported as an error with its line index.
#include <errno.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

using namespace ctr;

extern "C" {

// Parses complete lines from buf[0, len) until max_rows rows are produced or the buffer ends.
// Returns the number of rows written (>= 0) or a negative ctr_status; *consumed = bytes consumed
// (always ends just after a '\n', or == len when the buffer's last line has no newline and
// `final_chunk` != 0).  ids/vals: [max_rows, F] row-major; labels: [max_rows].
int64_t ctr_parse_libsvm(const char* buf, size_t len, int F, int64_t max_rows, int final_chunk, int32_t* ids,
                         float* vals, float* labels, size_t* consumed) {
  if (!buf || F <= 0 || max_rows < 0 || !ids || !vals || !labels || !consumed) {
    set_error("ctr_parse_libsvm: bad arguments");
    return CTR_ERR_INVALID_ARG;
  }
  size_t pos = 0;
  int64_t row = 0;
  while (row < max_rows && pos < len) {
    const char* nl = static_cast<const char*>(memchr(buf + pos, '\n', len - pos));
    size_t end;
    if (nl) end = (size_t)(nl - buf);
    else if (final_chunk) end = len;
    else break;  // incomplete last line: leave it for the next chunk
    const char* p = buf + pos;
    const char* e = buf + end;
    if (e > p && e[-1] == '\r') --e;
    while (p < e && *p == ' ') ++p;
    if (p == e) {  // blank line: TextLineDataset yields '', decode would fail; skip silently
      pos = nl ? end + 1 : end;
      continue;
    }
    char* q = nullptr;
    labels[row] = strtof(p, &q);
    if (q == p) {
      set_error("ctr_parse_libsvm: row %lld: label is not a number", (long long)row);
      return CTR_ERR_INVALID_ARG;
    }
    p = q;
    int f = 0;
    while (true) {
      while (p < e && *p == ' ') ++p;
      if (p >= e) break;
      if (f >= F) { f = F + 1; break; }
      long id = strtol(p, &q, 10);
      if (q == p || q >= e || *q != ':') {
        set_error("ctr_parse_libsvm: row %lld field %d: expected <id>:<val>", (long long)row, f);
        return CTR_ERR_INVALID_ARG;
      }
      if (id < (long)INT32_MIN || id > (long)INT32_MAX) {   // tf.string_to_number(out_type=int32) raises (DeepFM.py:74)
        set_error("ctr_parse_libsvm: row %lld field %d: id %ld is outside the int32 range", (long long)row, f, id);
        return CTR_ERR_INVALID_ARG;
      }
      p = q + 1;
      // strtof skips leading white space, newlines included: an empty value ("3:" at the end of a line, "3: 0.5")
      // must not swallow the next token / the next line's label
      if (p >= e || *p == ' ' || *p == '\t') {
        set_error("ctr_parse_libsvm: row %lld field %d: empty value", (long long)row, f);
        return CTR_ERR_INVALID_ARG;
      }
      const float v = strtof(p, &q);
      if (q == p || q > e) {
        set_error("ctr_parse_libsvm: row %lld field %d: value is not a number", (long long)row, f);
        return CTR_ERR_INVALID_ARG;
      }
      ids[row * F + f] = (int32_t)id;
      vals[row * F + f] = v;
      p = q;
      ++f;
    }
    if (f != F) {
      set_error("ctr_parse_libsvm: row %lld has %s%d id:val pairs, field_size is %d", (long long)row,
                f > F ? "more than " : "", f > F ? F : f, F);
      return CTR_ERR_INVALID_ARG;
    }
    ++row;
    pos = nl ? end + 1 : end;
  }
  *consumed = pos;
  return row;
}

// number of id:val pairs on the first non-blank line (the reference takes field_size from a flag)
int ctr_libsvm_count_fields(const char* buf, size_t len) {
  if (!buf) return CTR_ERR_INVALID_ARG;
  size_t pos = 0;
  while (pos < len && (buf[pos] == '\n' || buf[pos] == ' ' || buf[pos] == '\r')) ++pos;
  int colons = 0;
  for (; pos < len && buf[pos] != '\n'; ++pos) colons += buf[pos] == ':';
  return colons;
}

}  // extern "C"
