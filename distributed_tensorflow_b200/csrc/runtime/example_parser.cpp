// Batch parser for serialized tensorflow.Example records (the payload of TFRecord input files): the host-side input pipeline's
// hot loop, outside the interpreter.  Role of TF's C++ `ParseExample` kernel (fast_parse_example); the reference itself feeds
// numpy batches through placeholders (distributed_mnist.py:149-152) -- this is the same stage for record-file data sets.
//
//   Example  { 1: Features }          Features { 1: repeated entry { 1: key (string), 2: Feature } }       (a proto map)
//   Feature  { oneof 1: BytesList | 2: FloatList | 3: Int64List }          *List { 1: repeated value (floats / ints usually packed) }
//
// For every record and every wanted FIXED-LENGTH feature: float / int64 values are written into row `rec` of the feature's output
// array (exactly `count` values or an error); a bytes feature yields (absolute offset, length) of its single value inside `data`
// (zero copy: the caller slices).  `present[rec * nfeat + f]` tells which features the record had; unknown keys are skipped.
#include <cstdint>
#include <cstring>

namespace {

struct Span {
  const uint8_t* p;
  const uint8_t* end;
};

inline bool read_varint(Span& s, uint64_t* v) {
  uint64_t out = 0;
  for (int shift = 0; shift < 64 && s.p < s.end; shift += 7) {
    const uint8_t b = *s.p++;
    out |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) {
      *v = out;
      return true;
    }
  }
  return false;
}

// Next field of a message: number, wire type, and for length-delimited fields the payload span.  `scalar`: varint value / fixed bits.
inline bool next_field(Span& s, uint32_t* field, uint32_t* wire, Span* body, uint64_t* scalar) {
  uint64_t key;
  if (!read_varint(s, &key)) return false;
  *field = (uint32_t)(key >> 3);
  *wire = (uint32_t)(key & 7);
  switch (*wire) {
    case 0:
      return read_varint(s, scalar);
    case 1:
      if (s.end - s.p < 8) return false;
      memcpy(scalar, s.p, 8);
      s.p += 8;
      return true;
    case 5: {
      if (s.end - s.p < 4) return false;
      uint32_t v;
      memcpy(&v, s.p, 4);
      *scalar = v;
      s.p += 4;
      return true;
    }
    case 2: {
      uint64_t n;
      if (!read_varint(s, &n) || n > (uint64_t)(s.end - s.p)) return false;
      body->p = s.p;
      body->end = s.p + n;
      s.p += n;
      return true;
    }
    default:
      return false;          // groups: not produced by any Example writer
  }
}

enum { KIND_BYTES = 0, KIND_FLOAT = 1, KIND_INT64 = 2 };
enum { ERR_MALFORMED = 1, ERR_KIND = 2, ERR_COUNT = 3 };

// One Feature message into its destination.  Returns 0, -1 (the Feature is empty) or an ERR_* code.
int parse_feature(Span f, int kind, int64_t count, const uint8_t* base, void* dst) {
  uint32_t field, wire;
  Span list{nullptr, nullptr};
  uint64_t sc;
  int have = -1;
  while (f.p < f.end) {
    Span body{nullptr, nullptr};
    if (!next_field(f, &field, &wire, &body, &sc)) return ERR_MALFORMED;
    if (wire == 2 && field >= 1 && field <= 3) {
      have = (int)field - 1;
      list = body;
    }
  }
  if (have < 0) return -1;                                   // an empty Feature: the key counts as absent (TF uses the default)
  if (have != kind) return ERR_KIND;
  int64_t n = 0;
  while (list.p < list.end) {
    Span body{nullptr, nullptr};
    if (!next_field(list, &field, &wire, &body, &sc)) return ERR_MALFORMED;
    if (field != 1) continue;
    if (kind == KIND_BYTES) {
      if (wire != 2) return ERR_MALFORMED;
      if (n >= count) return ERR_COUNT;
      int64_t* o = static_cast<int64_t*>(dst) + 2 * n;
      o[0] = (int64_t)(body.p - base);
      o[1] = (int64_t)(body.end - body.p);
      ++n;
    } else if (kind == KIND_FLOAT) {
      float* o = static_cast<float*>(dst);
      if (wire == 2) {                                       // packed
        const int64_t m = (body.end - body.p) / 4;
        if ((body.end - body.p) % 4 || n + m > count) return (body.end - body.p) % 4 ? ERR_MALFORMED : ERR_COUNT;
        memcpy(o + n, body.p, (size_t)m * 4);
        n += m;
      } else if (wire == 5) {
        if (n >= count) return ERR_COUNT;
        const uint32_t bits = (uint32_t)sc;
        memcpy(o + n, &bits, 4);
        ++n;
      } else {
        return ERR_MALFORMED;
      }
    } else {
      int64_t* o = static_cast<int64_t*>(dst);
      if (wire == 2) {                                       // packed varints
        while (body.p < body.end) {
          uint64_t v;
          if (!read_varint(body, &v)) return ERR_MALFORMED;
          if (n >= count) return ERR_COUNT;
          o[n++] = (int64_t)v;
        }
      } else if (wire == 0) {
        if (n >= count) return ERR_COUNT;
        o[n++] = (int64_t)sc;
      } else {
        return ERR_MALFORMED;
      }
    }
  }
  return n == count ? 0 : ERR_COUNT;
}

}  // namespace

extern "C" {

// data: one buffer holding every record; record r = [rec_off[r], rec_off[r] + rec_len[r]).
// Feature f: key keys[f] (key_len[f] bytes), kind[f] (0 bytes, 1 float, 2 int64), count[f] values per record (bytes: values too),
// out[f]: float[nrec * count] / int64[nrec * count] / int64[nrec * count * 2] (offset, length pairs).
// present: uint8[nrec * nfeat] (zeroed here).  Returns 0, or -(record + 1) with *err_feature (-1: the record itself is malformed)
// and *err_code (1 malformed, 2 the feature holds another kind, 3 wrong number of values).
int64_t dtf_parse_examples(const uint8_t* data, const int64_t* rec_off, const int64_t* rec_len, int64_t nrec, int nfeat,
                           const char* const* keys, const int* key_len, const int* kind, const int64_t* count, void* const* out,
                           uint8_t* present, int* err_feature, int* err_code) {
  memset(present, 0, (size_t)(nrec * nfeat));
  for (int64_t r = 0; r < nrec; ++r) {
    Span ex{data + rec_off[r], data + rec_off[r] + rec_len[r]};
    uint32_t field, wire;
    uint64_t sc;
    while (ex.p < ex.end) {
      Span feats{nullptr, nullptr};
      if (!next_field(ex, &field, &wire, &feats, &sc)) { *err_feature = -1; *err_code = ERR_MALFORMED; return -(r + 1); }
      if (field != 1 || wire != 2) continue;
      while (feats.p < feats.end) {
        Span entry{nullptr, nullptr};
        if (!next_field(feats, &field, &wire, &entry, &sc)) { *err_feature = -1; *err_code = ERR_MALFORMED; return -(r + 1); }
        if (field != 1 || wire != 2) continue;
        Span key{nullptr, nullptr}, val{nullptr, nullptr};
        while (entry.p < entry.end) {
          Span body{nullptr, nullptr};
          if (!next_field(entry, &field, &wire, &body, &sc)) { *err_feature = -1; *err_code = ERR_MALFORMED; return -(r + 1); }
          if (wire != 2) continue;
          if (field == 1) key = body;
          if (field == 2) val = body;
        }
        if (key.p == nullptr) continue;
        const int64_t klen = key.end - key.p;
        for (int f = 0; f < nfeat; ++f) {
          if (key_len[f] != klen || memcmp(keys[f], key.p, (size_t)klen) != 0) continue;
          const int64_t per = count[f] * (kind[f] == KIND_BYTES ? 16 : kind[f] == KIND_FLOAT ? 4 : 8);
          void* dst = static_cast<uint8_t*>(out[f]) + r * per;
          const int rc = parse_feature(val.p ? val : Span{key.p, key.p}, kind[f], count[f], data, dst);
          if (rc > 0) { *err_feature = f; *err_code = rc; return -(r + 1); }
          if (rc == 0) present[r * nfeat + f] = 1;
          break;
        }
      }
    }
  }
  return 0;
}

}  // extern "C"
