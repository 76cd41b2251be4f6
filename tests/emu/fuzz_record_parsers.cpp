// Mutation fuzzer for the native record-file parsers (csrc/runtime/example_parser.cpp: dtf_parse_examples, csrc/runtime/bundle_io.cpp:
// dtf_tfrecord_scan), built with -fsanitize=address,undefined by tests/test_tfrecord.py: valid serialized Examples (seeds.bin, written
// by the test) are byte-flipped / truncated / extended and parsed from EXACT-size heap buffers, so any read past the end of a record
// or any write past an output row is a sanitizer report.  argv[1] = seeds file, argv[2] = iterations.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
extern "C" int64_t dtf_parse_examples(const uint8_t*, const int64_t*, const int64_t*, int64_t, int, const char* const*, const int*, const int*,
                                      const int64_t*, void* const*, uint8_t*, int*, int*);
extern "C" int64_t dtf_tfrecord_scan(const void*, int64_t, int64_t*, int64_t*, int64_t, int);
int main(int argc, char** argv) {
  FILE* f = fopen(argc > 1 ? argv[1] : "seeds.bin", "rb");
  if (!f) return 2;
  const int iters = argc > 2 ? atoi(argv[2]) : 400000;
  uint32_t n;
  if (fread(&n, 4, 1, f) != 1) return 2;
  std::vector<std::vector<uint8_t>> seeds(n);
  for (auto& s : seeds) {
    uint32_t l;
    if (fread(&l, 4, 1, f) != 1) return 2;
    s.resize(l);
    if (fread(s.data(), 1, l, f) != l) return 2;
  }
  const char* keys[3] = {"image_raw", "label", "weights"};
  int klen[3] = {9, 5, 7}, kind[3] = {0, 2, 1};
  int64_t count[3] = {1, 2, 3};
  srand(1234);
  long ok = 0, bad = 0;
  for (int it = 0; it < iters; ++it) {
    const int B = 1 + rand() % 4;
    // exact-size heap buffer so that ASAN sees any read past the end
    std::vector<int64_t> off(B), len(B);
    size_t total = 0;
    std::vector<std::vector<uint8_t>> recs(B);
    for (int b = 0; b < B; ++b) {
      recs[b] = seeds[rand() % n];
      const int muts = rand() % 4;
      for (int m = 0; m < muts && !recs[b].empty(); ++m) {
        const int kindm = rand() % 4;
        const size_t pos = rand() % recs[b].size();
        if (kindm == 0) recs[b][pos] = (uint8_t)rand();
        else if (kindm == 1) recs[b][pos] ^= (uint8_t)(1 << (rand() % 8));
        else if (kindm == 2) recs[b].resize(pos);
        else recs[b].insert(recs[b].begin() + pos, (uint8_t)rand());
      }
      off[b] = (int64_t)total; len[b] = (int64_t)recs[b].size(); total += recs[b].size();
    }
    uint8_t* data = (uint8_t*)malloc(total ? total : 1);
    for (int b = 0; b < B; ++b) memcpy(data + off[b], recs[b].data(), recs[b].size());
    std::vector<int64_t> o0(B * 2), o1(B * 2); std::vector<float> o2(B * 3); std::vector<uint8_t> present(B * 3);
    void* outs[3] = {o0.data(), o1.data(), o2.data()};
    int ef = 0, ec = 0;
    const int64_t rc = dtf_parse_examples(data, off.data(), len.data(), B, 3, keys, klen, kind, count, outs, present.data(), &ef, &ec);
    if (rc == 0) {
      ++ok;
      for (int b = 0; b < B; ++b) if (present[b * 3]) {        // located bytes must lie inside the buffer
        if (o0[b * 2] < 0 || o0[b * 2 + 1] < 0 || (size_t)(o0[b * 2] + o0[b * 2 + 1]) > total) { printf("BAD SPAN\n"); return 1; }
      }
    } else ++bad;
    // the record scanner on the same bytes (as if they were a file)
    std::vector<int64_t> so(total / 16 + 1), sl(total / 16 + 1);
    dtf_tfrecord_scan(data, (int64_t)total, so.data(), sl.data(), (int64_t)so.size(), 1);
    free(data);
  }
  printf("parsed ok %ld, rejected %ld\n", ok, bad);
  return 0;
}
