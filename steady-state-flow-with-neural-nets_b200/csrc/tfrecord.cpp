// Host-side reader / writer for the reference's dataset files: TFRecord framing around
// tf.train.Example protos holding two bytes features, 'boundary' (uint8 [H*W]) and 'sflow' (float32 [H*W*2]) --
// written by /root/reference/utils/createTFRecords.py:45-56, read by input/flow_input.py:14-30.  TensorFlow is not
// a dependency here: the framing (length, masked CRC32-C of the length, payload, masked CRC32-C of the payload) and
// the three proto messages involved (Example.features=1 -> Features.feature=1 map<string, Feature>, Feature.bytes_list=1
// -> BytesList.value=1) are restated from their public definitions (tensorflow/core/lib/io/record_writer.h,
// tensorflow/core/example/{example,feature}.proto).
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "flownet.h"

namespace {

uint32_t g_crc_table[8][256];
bool g_crc_ready = false;

void crc_init() {
  if (g_crc_ready) return;
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;      // CRC-32C (Castagnoli), reflected
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xFF];
  g_crc_ready = true;
}

#if defined(__x86_64__) && defined(__GNUC__)
// the CRC32 instruction of SSE4.2 computes exactly this polynomial
__attribute__((target("sse4.2"))) uint32_t crc32c_hw(const uint8_t* p, int64_t n) {
  uint64_t c = 0xFFFFFFFFu;
  while (n >= 8) {
    uint64_t w;
    memcpy(&w, p, 8);
    c = __builtin_ia32_crc32di(c, w);
    p += 8;
    n -= 8;
  }
  uint32_t c32 = (uint32_t)c;
  while (n-- > 0) c32 = __builtin_ia32_crc32qi(c32, *p++);
  return c32 ^ 0xFFFFFFFFu;
}
bool have_sse42() {
  static const bool v = __builtin_cpu_supports("sse4.2");
  return v;
}
#else
uint32_t crc32c_hw(const uint8_t*, int64_t) { return 0; }
bool have_sse42() { return false; }
#endif

uint32_t crc32c(const uint8_t* p, int64_t n) {
  if (have_sse42()) return crc32c_hw(p, n);
  crc_init();
  uint32_t c = 0xFFFFFFFFu;
  while (n >= 8) {          // slicing-by-8
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_table[7][w & 0xFF] ^ g_crc_table[6][(w >> 8) & 0xFF] ^ g_crc_table[5][(w >> 16) & 0xFF] ^
        g_crc_table[4][(w >> 24) & 0xFF] ^ g_crc_table[3][(w >> 32) & 0xFF] ^ g_crc_table[2][(w >> 40) & 0xFF] ^
        g_crc_table[1][(w >> 48) & 0xFF] ^ g_crc_table[0][(w >> 56) & 0xFF];
    p += 8;
    n -= 8;
  }
  while (n-- > 0) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xFF];
  return c ^ 0xFFFFFFFFu;
}

uint32_t masked_crc(const uint8_t* p, int64_t n) {
  const uint32_t c = crc32c(p, n);
  return ((c >> 15) | (c << 17)) + 0xa282ead8u;
}

// ---- minimal protobuf wire-format walking
bool read_varint(const uint8_t*& p, const uint8_t* end, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    v |= (uint64_t)(b & 0x7F) << shift;
    if (!(b & 0x80)) return true;
  }
  return false;
}
// next field of a message: returns false at the end or on a malformed field.  For length-delimited fields
// (wire type 2) [data, data+len) is the payload; other wire types are skipped (data = nullptr).
bool next_field(const uint8_t*& p, const uint8_t* end, uint32_t& field, const uint8_t*& data, uint64_t& len, bool& bad) {
  bad = false;
  if (p >= end) return false;
  uint64_t key;
  if (!read_varint(p, end, key)) { bad = true; return false; }
  field = (uint32_t)(key >> 3);
  data = nullptr;
  len = 0;
  switch (key & 7) {
    case 0: { uint64_t v; if (!read_varint(p, end, v)) { bad = true; return false; } return true; }
    case 1: if (end - p < 8) { bad = true; return false; } p += 8; return true;
    case 5: if (end - p < 4) { bad = true; return false; } p += 4; return true;
    case 2:
      if (!read_varint(p, end, len) || (uint64_t)(end - p) < len) { bad = true; return false; }
      data = p;
      p += len;
      return true;
    default: bad = true; return false;
  }
}

// Feature { BytesList bytes_list = 1 { repeated bytes value = 1 } }: first value
bool first_bytes_value(const uint8_t* p, const uint8_t* end, const uint8_t*& val, uint64_t& vlen) {
  uint32_t f; const uint8_t* d; uint64_t l; bool bad;
  while (next_field(p, end, f, d, l, bad)) {
    if (f == 1 && d) {                       // bytes_list
      const uint8_t* q = d; const uint8_t* qe = d + l;
      uint32_t f2; const uint8_t* d2; uint64_t l2;
      while (next_field(q, qe, f2, d2, l2, bad))
        if (f2 == 1 && d2) { val = d2; vlen = l2; return true; }
      return false;
    }
  }
  return false;
}

void put_varint(uint8_t*& p, uint64_t v) {
  while (v >= 0x80) { *p++ = (uint8_t)(v | 0x80); v >>= 7; }
  *p++ = (uint8_t)v;
}
int varint_size(uint64_t v) {
  int n = 1;
  while (v >= 0x80) { v >>= 7; ++n; }
  return n;
}
// size of one map entry  Features.feature[key] = Feature{bytes_list{value}}
uint64_t entry_size(uint64_t keylen, uint64_t vlen, uint64_t& feat, uint64_t& blist) {
  blist = 1 + varint_size(vlen) + vlen;               // BytesList: value = 1
  feat = 1 + varint_size(blist) + blist;              // Feature: bytes_list = 1
  return 1 + varint_size(keylen) + keylen + 1 + varint_size(feat) + feat;
}
void put_entry(uint8_t*& p, const char* key, const uint8_t* val, uint64_t vlen) {
  const uint64_t keylen = strlen(key);
  uint64_t feat, blist;
  const uint64_t esz = entry_size(keylen, vlen, feat, blist);
  *p++ = 0x0A; put_varint(p, esz);                    // Features.feature (field 1, map entry)
  *p++ = 0x0A; put_varint(p, keylen); memcpy(p, key, keylen); p += keylen;
  *p++ = 0x12; put_varint(p, feat);                   // entry.value (field 2) = Feature
  *p++ = 0x0A; put_varint(p, blist);                  // Feature.bytes_list
  *p++ = 0x0A; put_varint(p, vlen); memcpy(p, val, vlen); p += vlen;
}

}  // namespace

extern "C" {

int64_t fn_tfrecord_index(const uint8_t* buf, int64_t nbytes, int64_t* offsets, int64_t* lengths, int64_t max_records,
                          int32_t check_crc) {
  if (!buf || nbytes < 0) return FN_E_INVAL;
  int64_t pos = 0, n = 0;
  while (pos < nbytes) {
    if (nbytes - pos < 16) return FN_E_INVAL;                       // truncated header / footer: a record is >= 16 bytes
    uint64_t len;
    uint32_t lcrc;
    memcpy(&len, buf + pos, 8);
    memcpy(&lcrc, buf + pos + 8, 4);
    if (check_crc && masked_crc(buf + pos, 8) != lcrc) return FN_E_INVAL;
    if (len > (uint64_t)(nbytes - pos - 16)) return FN_E_INVAL;     // truncated payload
    if (check_crc) {
      uint32_t dcrc;
      memcpy(&dcrc, buf + pos + 12 + len, 4);
      if (masked_crc(buf + pos + 12, (int64_t)len) != dcrc) return FN_E_INVAL;
    }
    if (offsets && lengths && n < max_records) {
      offsets[n] = pos + 12;
      lengths[n] = (int64_t)len;
    }
    ++n;
    pos += 16 + (int64_t)len;
  }
  return n;
}

int fn_tfrecord_parse_flow_example(const uint8_t* payload, int64_t len, uint8_t* boundary, int64_t boundary_bytes, float* sflow,
                                   int64_t sflow_bytes) {
  if (!payload || len < 0 || !boundary || !sflow) return FN_E_INVAL;
  const uint8_t* p = payload; const uint8_t* end = payload + len;
  uint32_t f; const uint8_t* d; uint64_t l; bool bad;
  int found = 0;
  while (next_field(p, end, f, d, l, bad)) {
    if (f != 1 || !d) continue;                                     // Example.features
    const uint8_t* q = d; const uint8_t* qe = d + l;
    uint32_t f2; const uint8_t* d2; uint64_t l2;
    while (next_field(q, qe, f2, d2, l2, bad)) {
      if (f2 != 1 || !d2) continue;                                 // one map entry
      const uint8_t* r = d2; const uint8_t* re = d2 + l2;
      uint32_t f3; const uint8_t* d3; uint64_t l3;
      const uint8_t* key = nullptr; uint64_t keylen = 0; const uint8_t* feat = nullptr; uint64_t featlen = 0;
      while (next_field(r, re, f3, d3, l3, bad)) {
        if (f3 == 1 && d3) { key = d3; keylen = l3; }
        if (f3 == 2 && d3) { feat = d3; featlen = l3; }
      }
      if (bad || !key || !feat) continue;
      const uint8_t* val; uint64_t vlen;
      if (keylen == 8 && memcmp(key, "boundary", 8) == 0) {
        if (!first_bytes_value(feat, feat + featlen, val, vlen) || (int64_t)vlen != boundary_bytes) return FN_E_INVAL;
        memcpy(boundary, val, vlen);
        found |= 1;
      } else if (keylen == 5 && memcmp(key, "sflow", 5) == 0) {
        if (!first_bytes_value(feat, feat + featlen, val, vlen) || (int64_t)vlen != sflow_bytes) return FN_E_INVAL;
        memcpy(sflow, val, vlen);
        found |= 2;
      }
    }
    if (bad) return FN_E_INVAL;
  }
  if (bad) return FN_E_INVAL;
  return found == 3 ? 0 : FN_E_INVAL;
}

int64_t fn_tfrecord_write_flow_example(uint8_t* dst, int64_t capacity, const uint8_t* boundary, int64_t boundary_bytes,
                                       const float* sflow, int64_t sflow_bytes) {
  if (!boundary || !sflow || boundary_bytes < 0 || sflow_bytes < 0) return FN_E_INVAL;
  uint64_t f, b;
  const uint64_t e1 = entry_size(8, (uint64_t)boundary_bytes, f, b), e2 = entry_size(5, (uint64_t)sflow_bytes, f, b);
  const uint64_t features = 1 + varint_size(e1) + e1 + 1 + varint_size(e2) + e2;
  const uint64_t example = 1 + varint_size(features) + features;
  const int64_t total = 16 + (int64_t)example;
  if (!dst) return total;                                           // size query
  if (capacity < total) return FN_E_INVAL;
  uint8_t* p = dst;
  memcpy(p, &example, 8);
  const uint32_t lcrc = masked_crc(p, 8);
  memcpy(p + 8, &lcrc, 4);
  p += 12;
  uint8_t* body = p;
  *p++ = 0x0A; put_varint(p, features);                             // Example.features
  put_entry(p, "boundary", boundary, (uint64_t)boundary_bytes);
  put_entry(p, "sflow", reinterpret_cast<const uint8_t*>(sflow), (uint64_t)sflow_bytes);
  if ((uint64_t)(p - body) != example) return FN_E_INVAL;
  const uint32_t dcrc = masked_crc(body, (int64_t)example);
  memcpy(p, &dcrc, 4);
  return total;
}

/* n records of one file image -> consecutive slots of a batch (boundary: n x boundary_bytes, sflow: n x sflow_bytes),
 * parsed by up to nthreads host threads (each record is one 0.3 MB copy: memory-bandwidth work that one thread cannot
 * saturate).  Returns 0, or FN_E_INVAL if any record is not a {'boundary','sflow'} example of these sizes. */
int fn_tfrecord_parse_flow_batch(const uint8_t* buf, const int64_t* offsets, const int64_t* lengths, int64_t n, uint8_t* boundary,
                                 int64_t boundary_bytes, float* sflow, int64_t sflow_bytes, int32_t nthreads) {
  if (!buf || !offsets || !lengths || n < 0 || !boundary || !sflow) return FN_E_INVAL;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > n) nthreads = (int32_t)(n > 0 ? n : 1);
  std::atomic<int> err(0);
  auto work = [&](int t) {
    for (int64_t i = t; i < n; i += nthreads) {
      const int rc = fn_tfrecord_parse_flow_example(buf + offsets[i], lengths[i], boundary + i * boundary_bytes, boundary_bytes,
                                                    reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sflow) + i * sflow_bytes),
                                                    sflow_bytes);
      if (rc != 0) err.store(rc);
    }
  };
  if (nthreads == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  return err.load();
}

uint32_t fn_crc32c(const uint8_t* p, int64_t n) { return crc32c(p, n); }
/* the table-driven path, whatever the CPU offers (tests compare the two) */
uint32_t fn_crc32c_portable(const uint8_t* p, int64_t n) {
  crc_init();
  uint32_t c = 0xFFFFFFFFu;
  while (n-- > 0) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xFF];
  return c ^ 0xFFFFFFFFu;
}

}  // extern "C"
