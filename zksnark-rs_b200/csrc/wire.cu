// Flat binary layout of the three objects that cross the boundary -- QAP, CRS, Proof (include/zkb200.h: zkb_wire_*).
//
// The reference has no serialisation at all: `QAP`, `SigmaG1`, `SigmaG2`, `Proof` have private fields and no accessors
// (/root/reference/src/groth16/mod.rs:60-128), so a CRS cannot leave the process that ran setup().  A prover service on
// a GPU box needs exactly that (setup once, ship the CRS and the QAP, get proofs back), hence this format.
// Host-only code: no device is touched.
//
//   header, 64 bytes, little-endian:
//     0  magic "ZKB200W\0"        8  u32 version (1)      12  u32 kind (1 QAP, 2 CRS, 3 Proof)
//     16 u64 total bytes          24 u64 f0  32 u64 f1  40 u64 f2  48 u64 f3        56 u64 FNV-1a-64 of bytes [64, total)
//        QAP:  f0 n, f1 m, f2 n_input, f3 bit 0 = explicit roots follow        CRS: f0 n, f1 n_sum_gamma, f2 n_sum_delta
//   payload: arrays in the order below, each padded to a multiple of 8 bytes, elements as at the C ABI (4 x u64 LE limbs
//   of canonical residues; points affine, identity all-zero):
//     QAP    for u, v, w: row_ptr[m+1] u64 | gate[nnz] u32 | coeff[nnz][4] u64;   then roots[n][4] if f3 & 1
//     CRS    alpha1[8] beta1[8] delta1[8] beta2[16] gamma2[16] delta2[16] xi1[n][8] xi_t[n-1][8] sum_gamma[..][8]
//            sum_delta[..][8] xi2[n][16]
//     Proof  a[8] b[16] c[8]
// Readers return VIEWS: the pointers of the zkb_*_host struct point into the caller's buffer (8-byte aligned, kept alive).
#include <string.h>
#include "common.cuh"

namespace {

const char kMagic[8] = {'Z', 'K', 'B', '2', '0', '0', 'W', '\0'};
const uint32_t kVersion = 1;
const uint64_t kHeader = 64;

inline uint64_t pad8(uint64_t b) { return (b + 7) & ~(uint64_t)7; }
uint64_t fnv1a(const uint8_t* p, uint64_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
  return h;
}
void put_header(uint8_t* buf, uint32_t kind, uint64_t total, uint64_t f0, uint64_t f1, uint64_t f2, uint64_t f3) {
  memset(buf, 0, kHeader);
  memcpy(buf, kMagic, 8);
  memcpy(buf + 8, &kVersion, 4);
  memcpy(buf + 12, &kind, 4);
  uint64_t f[5] = {total, f0, f1, f2, f3};
  memcpy(buf + 16, f, 40);
}
void seal(uint8_t* buf, uint64_t total) {
  uint64_t h = fnv1a(buf + kHeader, total - kHeader);
  memcpy(buf + 56, &h, 8);
}
int open_header(const uint8_t* buf, uint64_t len, uint32_t want_kind, uint64_t f[5]) {
  using zkb::set_err;
  if (!buf || len < kHeader) return set_err(nullptr, ZKB_ERR_ARG, "wire: buffer shorter than the 64-byte header");
  if (((uintptr_t)buf & 7) != 0) return set_err(nullptr, ZKB_ERR_ARG, "wire: buffer must be 8-byte aligned (readers return views into it)");
  if (memcmp(buf, kMagic, 8) != 0) return set_err(nullptr, ZKB_ERR_ARG, "wire: bad magic");
  uint32_t ver, kind;
  memcpy(&ver, buf + 8, 4);
  memcpy(&kind, buf + 12, 4);
  if (ver != kVersion) return set_err(nullptr, ZKB_ERR_UNSUPPORTED, "wire: version %u (this library reads version %u)", ver, kVersion);
  if (want_kind && kind != want_kind) return set_err(nullptr, ZKB_ERR_ARG, "wire: object kind %u, expected %u", kind, want_kind);
  memcpy(f, buf + 16, 40);
  if (f[0] < kHeader || f[0] > len) return set_err(nullptr, ZKB_ERR_ARG, "wire: truncated (%llu bytes declared, %llu given)",
                                                   (unsigned long long)f[0], (unsigned long long)len);
  uint64_t sum;
  memcpy(&sum, buf + 56, 8);
  if (sum != fnv1a(buf + kHeader, f[0] - kHeader)) return set_err(nullptr, ZKB_ERR_ARG, "wire: checksum mismatch (corrupted payload)");
  return ZKB_OK;
}

uint64_t qap_bytes(const zkb_qap_host* q) {
  uint64_t b = kHeader;
  for (int t = 0; t < 3; t++) {
    const uint64_t nnz = q->row_ptr[t][q->m];
    b += (q->m + 1) * 8 + pad8(nnz * 4) + nnz * 32;
  }
  if (q->roots) b += q->n * 32;
  return b;
}
uint64_t crs_bytes(const zkb_crs_host* c) {
  return kHeader + (3 * 8 + 3 * 16 + c->n * 8 + (c->n ? c->n - 1 : 0) * 8 + c->n_sum_gamma * 8 + c->n_sum_delta * 8 + c->n * 16) * 8;
}

}  // namespace

using zkb::set_err;

extern "C" {

int zkb_wire_kind(const uint8_t* buf, uint64_t len, int* kind, uint64_t* total_bytes) {
  uint64_t f[5];
  ZKB_TRY(open_header(buf, len, 0, f));
  uint32_t k;
  memcpy(&k, buf + 12, 4);
  if (kind) *kind = (int)k;
  if (total_bytes) *total_bytes = f[0];
  return ZKB_OK;
}

int zkb_wire_size_qap(const zkb_qap_host* q, uint64_t* bytes) {
  if (!q || !bytes) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_size_qap: NULL argument");
  for (int t = 0; t < 3; t++)
    if (!q->row_ptr[t]) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_size_qap: NULL row_ptr");
  *bytes = qap_bytes(q);
  return ZKB_OK;
}

int zkb_wire_write_qap(const zkb_qap_host* q, uint8_t* buf, uint64_t cap) {
  uint64_t total;
  ZKB_TRY(zkb_wire_size_qap(q, &total));
  if (!buf || cap < total) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_write_qap: buffer too small (%llu needed)", (unsigned long long)total);
  put_header(buf, 1, total, q->n, q->m, q->n_input, q->roots ? 1 : 0);
  uint8_t* p = buf + kHeader;
  for (int t = 0; t < 3; t++) {
    const uint64_t nnz = q->row_ptr[t][q->m];
    if (nnz && (!q->gate[t] || !q->coeff[t])) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_write_qap: NULL rows");
    memcpy(p, q->row_ptr[t], (q->m + 1) * 8); p += (q->m + 1) * 8;
    memset(p, 0, pad8(nnz * 4));
    if (nnz) memcpy(p, q->gate[t], nnz * 4);
    p += pad8(nnz * 4);
    if (nnz) memcpy(p, q->coeff[t], nnz * 32);
    p += nnz * 32;
  }
  if (q->roots) { memcpy(p, q->roots, q->n * 32); p += q->n * 32; }
  seal(buf, total);
  return ZKB_OK;
}

int zkb_wire_read_qap(const uint8_t* buf, uint64_t len, zkb_qap_host* out) {
  if (!out) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_read_qap: NULL argument");
  uint64_t f[5];
  ZKB_TRY(open_header(buf, len, 1, f));
  memset(out, 0, sizeof *out);
  out->n = f[1]; out->m = f[2]; out->n_input = f[3];
  const uint64_t total = f[0];
  if (out->m == 0 || out->m > ((uint64_t)1 << 32)) return set_err(nullptr, ZKB_ERR_ARG, "wire: qap.m out of range");
  uint64_t off = kHeader;
  for (int t = 0; t < 3; t++) {
    if (off + (out->m + 1) * 8 > total) return set_err(nullptr, ZKB_ERR_ARG, "wire: qap rows run past the end");
    const uint64_t* rp = reinterpret_cast<const uint64_t*>(buf + off);
    out->row_ptr[t] = rp;
    off += (out->m + 1) * 8;
    const uint64_t nnz = rp[out->m];
    if (nnz > (total - off) / 32) return set_err(nullptr, ZKB_ERR_ARG, "wire: qap rows run past the end");
    out->gate[t] = reinterpret_cast<const uint32_t*>(buf + off);
    off += pad8(nnz * 4);
    out->coeff[t] = reinterpret_cast<const uint64_t*>(buf + off);
    off += nnz * 32;
    if (off > total) return set_err(nullptr, ZKB_ERR_ARG, "wire: qap rows run past the end");
  }
  if (f[4] & 1) {
    if (out->n > (total - off) / 32) return set_err(nullptr, ZKB_ERR_ARG, "wire: qap roots run past the end");
    out->roots = reinterpret_cast<const uint64_t*>(buf + off);
    off += out->n * 32;
  }
  if (off != total) return set_err(nullptr, ZKB_ERR_ARG, "wire: %llu trailing bytes", (unsigned long long)(total - off));
  return ZKB_OK;
}

int zkb_wire_size_crs(const zkb_crs_host* c, uint64_t* bytes) {
  if (!c || !bytes) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_size_crs: NULL argument");
  *bytes = crs_bytes(c);
  return ZKB_OK;
}

int zkb_wire_write_crs(const zkb_crs_host* c, uint8_t* buf, uint64_t cap) {
  uint64_t total;
  ZKB_TRY(zkb_wire_size_crs(c, &total));
  if (!buf || cap < total) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_write_crs: buffer too small (%llu needed)", (unsigned long long)total);
  const uint64_t nxt = c->n ? c->n - 1 : 0;
  const void* src[11] = {c->alpha1, c->beta1, c->delta1, c->beta2, c->gamma2, c->delta2, c->xi1, c->xi_t, c->sum_gamma, c->sum_delta, c->xi2};
  const uint64_t cnt[11] = {8, 8, 8, 16, 16, 16, c->n * 8, nxt * 8, c->n_sum_gamma * 8, c->n_sum_delta * 8, c->n * 16};
  put_header(buf, 2, total, c->n, c->n_sum_gamma, c->n_sum_delta, 0);
  uint8_t* p = buf + kHeader;
  for (int k = 0; k < 11; k++) {
    if (cnt[k] && !src[k]) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_write_crs: NULL vector");
    if (cnt[k]) memcpy(p, src[k], cnt[k] * 8);
    p += cnt[k] * 8;
  }
  seal(buf, total);
  return ZKB_OK;
}

int zkb_wire_read_crs(const uint8_t* buf, uint64_t len, zkb_crs_host* out) {
  if (!out) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_read_crs: NULL argument");
  uint64_t f[5];
  ZKB_TRY(open_header(buf, len, 2, f));
  memset(out, 0, sizeof *out);
  out->n = f[1]; out->n_sum_gamma = f[2]; out->n_sum_delta = f[3];
  if (out->n > ((uint64_t)1 << 32) || out->n_sum_gamma > ((uint64_t)1 << 32) || out->n_sum_delta > ((uint64_t)1 << 32) ||
      crs_bytes(out) != f[0])
    return set_err(nullptr, ZKB_ERR_ARG, "wire: CRS sizes do not match the declared length");
  const uint64_t nxt = out->n ? out->n - 1 : 0;
  const uint64_t* p = reinterpret_cast<const uint64_t*>(buf + kHeader);
  out->alpha1 = p; p += 8;
  out->beta1 = p; p += 8;
  out->delta1 = p; p += 8;
  out->beta2 = p; p += 16;
  out->gamma2 = p; p += 16;
  out->delta2 = p; p += 16;
  out->xi1 = p; p += out->n * 8;
  out->xi_t = p; p += nxt * 8;
  out->sum_gamma = p; p += out->n_sum_gamma * 8;
  out->sum_delta = p; p += out->n_sum_delta * 8;
  out->xi2 = p;
  return ZKB_OK;
}

int zkb_wire_write_proof(const zkb_proof* pr, uint8_t* buf, uint64_t cap) {
  if (!pr || !buf || cap < ZKB_WIRE_PROOF_BYTES) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_write_proof: bad argument");
  put_header(buf, 3, ZKB_WIRE_PROOF_BYTES, 0, 0, 0, 0);
  memcpy(buf + kHeader, pr, sizeof(zkb_proof));
  seal(buf, ZKB_WIRE_PROOF_BYTES);
  return ZKB_OK;
}

int zkb_wire_read_proof(const uint8_t* buf, uint64_t len, zkb_proof* out) {
  if (!out) return set_err(nullptr, ZKB_ERR_ARG, "zkb_wire_read_proof: NULL argument");
  uint64_t f[5];
  ZKB_TRY(open_header(buf, len, 3, f));
  if (f[0] != ZKB_WIRE_PROOF_BYTES) return set_err(nullptr, ZKB_ERR_ARG, "wire: proof record has the wrong length");
  memcpy(out, buf + kHeader, sizeof(zkb_proof));
  return ZKB_OK;
}

}  // extern "C"
