// zkb200.hpp -- C++ host-side mirror of the reference's Groth16 interface over the C ABI (zkb200.h).
//
// The reference is compiled code (Rust) and its toolchain is absent from the build image, so the host
// side above the C ABI is written in C++ with the reference's names, argument meaning and error
// behaviour (/root/reference/src/groth16/mod.rs):
//
//   groth16::setup(&qap)                         mod.rs:134-197   ->  zkb200::groth16::setup(ctx, qap)
//   groth16::prove(&qap, (&s1, &s2), &weights)   mod.rs:213-296   ->  zkb200::groth16::prove(ctx, qap, sigma, weights)
//   groth16::verify((s1, s2), &inputs, proof)    mod.rs:299-320   ->  zkb200::groth16::verify(ctx, sigma, inputs, proof)
//   QAP::from(root_representation)               fr.rs:140-173    ->  zkb200::QAP::from(ctx, rep)
//   DummyRep { u, v, w, roots, input }           circuit/dummy_rep.rs:7-13 -> zkb200::RootRepresentation
//   FrLocal (+ - * /, From<usize>, from_str, random_elem)  fr.rs:18-99 -> zkb200::Fr
//
// The reference panics on every failure (fr.rs:54, field/mod.rs:440); this mirror throws zkb200::Error.
// (SigmaG1, SigmaG2) stay resident on the device as one `Sigma` handle: the reference's types have
// private fields and no accessors, so nothing else could be done with them anyway.
//
// Everything that computes a proof, a CRS or a verdict runs on the device through libzkb200.so.  The
// only host arithmetic here is `Fr` itself (callers build their witnesses with it, exactly as the
// reference's tests do with FrLocal): plain 256-bit modular arithmetic, never used by the library.
#pragma once
#include <array>
#include <cerrno>
#include <cstdio>
#include <cstring>
#if defined(__linux__)
#include <sys/random.h>
#endif
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "zkb200.h"

namespace zkb200 {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// ---- FrLocal (fr.rs:9-99): canonical residue mod r, 4 x u64 little-endian ------------------------
struct Fr {
  std::array<uint64_t, 4> l{};
  static constexpr std::array<uint64_t, 4> R = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                                                0x30644e72e131a029ull};
  Fr() = default;
  Fr(uint64_t x) { l[0] = x; }  // From<usize> (fr.rs:73-77)
  static Fr zero() { return Fr(); }
  static Fr one() { return Fr(1); }
  static Fr from_str(const std::string& dec) {  // decimal (fr.rs:79-88)
    Fr r;
    for (char ch : dec) {
      if (ch < '0' || ch > '9') throw Error(ZKB_ERR_ARG, "Fr::from_str: not a decimal string");
      r = r * Fr(10) + Fr((uint64_t)(ch - '0'));
    }
    return r;
  }
  // uniform and never zero (fr.rs:90-99).  The reference draws from rand::thread_rng (an OS-seeded CSPRNG); these
  // values are the toxic waste of setup() and the blinding r, s of prove(), so every candidate comes straight from
  // the kernel CSPRNG (getrandom(2), /dev/urandom as the fallback) -- never from a seeded userspace generator --
  // with rejection sampling on the 254-bit candidates.  Throws when the OS source is unavailable.
  static void os_random(void* buf, size_t len) {
    unsigned char* p = static_cast<unsigned char*>(buf);
    size_t got = 0;
#if defined(__linux__)
    while (got < len) {
      ssize_t k = ::getrandom(p + got, len - got, 0);
      if (k < 0) {
        if (errno == EINTR) continue;
        break;  // ENOSYS etc.: try /dev/urandom
      }
      got += (size_t)k;
    }
#endif
    if (got < len) {
      std::FILE* f = std::fopen("/dev/urandom", "rb");
      if (f) {
        got += std::fread(p + got, 1, len - got, f);
        std::fclose(f);
      }
    }
    if (got < len) throw Error(ZKB_ERR_UNSUPPORTED, "Fr::random_elem: no OS random source (getrandom, /dev/urandom)");
  }
  static Fr random_elem() {
    for (;;) {
      Fr r;
      os_random(r.l.data(), sizeof r.l);
      r.l[3] &= 0x3fffffffffffffffull;
      if (!geq(r.l, R) && !r.is_zero()) return r;
    }
  }
  bool is_zero() const { return !(l[0] | l[1] | l[2] | l[3]); }
  bool operator==(const Fr& o) const { return l == o.l; }
  bool operator!=(const Fr& o) const { return !(l == o.l); }
  Fr operator+(const Fr& o) const {
    Fr r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; i++) { c += (unsigned __int128)l[i] + o.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    if (c || geq(r.l, R)) sub_in_place(r.l, R);
    return r;
  }
  Fr operator-() const { return is_zero() ? *this : Fr::from_limbs(R) - *this; }
  Fr operator-(const Fr& o) const {
    Fr r = *this;
    if (sub_in_place(r.l, o.l)) {  // borrowed: add r back
      unsigned __int128 c = 0;
      for (int i = 0; i < 4; i++) { c += (unsigned __int128)r.l[i] + R[i]; r.l[i] = (uint64_t)c; c >>= 64; }
    }
    return r;
  }
  Fr operator*(const Fr& o) const {
    uint64_t t[8] = {};
    for (int i = 0; i < 4; i++) {
      unsigned __int128 c = 0;
      for (int j = 0; j < 4; j++) { c += (unsigned __int128)l[i] * o.l[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
      t[i + 4] = (uint64_t)c;
    }
    // binary reduction of the 512-bit product (host-side convenience arithmetic: clarity over speed)
    Fr r;
    for (int bit = 511; bit >= 0; bit--) {
      uint64_t top = r.l[3] >> 63;
      for (int i = 3; i > 0; i--) r.l[i] = (r.l[i] << 1) | (r.l[i - 1] >> 63);
      r.l[0] = (r.l[0] << 1) | ((t[bit >> 6] >> (bit & 63)) & 1);
      if (top || geq(r.l, R)) sub_in_place(r.l, R);
    }
    return r;
  }
  Fr pow(const std::array<uint64_t, 4>& e) const {
    Fr acc = one();
    for (int bit = 255; bit >= 0; bit--) {
      acc = acc * acc;
      if ((e[bit >> 6] >> (bit & 63)) & 1) acc = acc * *this;
    }
    return acc;
  }
  Fr mul_inv() const {  // fr.rs:67-70; the reference panics on zero (fr.rs:54)
    if (is_zero()) throw Error(ZKB_ERR_DIV_ZERO, "Fr: inverse of zero");
    std::array<uint64_t, 4> e = R;
    e[0] -= 2;
    return pow(e);
  }
  Fr operator/(const Fr& o) const { return *this * o.mul_inv(); }
  static Fr from_limbs(const std::array<uint64_t, 4>& a) { Fr r; r.l = a; return r; }

 private:
  static bool geq(const std::array<uint64_t, 4>& a, const std::array<uint64_t, 4>& b) {
    for (int i = 3; i >= 0; i--) {
      if (a[i] > b[i]) return true;
      if (a[i] < b[i]) return false;
    }
    return true;
  }
  static bool sub_in_place(std::array<uint64_t, 4>& a, const std::array<uint64_t, 4>& b) {  // returns the borrow
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) {
      unsigned __int128 d = (unsigned __int128)a[i] - b[i] - br;
      a[i] = (uint64_t)d;
      br = (d >> 64) & 1;
    }
    return br != 0;
  }
};

// G1Local / G2Local (fr.rs:12-16) as affine canonical coordinates; the identity is all-zero
struct G1 { uint64_t v[8] = {}; bool operator==(const G1& o) const { return !memcmp(v, o.v, sizeof v); } };
struct G2 { uint64_t v[16] = {}; bool operator==(const G2& o) const { return !memcmp(v, o.v, sizeof v); } };

// Proof<G1Local, G2Local> (mod.rs:124-128)
struct Proof {
  G1 a; G2 b; G1 c;
  bool operator==(const Proof& o) const { return a == o.a && b == o.b && c == o.c; }
};

// the DummyRep data model (circuit/dummy_rep.rs:7-13): per wire, the (root, value) pairs where the row is non-zero
struct RootRepresentation {
  std::vector<std::vector<std::pair<Fr, Fr>>> u, v, w;
  std::vector<Fr> roots;
  size_t input = 0;
};

class Context {
 public:
  explicit Context(int device = 0) { check(nullptr, zkb_ctx_create(&h_, device), "zkb_ctx_create"); }
  ~Context() { zkb_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  zkb_ctx* get() const { return h_; }
  void check(int rc, const char* what) const { check(h_, rc, what); }
  static void check(const zkb_ctx* h, int rc, const char* what) {
    if (rc != ZKB_OK) throw Error(rc, std::string(what) + ": " + zkb_last_error(h));
  }

 private:
  zkb_ctx* h_ = nullptr;
};

// QAP<CoefficientPoly<FrLocal>> (mod.rs:60-67), device resident
class QAP {
 public:
  // `impl From<RootRepresentation> for QAP` (fr.rs:140-173).  Roots omega^0 .. omega^(n-1) in order (n a power of
  // two >= 2) take the NTT path; any other pairwise-distinct roots (ASTParser's 1..=n, circuit/mod.rs:517) the
  // dense device path (n <= 32768).
  static QAP from(Context& ctx, const RootRepresentation& rep) {
    const size_t n = rep.roots.size(), m = rep.u.size();
    if (rep.v.size() != m || rep.w.size() != m) throw Error(ZKB_ERR_ARG, "QAP: u, v, w must have the same number of rows");  // fr.rs:157-158
    bool fast = n >= 2 && (n & (n - 1)) == 0;
    if (fast) {
      int log_n = 0;
      while (((size_t)1 << log_n) < n) log_n++;
      Fr w = omega(log_n), acc = Fr::one();
      for (size_t k = 0; k < n && fast; k++) { fast = rep.roots[k] == acc; acc = acc * w; }
    }
    std::vector<uint64_t> ptr[3], coeff[3], roots;
    std::vector<uint32_t> gate[3];
    const std::vector<std::vector<std::pair<Fr, Fr>>>* mats[3] = {&rep.u, &rep.v, &rep.w};
    for (int t = 0; t < 3; t++) {
      ptr[t].push_back(0);
      for (const auto& row : *mats[t]) {
        for (const auto& e : row) {
          size_t k = 0;
          while (k < n && rep.roots[k] != e.first) k++;
          if (k == n) throw Error(ZKB_ERR_ARG, "QAP: a row entry is not at one of the roots");
          gate[t].push_back((uint32_t)k);
          coeff[t].insert(coeff[t].end(), e.second.l.begin(), e.second.l.end());
        }
        ptr[t].push_back(gate[t].size());
      }
    }
    zkb_qap_host h{};
    h.n = n; h.m = m; h.n_input = rep.input;
    for (int t = 0; t < 3; t++) { h.row_ptr[t] = ptr[t].data(); h.gate[t] = gate[t].data(); h.coeff[t] = coeff[t].data(); }
    if (!fast) {
      for (const Fr& r : rep.roots) roots.insert(roots.end(), r.l.begin(), r.l.end());
      h.roots = roots.data();
    }
    QAP q(ctx);
    q.n_ = n; q.m_ = m; q.input_ = rep.input;
    ctx.check(zkb_qap_upload(ctx.get(), &h, &q.h_), "zkb_qap_upload");
    return q;
  }
  // 5^((r-1)/2^log_n): the primitive 2^log_n-th root of unity the fast domain is built on
  static Fr omega(int log_n) {
    std::array<uint64_t, 4> e = Fr::R;
    e[0] -= 1;
    for (int s = 0; s < log_n; s++) {  // e >>= 1
      for (int i = 0; i < 3; i++) e[i] = (e[i] >> 1) | (e[i + 1] << 63);
      e[3] >>= 1;
    }
    return Fr(5).pow(e);
  }
  QAP(QAP&& o) noexcept : ctx_(o.ctx_), h_(o.h_), n_(o.n_), m_(o.m_), input_(o.input_) { o.h_ = nullptr; }
  QAP(const QAP&) = delete;
  ~QAP() { if (h_) zkb_qap_free(ctx_->get(), h_); }
  size_t degree() const { return n_; }
  size_t rows() const { return m_; }
  size_t input() const { return input_; }
  const zkb_qap* get() const { return h_; }

 private:
  explicit QAP(Context& c) : ctx_(&c) {}
  Context* ctx_;
  zkb_qap* h_ = nullptr;
  size_t n_ = 0, m_ = 0, input_ = 0;
};

// (SigmaG1<G1Local>, SigmaG2<G2Local>) (mod.rs:105-121), device resident (window-expanded MSM tables)
class Sigma {
 public:
  Sigma(Context& c, zkb_crs* h) : ctx_(&c), h_(h) {}
  Sigma(Sigma&& o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
  Sigma(const Sigma&) = delete;
  ~Sigma() { if (h_) zkb_crs_free(ctx_->get(), h_); }
  const zkb_crs* get() const { return h_; }

 private:
  Context* ctx_;
  zkb_crs* h_;
};

namespace groth16 {

// setup with the five secrets (alpha, beta, gamma, delta, x) injected -- the seam parity tests use
inline Sigma setup_with(Context& ctx, const QAP& qap, const std::array<Fr, 5>& toxic) {
  uint64_t t[20];
  for (int i = 0; i < 5; i++) memcpy(t + 4 * i, toxic[i].l.data(), 32);
  zkb_crs* h = nullptr;
  ctx.check(zkb_setup(ctx.get(), qap.get(), t, 0, 1, &h), "zkb_setup");
  return Sigma(ctx, h);
}
// groth16::setup (mod.rs:134-197): the toxic waste is drawn with random_elem (mod.rs:139-145)
inline Sigma setup(Context& ctx, const QAP& qap) {
  return setup_with(ctx, qap, {Fr::random_elem(), Fr::random_elem(), Fr::random_elem(), Fr::random_elem(), Fr::random_elem()});
}

inline Proof prove_with_rs(Context& ctx, const QAP& qap, const Sigma& sigma, const std::vector<Fr>& weights, const Fr& r, const Fr& s) {
  // every zip in prove() ends at the shorter side (mod.rs:237..288): extra weights are ignored, missing ones are zero
  std::vector<uint64_t> w(4 * qap.rows(), 0);
  for (size_t i = 0; i < weights.size() && i < qap.rows(); i++) memcpy(&w[4 * i], weights[i].l.data(), 32);
  zkb_proof out;
  ctx.check(zkb_prove(ctx.get(), qap.get(), sigma.get(), w.data(), r.l.data(), s.l.data(), &out), "zkb_prove");
  Proof p;
  memcpy(p.a.v, out.a, sizeof out.a); memcpy(p.b.v, out.b, sizeof out.b); memcpy(p.c.v, out.c, sizeof out.c);
  return p;
}
// groth16::prove (mod.rs:213-296); r, s drawn with random_elem (mod.rs:231)
inline Proof prove(Context& ctx, const QAP& qap, const Sigma& sigma, const std::vector<Fr>& weights) {
  return prove_with_rs(ctx, qap, sigma, weights, Fr::random_elem(), Fr::random_elem());
}

// Throughput mode: one proof per (weights[i], rs[i]) against the same QAP / CRS, several in flight on the device
// (zkb_prove_batch).  Same results as prove_with_rs in a loop.
inline std::vector<Proof> prove_many(Context& ctx, const QAP& qap, const Sigma& sigma, const std::vector<std::vector<Fr>>& weights,
                                     const std::vector<std::pair<Fr, Fr>>& rs) {
  if (weights.size() != rs.size()) throw Error(ZKB_ERR_ARG, "prove_many: weights and (r, s) pairs differ in number");
  const size_t count = weights.size();
  std::vector<std::vector<uint64_t>> w(count, std::vector<uint64_t>(4 * qap.rows(), 0));
  std::vector<const uint64_t*> ptrs(count);
  std::vector<uint64_t> r(4 * count), s(4 * count);
  for (size_t k = 0; k < count; k++) {
    for (size_t i = 0; i < weights[k].size() && i < qap.rows(); i++) memcpy(&w[k][4 * i], weights[k][i].l.data(), 32);
    ptrs[k] = w[k].data();
    memcpy(&r[4 * k], rs[k].first.l.data(), 32);
    memcpy(&s[4 * k], rs[k].second.l.data(), 32);
  }
  std::vector<zkb_proof> out(count);
  if (count) ctx.check(zkb_prove_batch(ctx.get(), qap.get(), sigma.get(), ptrs.data(), 0, r.data(), s.data(), count, out.data()), "zkb_prove_batch");
  std::vector<Proof> ps(count);
  for (size_t k = 0; k < count; k++) {
    memcpy(ps[k].a.v, out[k].a, sizeof out[k].a); memcpy(ps[k].b.v, out[k].b, sizeof out[k].b); memcpy(ps[k].c.v, out[k].c, sizeof out[k].c);
  }
  return ps;
}

// groth16::verify (mod.rs:299-320); the CRS is borrowed (the reference consumes it)
inline bool verify(Context& ctx, const Sigma& sigma, const std::vector<Fr>& inputs, const Proof& proof) {
  std::vector<uint64_t> in(4 * inputs.size() + 4, 0);
  for (size_t i = 0; i < inputs.size(); i++) memcpy(&in[4 * i], inputs[i].l.data(), 32);
  zkb_proof pc;
  memcpy(pc.a, proof.a.v, sizeof pc.a); memcpy(pc.b, proof.b.v, sizeof pc.b); memcpy(pc.c, proof.c.v, sizeof pc.c);
  int ok = 0;
  ctx.check(zkb_verify(ctx.get(), sigma.get(), in.data(), inputs.size(), &pc, &ok), "zkb_verify");
  return ok == 1;
}

// `weights()` / `evaluate()` (circuit/mod.rs:529-656) on the device, the parse already done: `input_wires` are the
// positions of the program's `(in ...)` variables in the weight vector, in declaration order.  The plan levelises the
// circuit once; weights(values) evaluates it for one assignment of the inputs -> [1, every wire's value in wire order].
// program_order = true rejects what the reference's sequential walk rejects (a gate reading a wire a later gate assigns).
class WitnessPlan {
 public:
  WitnessPlan(Context& ctx, const QAP& qap, const std::vector<uint32_t>& input_wires, bool program_order = true)
      : ctx_(&ctx), m_(qap.rows()), n_in_(input_wires.size()) {
    ctx.check(zkb_witness_plan_create(ctx.get(), qap.get(), input_wires.data(), input_wires.size(),
                                      program_order ? ZKB_WITNESS_PROGRAM_ORDER : 0, &h_), "zkb_witness_plan_create");
  }
  WitnessPlan(WitnessPlan&& o) noexcept : ctx_(o.ctx_), h_(o.h_), m_(o.m_), n_in_(o.n_in_) { o.h_ = nullptr; }
  WitnessPlan(const WitnessPlan&) = delete;
  ~WitnessPlan() { if (h_) zkb_witness_plan_free(ctx_->get(), h_); }
  std::vector<Fr> weights(const std::vector<Fr>& values) const {
    std::vector<uint64_t> in(4 * values.size() + 4, 0), out(4 * m_);
    for (size_t i = 0; i < values.size(); i++) memcpy(&in[4 * i], values[i].l.data(), 32);
    ctx_->check(zkb_witness_generate(ctx_->get(), h_, in.data(), values.size(), 0, out.data(), 0), "zkb_witness_generate");
    std::vector<Fr> w(m_);
    for (size_t i = 0; i < m_; i++) memcpy(w[i].l.data(), &out[4 * i], 32);
    return w;
  }
  uint64_t levels() const { uint64_t l = 0; zkb_witness_plan_info(h_, nullptr, &l, nullptr, nullptr); return l; }
  const zkb_witness_plan* get() const { return h_; }

 private:
  Context* ctx_;
  zkb_witness_plan* h_ = nullptr;
  size_t m_, n_in_;
};

// One verdict per (inputs[i], proofs[i]) against the same CRS, every inputs[i] of the same length (zkb_verify_batch).
inline std::vector<bool> verify_many(Context& ctx, const Sigma& sigma, const std::vector<std::vector<Fr>>& inputs,
                                     const std::vector<Proof>& proofs) {
  if (inputs.size() != proofs.size()) throw Error(ZKB_ERR_ARG, "verify_many: inputs and proofs differ in number");
  const size_t count = proofs.size(), k = count ? inputs[0].size() : 0;
  std::vector<uint64_t> in(4 * k * count + 4, 0);
  std::vector<zkb_proof> pcs(count);
  for (size_t i = 0; i < count; i++) {
    if (inputs[i].size() != k) throw Error(ZKB_ERR_ARG, "verify_many: every proof needs the same number of inputs");
    for (size_t j = 0; j < k; j++) memcpy(&in[4 * (i * k + j)], inputs[i][j].l.data(), 32);
    memcpy(pcs[i].a, proofs[i].a.v, sizeof pcs[i].a); memcpy(pcs[i].b, proofs[i].b.v, sizeof pcs[i].b);
    memcpy(pcs[i].c, proofs[i].c.v, sizeof pcs[i].c);
  }
  std::vector<int> ok(count, 0);
  if (count) ctx.check(zkb_verify_batch(ctx.get(), sigma.get(), in.data(), k, pcs.data(), count, ok.data()), "zkb_verify_batch");
  std::vector<bool> res(count);
  for (size_t i = 0; i < count; i++) res[i] = ok[i] == 1;
  return res;
}

}  // namespace groth16
}  // namespace zkb200
